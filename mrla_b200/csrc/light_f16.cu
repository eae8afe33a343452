// MRLA-light tail, fp16 activations (fp32 accumulate) — DeiT's autocast dtype (deit/engine.py:37).
#include "light_launch.cuh"
namespace mrla {
template int light_forward_t<__half>(const MrlaLightArgs&, cudaStream_t);
template int light_backward_t<__half>(const MrlaLightArgs&, cudaStream_t);
}  // namespace mrla
