// MRLA-light tail, bf16 activations (fp32 accumulate).
#include "light_launch.cuh"
namespace mrla {
template int light_forward_t<__nv_bfloat16>(const MrlaLightArgs&, cudaStream_t);
template int light_backward_t<__nv_bfloat16>(const MrlaLightArgs&, cudaStream_t);
}  // namespace mrla
