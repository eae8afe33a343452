// MRLA-light tail, fp32 activations.
#include "light_launch.cuh"
namespace mrla {
template int light_forward_t<float>(const MrlaLightArgs&, cudaStream_t);
template int light_backward_t<float>(const MrlaLightArgs&, cudaStream_t);
}  // namespace mrla
