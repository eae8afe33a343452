// MRLA-base tail, f32 activations.
#include "base_launch.cuh"
namespace mrla {
template int base_forward_t<float>(const MrlaBaseArgs&, cudaStream_t);
template int base_backward_t<float>(const MrlaBaseArgs&, cudaStream_t);
}  // namespace mrla
