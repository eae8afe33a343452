// MRLA-base tail, f16 activations.
#include "base_launch.cuh"
namespace mrla {
template int base_forward_t<__half>(const MrlaBaseArgs&, cudaStream_t);
template int base_backward_t<__half>(const MrlaBaseArgs&, cudaStream_t);
}  // namespace mrla
