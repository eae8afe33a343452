// MRLA-base tail, bf16 activations.
#include "base_launch.cuh"
namespace mrla {
template int base_forward_t<__nv_bfloat16>(const MrlaBaseArgs&, cudaStream_t);
template int base_backward_t<__nv_bfloat16>(const MrlaBaseArgs&, cudaStream_t);
}  // namespace mrla
