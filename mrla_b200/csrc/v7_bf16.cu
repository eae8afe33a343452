// v7 sweeps of the MRLA-light tail, __nv_bfloat16 activations (explicit instantiations; light_launch.cuh declares them extern).
#include "light_v7_launch.cuh"
namespace mrla {
template int v7_launch_fwd<__nv_bfloat16, 0>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_fwd<__nv_bfloat16, 1>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_fwd<__nv_bfloat16, 2>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_bwd<__nv_bfloat16>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, bool, const void*, int64_t, float*, float*, float*);
}  // namespace mrla
