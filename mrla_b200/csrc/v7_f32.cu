// v7 sweeps of the MRLA-light tail, float activations (explicit instantiations; light_launch.cuh declares them extern).
#include "light_v7_launch.cuh"
namespace mrla {
template int v7_launch_fwd<float, 0>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_fwd<float, 1>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_fwd<float, 2>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t, float*, int, int);
template int v7_launch_bwd<float>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, bool, const void*, int64_t, float*, float*, float*);
}  // namespace mrla
