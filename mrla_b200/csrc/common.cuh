// Shared device helpers for the MRLA sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <nvtx3/nvToolsExt.h>

namespace mrla {

// NVTX range around a C-ABI entry point (SURVEY.md section 5 "tracing"): nsys / ncu timelines show one named range per
// call of the library with the kernels it enqueued inside.  Header-only NVTX v3: a no-op when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// ------------------------------------------------------------------ dtype conversion
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// CV consecutive elements moved as one aligned vector access (2..16 bytes).
template <typename T, int CV> struct alignas(sizeof(T) * CV) Pack { T v[CV]; };

template <typename T, int CV>
__device__ __forceinline__ void ld_vec(const T* __restrict__ p, float (&out)[CV]) {
  Pack<T, CV> pk = *reinterpret_cast<const Pack<T, CV>*>(p);
#pragma unroll
  for (int i = 0; i < CV; ++i) out[i] = to_f<T>(pk.v[i]);
}
template <typename T, int CV>
__device__ __forceinline__ void ld_vec_pred(const T* __restrict__ p, bool pred, float (&out)[CV]) {
  if (pred) {
    ld_vec<T, CV>(p, out);
  } else {
#pragma unroll
    for (int i = 0; i < CV; ++i) out[i] = 0.f;
  }
}
template <typename T, int CV>
__device__ __forceinline__ void st_vec(T* __restrict__ p, const float (&in)[CV]) {
  Pack<T, CV> pk;
#pragma unroll
  for (int i = 0; i < CV; ++i) pk.v[i] = from_f<T>(in[i]);
  *reinterpret_cast<Pack<T, CV>*>(p) = pk;
}
// fp32 side arrays ([B,C] coefficients): CV consecutive floats
template <int CV>
__device__ __forceinline__ void ld_f32(const float* __restrict__ p, float (&out)[CV]) {
  Pack<float, CV> pk = *reinterpret_cast<const Pack<float, CV>*>(p);
#pragma unroll
  for (int i = 0; i < CV; ++i) out[i] = pk.v[i];
}

// ------------------------------------------------------------------ activation on V
// ACT = 0: identity.  ACT = 1: exact (erf) GELU, as nn.GELU() in deit/deit_mrla_light.py:153.
template <int ACT> __device__ __forceinline__ float act_fwd(float u) {
  if (ACT == 1) return 0.5f * u * (1.f + erff(u * 0.70710678118654752440f));
  return u;
}
template <int ACT> __device__ __forceinline__ float act_grad(float u) {
  if (ACT == 1) {
    const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * u * u);
    return cdf + u * pdf;
  }
  return 1.f;
}

// ------------------------------------------------------------------ warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// How the threads of a "marching" CTA are laid out (see light_sweeps.cuh).
//   NCHW: tid = slot*W + w   (slot = plane within the CTA),  channel vector CV = 1
//   NHWC: tid = w*LP + slot  (slot = channel-vector lane),   CV channels per thread
struct March {
  int c;        // first channel owned by this thread
  int w;        // image column owned by this thread
  int slot;     // reduction slot (plane / channel-vector lane)
  int valid;    // thread maps to real data
  int tW;       // tid distance between columns w and w+1 of the same slot
  int tS;       // tid distance between slots
  int nslots;   // slots in this CTA
  int64_t sC, sH, sW;  // element strides of channel / row / column inside one sample
};

template <int LAYOUT>
__device__ __forceinline__ March make_march(int C, int H, int W, int CV, int slots_per_cta) {
  March m;
  const int tid = threadIdx.x;
  m.nslots = slots_per_cta;
  if (LAYOUT == 0) {  // NCHW
    m.slot = tid / W;
    m.w = tid - m.slot * W;
    m.c = blockIdx.x * slots_per_cta + m.slot;
    m.tW = 1;
    m.tS = W;
    m.sC = (int64_t)H * W;
    m.sH = W;
    m.sW = 1;
    m.valid = (m.slot < slots_per_cta) && (m.c < C);
  } else {  // NHWC
    m.w = tid / slots_per_cta;
    m.slot = tid - m.w * slots_per_cta;
    m.c = (blockIdx.x * slots_per_cta + m.slot) * CV;
    m.tW = slots_per_cta;
    m.tS = 1;
    m.sC = 1;
    m.sH = (int64_t)W * C;
    m.sW = C;
    m.valid = (m.w < W) && (m.c < C);
  }
  return m;
}

// Deterministic reduction over the W column-threads of every slot.
// Each thread contributes NV floats; `emit(slot, i, sum)` is called once per (slot, i).
template <int NV, typename Emit>
__device__ __forceinline__ void reduce_over_columns(const float (&v)[NV], float* sm, const March& m, int W, Emit emit) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[tid * NV + i] = v[i];
  __syncthreads();
  const int total = m.nslots * NV;
  for (int idx = tid; idx < total; idx += blockDim.x) {
    const int slot = idx / NV;
    const int i = idx - slot * NV;
    float s = 0.f;
    const float* p = sm + (size_t)(slot * m.tS) * NV + i;
    const int step = m.tW * NV;
    for (int w = 0; w < W; ++w) s += p[(size_t)w * step];
    emit(slot, i, s);
  }
  __syncthreads();
}


// where the last negative return code of this thread was raised ("file:line"): mrla_last_error_site()
extern thread_local const char* g_err_site;
#define MRLA_STR2_(x) #x
#define MRLA_STR_(x) MRLA_STR2_(x)
#define MRLA_FAIL(code) (::mrla::g_err_site = __FILE__ ":" MRLA_STR_(__LINE__), (code))

}  // namespace mrla
