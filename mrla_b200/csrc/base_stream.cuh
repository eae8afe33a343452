// MRLA-base, NHWC fast path: the four element-wise sweeps (F2 mix, F4 apply, B0 moments, B2 scatter) as
// vectorised streaming kernels.  None of them has a stencil, so no tiles / halos are needed: a thread owns
// 8 consecutive channels (one 16-byte vector of bf16, two of fp32) of a (sample, channel-block) slab and walks
// the pixels with PL-way pixel parallelism and 2x unrolling (independent 128-bit loads in flight); per-(b,c)
// sums are reduced across the pixel lanes through shared memory once per slab (deterministic, no atomics).
// Requires C % 8 == 0 and d % 8 == 0 (the 8 channels of a thread then belong to one head, so the softmax
// weights are per-thread scalars).  Everything else takes the generic kernels in base_kernels.cuh.
#pragma once
#include "base_kernels.cuh"

namespace mrla {

constexpr int kSV = 8;  // channels per thread

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[kSV]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[kSV]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ float rnd(float v) { return v; }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&o)[kSV]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[kSV]) {
    uint4 u;
    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float (&o)[kSV]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void st(__half* p, const float (&v)[kSV]) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
};

struct StreamShape {
  int B, C, HW;
  int t, d;
  int items;   // B * (C / (CL*8))
};

// thread map of a 256-thread CTA: CL channel lanes x PL pixel lanes, channel lane fastest (coalesced 16 B vectors)
template <int PL>
struct StreamMap {
  static constexpr int CL = 256 / PL;
  static constexpr int CBS = CL * kSV;
  int cl, pl;
  __device__ __forceinline__ StreamMap() : cl(threadIdx.x % CL), pl(threadIdx.x / CL) {}
};

// sum NV per-thread values over the PL pixel lanes; emit(channel_lane, i, sum) for i < NV
template <int PL, int NV, typename Emit>
__device__ __forceinline__ void reduce_pixels(const float (&v)[NV], float* sm, const StreamMap<PL>& m, Emit emit) {
  constexpr int CL = StreamMap<PL>::CL;
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[((size_t)m.pl * CL + m.cl) * NV + i] = v[i];
  __syncthreads();
  for (int idx = threadIdx.x; idx < CL * NV; idx += 256) {
    const int cl = idx / NV, i = idx - cl * NV;
    float s = 0.f;
#pragma unroll 4
    for (int p = 0; p < PL; ++p) s += sm[((size_t)p * CL + cl) * NV + i];
    emit(cl, i, s);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------ F2: S = Σ_j p_j V_j ; ΣS, ΣS²
template <typename T, int PL>
__global__ void __launch_bounds__(256) k_base_mix_nhwc(const T* __restrict__ v, T* __restrict__ sout,
                                                       const float* __restrict__ p, float* __restrict__ smom,
                                                       StreamShape s, int64_t bs_v, int64_t ts_v, int64_t bs_s) {
  extern __shared__ float smem[];
  const StreamMap<PL> m;
  constexpr int CBS = StreamMap<PL>::CBS;
  const int ncb = s.C / CBS;
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  for (int item = blockIdx.x; item < s.items; item += gridDim.x) {
    const int b = item / ncb, cb = item - b * ncb;
    const int c = cb * CBS + m.cl * kSV;
    const T* vb = v + (int64_t)b * bs_v + c;
    T* sb = sout + (int64_t)b * bs_s + c;
    const float* pb = p + ((int64_t)b * g + c / s.d) * s.t;
    float acc[2 * kSV];
#pragma unroll
    for (int i = 0; i < 2 * kSV; ++i) acc[i] = 0.f;
    for (int px = m.pl; px < s.HW; px += PL) {
      float sv[kSV];
#pragma unroll
      for (int i = 0; i < kSV; ++i) sv[i] = 0.f;
#pragma unroll 2
      for (int j = 0; j < s.t; ++j) {
        float vv[kSV];
        Vec8<T>::ld(vb + (int64_t)j * ts_v + (int64_t)px * s.C, vv);
        const float pj = pb[j];
#pragma unroll
        for (int i = 0; i < kSV; ++i) sv[i] = fmaf(pj, vv[i], sv[i]);
      }
      Vec8<T>::st(sb + (int64_t)px * s.C, sv);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float r = Vec8<T>::rnd(sv[i]);   // statistics of the values BatchNorm will read
        acc[i] += r;
        acc[kSV + i] = fmaf(r, r, acc[kSV + i]);
      }
    }
    reduce_pixels<PL, 2 * kSV>(acc, smem, m, [&](int cl, int i, float sum) {
      const int mi = i / kSV, ci = i - mi * kSV;
      smom[(int64_t)mi * BC + (int64_t)b * s.C + cb * CBS + cl * kSV + ci] = sum;
    });
  }
}

// ------------------------------------------------------------------------ F4: Y = res*X + m_b*act(cA*S + cD)
template <typename T, int PL>
__global__ void __launch_bounds__(256) k_base_apply_nhwc(const T* __restrict__ x, const T* __restrict__ sin,
                                                         T* __restrict__ y, const float* __restrict__ chan,
                                                         const float* __restrict__ drop_scale, StreamShape s,
                                                         int64_t bs_x, int64_t bs_s, int64_t bs_y, float res, int relu) {
  const StreamMap<PL> m;
  constexpr int CBS = StreamMap<PL>::CBS;
  const int ncb = s.C / CBS;
  for (int item = blockIdx.x; item < s.items; item += gridDim.x) {
    const int b = item / ncb, cb = item - b * ncb;
    const int c = cb * CBS + m.cl * kSV;
    float cA[kSV], cD[kSV];
#pragma unroll
    for (int i = 0; i < kSV; ++i) { cA[i] = chan[c + i]; cD[i] = chan[s.C + c + i]; }
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    const T* xb = x + (int64_t)b * bs_x + c;
    const T* sb = sin + (int64_t)b * bs_s + c;
    T* yb = y + (int64_t)b * bs_y + c;
#pragma unroll 2
    for (int px = m.pl; px < s.HW; px += PL) {
      float xv[kSV], sv[kSV], out[kSV];
      Vec8<T>::ld(xb + (int64_t)px * s.C, xv);
      Vec8<T>::ld(sb + (int64_t)px * s.C, sv);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        float z = fmaf(cA[i], sv[i], cD[i]);
        if (relu) z = fmaxf(z, 0.f);
        out[i] = fmaf(res, xv[i], mb * z);
      }
      Vec8<T>::st(yb + (int64_t)px * s.C, out);
    }
  }
}

// ------------------------------------------------------------------------ B0: ΣdZ, ΣdZ·S
template <typename T, int PL>
__global__ void __launch_bounds__(256) k_base_mom_bwd_nhwc(const T* __restrict__ dy, const T* __restrict__ sin,
                                                           const float* __restrict__ chan,
                                                           const float* __restrict__ drop_scale,
                                                           float* __restrict__ gmom, StreamShape s, int64_t bs_dy,
                                                           int64_t bs_s, int relu) {
  extern __shared__ float smem[];
  const StreamMap<PL> m;
  constexpr int CBS = StreamMap<PL>::CBS;
  const int ncb = s.C / CBS;
  const int64_t BC = (int64_t)s.B * s.C;
  for (int item = blockIdx.x; item < s.items; item += gridDim.x) {
    const int b = item / ncb, cb = item - b * ncb;
    const int c = cb * CBS + m.cl * kSV;
    float cA[kSV], cD[kSV];
#pragma unroll
    for (int i = 0; i < kSV; ++i) { cA[i] = chan[c + i]; cD[i] = chan[s.C + c + i]; }
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    const T* gb = dy + (int64_t)b * bs_dy + c;
    const T* sb = sin + (int64_t)b * bs_s + c;
    float acc[2 * kSV];
#pragma unroll
    for (int i = 0; i < 2 * kSV; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int px = m.pl; px < s.HW; px += PL) {
      float gv[kSV], sv[kSV];
      Vec8<T>::ld(gb + (int64_t)px * s.C, gv);
      Vec8<T>::ld(sb + (int64_t)px * s.C, sv);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float z = fmaf(cA[i], sv[i], cD[i]);
        const float dz = (relu && z <= 0.f) ? 0.f : mb * gv[i];
        acc[i] += dz;
        acc[kSV + i] = fmaf(dz, sv[i], acc[kSV + i]);
      }
    }
    reduce_pixels<PL, 2 * kSV>(acc, smem, m, [&](int cl, int i, float sum) {
      const int mi = i / kSV, ci = i - mi * kSV;
      gmom[(int64_t)mi * BC + (int64_t)b * s.C + cb * CBS + cl * kSV + ci] = sum;
    });
  }
}

// ------------------------------------------------------------------------ B2: dV_j (+)= p_j dS ; dpm[j] = Σ dS·V_j
// One pass handles cache slots [j0, j0+kBaseChunk).  The 8 channels of a thread belong to one head, so the
// head sum Σ_{c in h} Σ_hw dS V_j is accumulated in ONE register per slot; it is stored on the first of the 8
// channels (zeros on the other 7) so the [t,B,C] layout the attention-backward kernel reads stays unchanged.
template <typename T, int PL>
__global__ void __launch_bounds__(256) k_base_scatter_nhwc(const T* __restrict__ dy, const T* __restrict__ sin,
                                                           const T* __restrict__ v, T* __restrict__ dv,
                                                           const float* __restrict__ p, const float* __restrict__ chan,
                                                           const float* __restrict__ bchan,
                                                           const float* __restrict__ drop_scale,
                                                           float* __restrict__ dpm, StreamShape s, int j0,
                                                           int accumulate, int64_t bs_dy, int64_t bs_s, int64_t bs_v,
                                                           int64_t ts_v, int64_t bs_dv, int64_t ts_dv, int relu) {
  extern __shared__ float smem[];
  const StreamMap<PL> m;
  constexpr int CBS = StreamMap<PL>::CBS;
  const int ncb = s.C / CBS;
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const int nj = min(kBaseChunk, s.t - j0);
  for (int item = blockIdx.x; item < s.items; item += gridDim.x) {
    const int b = item / ncb, cb = item - b * ncb;
    const int c = cb * CBS + m.cl * kSV;
    float cA[kSV], cD[kSV], e1[kSV], e0[kSV], e2[kSV];
#pragma unroll
    for (int i = 0; i < kSV; ++i) {
      cA[i] = chan[c + i]; cD[i] = chan[s.C + c + i];
      e1[i] = bchan[c + i]; e0[i] = bchan[s.C + c + i]; e2[i] = bchan[2 * s.C + c + i];
    }
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    const float* pb = p + ((int64_t)b * g + c / s.d) * s.t + j0;
    float pj[kBaseChunk], acc[kBaseChunk];
#pragma unroll
    for (int j = 0; j < kBaseChunk; ++j) { pj[j] = j < nj ? pb[j] : 0.f; acc[j] = 0.f; }
    const T* gb = dy + (int64_t)b * bs_dy + c;
    const T* sb = sin + (int64_t)b * bs_s + c;
    const T* vb = v + (int64_t)b * bs_v + (int64_t)j0 * ts_v + c;
    T* dvb = dv + (int64_t)b * bs_dv + (int64_t)j0 * ts_dv + c;
    for (int px = m.pl; px < s.HW; px += PL) {
      float gv[kSV], sv[kSV], ds[kSV];
      Vec8<T>::ld(gb + (int64_t)px * s.C, gv);
      Vec8<T>::ld(sb + (int64_t)px * s.C, sv);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float z = fmaf(cA[i], sv[i], cD[i]);
        const float dz = (relu && z <= 0.f) ? 0.f : mb * gv[i];
        ds[i] = fmaf(e1[i], dz, fmaf(e2[i], sv[i], e0[i]));
      }
#pragma unroll
      for (int j = 0; j < kBaseChunk; ++j) {
        if (j < nj) {
          float vv[kSV], dvv[kSV];
          Vec8<T>::ld(vb + (int64_t)j * ts_v + (int64_t)px * s.C, vv);
          T* dst = dvb + (int64_t)j * ts_dv + (int64_t)px * s.C;
          if (accumulate) Vec8<T>::ld(dst, dvv);
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < kSV; ++i) {
            dot = fmaf(ds[i], vv[i], dot);
            dvv[i] = accumulate ? fmaf(pj[j], ds[i], dvv[i]) : pj[j] * ds[i];
          }
          acc[j] += dot;
          Vec8<T>::st(dst, dvv);
        }
      }
    }
    reduce_pixels<PL, kBaseChunk>(acc, smem, m, [&](int cl, int j, float sum) {
      if (j < nj) {
        float* dst = dpm + (int64_t)(j0 + j) * BC + (int64_t)b * s.C + cb * CBS + cl * kSV;
        *reinterpret_cast<float4*>(dst) = make_float4(sum, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    });
  }
}

}  // namespace mrla
