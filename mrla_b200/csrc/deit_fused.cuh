// Fused DeiT MRLA-light module (token layout): ONE kernel per direction, one CTA per sample.
//
//   xn = LN_x(x) ; on = LN_o(o)                                     (deit/deit_mrla_light.py:195-196, eps 1e-6)
//   img = xn[:, 1:, :] viewed [B, C, S, S] ; y = mean_hw img ; Q, K = xcorr1d(y, wq / wk) ; a[b,h] = sigmoid(Q.K / sqrt(d))
//   out[:, 0] = xn[:, 0] ;  out[:, 1 + t] = a * GELU(dwconv3x3(img))[t] + lambda * on[:, 1 + t]      (:157-180, 199-207)
//
// Nothing couples two samples (no BatchNorm on this branch), and one sample is 197 x 192 values = 76 KB in bf16, so the
// whole module — both LayerNorms, the cls pass-through, GAP, the ECA gate, the depthwise conv with GELU, the lambda
// recurrence — runs out of one CTA's shared memory: x and o are read once, out is written once, and there is no [B,C]
// side tensor round trip at all.  The reference issues ~25 ATen launches for this (two LayerNorms, slicing / reshape /
// permute copies, conv, GELU, conv1d x2, einsum, sigmoid, mul, add, cat) on a tensor that fits in L2, i.e. it is pure
// launch latency; round 1 of this repo still ran both LayerNorms and the cat as library calls.
// Backward is the same shape: recompute xn, one pass for the gate gradient (needs sum dS*V), one pass for dU -> conv^T
// -> d(xn) (scatter form, T halo recomputed across the two column groups), then both LayerNorm backwards token by token.
// Parameter gradients leave as per-sample partials [B, *] and are summed by k_deit_reduce.
#pragma once
#include "light_v7.cuh"

namespace mrla {

struct DeitParams {
  int B, n, C, S, d, k;        // tokens n = S*S + 1, heads g = C/d, ECA taps k
  int NQ;                      // column groups of 7
  float eps;
  const void* x; const void* o; void* out;
  const float* gx; const float* bx; const float* go; const float* bo;   // LayerNorm weights / biases [C]
  const float* wq; const float* wk; const float* wv; const float* lam;
  float* stats_x; float* stats_o;     // [B, n, 2] (mean, rstd)   saved
  float* gate;                        // [B, g]                   saved
  // backward
  const void* dout; void* dx; void* dox;
  float* part;                        // [B, PW] per-sample parameter-gradient partials
  int PW;                             // = 14*C + 2*k : dWv[9C] | dlam[C] | dgx | dbx | dgo | dbo [C each] | dwq[k] | dwk[k]
};

constexpr int kDeitMaxCJ = 6;   // C / 64 <= 6  (C <= 384)

template <typename T> __device__ __forceinline__ void sts_pair_t(uint32_t saddr, float2 v) { sts_raw(saddr, pack_pair<T>(v)); }

template <int N>
__device__ __forceinline__ float2 conv9n(const float2 (&top)[N], const float2 (&mid)[N], const float2 (&bot)[N],
                                         const float2 (&w9)[9], int j) {
  float2 s = fmul2(w9[0], top[j]);
  s = ffma2(w9[1], top[j + 1], s); s = ffma2(w9[2], top[j + 2], s);
  s = ffma2(w9[3], mid[j], s); s = ffma2(w9[4], mid[j + 1], s); s = ffma2(w9[5], mid[j + 2], s);
  s = ffma2(w9[6], bot[j], s); s = ffma2(w9[7], bot[j + 1], s); s = ffma2(w9[8], bot[j + 2], s);
  return s;
}

// LayerNorm statistics of one token held as CJ pairs per lane
__device__ __forceinline__ void ln_stats(const float2 (&v)[kDeitMaxCJ], int CJ, int C, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j)
    if (j < CJ) s += v[j].x + v[j].y;
  s = warp_sum(s);
  mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j)
    if (j < CJ) {
      const float a = v[j].x - mean, b = v[j].y - mean;
      q = fmaf(a, a, fmaf(b, b, q));
    }
  q = warp_sum(q);
  rstd = rsqrtf(q / (float)C + eps);
}

// shared memory: xn tile [n][C] (T) | ys[C] | qk[C] | kq[2][C] | gate[g<=64] | sx[n] float2 | so[n] float2 | red scratch
// MAXT = 192: two CTAs per SM (all 256 samples of a DeiT-tiny batch are resident at once and the memory-latency-bound
// LayerNorm phases of one CTA overlap the arithmetic phases of the other); MAXT = 384: wider models, one CTA per SM
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 192 ? 2 : 1) k_deit_light_fwd(const DeitParams P) {
  constexpr int ES = sizeof(T);
  constexpr int K = kV7;
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int C = P.C, n = P.n, S = P.S, NP = C / 2, CJ = C / 64, g = C / P.d;
  unsigned char* xn = smem;
  float* ys = reinterpret_cast<float*>(smem + (size_t)n * C * ES);
  float* qk = ys + C;
  float* gate = qk + C;
  float2* so = reinterpret_cast<float2*>(gate + 64);
  const uint32_t xn_s = smem_u32(xn);
  const T* xb = static_cast<const T*>(P.x) + (int64_t)b * n * C;
  const T* ob = static_cast<const T*>(P.o) + (int64_t)b * n * C;
  T* outb = static_cast<T*>(P.out) + (int64_t)b * n * C;

  // ---- phase 1: LayerNorm of every token of x (normalised values -> shared memory) and statistics of o.
  // A warp owns tokens warp, warp+nw, ...; the loads of two tokens (x and o) are issued before any reduction so that
  // eight 128-byte requests per lane are in flight, and the GAP sums of the image tokens are taken from the registers.
  float* red = reinterpret_cast<float*>(so + n);   // [nw][C] GAP partials
  float2 gsum[kDeitMaxCJ];
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j) gsum[j] = f2(0.f, 0.f);
  for (int t0 = warp; t0 < n; t0 += 2 * nw) {
    float2 vx[2][kDeitMaxCJ], vo[2][kDeitMaxCJ];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * nw;
#pragma unroll
      for (int j = 0; j < kDeitMaxCJ; ++j)
        if (j < CJ && t < n) {
          vx[u][j] = ldg_pair<T>(xb + (int64_t)t * C + 2 * (j * 32 + lane));
          vo[u][j] = ldg_pair<T>(ob + (int64_t)t * C + 2 * (j * 32 + lane));
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * nw;
      if (t < n) {
        float mean, rstd, mo, ro;
        ln_stats(vx[u], CJ, C, P.eps, mean, rstd);
        ln_stats(vo[u], CJ, C, P.eps, mo, ro);
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ) {
            const int c = 2 * (j * 32 + lane);
            const float2 gm = *reinterpret_cast<const float2*>(P.gx + c), bt = *reinterpret_cast<const float2*>(P.bx + c);
            const float2 y = f2((vx[u][j].x - mean) * rstd * gm.x + bt.x, (vx[u][j].y - mean) * rstd * gm.y + bt.y);
            const typename RawPair<T>::type yr = pack_pair<T>(y);
            sts_raw(xn_s + (uint32_t)(t * C + c) * ES, yr);
            if (t == 0) {
              stg_raw<T>(outb + c, yr);   // cls row of the output is LN_x(x)[:, 0] (deit_mrla_light.py:207)
            } else {
              const float2 ys_ = unpack_pair<T>(yr);   // the GAP averages the STORED (rounded) values, like the reference
              gsum[j] = f2(gsum[j].x + ys_.x, gsum[j].y + ys_.y);
            }
          }
        if (lane == 0) {
          so[t] = f2(mo, ro);
          *reinterpret_cast<float2*>(P.stats_x + ((int64_t)b * n + t) * 2) = f2(mean, rstd);
          *reinterpret_cast<float2*>(P.stats_o + ((int64_t)b * n + t) * 2) = f2(mo, ro);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j)
    if (j < CJ) *reinterpret_cast<float2*>(red + (size_t)warp * C + 2 * (j * 32 + lane)) = gsum[j];
  __syncthreads();
  // ---- phase 2: GAP over the S*S image tokens, ECA conv1d pair over the channel axis, per-head sigmoid gate
  for (int c = tid; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[(size_t)w * C + c];
    ys[c] = s / (float)(n - 1);
  }
  __syncthreads();
  {
    const int pad = (P.k - 1) / 2;
    for (int c = tid; c < C; c += blockDim.x) {
      float q = 0.f, kk = 0.f;
      for (int j = 0; j < P.k; ++j) {
        const int cc = c + j - pad;
        const float yv = (cc >= 0 && cc < C) ? ys[cc] : 0.f;
        q = fmaf(P.wq[j], yv, q);
        kk = fmaf(P.wk[j], yv, kk);
      }
      qk[c] = q * kk;
    }
  }
  __syncthreads();
  for (int h = tid; h < g; h += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < P.d; ++i) acc += qk[h * P.d + i];
    const float a = 1.f / (1.f + __expf(-acc * rsqrtf((float)P.d)));
    gate[h] = a;
    P.gate[(int64_t)b * g + h] = a;
  }
  __syncthreads();
  // ---- phase 3: thread = (channel pair p, column group q) marches down the S rows of the token image
  const int p = tid % NP, q = tid / NP;
  if (q >= P.NQ) return;
  const int c = 2 * p;
  float2 w9[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w9[i] = f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]);
  const float2 a2 = f2(gate[c / P.d], gate[(c + 1) / P.d]);
  const float2 lm = *reinterpret_cast<const float2*>(P.lam + c);
  const float2 go2 = *reinterpret_cast<const float2*>(P.go + c), bo2 = *reinterpret_cast<const float2*>(P.bo + c);
  float2 win[3][K + 2];
#pragma unroll
  for (int j = 0; j < K + 2; ++j) { win[0][j] = f2(0.f, 0.f); win[1][j] = f2(0.f, 0.f); win[2][j] = f2(0.f, 0.f); }
  auto load_row = [&](int h, float2 (&dst)[K + 2]) {
#pragma unroll
    for (int j = 0; j < K + 2; ++j) {
      const int w = q * K - 1 + j;
      dst[j] = (h >= 0 && h < S && w >= 0 && w < S) ? lds_pair<T>(xn_s + (uint32_t)((1 + h * S + w) * C + c) * ES) : f2(0.f, 0.f);
    }
  };
  load_row(0, win[1]);
  float2 onext[K];
  auto load_o = [&](int h, float2 (&dst)[K]) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int w = q * K + j;
      dst[j] = (h < S && w < S) ? ldg_pair<T>(ob + (int64_t)(1 + h * S + w) * C + c) : f2(0.f, 0.f);
    }
  };
  load_o(0, onext);
  for (int h = 0; h < S; ++h) {
    load_row(h + 1, win[2]);
    float2 ocur[K];
#pragma unroll
    for (int j = 0; j < K; ++j) ocur[j] = onext[j];
    load_o(h + 1, onext);   // in flight while this row is computed
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int w = q * K + j;
      if (w < S) {
        const float2 u = conv9n<K + 2>(win[0], win[1], win[2], w9, j);
        const float2 v = f2(act_fwd<1>(u.x), act_fwd<1>(u.y));
        const int t = 1 + h * S + w;
        const float2 ov = ocur[j];
        const float2 st = so[t];
        const float2 on = f2((ov.x - st.x) * st.y * go2.x + bo2.x, (ov.y - st.x) * st.y * go2.y + bo2.y);
        stg_pair<T>(outb + (int64_t)t * C + c, f2(fmaf(a2.x, v.x, lm.x * on.x), fmaf(a2.y, v.y, lm.y * on.y)));
      }
    }
#pragma unroll
    for (int j = 0; j < K + 2; ++j) { win[0][j] = win[1][j]; win[1][j] = win[2][j]; }
  }
}

// =====================================================================================================
// backward.  d_out [B,n,C] -> dx, do [B,n,C] and per-sample parameter partials.
// =====================================================================================================
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 192 ? 2 : 1) k_deit_light_bwd(const DeitParams P) {
  constexpr int ES = sizeof(T);
  constexpr int K = kV7;
  constexpr int KT = K + 2, KX = K + 4;
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int C = P.C, n = P.n, S = P.S, NP = C / 2, CJ = C / 64, g = C / P.d, k = P.k, pad = (P.k - 1) / 2;
  unsigned char* xn = smem;                                             // [n][C] T : xn, later overwritten by d(xn)
  float* ys = reinterpret_cast<float*>(smem + (size_t)n * C * ES);      // [C] GAP
  float* qv = ys + C;                                                   // [C] Q
  float* kv = qv + C;                                                   // [C] K
  float* da = kv + C;                                                   // [C] per-channel sum dS*V, later dQ
  float* dk = da + C;                                                   // [C] dK
  float* gy = dk + C;                                                   // [C] GAP gradient / (n-1)
  float* gate = gy + C;                                                 // [64]
  float* dlog = gate + 64;                                              // [64]
  float* red = dlog + 64;                                               // [NQ][10][C] scratch for cross-group sums
  const uint32_t xn_s = smem_u32(xn);
  const T* xb = static_cast<const T*>(P.x) + (int64_t)b * n * C;
  const T* ob = static_cast<const T*>(P.o) + (int64_t)b * n * C;
  const T* gb = static_cast<const T*>(P.dout) + (int64_t)b * n * C;
  T* dxb = static_cast<T*>(P.dx) + (int64_t)b * n * C;
  T* dob = static_cast<T*>(P.dox) + (int64_t)b * n * C;
  const float* sxg = P.stats_x + (int64_t)b * n * 2;
  const float* sog = P.stats_o + (int64_t)b * n * 2;
  float* part = P.part + (int64_t)b * P.PW;

  // ---- phase 1: recompute xn = LN_x(x) from the saved statistics (two tokens in flight per warp), GAP from registers
  {
    float2 gsum[kDeitMaxCJ];
#pragma unroll
    for (int j = 0; j < kDeitMaxCJ; ++j) gsum[j] = f2(0.f, 0.f);
    for (int t0 = warp; t0 < n; t0 += 2 * nw) {
      float2 vx[2][kDeitMaxCJ];
      float2 st[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * nw;
        st[u] = (t < n) ? *reinterpret_cast<const float2*>(sxg + 2 * t) : f2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ && t < n) vx[u][j] = ldg_pair<T>(xb + (int64_t)t * C + 2 * (j * 32 + lane));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * nw;
        if (t < n) {
#pragma unroll
          for (int j = 0; j < kDeitMaxCJ; ++j)
            if (j < CJ) {
              const int c = 2 * (j * 32 + lane);
              const float2 gm = *reinterpret_cast<const float2*>(P.gx + c), bt = *reinterpret_cast<const float2*>(P.bx + c);
              const typename RawPair<T>::type yr = pack_pair<T>(f2((vx[u][j].x - st[u].x) * st[u].y * gm.x + bt.x,
                                                                   (vx[u][j].y - st[u].x) * st[u].y * gm.y + bt.y));
              sts_raw(xn_s + (uint32_t)(t * C + c) * ES, yr);
              if (t >= 1) {
                const float2 ys_ = unpack_pair<T>(yr);
                gsum[j] = f2(gsum[j].x + ys_.x, gsum[j].y + ys_.y);
              }
            }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kDeitMaxCJ; ++j)
      if (j < CJ) *reinterpret_cast<float2*>(red + (size_t)warp * C + 2 * (j * 32 + lane)) = gsum[j];
  }
  if (tid < g) gate[tid] = P.gate[(int64_t)b * g + tid];
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[(size_t)w * C + c];
    ys[c] = s / (float)(n - 1);
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float q = 0.f, kk = 0.f;
    for (int j = 0; j < k; ++j) {
      const int cc = c + j - pad;
      const float yv = (cc >= 0 && cc < C) ? ys[cc] : 0.f;
      q = fmaf(P.wq[j], yv, q);
      kk = fmaf(P.wk[j], yv, kk);
    }
    qv[c] = q;
    kv[c] = kk;
  }
  __syncthreads();   // the GAP partials in `red` have been consumed: phase 2 reuses the scratch
  // ---- phase 2: sum_t dS*V per channel (V recomputed) -> gate gradient -> GAP gradient
  const int p = tid % NP, q = tid / NP;
  const bool worker = q < P.NQ;
  const int c = 2 * p;
  float2 w9[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w9[i] = worker ? f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]) : f2(0.f, 0.f);
  auto xrow = [&](int h, int w) -> float2 {
    return (h >= 0 && h < S && w >= 0 && w < S) ? lds_pair<T>(xn_s + (uint32_t)((1 + h * S + w) * C + c) * ES) : f2(0.f, 0.f);
  };
  auto grow = [&](int h, int w) -> float2 {   // dS = d_out of an image token
    return (h >= 0 && h < S && w >= 0 && w < S) ? ldg_pair<T>(gb + (int64_t)(1 + h * S + w) * C + c) : f2(0.f, 0.f);
  };
  if (worker) {
    float2 acc = f2(0.f, 0.f);
    float2 win[3][K + 2];
#pragma unroll
    for (int j = 0; j < K + 2; ++j) { win[0][j] = f2(0.f, 0.f); win[1][j] = xrow(0, q * K - 1 + j); }
    float2 gnext[K];
#pragma unroll
    for (int j = 0; j < K; ++j) gnext[j] = grow(0, q * K + j);
    for (int h = 0; h < S; ++h) {
#pragma unroll
      for (int j = 0; j < K + 2; ++j) win[2][j] = xrow(h + 1, q * K - 1 + j);
      float2 gcur[K];
#pragma unroll
      for (int j = 0; j < K; ++j) { gcur[j] = gnext[j]; gnext[j] = grow(h + 1, q * K + j); }   // next row in flight
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int w = q * K + j;
        if (w < S) {
          const float2 u = conv9n<K + 2>(win[0], win[1], win[2], w9, j);
          const float2 ds = gcur[j];
          acc.x = fmaf(ds.x, act_fwd<1>(u.x), acc.x);
          acc.y = fmaf(ds.y, act_fwd<1>(u.y), acc.y);
        }
      }
#pragma unroll
      for (int j = 0; j < K + 2; ++j) { win[0][j] = win[1][j]; win[1][j] = win[2][j]; }
    }
    red[(q * 10 + 0) * C + c] = acc.x;
    red[(q * 10 + 0) * C + c + 1] = acc.y;
  }
  __syncthreads();
  for (int cc = tid; cc < C; cc += blockDim.x) {
    float s = 0.f;
    for (int qq = 0; qq < P.NQ; ++qq) s += red[(qq * 10 + 0) * C + cc];
    da[cc] = s;
  }
  __syncthreads();
  for (int h = tid; h < g; h += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < P.d; ++i) s += da[h * P.d + i];
    const float a = gate[h];
    dlog[h] = s * a * (1.f - a) * rsqrtf((float)P.d);
  }
  __syncthreads();
  for (int cc = tid; cc < C; cc += blockDim.x) {
    const float dl = dlog[cc / P.d];
    da[cc] = dl * kv[cc];   // dQ
    dk[cc] = dl * qv[cc];   // dK
  }
  __syncthreads();
  for (int cc = tid; cc < C; cc += blockDim.x) {
    float s = 0.f;   // dy[c] = sum_j wq[j]*dQ[c-j+pad] + wk[j]*dK[c-j+pad]
    for (int j = 0; j < k; ++j) {
      const int ci = cc - j + pad;
      if (ci >= 0 && ci < C) s = fmaf(P.wq[j], da[ci], fmaf(P.wk[j], dk[ci], s));
    }
    gy[cc] = s / (float)(n - 1);
  }
  if (tid < 2 * k) {   // dwq[j] = sum_c dQ[c]*y[c+j-pad] ; dwk likewise
    const int j = tid % k;
    const float* src = tid < k ? da : dk;
    float s = 0.f;
    for (int cc = 0; cc < C; ++cc) {
      const int ci = cc + j - pad;
      if (ci >= 0 && ci < C) s = fmaf(src[cc], ys[ci], s);
    }
    part[14 * C + tid] = s;
  }
  __syncthreads();
  // ---- phase 3: dU = a*dS*gelu'(U) on K+2 columns (halo recomputed), conv^T in scatter form, dWv, d(xn) -> shared
  // (d(xn) row h-1 replaces xn row h-1 once nobody needs that xn row any more: one CTA barrier per image row)
  float2 dwv[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) dwv[i] = f2(0.f, 0.f);
  float2 dlam = f2(0.f, 0.f);
  {
    const float2 a2 = worker ? f2(gate[c / P.d], gate[(c + 1) / P.d]) : f2(0.f, 0.f);
    const float2 gy2 = worker ? f2(gy[c], gy[c + 1]) : f2(0.f, 0.f);
    float2 xw[3][KX];
    float2 acc[3][K];
#pragma unroll
    for (int j = 0; j < KX; ++j) { xw[0][j] = f2(0.f, 0.f); xw[1][j] = f2(0.f, 0.f); xw[2][j] = f2(0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < K; ++j) { acc[0][j] = f2(0.f, 0.f); acc[1][j] = f2(0.f, 0.f); acc[2][j] = f2(0.f, 0.f); }
    if (worker) {
#pragma unroll
      for (int j = 0; j < KX; ++j) xw[1][j] = xrow(0, q * K - 2 + j);
    }
    // step s: fetch x row s; dU row s-1; scatter; dX row s-2 complete
    float2 gnext[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) gnext[j] = worker ? grow(0, q * K - 1 + j) : f2(0.f, 0.f);
    for (int s = 1; s <= S + 1; ++s) {
      float2 done[K];
      if (worker) {
#pragma unroll
        for (int j = 0; j < KX; ++j) xw[2][j] = xrow(s, q * K - 2 + j);
        float2 tt[KT], gcur[KT];
        const int t = s - 1;
#pragma unroll
        for (int j = 0; j < KT; ++j) { gcur[j] = gnext[j]; gnext[j] = grow(s, q * K - 1 + j); }   // row t+1 in flight
#pragma unroll
        for (int j = 0; j < KT; ++j) {
          const int w = q * K - 1 + j;
          tt[j] = f2(0.f, 0.f);
          if (t < S && w >= 0 && w < S) {
            float2 u = fmul2(w9[0], xw[0][j]);
            u = ffma2(w9[1], xw[0][j + 1], u); u = ffma2(w9[2], xw[0][j + 2], u);
            u = ffma2(w9[3], xw[1][j], u); u = ffma2(w9[4], xw[1][j + 1], u); u = ffma2(w9[5], xw[1][j + 2], u);
            u = ffma2(w9[6], xw[2][j], u); u = ffma2(w9[7], xw[2][j + 1], u); u = ffma2(w9[8], xw[2][j + 2], u);
            const float2 ds = gcur[j];
            tt[j] = f2(a2.x * ds.x * act_grad<1>(u.x), a2.y * ds.y * act_grad<1>(u.y));
            if (j >= 1 && j <= K) {
              dwv[0] = ffma2(tt[j], xw[0][j], dwv[0]); dwv[1] = ffma2(tt[j], xw[0][j + 1], dwv[1]); dwv[2] = ffma2(tt[j], xw[0][j + 2], dwv[2]);
              dwv[3] = ffma2(tt[j], xw[1][j], dwv[3]); dwv[4] = ffma2(tt[j], xw[1][j + 1], dwv[4]); dwv[5] = ffma2(tt[j], xw[1][j + 2], dwv[5]);
              dwv[6] = ffma2(tt[j], xw[2][j], dwv[6]); dwv[7] = ffma2(tt[j], xw[2][j + 1], dwv[7]); dwv[8] = ffma2(tt[j], xw[2][j + 2], dwv[8]);
            }
          }
        }
        // acc[0] = dX row t-1, acc[1] = row t, acc[2] = row t+1
#pragma unroll
        for (int jo = 0; jo < K; ++jo) {
          acc[0][jo] = ffma2(w9[0], tt[jo + 2], acc[0][jo]); acc[0][jo] = ffma2(w9[1], tt[jo + 1], acc[0][jo]); acc[0][jo] = ffma2(w9[2], tt[jo], acc[0][jo]);
          acc[1][jo] = ffma2(w9[3], tt[jo + 2], acc[1][jo]); acc[1][jo] = ffma2(w9[4], tt[jo + 1], acc[1][jo]); acc[1][jo] = ffma2(w9[5], tt[jo], acc[1][jo]);
          acc[2][jo] = ffma2(w9[6], tt[jo + 2], acc[2][jo]); acc[2][jo] = ffma2(w9[7], tt[jo + 1], acc[2][jo]); acc[2][jo] = ffma2(w9[8], tt[jo], acc[2][jo]);
          done[jo] = f2(acc[0][jo].x + gy2.x, acc[0][jo].y + gy2.y);
        }
      }
      __syncthreads();   // every thread has read xn rows <= s for this step: row s-2 may now be replaced by d(xn)
      if (worker && s >= 2) {
        const int h = s - 2;
#pragma unroll
        for (int jo = 0; jo < K; ++jo) {
          const int w = q * K + jo;
          if (w < S) sts_pair_t<T>(xn_s + (uint32_t)((1 + h * S + w) * C + c) * ES, done[jo]);
        }
      }
      if (worker) {
#pragma unroll
        for (int jo = 0; jo < K; ++jo) { acc[0][jo] = acc[1][jo]; acc[1][jo] = acc[2][jo]; acc[2][jo] = f2(0.f, 0.f); }
#pragma unroll
        for (int j = 0; j < KX; ++j) { xw[0][j] = xw[1][j]; xw[1][j] = xw[2][j]; }
      }
    }
  }
  __syncthreads();
  // cls token: d(xn[0]) = d_out[0]
  for (int cc = tid; cc < NP; cc += blockDim.x)
    sts_pair_t<T>(xn_s + (uint32_t)(2 * cc) * ES, ldg_pair<T>(gb + 2 * cc));
  // dWv partial: reduce over the column groups
  if (worker) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      red[(q * 10 + i) * C + c] = dwv[i].x;
      red[(q * 10 + i) * C + c + 1] = dwv[i].y;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 9 * C; idx += blockDim.x) {
    const int i = idx / C, cc = idx - i * C;
    float s = 0.f;
    for (int qq = 0; qq < P.NQ; ++qq) s += red[(qq * 10 + i) * C + cc];
    part[cc * 9 + i] = s;
  }
  __syncthreads();
  // ---- phase 4: both LayerNorm backwards, one token per warp (the loads of two tokens are issued together); per-lane
  // channel sums for d(gamma), d(beta), d(lambda)
  float2 sgx[kDeitMaxCJ], sbx[kDeitMaxCJ], sgo[kDeitMaxCJ], sbo[kDeitMaxCJ], slm[kDeitMaxCJ];
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j) { sgx[j] = sbx[j] = sgo[j] = sbo[j] = slm[j] = f2(0.f, 0.f); }
  const float invC = 1.f / (float)C;
  for (int t0 = warp; t0 < n; t0 += 2 * nw) {
    float2 vx[2][kDeitMaxCJ], vo[2][kDeitMaxCJ], vg[2][kDeitMaxCJ];
    float2 stx[2], sto[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * nw;
      stx[u] = sto[u] = f2(0.f, 0.f);
      if (t < n) {
        stx[u] = *reinterpret_cast<const float2*>(sxg + 2 * t);
        sto[u] = *reinterpret_cast<const float2*>(sog + 2 * t);
      }
#pragma unroll
      for (int j = 0; j < kDeitMaxCJ; ++j)
        if (j < CJ && t < n) {
          const int cc = 2 * (j * 32 + lane);
          vx[u][j] = ldg_pair<T>(xb + (int64_t)t * C + cc);
          vo[u][j] = ldg_pair<T>(ob + (int64_t)t * C + cc);
          vg[u][j] = (t >= 1) ? ldg_pair<T>(gb + (int64_t)t * C + cc) : f2(0.f, 0.f);
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * nw;
      if (t < n) {
        // LN_x
        float2 xh[kDeitMaxCJ], gg[kDeitMaxCJ];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ) {
            const int cc = 2 * (j * 32 + lane);
            const float2 dxn = lds_pair<T>(xn_s + (uint32_t)(t * C + cc) * ES);
            const float2 gm = *reinterpret_cast<const float2*>(P.gx + cc);
            xh[j] = f2((vx[u][j].x - stx[u].x) * stx[u].y, (vx[u][j].y - stx[u].x) * stx[u].y);
            gg[j] = f2(dxn.x * gm.x, dxn.y * gm.y);
            s1 += gg[j].x + gg[j].y;
            s2 = fmaf(gg[j].x, xh[j].x, fmaf(gg[j].y, xh[j].y, s2));
            sgx[j] = f2(fmaf(dxn.x, xh[j].x, sgx[j].x), fmaf(dxn.y, xh[j].y, sgx[j].y));
            sbx[j] = f2(sbx[j].x + dxn.x, sbx[j].y + dxn.y);
          }
        s1 = warp_sum(s1) * invC;
        s2 = warp_sum(s2) * invC;
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ) {
            const int cc = 2 * (j * 32 + lane);
            stg_pair<T>(dxb + (int64_t)t * C + cc, f2(stx[u].y * (gg[j].x - s1 - xh[j].x * s2), stx[u].y * (gg[j].y - s1 - xh[j].y * s2)));
          }
        // LN_o : d(on) = lambda * dS for image tokens, 0 for the cls token (its `on` row is unused, deit_mrla_light.py:204-207)
        float o1 = 0.f, o2 = 0.f;
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ) {
            const int cc = 2 * (j * 32 + lane);
            const float2 ds = vg[u][j];
            const float2 lm = *reinterpret_cast<const float2*>(P.lam + cc);
            const float2 gm = *reinterpret_cast<const float2*>(P.go + cc), bt = *reinterpret_cast<const float2*>(P.bo + cc);
            xh[j] = f2((vo[u][j].x - sto[u].x) * sto[u].y, (vo[u][j].y - sto[u].x) * sto[u].y);
            const float2 don = f2(lm.x * ds.x, lm.y * ds.y);
            gg[j] = f2(don.x * gm.x, don.y * gm.y);
            o1 += gg[j].x + gg[j].y;
            o2 = fmaf(gg[j].x, xh[j].x, fmaf(gg[j].y, xh[j].y, o2));
            sgo[j] = f2(fmaf(don.x, xh[j].x, sgo[j].x), fmaf(don.y, xh[j].y, sgo[j].y));
            sbo[j] = f2(sbo[j].x + don.x, sbo[j].y + don.y);
            slm[j] = f2(fmaf(ds.x, xh[j].x * gm.x + bt.x, slm[j].x), fmaf(ds.y, xh[j].y * gm.y + bt.y, slm[j].y));
          }
        o1 = warp_sum(o1) * invC;
        o2 = warp_sum(o2) * invC;
#pragma unroll
        for (int j = 0; j < kDeitMaxCJ; ++j)
          if (j < CJ) {
            const int cc = 2 * (j * 32 + lane);
            stg_pair<T>(dob + (int64_t)t * C + cc, f2(sto[u].y * (gg[j].x - o1 - xh[j].x * o2), sto[u].y * (gg[j].y - o1 - xh[j].y * o2)));
          }
      }
    }
  }
  // cross-warp reduction of the five per-channel sums: red[warp][5][C]  (the scratch holds nw*5*C <= NQ*10*C floats)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kDeitMaxCJ; ++j) if (j < CJ) {
    const int cc = 2 * (j * 32 + lane);
    float* r0 = red + (size_t)warp * 5 * C;
    *reinterpret_cast<float2*>(r0 + 0 * C + cc) = slm[j];
    *reinterpret_cast<float2*>(r0 + 1 * C + cc) = sgx[j];
    *reinterpret_cast<float2*>(r0 + 2 * C + cc) = sbx[j];
    *reinterpret_cast<float2*>(r0 + 3 * C + cc) = sgo[j];
    *reinterpret_cast<float2*>(r0 + 4 * C + cc) = sbo[j];
  }
  __syncthreads();
  for (int idx = tid; idx < 5 * C; idx += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[(size_t)w * 5 * C + idx];
    part[9 * C + idx] = s;
  }
}

// sum the per-sample partials over the batch: part [B, PW] -> out [PW]   (deterministic order)
static __global__ void k_deit_reduce(const float* __restrict__ part, int B, int PW, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PW) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += part[(int64_t)b * PW + i];
  out[i] = s;
}


// =====================================================================================================
// Token LayerNorm (one warp per token) for the modules that keep the generic tail kernels: DeiT MRLA-base
// (deit/deit_mrla_base.py:224-229 `xn = self.normx(xt)`), replacing at::layer_norm and its backward.
//   forward : xn[b,t,:] = (x - mean) * rstd * gamma + beta ; stats [B,n,2] ; the cls row (t = 0) is ALSO written to
//             cls_out[b, 0, :] (the module's output buffer: out[:, 0] = xn[:, 0], :242), so no torch.cat is needed.
//   backward: d(xn) of token 0 comes from g_cls[b, 0, :], of token t >= 1 from g_img[b, t-1, :] (the gradient of the
//             token image as the tail's backward left it) -> dx ; per-CTA partial sums of d(gamma), d(beta).
// =====================================================================================================
struct LnParams {
  int B, n, C;
  float eps;
  const void* x; void* xn; void* cls_out; int64_t bs_cls;    // bs_* = batch strides in elements
  const float* gamma; const float* beta;
  float* stats;
  const void* g_cls; int64_t bs_gcls; const void* g_img; int64_t bs_gimg;
  void* dx; float* part;                                       // part [gridDim.x, 2, C]
};

template <typename T>
__global__ void __launch_bounds__(256) k_ln_tokens_fwd(const LnParams P) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t ntok = (int64_t)P.B * P.n;
  for (int64_t tok = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); tok < ntok; tok += (int64_t)gridDim.x * wpb) {
    const T* xr = static_cast<const T*>(P.x) + tok * P.C;
    float s = 0.f;
    for (int c = 2 * lane; c < P.C; c += 64) { const float2 v = ldg_pair<T>(xr + c); s += v.x + v.y; }
    const float mean = warp_sum(s) / (float)P.C;
    float q = 0.f;
    for (int c = 2 * lane; c < P.C; c += 64) {
      const float2 v = ldg_pair<T>(xr + c);
      q = fmaf(v.x - mean, v.x - mean, fmaf(v.y - mean, v.y - mean, q));
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)P.C + P.eps);
    const int64_t b = tok / P.n;
    const int t = (int)(tok - b * P.n);
    T* yr = static_cast<T*>(P.xn) + tok * P.C;
    T* cr = (t == 0 && P.cls_out) ? static_cast<T*>(P.cls_out) + b * P.bs_cls : nullptr;
    for (int c = 2 * lane; c < P.C; c += 64) {
      const float2 v = ldg_pair<T>(xr + c);
      const float2 gm = *reinterpret_cast<const float2*>(P.gamma + c), bt = *reinterpret_cast<const float2*>(P.beta + c);
      const float2 y = f2((v.x - mean) * rstd * gm.x + bt.x, (v.y - mean) * rstd * gm.y + bt.y);
      stg_pair<T>(yr + c, y);
      if (cr) stg_pair<T>(cr + c, y);
    }
    if (lane == 0) *reinterpret_cast<float2*>(P.stats + tok * 2) = f2(mean, rstd);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_ln_tokens_bwd(const LnParams P) {
  extern __shared__ float sm[];   // [2][C] per-CTA sums (shared-memory atomics would be non-deterministic: warps take turns)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * P.C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int64_t ntok = (int64_t)P.B * P.n;
  const float invC = 1.f / (float)P.C;
  // every warp keeps its own channel sums in registers for up to 768 channels (12 pairs per lane)
  float2 sg[12], sb[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) sg[j] = sb[j] = f2(0.f, 0.f);
  for (int64_t tok = (int64_t)blockIdx.x * wpb + warp; tok < ntok; tok += (int64_t)gridDim.x * wpb) {
    const int64_t b = tok / P.n;
    const int t = (int)(tok - b * P.n);
    const T* xr = static_cast<const T*>(P.x) + tok * P.C;
    const T* gr = (t == 0) ? static_cast<const T*>(P.g_cls) + b * P.bs_gcls
                           : static_cast<const T*>(P.g_img) + b * P.bs_gimg + (int64_t)(t - 1) * P.C;
    const float2 st = *reinterpret_cast<const float2*>(P.stats + tok * 2);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int c = 2 * lane + 64 * j;
      if (c < P.C) {
        const float2 v = ldg_pair<T>(xr + c), g = ldg_pair<T>(gr + c);
        const float2 gm = *reinterpret_cast<const float2*>(P.gamma + c);
        const float2 xh = f2((v.x - st.x) * st.y, (v.y - st.x) * st.y);
        const float2 gg = f2(g.x * gm.x, g.y * gm.y);
        s1 += gg.x + gg.y;
        s2 = fmaf(gg.x, xh.x, fmaf(gg.y, xh.y, s2));
        sg[j] = f2(fmaf(g.x, xh.x, sg[j].x), fmaf(g.y, xh.y, sg[j].y));
        sb[j] = f2(sb[j].x + g.x, sb[j].y + g.y);
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    T* dr = static_cast<T*>(P.dx) + tok * P.C;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const int c = 2 * lane + 64 * j;
      if (c < P.C) {
        const float2 v = ldg_pair<T>(xr + c), g = ldg_pair<T>(gr + c);
        const float2 gm = *reinterpret_cast<const float2*>(P.gamma + c);
        const float2 xh = f2((v.x - st.x) * st.y, (v.y - st.x) * st.y);
        stg_pair<T>(dr + c, f2(st.y * (g.x * gm.x - s1 - xh.x * s2), st.y * (g.y * gm.y - s1 - xh.y * s2)));
      }
    }
  }
  // deterministic CTA reduction: warps add their sums one after the other
  for (int w = 0; w < wpb; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int c = 2 * lane + 64 * j;
        if (c < P.C) {
          sm[c] += sg[j].x; sm[c + 1] += sg[j].y;
          sm[P.C + c] += sb[j].x; sm[P.C + c + 1] += sb[j].y;
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * P.C; i += blockDim.x) P.part[(int64_t)blockIdx.x * 2 * P.C + i] = sm[i];
}

}  // namespace mrla
