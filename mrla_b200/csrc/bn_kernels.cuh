// Channels-last BatchNorm2d (+ optional ReLU) for the bottleneck's bn1 / bn2 / bn3 and the stem — the producers
// on either side of the MRLA tail.  PyTorch routes bf16 channels_last batch norm to its
// native kernels (cuDNN BN is skipped for bf16); they take ~31 ms of a 60 ms resnet50_mrlal step on B200.
// This is the same two-pass structure as the tail's moments / apply sweeps without the stencil:
//   forward   k_bn_stats   R1     per-channel Σx, Σx² (per-CTA partials, fp64 finish)      -> mean, rstd, a, b
//             k_bn_apply   R1 W1  y = act(a_c x + b_c)
//   backward  k_bn_bwd_red R2     Σdz, Σdz·(x−μ) (dz = dy·[a x + b > 0] when ReLU is fused) -> dγ dβ, A B C
//             k_bn_bwd_app R2 W1  dx = A_c dz + B_c x + C_c
// View: x is [M = B·H·W rows, C channels] row-major (NHWC).  A thread owns 8 consecutive channels (one 128-bit
// vector of bf16) and strides over rows; a 256-thread CTA is CL = C/8 channel lanes x RL row lanes.
#pragma once
#include "base_stream.cuh"

namespace mrla {

struct BnShape {
  int64_t M;     // rows (B*H*W)
  int C;
  int CL, RL;    // channel lanes (C/8), row lanes (256/CL)
  int nparts;    // gridDim.x of the reduction kernels
};

template <int NV>
__device__ __forceinline__ void bn_reduce_rows(const float (&v)[NV], float* sm, bool active, int cl, int rl, int CL,
                                               int RL, float* out_base, int C, int c) {
  // sum over the RL row lanes of this CTA, write NV/8 vectors of 8 channels: out_base[k*C + c + i]
  if (active) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[((size_t)rl * CL + cl) * NV + i] = v[i];
  }
  __syncthreads();
  if (active && rl == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = 0.f;
      for (int r = 0; r < RL; ++r) s += sm[((size_t)r * CL + cl) * NV + i];
      out_base[(size_t)(i / kSV) * C + c + (i % kSV)] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------- forward statistics
// Sums are taken of (x - pivot_c), pivot_c = x[0, c]: E[x^2] - mean^2 on raw values cancels catastrophically when
// |mean| >> std (mean 50, std 0.1 leaves no correct bits of the variance in fp32); shifted by any sample of the channel the
// two terms are of the size of the variance itself.  pivot [C] is written by CTA 0 for k_bn_finalize.
template <typename T>
__global__ void __launch_bounds__(256) k_bn_stats(const T* __restrict__ x, float* __restrict__ part,
                                                  float* __restrict__ pivot, BnShape s) {
  extern __shared__ float smem[];
  const int cl = threadIdx.x % s.CL, rl = threadIdx.x / s.CL;
  const bool active = rl < s.RL;
  const int c = cl * kSV;
  float acc[2 * kSV];
#pragma unroll
  for (int i = 0; i < 2 * kSV; ++i) acc[i] = 0.f;
  if (active) {
    float pv[kSV];
    Vec8<T>::ld(x + c, pv);
    if (blockIdx.x == 0 && rl == 0) {
#pragma unroll
      for (int i = 0; i < kSV; ++i) pivot[c + i] = pv[i];
    }
    const int64_t stride = (int64_t)gridDim.x * s.RL;
    int64_t row = (int64_t)blockIdx.x * s.RL + rl;
    for (; row + 3 * stride < s.M; row += 4 * stride) {   // four independent 128-bit loads in flight
      float a[kSV], b[kSV], e[kSV], f[kSV];
      Vec8<T>::ld(x + row * s.C + c, a);
      Vec8<T>::ld(x + (row + stride) * s.C + c, b);
      Vec8<T>::ld(x + (row + 2 * stride) * s.C + c, e);
      Vec8<T>::ld(x + (row + 3 * stride) * s.C + c, f);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float da = a[i] - pv[i], db = b[i] - pv[i], de = e[i] - pv[i], df = f[i] - pv[i];
        acc[i] += (da + db) + (de + df);
        acc[kSV + i] = fmaf(da, da, fmaf(db, db, fmaf(de, de, fmaf(df, df, acc[kSV + i]))));
      }
    }
    for (; row + stride < s.M; row += 2 * stride) {
      float a[kSV], b[kSV];
      Vec8<T>::ld(x + row * s.C + c, a);
      Vec8<T>::ld(x + (row + stride) * s.C + c, b);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float da = a[i] - pv[i], db = b[i] - pv[i];
        acc[i] += da + db;
        acc[kSV + i] = fmaf(da, da, fmaf(db, db, acc[kSV + i]));
      }
    }
    if (row < s.M) {
      float a[kSV];
      Vec8<T>::ld(x + row * s.C + c, a);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float da = a[i] - pv[i];
        acc[i] += da;
        acc[kSV + i] = fmaf(da, da, acc[kSV + i]);
      }
    }
  }
  bn_reduce_rows<2 * kSV>(acc, smem, active, cl, rl, s.CL, s.RL, part + (size_t)blockIdx.x * 2 * s.C, s.C, c);
}

// part [nparts, 2, C] -> mean, rstd (saved) ; coef [2,C]: a = γ r, b = β − γ r μ ; running stats
static __global__ void __launch_bounds__(1024) k_bn_finalize(const float* __restrict__ part, const float* __restrict__ pivot,
                                                            int nparts, int C, double n,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ running_mean, float* __restrict__ running_var,
                                                            float* __restrict__ stats, float* __restrict__ coef, float eps,
                                                            float momentum, int training, int update_running) {
  __shared__ double r1[32][33], r2[32][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C && training)
#pragma unroll 4
    for (int p = pl; p < nparts; p += 32) {
      s1 += (double)part[((size_t)p * 2) * C + c];
      s2 += (double)part[((size_t)p * 2 + 1) * C + c];
    }
  r1[pl][cl] = s1;
  r2[pl][cl] = s2;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  double mu, r;
  if (training) {
    double a1 = 0.0, a2 = 0.0;
    for (int j = 0; j < 32; ++j) { a1 += r1[j][cl]; a2 += r2[j][cl]; }
    const double ms = a1 / n;            // mean of the shifted values
    mu = (double)pivot[c] + ms;
    double var = a2 / n - ms * ms;
    if (var < 0.0) var = 0.0;
    r = 1.0 / sqrt(var + (double)eps);
    if (update_running && running_mean != nullptr) {
      const double unb = var * (n / fmax(n - 1.0, 1.0));
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + (double)momentum * mu);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + (double)momentum * unb);
    }
  } else {
    mu = running_mean[c];
    r = 1.0 / sqrt((double)running_var[c] + (double)eps);
  }
  const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
  stats[c] = (float)mu;
  stats[C + c] = (float)r;
  coef[c] = (float)(ga * r);
  coef[C + c] = (float)(be - ga * r * mu);
}

// ------------------------------------------------------------------------------------------- forward apply
template <typename T, bool RELU>
__global__ void __launch_bounds__(256) k_bn_apply(const T* __restrict__ x, T* __restrict__ y,
                                                  const float* __restrict__ coef, BnShape s) {
  const int cl = threadIdx.x % s.CL, rl = threadIdx.x / s.CL;
  if (rl >= s.RL) return;
  const int c = cl * kSV;
  float a[kSV], b[kSV];
#pragma unroll
  for (int i = 0; i < kSV; ++i) { a[i] = coef[c + i]; b[i] = coef[s.C + c + i]; }
  const int64_t stride = (int64_t)gridDim.x * s.RL;
#pragma unroll 2
  for (int64_t row = (int64_t)blockIdx.x * s.RL + rl; row < s.M; row += stride) {
    float v[kSV];
    Vec8<T>::ld(x + row * s.C + c, v);
#pragma unroll
    for (int i = 0; i < kSV; ++i) {
      v[i] = fmaf(a[i], v[i], b[i]);
      if (RELU) v[i] = fmaxf(v[i], 0.f);
    }
    Vec8<T>::st(y + row * s.C + c, v);
  }
}

// ------------------------------------------------------------------------------------------- backward reduce
template <typename T, bool RELU>
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const T* __restrict__ dy, const T* __restrict__ x,
                                                       const float* __restrict__ coef, const float* __restrict__ stats,
                                                       float* __restrict__ part, BnShape s) {
  extern __shared__ float smem[];
  const int cl = threadIdx.x % s.CL, rl = threadIdx.x / s.CL;
  const bool active = rl < s.RL;
  const int c = cl * kSV;
  float a[kSV], b[kSV], mu[kSV];
#pragma unroll
  for (int i = 0; i < kSV; ++i) {
    a[i] = active ? coef[c + i] : 0.f; b[i] = active ? coef[s.C + c + i] : 0.f;
    mu[i] = active ? stats[c + i] : 0.f;
  }
  float acc[2 * kSV];
#pragma unroll
  for (int i = 0; i < 2 * kSV; ++i) acc[i] = 0.f;
  if (active) {
    const int64_t stride = (int64_t)gridDim.x * s.RL;
#pragma unroll 2
    for (int64_t row = (int64_t)blockIdx.x * s.RL + rl; row < s.M; row += stride) {
      float g[kSV], v[kSV];
      Vec8<T>::ld(dy + row * s.C + c, g);
      Vec8<T>::ld(x + row * s.C + c, v);
#pragma unroll
      for (int i = 0; i < kSV; ++i) {
        const float dz = (RELU && fmaf(a[i], v[i], b[i]) <= 0.f) ? 0.f : g[i];
        acc[i] += dz;
        acc[kSV + i] = fmaf(dz, v[i] - mu[i], acc[kSV + i]);   // centred: no sum dz*x - mean*sum dz cancellation
      }
    }
  }
  bn_reduce_rows<2 * kSV>(acc, smem, active, cl, rl, s.CL, s.RL, part + (size_t)blockIdx.x * 2 * s.C, s.C, c);
}

// part [nparts,2,C] (Σdz, Σdz·x) -> dγ dβ ; bcoef [3,C]: dx = A dz + B x + Cc
//   train: dx = γ r (dz − m1 − x̂ m2), x̂ = r (x − μ), m1 = Σdz/n, m2 = Σdz x̂ / n ;  eval: dx = γ r dz
static __global__ void __launch_bounds__(1024) k_bn_bwd_finalize(const float* __restrict__ part, int nparts, int C,
                                                                double n, const float* __restrict__ gamma,
                                                                const float* __restrict__ stats,
                                                                float* __restrict__ bcoef, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, int training, int centred) {
  __shared__ double r1[32][33], r2[32][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C)
#pragma unroll 4
    for (int p = pl; p < nparts; p += 32) {
      s1 += (double)part[((size_t)p * 2) * C + c];
      s2 += (double)part[((size_t)p * 2 + 1) * C + c];
    }
  r1[pl][cl] = s1;
  r2[pl][cl] = s2;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  double a1 = 0.0, a2 = 0.0;
  for (int j = 0; j < 32; ++j) { a1 += r1[j][cl]; a2 += r2[j][cl]; }
  const double mu = stats[c], r = stats[C + c], ga = gamma ? (double)gamma[c] : 1.0;
  // Σdz x̂ : the second sum is Σdz (x - mean) (centred, k_bn_bwd_reduce) or Σdz x (raw, from the MRLA tail's sweep B)
  const double dbe = a1, dga = centred ? r * a2 : r * (a2 - mu * a1);
  if (dbeta) dbeta[c] = (float)dbe;
  if (dgamma) dgamma[c] = (float)dga;
  const double m1 = training ? dbe / n : 0.0, m2 = training ? dga / n : 0.0;
  bcoef[c] = (float)(ga * r);
  bcoef[C + c] = (float)(-ga * r * r * m2);
  bcoef[2 * C + c] = (float)(ga * r * (-m1 + m2 * r * mu));
}

// ------------------------------------------------------------------------------------------- backward apply
template <typename T, bool RELU>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const T* __restrict__ dy, const T* __restrict__ x,
                                                      T* __restrict__ dx, const float* __restrict__ coef,
                                                      const float* __restrict__ bcoef, BnShape s) {
  const int cl = threadIdx.x % s.CL, rl = threadIdx.x / s.CL;
  if (rl >= s.RL) return;
  const int c = cl * kSV;
  float a[kSV], b[kSV], A[kSV], Bc[kSV], Cc[kSV];
#pragma unroll
  for (int i = 0; i < kSV; ++i) {
    a[i] = coef[c + i]; b[i] = coef[s.C + c + i];
    A[i] = bcoef[c + i]; Bc[i] = bcoef[s.C + c + i]; Cc[i] = bcoef[2 * s.C + c + i];
  }
  const int64_t stride = (int64_t)gridDim.x * s.RL;
#pragma unroll 2
  for (int64_t row = (int64_t)blockIdx.x * s.RL + rl; row < s.M; row += stride) {
    float g[kSV], v[kSV], o[kSV];
    Vec8<T>::ld(dy + row * s.C + c, g);
    Vec8<T>::ld(x + row * s.C + c, v);
#pragma unroll
    for (int i = 0; i < kSV; ++i) {
      const float dz = (RELU && fmaf(a[i], v[i], b[i]) <= 0.f) ? 0.f : g[i];
      o[i] = fmaf(A[i], dz, fmaf(Bc[i], v[i], Cc[i]));
    }
    Vec8<T>::st(dx + row * s.C + c, o);
  }
}

}  // namespace mrla
