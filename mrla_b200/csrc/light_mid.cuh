// The [B,C]-sized "mid" kernels between the N-sized sweeps of the MRLA-light tail
// (SURVEY.md §8a closed form).  All of this is < 0.1 % of the traffic; it replaces the
// launch-latency-bound conv1d / bmm / sigmoid / batch-norm-statistics ATen calls of
// resnet/models/modules/mrla_light_module.py:59-70 and nn.BatchNorm2d.
#pragma once
#include "common.cuh"

namespace mrla {

struct MidShape {
  int B, C, HW, d, k;   // d = channels per head, k = ECA kernel size
  int bn_mode;          // 0 none, 1 train, 2 eval
  int has_o;            // lambda*o term present
  int full_mom;         // forward moments 1..5 are valid (BN train)
  int update_running;
  float eps, momentum;
  int da_summed;        // backward: bcoef[6] already holds da summed over the head (k_light_mid_bwd)
};

// ---------------------------------------------------------------------------- gate (one CTA per b)
// y = Σx/HW ; Q = xcorr(y,wq) ; K = xcorr(y,wk) ; a[b,h] = sigmoid(Σ_{c in h} Q K / sqrt(d))
static __global__ void __launch_bounds__(1024) k_light_gate(const float* __restrict__ mom, const float* __restrict__ wq,
                                                     const float* __restrict__ wk, float* __restrict__ gate,
                                                     MidShape s) {
  extern __shared__ float sm[];  // y[C] | qk[C]
  float* ys = sm;
  float* qk = sm + s.C;
  const int b = blockIdx.x;
  const float inv_hw = 1.f / (float)s.HW;
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) ys[c] = mom[(int64_t)b * s.C + c] * inv_hw;
  __syncthreads();
  const int pad = (s.k - 1) / 2;
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
    float q = 0.f, kk = 0.f;
    for (int j = 0; j < s.k; ++j) {
      const int cc = c + j - pad;
      const float yv = (cc >= 0 && cc < s.C) ? ys[cc] : 0.f;
      q = fmaf(wq[j], yv, q);
      kk = fmaf(wk[j], yv, kk);
    }
    qk[c] = q * kk;
  }
  __syncthreads();
  const int g = s.C / s.d;
  const float norm = rsqrtf((float)s.d);
  for (int h = threadIdx.x; h < g; h += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < s.d; ++i) acc += qk[h * s.d + i];
    const float logit = acc * norm;
    gate[(int64_t)b * g + h] = 1.f / (1.f + __expf(-logit));
  }
}

// Tiling of the per-channel kernels: kMidCPB channels x kMidBL batch lanes per 1024-thread CTA.  A warp holds 4 batch
// lanes of the 8 channels (32-byte coalesced segments of the [B,C] side tensors); sums over the batch go through two
// shuffles, one shared-memory slot per warp and a 32-term serial tail — C/8 CTAs instead of C/32, B/128 dependent
// iterations per thread instead of B/32.
constexpr int kMidCPB = 8;
constexpr int kMidBL = 128;

// sum `v` over all threads of the CTA that share channel lane `cl` (threadIdx.x & 7); result valid where bl == 0
__device__ __forceinline__ double mid_channel_sum(double v, double (*red)[kMidCPB], int cl, int bl) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  __syncthreads();                                  // previous use of `red` is over
  if ((threadIdx.x & 31) < kMidCPB) red[threadIdx.x >> 5][cl] = v;
  __syncthreads();
  double a = 0.0;
  if (bl == 0)
    for (int j = 0; j < 32; ++j) a += red[j][cl];
  return a;
}

// ------------------------------------------------- BN statistics + forward coefficients (8 channels / CTA)
// mean_c = Σ_b (a ΣV + λ Σo)/n ; E[s²]_c = Σ_b (a² ΣV² + 2aλ ΣVo + λ² Σo²)/n  (accumulated in fp64)
// coef = [3,B,C]:  A = m_b γ r a ,  L = m_b γ r λ ,  D = m_b (β − γ r μ)
static __global__ void __launch_bounds__(1024) k_light_bn_coef(const float* __restrict__ mom, const float* __restrict__ gate,
                                                        const float* __restrict__ lam, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ running_mean,
                                                        float* __restrict__ running_var,
                                                        const float* __restrict__ drop_scale, float* __restrict__ mean,
                                                        float* __restrict__ rstd, float* __restrict__ coef, MidShape s) {
  __shared__ double red[32][kMidCPB];
  __shared__ float s_mu[kMidCPB], s_r[kMidCPB];
  const int cl = threadIdx.x & (kMidCPB - 1), bl = threadIdx.x / kMidCPB;
  const int c = blockIdx.x * kMidCPB + cl;
  const bool cok = c < s.C;
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const float lm = (cok && s.has_o) ? lam[c] : 0.f;
  const double n = (double)s.B * (double)s.HW;

  if (s.bn_mode == 1) {
    double s1 = 0.0, s2 = 0.0;
    if (cok) {
#pragma unroll 2
      for (int b = bl; b < s.B; b += kMidBL) {
        const int64_t i = (int64_t)b * s.C + c;
        const double a = gate[(int64_t)b * g + c / s.d];
        const double sv = mom[BC + i], svv = mom[2 * BC + i];
        double t1 = a * sv, t2 = a * a * svv;
        if (s.has_o) {
          const double svo = mom[3 * BC + i], so = mom[4 * BC + i], soo = mom[5 * BC + i];
          t1 += (double)lm * so;
          t2 += 2.0 * a * lm * svo + (double)lm * lm * soo;
        }
        s1 += t1;
        s2 += t2;
      }
    }
    const double a1 = mid_channel_sum(s1, red, cl, bl);
    const double a2 = mid_channel_sum(s2, red, cl, bl);
    if (bl == 0) {
      const double mu = a1 / n;
      double var = a2 / n - mu * mu;
      if (var < 0.0) var = 0.0;
      const double r = 1.0 / sqrt(var + (double)s.eps);
      s_mu[cl] = (float)mu;
      s_r[cl] = (float)r;
      if (cok) {
        mean[c] = (float)mu;
        rstd[c] = (float)r;
        if (s.update_running && running_mean != nullptr) {
          const double unb = var * (n / fmax(n - 1.0, 1.0));
          running_mean[c] = (float)((1.0 - s.momentum) * (double)running_mean[c] + (double)s.momentum * mu);
          running_var[c] = (float)((1.0 - s.momentum) * (double)running_var[c] + (double)s.momentum * unb);
        }
      }
    }
    __syncthreads();
  } else {
    if (bl == 0) {
      float mu = 0.f, r = 1.f;
      if (s.bn_mode == 2 && cok) {
        mu = running_mean[c];
        r = (float)(1.0 / sqrt((double)running_var[c] + (double)s.eps));
      }
      s_mu[cl] = mu;
      s_r[cl] = r;
      if (cok) { mean[c] = mu; rstd[c] = r; }
    }
    __syncthreads();
  }
  if (!cok) return;
  const float ga = (s.bn_mode != 0) ? gamma[c] : 1.f;
  const float be = (s.bn_mode != 0) ? beta[c] : 0.f;
  const float gr = ga * s_r[cl];
  const float dterm = be - gr * s_mu[cl];
  for (int b = bl; b < s.B; b += kMidBL) {
    const int64_t i = (int64_t)b * s.C + c;
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    const float a = gate[(int64_t)b * g + c / s.d];
    coef[i] = mb * gr * a;
    coef[BC + i] = mb * gr * lm;
    coef[2 * BC + i] = mb * dterm;
  }
}

// ------------------------------------------------------- backward, per-channel part (8 channels / CTA)
// G' = m_b dY.  dβ = Σ_b ΣG' ; dγ = r Σ_b (a ΣG'V + λ ΣG'o − μ ΣG') ; m1 = dβ/n ; m2 = dγ/n  (train only)
// dλ = γ r Σ_b (ΣG'o − m1 Σo − m2 ΣŜo) ; da[b,c] = γ r (ΣG'V − m1 ΣV − m2 ΣŜV)
// bcoef = [7,B,C]: Q0,Q1,Q2,Q3, Ta, dyc (filled by k_light_bwd_gate), da[b,c]
static __global__ void __launch_bounds__(1024) k_light_bwd_chan(const float* __restrict__ mom, const float* __restrict__ gmom,
                                                         const float* __restrict__ gate, const float* __restrict__ lam,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ drop_scale,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         float* __restrict__ bcoef, float* __restrict__ dlam,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         MidShape s) {
  __shared__ double red[32][kMidCPB];
  __shared__ double s_m1[kMidCPB], s_m2[kMidCPB];
  const int cl = threadIdx.x & (kMidCPB - 1), bl = threadIdx.x / kMidCPB;
  const int c = blockIdx.x * kMidCPB + cl;
  const bool cok = c < s.C;
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const double lm = (cok && s.has_o) ? (double)lam[c] : 0.0;
  const double mu = cok ? (double)mean[c] : 0.0;
  const double r = cok ? (double)rstd[c] : 1.0;
  const double ga = (cok && s.bn_mode != 0) ? (double)gamma[c] : 1.0;
  const double n = (double)s.B * (double)s.HW;

  // pass 1: dβ, dγ
  double s1 = 0.0, s2 = 0.0;
  if (cok && s.bn_mode != 0) {
#pragma unroll 2
    for (int b = bl; b < s.B; b += kMidBL) {
      const int64_t i = (int64_t)b * s.C + c;
      const double mb = drop_scale ? (double)drop_scale[b] : 1.0;
      const double a = gate[(int64_t)b * g + c / s.d];
      const double g1 = mb * gmom[i], gv = mb * gmom[BC + i];
      const double go = s.has_o ? mb * gmom[2 * BC + i] : 0.0;
      s1 += g1;
      s2 += a * gv + lm * go - mu * g1;
    }
  }
  const double a1 = mid_channel_sum(s1, red, cl, bl);
  double a2 = mid_channel_sum(s2, red, cl, bl);
  if (bl == 0) {
    a2 *= r;
    if (cok && s.bn_mode != 0) {
      if (dbeta) dbeta[c] = (float)a1;
      if (dgamma) dgamma[c] = (float)a2;
    }
    const bool tr = (s.bn_mode == 1);
    s_m1[cl] = tr ? a1 / n : 0.0;
    s_m2[cl] = tr ? a2 / n : 0.0;
  }
  __syncthreads();
  const double m1 = s_m1[cl], m2 = s_m2[cl];
  const double gr = ga * r;

  // pass 2: dλ, da[b,c], sweep-B coefficients
  double sl = 0.0;
  if (cok) {
#pragma unroll 2
    for (int b = bl; b < s.B; b += kMidBL) {
      const int64_t i = (int64_t)b * s.C + c;
      const double mb = drop_scale ? (double)drop_scale[b] : 1.0;
      const double a = gate[(int64_t)b * g + c / s.d];
      const double gv = mb * gmom[BC + i];
      const double go = s.has_o ? mb * gmom[2 * BC + i] : 0.0;
      double da = gv, dl = go;
      if (s.bn_mode == 1) {
        const double sv = mom[BC + i], svv = mom[2 * BC + i];
        double svo = 0.0, so = 0.0, soo = 0.0;
        if (s.has_o) { svo = mom[3 * BC + i]; so = mom[4 * BC + i]; soo = mom[5 * BC + i]; }
        const double shv = r * (a * svv + lm * svo - mu * sv);  // Σ Ŝ V
        const double sho = r * (a * svo + lm * soo - mu * so);  // Σ Ŝ o
        da = gv - m1 * sv - m2 * shv;
        dl = go - m1 * so - m2 * sho;
      }
      sl += gr * dl;
      bcoef[6 * BC + i] = (float)(gr * da);
      bcoef[0 * BC + i] = (float)(gr * (-m1 + m2 * r * mu));
      bcoef[1 * BC + i] = (float)(gr * mb);
      bcoef[2 * BC + i] = (float)(-gr * m2 * r * a);
      bcoef[3 * BC + i] = (float)(-gr * m2 * r * lm);
      bcoef[4 * BC + i] = (float)a;
    }
  }
  if (dlam != nullptr) {
    const double al = mid_channel_sum(sl, red, cl, bl);
    if (bl == 0 && cok) dlam[c] = (float)al;
  }
}

// --------------------------------------------------------------- backward, gate part (one CTA per b)
// da[b,h] = Σ_{c in h} da[b,c] ; dlogit = da a (1−a) / sqrt(d) ; dQ = dlogit K ; dK = dlogit Q
// dy = xcorrᵀ(dQ,wq) + xcorrᵀ(dK,wk) -> bcoef[5] = dy/HW ;  per-b partials of dwq, dwk -> wqk_part[B,2k]
static __global__ void __launch_bounds__(1024) k_light_bwd_gate(const float* __restrict__ mom, const float* __restrict__ wq,
                                                         const float* __restrict__ wk, const float* __restrict__ gate,
                                                         float* __restrict__ bcoef, float* __restrict__ wqk_part,
                                                         MidShape s) {
  extern __shared__ float sm[];  // y[C] | Q[C] | K[C] | dQ[C] | dK[C] | dlogit[g]
  float* ys = sm;
  float* qs = sm + s.C;
  float* ks = sm + 2 * s.C;
  float* dq = sm + 3 * s.C;
  float* dk = sm + 4 * s.C;
  float* dl = sm + 5 * s.C;
  const int b = blockIdx.x;
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const float inv_hw = 1.f / (float)s.HW;
  const int pad = (s.k - 1) / 2;
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) ys[c] = mom[(int64_t)b * s.C + c] * inv_hw;
  __syncthreads();
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
    float q = 0.f, kk = 0.f;
    for (int j = 0; j < s.k; ++j) {
      const int cc = c + j - pad;
      const float yv = (cc >= 0 && cc < s.C) ? ys[cc] : 0.f;
      q = fmaf(wq[j], yv, q);
      kk = fmaf(wk[j], yv, kk);
    }
    qs[c] = q;
    ks[c] = kk;
  }
  const float norm = rsqrtf((float)s.d);
  for (int h = threadIdx.x; h < g; h += blockDim.x) {
    float acc = 0.f;
    if (s.da_summed) acc = bcoef[6 * BC + (int64_t)b * s.C + h * s.d];
    else
      for (int i = 0; i < s.d; ++i) acc += bcoef[6 * BC + (int64_t)b * s.C + h * s.d + i];
    const float a = gate[(int64_t)b * g + h];
    dl[h] = acc * a * (1.f - a) * norm;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
    const float d_ = dl[c / s.d];
    dq[c] = d_ * ks[c];
    dk[c] = d_ * qs[c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < s.k; ++j) {
      const int cc = c - j + pad;  // Q[cc] used y[cc + j - pad] = y[c]
      if (cc >= 0 && cc < s.C) acc = fmaf(wq[j], dq[cc], fmaf(wk[j], dk[cc], acc));
    }
    bcoef[5 * BC + (int64_t)b * s.C + c] = acc * inv_hw;
  }
  // dwq[j] = Σ_c y[c+j-pad] dQ[c], dwk likewise: one warp per tap (k <= 15), no block barriers
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < 2 * s.k; j += nw) {
    const int jj = (j < s.k) ? j : j - s.k;
    const float* src = (j < s.k) ? dq : dk;
    float acc = 0.f;
    for (int c = lane; c < s.C; c += 32) {
      const int cc = c + jj - pad;
      if (cc >= 0 && cc < s.C) acc = fmaf(ys[cc], src[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) wqk_part[(int64_t)b * 2 * s.k + j] = acc;
  }
}

// --------------------------------------------------------------------------------- final reductions
// dwv[c,tap] = Σ_p wv_part[p,c,tap] (one thread per output) ; dwq[j] = Σ_b wqk_part[b,j], dwk[j] = Σ_b wqk_part[b,k+j]
// (one warp per tap: strided loads + shuffle tree instead of a B-long dependent chain)
static __global__ void k_light_finish(const float* __restrict__ wv_part, int nparts, const float* __restrict__ wqk_part,
                                      float* __restrict__ dwv, float* __restrict__ dwq, float* __restrict__ dwk, int B,
                                      int C, int k) {
  const int64_t n1 = (int64_t)C * 9;
  const int nb1 = (int)((n1 + blockDim.x - 1) / blockDim.x);
  if ((int)blockIdx.x < nb1) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n1 && dwv != nullptr) {
      float acc = 0.f;
      for (int p = 0; p < nparts; ++p) acc += wv_part[(int64_t)p * n1 + idx];
      dwv[idx] = acc;
    }
    return;
  }
  const int warps_per_block = blockDim.x >> 5;
  const int j = ((int)blockIdx.x - nb1) * warps_per_block + (threadIdx.x >> 5);
  if (j >= 2 * k) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int b = lane; b < B; b += 32) acc += wqk_part[(int64_t)b * 2 * k + j];
  acc = warp_sum(acc);
  if (lane == 0) {
    if (j < k) { if (dwq) dwq[j] = acc; }
    else { if (dwk) dwk[j - k] = acc; }
  }
}

}  // namespace mrla
