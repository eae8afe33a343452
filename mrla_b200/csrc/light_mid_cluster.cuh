// Cluster versions of the [B,C]-sized "mid" step of the MRLA-light tail (SURVEY.md §8a closed form): one launch forward
// (gate + BN statistics + sweep-2 coefficients) and one launch for the per-channel half of backward, replacing
// k_light_gate + k_light_bn_coef / k_light_bwd_chan of light_mid.cuh when a CTA can own whole heads (32 % d == 0,
// C % 32 == 0, k <= 15).
//
// Tiling: CTA = 32 consecutive channels (one 128-byte row segment of the [B,C] side tensors per warp load) x a slice
// of the batch; the NS CTAs that share a channel group form one thread-block cluster (grid (C/32, NS), cluster
// (1, NS, 1)).  Sums over the batch are reduced inside the CTA through shared memory, across the cluster through
// distributed shared memory (every CTA adds the NS partials in rank order, so all of them hold bit-identical
// statistics) and never touch global memory or atomics.  The gate needs Σx of its head's channels plus the k-tap
// halo only, so each warp evaluates it for its own sample in registers (shuffle butterfly over the d lanes of a head).
// Reference arithmetic: resnet/models/modules/mrla_light_module.py:56-70, nn.BatchNorm2d (resnet_mrla_light.py:85).
#pragma once
#include <cooperative_groups.h>

#include "light_mid.cuh"

namespace mrla {

namespace cg = cooperative_groups;

constexpr int kMcW = 32;       // channels per CTA
constexpr int kMcWarps = 16;   // batch lanes (warps) per CTA
constexpr int kMcMaxK = 15;    // ECA taps supported by the fused path
constexpr int kMcMaxNS = 8;    // portable cluster size

inline bool mid_cluster_ok(int C, int d, int k) { return d >= 1 && d <= 32 && (32 % d) == 0 && (C % kMcW) == 0 && k <= kMcMaxK && (k & 1); }
inline int mid_cluster_ns(int B) {
  int ns = 1;
  while (ns < kMcMaxNS && B >= 32 * ns * 2) ns *= 2;   // >= 32 samples per CTA
  return ns;
}

// a[b, head(c)] for the warp's sample `b`: every lane ends up with its own head's gate
__device__ __forceinline__ float mid_gate_row(const float* __restrict__ xsum_row, float* ys, const float* swq, const float* swk,
                                              int c0, int lane, int C, int d, int k, float inv_hw, float norm,
                                              float* q_out = nullptr, float* k_out = nullptr) {
  const int pad = (k - 1) / 2;
  ys[pad + lane] = xsum_row[c0 + lane] * inv_hw;
  if (lane < 2 * pad) {
    const int cc = (lane < pad) ? c0 - pad + lane : c0 + kMcW + (lane - pad);
    ys[(lane < pad) ? lane : kMcW + lane] = (cc >= 0 && cc < C) ? xsum_row[cc] * inv_hw : 0.f;
  }
  __syncwarp();
  float q = 0.f, kk = 0.f;
  for (int j = 0; j < k; ++j) {
    const float yv = ys[lane + j];
    q = fmaf(swq[j], yv, q);
    kk = fmaf(swk[j], yv, kk);
  }
  __syncwarp();   // ys is rewritten by the next sample of this warp
  if (q_out) { *q_out = q; *k_out = kk; }
  float qk = q * kk;
  for (int off = d >> 1; off > 0; off >>= 1) qk += __shfl_xor_sync(0xffffffffu, qk, off);
  return 1.f / (1.f + __expf(-qk * norm));
}

// sum v over the kMcWarps warps of the CTA for every channel lane; result valid in warp 0
__device__ __forceinline__ double mid_cta_sum(double v, double (*red)[kMcW], int w, int lane) {
  __syncthreads();            // previous use of `red` is over
  red[w][lane] = v;
  __syncthreads();
  double a = 0.0;
  if (w == 0)
    for (int j = 0; j < kMcWarps; ++j) a += red[j][lane];
  return a;
}

// ------------------------------------------------------------------ forward: gate + BN statistics + coefficients
static __global__ void __launch_bounds__(kMcW * kMcWarps) k_light_mid_fwd(
    const float* __restrict__ mom, const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ lam,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
    float* __restrict__ running_var, const float* __restrict__ drop_scale, float* __restrict__ gate,
    float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ coef, MidShape s) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double part[2][kMcW];                 // this CTA's partial sums, read by the cluster peers
  __shared__ double red[kMcWarps][kMcW];
  __shared__ float ysm[kMcWarps][kMcW + kMcMaxK + 1];
  __shared__ float swq[kMcMaxK + 1], swk[kMcMaxK + 1];
  __shared__ float s_mu[kMcW], s_r[kMcW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * kMcW, c = c0 + lane;
  const int ns = gridDim.y, rank = blockIdx.y;
  const int per = (s.B + ns - 1) / ns;
  const int b_beg = rank * per, b_end = min(s.B, b_beg + per);
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const float inv_hw = 1.f / (float)s.HW, norm = rsqrtf((float)s.d);
  const float lm = s.has_o ? lam[c] : 0.f;
  const double n = (double)s.B * (double)s.HW;
  if (threadIdx.x < s.k) { swq[threadIdx.x] = wq[threadIdx.x]; swk[threadIdx.x] = wk[threadIdx.x]; }
  __syncthreads();

  double s1 = 0.0, s2 = 0.0;
  for (int b = b_beg + w; b < b_end; b += kMcWarps) {
    const int64_t i = (int64_t)b * s.C + c;
    // issue the moment loads before the gate arithmetic
    float sv = 0.f, svv = 0.f, svo = 0.f, so = 0.f, soo = 0.f;
    if (s.bn_mode == 1) {
      sv = mom[BC + i]; svv = mom[2 * BC + i];
      if (s.has_o) { svo = mom[3 * BC + i]; so = mom[4 * BC + i]; soo = mom[5 * BC + i]; }
    }
    const float af = mid_gate_row(mom + (int64_t)b * s.C, ysm[w], swq, swk, c0, lane, s.C, s.d, s.k, inv_hw, norm);
    if ((lane % s.d) == 0) gate[(int64_t)b * g + c / s.d] = af;
    if (s.bn_mode == 1) {
      const double a = af;
      double t1 = a * sv, t2 = a * a * svv;
      if (s.has_o) {
        t1 += (double)lm * so;
        t2 += 2.0 * a * lm * svo + (double)lm * lm * soo;
      }
      s1 += t1;
      s2 += t2;
    }
  }
  if (s.bn_mode == 1) {
    const double a1 = mid_cta_sum(s1, red, w, lane);
    const double a2 = mid_cta_sum(s2, red, w, lane);
    if (w == 0) { part[0][lane] = a1; part[1][lane] = a2; }
    cluster.sync();
    if (w == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int r = 0; r < ns; ++r) {
        const double* rp = cluster.map_shared_rank(&part[0][0], r);
        t1 += rp[lane];
        t2 += rp[kMcW + lane];
      }
      const double mu = t1 / n;
      double var = t2 / n - mu * mu;
      if (var < 0.0) var = 0.0;
      const double r = 1.0 / sqrt(var + (double)s.eps);
      s_mu[lane] = (float)mu;
      s_r[lane] = (float)r;
      if (rank == 0) {
        mean[c] = (float)mu;
        rstd[c] = (float)r;
        if (s.update_running && running_mean != nullptr) {
          const double unb = var * (n / fmax(n - 1.0, 1.0));
          running_mean[c] = (float)((1.0 - s.momentum) * (double)running_mean[c] + (double)s.momentum * mu);
          running_var[c] = (float)((1.0 - s.momentum) * (double)running_var[c] + (double)s.momentum * unb);
        }
      }
    }
  } else if (w == 0) {
    float mu = 0.f, r = 1.f;
    if (s.bn_mode == 2) {
      mu = running_mean[c];
      r = (float)(1.0 / sqrt((double)running_var[c] + (double)s.eps));
    }
    s_mu[lane] = mu;
    s_r[lane] = r;
    if (rank == 0) { mean[c] = mu; rstd[c] = r; }
  }
  __syncthreads();   // s_mu / s_r, and this CTA's gate[] rows, are visible to all its threads
  const float ga = (s.bn_mode != 0) ? gamma[c] : 1.f;
  const float be = (s.bn_mode != 0) ? beta[c] : 0.f;
  const float gr = ga * s_r[lane];
  const float dterm = be - gr * s_mu[lane];
  for (int b = b_beg + w; b < b_end; b += kMcWarps) {
    const int64_t i = (int64_t)b * s.C + c;
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    const float a = gate[(int64_t)b * g + c / s.d];
    coef[i] = mb * gr * a;
    coef[BC + i] = mb * gr * lm;
    coef[2 * BC + i] = mb * dterm;
  }
  cluster.sync();   // no CTA leaves while a peer may still read its `part`
}

// ------------------------------------------------------------------ backward, per-channel half
// Same arithmetic as k_light_bwd_chan; bcoef[6] receives da summed over the head (every channel of the head holds the
// head's sum), which k_light_bwd_gate reads with MidShape::da_summed = 1.
static __global__ void __launch_bounds__(kMcW * kMcWarps) k_light_mid_bwd(
    const float* __restrict__ mom, const float* __restrict__ gmom, const float* __restrict__ gate,
    const float* __restrict__ lam, const float* __restrict__ gamma, const float* __restrict__ drop_scale,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ bcoef, float* __restrict__ dlam,
    float* __restrict__ dgamma, float* __restrict__ dbeta, MidShape s) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double part[2][kMcW];
  __shared__ double part_l[kMcW];
  __shared__ double red[kMcWarps][kMcW];
  __shared__ double s_m1[kMcW], s_m2[kMcW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * kMcW + lane;
  const int ns = gridDim.y, rank = blockIdx.y;
  const int per = (s.B + ns - 1) / ns;
  const int b_beg = rank * per, b_end = min(s.B, b_beg + per);
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const double lm = s.has_o ? (double)lam[c] : 0.0;
  const double mu = (double)mean[c];
  const double r = (double)rstd[c];
  const double ga = (s.bn_mode != 0) ? (double)gamma[c] : 1.0;
  const double n = (double)s.B * (double)s.HW;

  // pass 1: dβ, dγ
  double s1 = 0.0, s2 = 0.0;
  if (s.bn_mode != 0) {
    for (int b = b_beg + w; b < b_end; b += kMcWarps) {
      const int64_t i = (int64_t)b * s.C + c;
      const double mb = drop_scale ? (double)drop_scale[b] : 1.0;
      const double a = gate[(int64_t)b * g + c / s.d];
      const double g1 = mb * gmom[i], gv = mb * gmom[BC + i];
      const double go = s.has_o ? mb * gmom[2 * BC + i] : 0.0;
      s1 += g1;
      s2 += a * gv + lm * go - mu * g1;
    }
  }
  const double p1 = mid_cta_sum(s1, red, w, lane);
  const double p2 = mid_cta_sum(s2, red, w, lane);
  if (w == 0) { part[0][lane] = p1; part[1][lane] = p2; }
  cluster.sync();
  if (w == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (int rr = 0; rr < ns; ++rr) {
      const double* rp = cluster.map_shared_rank(&part[0][0], rr);
      a1 += rp[lane];
      a2 += rp[kMcW + lane];
    }
    a2 *= r;
    if (rank == 0 && s.bn_mode != 0) {
      if (dbeta) dbeta[c] = (float)a1;
      if (dgamma) dgamma[c] = (float)a2;
    }
    const bool tr = (s.bn_mode == 1);
    s_m1[lane] = tr ? a1 / n : 0.0;
    s_m2[lane] = tr ? a2 / n : 0.0;
  }
  __syncthreads();
  const double m1 = s_m1[lane], m2 = s_m2[lane];
  const double gr = ga * r;

  // pass 2: dλ, head-summed da, sweep-B coefficients
  double sl = 0.0;
  for (int b = b_beg + w; b < b_end; b += kMcWarps) {
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
    float dah = (float)(gr * da);
    for (int off = s.d >> 1; off > 0; off >>= 1) dah += __shfl_xor_sync(0xffffffffu, dah, off);
    bcoef[6 * BC + i] = dah;
    bcoef[0 * BC + i] = (float)(gr * (-m1 + m2 * r * mu));
    bcoef[1 * BC + i] = (float)(gr * mb);
    bcoef[2 * BC + i] = (float)(-gr * m2 * r * a);
    bcoef[3 * BC + i] = (float)(-gr * m2 * r * lm);
    bcoef[4 * BC + i] = (float)a;
  }
  if (dlam != nullptr) {
    const double pl = mid_cta_sum(sl, red, w, lane);
    if (w == 0) part_l[lane] = pl;
    cluster.sync();
    if (w == 0 && rank == 0) {
      double al = 0.0;
      for (int rr = 0; rr < ns; ++rr) al += cluster.map_shared_rank(&part_l[0], rr)[lane];
      dlam[c] = (float)al;
    }
  }
  cluster.sync();   // no CTA leaves while a peer may still read its partials
}

// launch helper: grid (C/32, ns), cluster (1, ns, 1)
template <typename K, typename... Args>
inline cudaError_t launch_mid_cluster(K kernel, int C, int ns, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C / kMcW, ns, 1);
  cfg.blockDim = dim3(kMcW * kMcWarps, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = ns;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace mrla
