// MRLA-base block tail (softmax over the accumulated layer keys of a stage), SURVEY.md §8a rows A5/A6/A8.
//
//   forward   F0  v_t = dwconv3x3(x) -> cache slot t-1 ; Σx               (k_light_mom_fwd<FULL=0> + k_base_conv)
//             F1  q, k_t = xcorr(gap(x)) ; p = softmax_j(q·K_j/sqrt(d))    (k_base_attn)
//             F2  S = Σ_j p_j V_j ; ΣS, ΣS² per (b,c)                       (k_base_mix)
//             F3  BN statistics -> per-channel scale/shift                  (k_base_bn)
//             F4  Y = res*X + m_b * relu(cA*S + cD)                         (k_base_apply)
//   backward  B0  ΣdZ, ΣdZ·S per (b,c)        (dZ = m_b dY [Z>0])           (k_base_mom_bwd)
//             B1  dβ dγ, dS = E1 dZ + E0 + E2 S                             (k_base_bwd_chan)
//             B2  dV_j (+)= p_j dS ; dpm[j] = Σ dS V_j                      (k_base_scatter)
//             B3  softmax backward -> dq, dK (+)=, GAP grad, dwq/dwk        (k_base_bwd_attn)
//             B4  dX = res*dY + dwconv^T(dV_t) + dyc ; dWv partials         (k_base_dx)
// The V cache is never concatenated (the reference's torch.cat re-copies it every block,
// resnet/models/modules/mrla_base_module.py:65-70): slots are written in place and gradients w.r.t. the cached
// V_j / K_j are accumulated in place by the later blocks of the stage.
#pragma once
#include "common.cuh"
#include "light_sweeps.cuh"

namespace mrla {

constexpr int kBaseChunk = 8;  // cache slots handled per pass in k_base_scatter

struct BaseShape {
  int B, C, H, W, slots;
  int t;        // cache depth (slots 0..t-1 valid, this block wrote slot t-1)
  int d;        // channels per head
};

// ------------------------------------------------------------------------ F0: v_t = dwconv3x3(x)
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_conv(const T* __restrict__ x, T* __restrict__ v,
                                                   const float* __restrict__ wv, BaseShape s, int64_t bs_x,
                                                   int64_t bs_v) {
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float w9[9][CV];
  load_wv<CV>(wv, m, w9);
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    T* vb = v + (int64_t)b * bs_v + (int64_t)m.c * m.sC;
    RowTriple<T, CV> r0, r1, r2;
    load_row<T, CV>(xb, m, -1, s.H, s.W, r0);
    load_row<T, CV>(xb, m, 0, s.H, s.W, r1);
    for (int h = 0; h < s.H; ++h) {
      load_row<T, CV>(xb, m, h + 1, s.H, s.W, r2);
      float u[CV];
      conv_window<T, CV>(r0, r1, r2, w9, u);
      if (m.valid) st_vec<T, CV>(vb + (int64_t)h * m.sH + (int64_t)m.w * m.sW, u);
      r0 = r1; r1 = r2;
    }
  }
}

// ------------------------------------------------------------------------ F1: attention weights (one CTA per b)
// sx [B,C] ; q [B,C] ; kcache [B,t_cap,C] (row t-1 written here) ; p [B,g,t]
static __global__ void __launch_bounds__(1024) k_base_attn(const float* __restrict__ sx, const float* __restrict__ wq,
                                                           const float* __restrict__ wk, float* __restrict__ q,
                                                           float* __restrict__ kcache, float* __restrict__ p, int B,
                                                           int C, int HW, int d, int k, int t, int t_cap) {
  extern __shared__ float sm[];  // y[C] | q[C]
  float* ys = sm;
  float* qs = sm + C;
  const int b = blockIdx.x;
  const float inv_hw = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) ys[c] = sx[(int64_t)b * C + c] * inv_hw;
  __syncthreads();
  const int pad = (k - 1) / 2;
  float* krow = kcache + ((int64_t)b * t_cap + (t - 1)) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float qq = 0.f, kk = 0.f;
    for (int j = 0; j < k; ++j) {
      const int cc = c + j - pad;
      const float yv = (cc >= 0 && cc < C) ? ys[cc] : 0.f;
      qq = fmaf(wq[j], yv, qq);
      kk = fmaf(wk[j], yv, kk);
    }
    qs[c] = qq;
    q[(int64_t)b * C + c] = qq;
    krow[c] = kk;
  }
  __syncthreads();
  const int g = C / d;
  const float norm = rsqrtf((float)d);
  for (int h = threadIdx.x; h < g; h += blockDim.x) {
    float* ph = p + ((int64_t)b * g + h) * t;
    float mx = -INFINITY;
    for (int j = 0; j < t; ++j) {
      const float* kr = kcache + ((int64_t)b * t_cap + j) * C + h * d;
      float acc = 0.f;
      for (int i = 0; i < d; ++i) acc = fmaf(qs[h * d + i], kr[i], acc);
      acc *= norm;
      ph[j] = acc;
      mx = fmaxf(mx, acc);
    }
    float den = 0.f;
    for (int j = 0; j < t; ++j) {
      const float e = expf(ph[j] - mx);
      ph[j] = e;
      den += e;
    }
    const float inv = 1.f / den;
    for (int j = 0; j < t; ++j) ph[j] *= inv;
  }
}

// ------------------------------------------------------------------------ F2: S = Σ_j p_j V_j, moments of S
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_mix(const T* __restrict__ v, T* __restrict__ sout,
                                                  const float* __restrict__ p, float* __restrict__ smom, BaseShape s,
                                                  int64_t bs_v, int64_t ts_v, int64_t bs_s) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* vb = v + (int64_t)b * bs_v + (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    T* sb = sout + (int64_t)b * bs_s + (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    float acc[2 * CV];
#pragma unroll
    for (int i = 0; i < 2 * CV; ++i) acc[i] = 0.f;
    for (int h = 0; h < s.H; ++h) {
      float sv[CV];
#pragma unroll
      for (int i = 0; i < CV; ++i) sv[i] = 0.f;
      if (m.valid) {
        for (int j = 0; j < s.t; ++j) {
          float vv[CV];
          ld_vec<T, CV>(vb + (int64_t)j * ts_v + (int64_t)h * m.sH, vv);
#pragma unroll
          for (int i = 0; i < CV; ++i)
            sv[i] = fmaf(p[((int64_t)b * g + (m.c + i) / s.d) * s.t + j], vv[i], sv[i]);
        }
        st_vec<T, CV>(sb + (int64_t)h * m.sH, sv);
        // statistics are taken over the values BatchNorm will actually read (rounded to T)
#pragma unroll
        for (int i = 0; i < CV; ++i) {
          const float r = to_f<T>(from_f<T>(sv[i]));
          acc[i] += r;
          acc[CV + i] = fmaf(r, r, acc[CV + i]);
        }
      }
    }
    const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
    reduce_over_columns<2 * CV>(acc, smem, m, s.W, [&](int slot, int i, float sum) {
      const int mi = i / CV, vi = i - mi * CV;
      const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + vi);
      if (c < s.C) smom[(int64_t)mi * BC + (int64_t)b * s.C + c] = sum;
    });
  }
}

// ------------------------------------------------------------------------ F3: BN statistics of S (32 channels / CTA)
// chan = [4,C]: cA = γ r, cD = β − γ r μ, mean, rstd
static __global__ void __launch_bounds__(1024) k_base_bn(const float* __restrict__ smom, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float* __restrict__ chan, int B,
                                                         int C, int HW, int bn_mode, int update_running, float eps,
                                                         float momentum) {
  __shared__ double red1[32][33];
  __shared__ double red2[32][33];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool cok = c < C;
  const int64_t BC = (int64_t)B * C;
  double s1 = 0.0, s2 = 0.0;
  if (cok && bn_mode == 1) {
    for (int b = bl; b < B; b += 32) {
      s1 += (double)smom[(int64_t)b * C + c];
      s2 += (double)smom[BC + (int64_t)b * C + c];
    }
  }
  red1[bl][cl] = s1;
  red2[bl][cl] = s2;
  __syncthreads();
  if (bl != 0 || !cok) return;
  double mu = 0.0, r = 1.0;
  if (bn_mode == 1) {
    double a1 = 0.0, a2 = 0.0;
    for (int j = 0; j < 32; ++j) { a1 += red1[j][cl]; a2 += red2[j][cl]; }
    const double n = (double)B * (double)HW;
    mu = a1 / n;
    double var = a2 / n - mu * mu;
    if (var < 0.0) var = 0.0;
    r = 1.0 / sqrt(var + (double)eps);
    if (update_running && running_mean != nullptr) {
      const double unb = var * (n / fmax(n - 1.0, 1.0));
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + (double)momentum * mu);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + (double)momentum * unb);
    }
  } else if (bn_mode == 2) {
    mu = running_mean[c];
    r = 1.0 / sqrt((double)running_var[c] + (double)eps);
  }
  const double ga = bn_mode ? (double)gamma[c] : 1.0, be = bn_mode ? (double)beta[c] : 0.0;
  chan[c] = (float)(ga * r);
  chan[C + c] = (float)(be - ga * r * mu);
  chan[2 * C + c] = (float)mu;
  chan[3 * C + c] = (float)r;
}

// ------------------------------------------------------------------------ F4: Y = res*X + m_b*act(cA*S + cD)
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_apply(const T* __restrict__ x, const T* __restrict__ sin,
                                                    T* __restrict__ y, const float* __restrict__ chan,
                                                    const float* __restrict__ drop_scale, BaseShape s, int64_t bs_x,
                                                    int64_t bs_s, int64_t bs_y, float res, int relu) {
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float cA[CV], cD[CV];
#pragma unroll
  for (int i = 0; i < CV; ++i) {
    cA[i] = m.valid ? chan[m.c + i] : 0.f;
    cD[i] = m.valid ? chan[s.C + m.c + i] : 0.f;
  }
  if (!m.valid) return;
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const int64_t off = (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    const T* xb = x + (int64_t)b * bs_x + off;
    const T* sb = sin + (int64_t)b * bs_s + off;
    T* yb = y + (int64_t)b * bs_y + off;
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    for (int h = 0; h < s.H; ++h) {
      float xv[CV], sv[CV], out[CV];
      ld_vec<T, CV>(xb + (int64_t)h * m.sH, xv);
      ld_vec<T, CV>(sb + (int64_t)h * m.sH, sv);
#pragma unroll
      for (int i = 0; i < CV; ++i) {
        float z = fmaf(cA[i], sv[i], cD[i]);
        if (relu) z = fmaxf(z, 0.f);
        out[i] = fmaf(res, xv[i], mb * z);
      }
      st_vec<T, CV>(yb + (int64_t)h * m.sH, out);
    }
  }
}

// ------------------------------------------------------------------------ B0: ΣdZ, ΣdZ·S  (dZ = m_b dY [Z>0])
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_mom_bwd(const T* __restrict__ dy, const T* __restrict__ sin,
                                                      const float* __restrict__ chan,
                                                      const float* __restrict__ drop_scale, float* __restrict__ gmom,
                                                      BaseShape s, int64_t bs_dy, int64_t bs_s, int relu) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float cA[CV], cD[CV];
#pragma unroll
  for (int i = 0; i < CV; ++i) {
    cA[i] = m.valid ? chan[m.c + i] : 0.f;
    cD[i] = m.valid ? chan[s.C + m.c + i] : 0.f;
  }
  const int64_t BC = (int64_t)s.B * s.C;
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const int64_t off = (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    const T* gb = dy + (int64_t)b * bs_dy + off;
    const T* sb = sin + (int64_t)b * bs_s + off;
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    float acc[2 * CV];
#pragma unroll
    for (int i = 0; i < 2 * CV; ++i) acc[i] = 0.f;
    if (m.valid) {
      for (int h = 0; h < s.H; ++h) {
        float gv[CV], sv[CV];
        ld_vec<T, CV>(gb + (int64_t)h * m.sH, gv);
        ld_vec<T, CV>(sb + (int64_t)h * m.sH, sv);
#pragma unroll
        for (int i = 0; i < CV; ++i) {
          const float z = fmaf(cA[i], sv[i], cD[i]);
          const float dz = (relu && z <= 0.f) ? 0.f : mb * gv[i];
          acc[i] += dz;
          acc[CV + i] = fmaf(dz, sv[i], acc[CV + i]);
        }
      }
    }
    const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
    reduce_over_columns<2 * CV>(acc, smem, m, s.W, [&](int slot, int i, float sum) {
      const int mi = i / CV, vi = i - mi * CV;
      const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + vi);
      if (c < s.C) gmom[(int64_t)mi * BC + (int64_t)b * s.C + c] = sum;
    });
  }
}

// ------------------------------------------------------------------------ B1: per-channel BN backward (32 ch / CTA)
// bchan = [3,C]: E1 (dZ coef), E0 (const), E2 (S coef)   so that dS = E1*dZ + E0 + E2*S
static __global__ void __launch_bounds__(1024) k_base_bwd_chan(const float* __restrict__ gmom,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ chan, float* __restrict__ bchan,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               int B, int C, int HW, int bn_mode) {
  __shared__ double red1[32][33];
  __shared__ double red2[32][33];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool cok = c < C;
  const int64_t BC = (int64_t)B * C;
  double s1 = 0.0, s2 = 0.0;
  if (cok && bn_mode != 0) {
    for (int b = bl; b < B; b += 32) {
      s1 += (double)gmom[(int64_t)b * C + c];
      s2 += (double)gmom[BC + (int64_t)b * C + c];
    }
  }
  red1[bl][cl] = s1;
  red2[bl][cl] = s2;
  __syncthreads();
  if (bl != 0 || !cok) return;
  if (bn_mode == 0) {
    bchan[c] = 1.f; bchan[C + c] = 0.f; bchan[2 * C + c] = 0.f;
    return;
  }
  double a1 = 0.0, a2 = 0.0;
  for (int j = 0; j < 32; ++j) { a1 += red1[j][cl]; a2 += red2[j][cl]; }
  const double mu = chan[2 * C + c], r = chan[3 * C + c], ga = gamma[c];
  const double dbe = a1, dga = r * (a2 - mu * a1);
  if (dbeta) dbeta[c] = (float)dbe;
  if (dgamma) dgamma[c] = (float)dga;
  const double n = (double)B * (double)HW;
  const double m1 = bn_mode == 1 ? dbe / n : 0.0, m2 = bn_mode == 1 ? dga / n : 0.0;
  bchan[c] = (float)(ga * r);
  bchan[C + c] = (float)(ga * r * (-m1 + m2 * r * mu));
  bchan[2 * C + c] = (float)(-ga * r * r * m2);
}

// ------------------------------------------------------------------------ B2: dV_j (+)= p_j dS ; dpm[j] = Σ dS V_j
// one pass handles cache slots [j0, j0+kBaseChunk)
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_scatter(const T* __restrict__ dy, const T* __restrict__ sin,
                                                      const T* __restrict__ v, T* __restrict__ dv,
                                                      const float* __restrict__ p, const float* __restrict__ chan,
                                                      const float* __restrict__ bchan,
                                                      const float* __restrict__ drop_scale, float* __restrict__ dpm,
                                                      BaseShape s, int j0, int accumulate, int64_t bs_dy, int64_t bs_s,
                                                      int64_t bs_v, int64_t ts_v, int64_t bs_dv, int64_t ts_dv,
                                                      int relu) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  const int g = s.C / s.d;
  const int64_t BC = (int64_t)s.B * s.C;
  const int nj = min(kBaseChunk, s.t - j0);
  float cA[CV], cD[CV], e1[CV], e0[CV], e2[CV];
#pragma unroll
  for (int i = 0; i < CV; ++i) {
    cA[i] = m.valid ? chan[m.c + i] : 0.f;
    cD[i] = m.valid ? chan[s.C + m.c + i] : 0.f;
    e1[i] = m.valid ? bchan[m.c + i] : 0.f;
    e0[i] = m.valid ? bchan[s.C + m.c + i] : 0.f;
    e2[i] = m.valid ? bchan[2 * s.C + m.c + i] : 0.f;
  }
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const int64_t off = (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    const T* gb = dy + (int64_t)b * bs_dy + off;
    const T* sb = sin + (int64_t)b * bs_s + off;
    const T* vb = v + (int64_t)b * bs_v + off;
    T* dvb = dv + (int64_t)b * bs_dv + off;
    const float mb = drop_scale ? drop_scale[b] : 1.f;
    float pj[kBaseChunk][CV];
    float acc[kBaseChunk * CV];
#pragma unroll
    for (int j = 0; j < kBaseChunk; ++j)
#pragma unroll
      for (int i = 0; i < CV; ++i) {
        pj[j][i] = (m.valid && j < nj) ? p[((int64_t)b * g + (m.c + i) / s.d) * s.t + j0 + j] : 0.f;
        acc[j * CV + i] = 0.f;
      }
    if (m.valid) {
      for (int h = 0; h < s.H; ++h) {
        float gv[CV], sv[CV], ds[CV];
        ld_vec<T, CV>(gb + (int64_t)h * m.sH, gv);
        ld_vec<T, CV>(sb + (int64_t)h * m.sH, sv);
#pragma unroll
        for (int i = 0; i < CV; ++i) {
          const float z = fmaf(cA[i], sv[i], cD[i]);
          const float dz = (relu && z <= 0.f) ? 0.f : mb * gv[i];
          ds[i] = fmaf(e1[i], dz, fmaf(e2[i], sv[i], e0[i]));
        }
#pragma unroll
        for (int j = 0; j < kBaseChunk; ++j) {
          if (j < nj) {
            float vv[CV], dvv[CV];
            ld_vec<T, CV>(vb + (int64_t)(j0 + j) * ts_v + (int64_t)h * m.sH, vv);
            T* dst = dvb + (int64_t)(j0 + j) * ts_dv + (int64_t)h * m.sH;
            if (accumulate) ld_vec<T, CV>(dst, dvv);
#pragma unroll
            for (int i = 0; i < CV; ++i) {
              acc[j * CV + i] = fmaf(ds[i], vv[i], acc[j * CV + i]);
              dvv[i] = accumulate ? fmaf(pj[j][i], ds[i], dvv[i]) : pj[j][i] * ds[i];
            }
            st_vec<T, CV>(dst, dvv);
          }
        }
      }
    }
    const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
    reduce_over_columns<kBaseChunk * CV>(acc, smem, m, s.W, [&](int slot, int i, float sum) {
      const int j = i / CV, vi = i - j * CV;
      const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + vi);
      if (c < s.C && j < nj) dpm[(int64_t)(j0 + j) * BC + (int64_t)b * s.C + c] = sum;
    });
  }
}

// ------------------------------------------------------------------------ B3: softmax backward (one CTA per b)
// dp[h,j] = Σ_{c in h} dpm[j,b,c] ; dl_j = p_j (dp_j − Σ_i p_i dp_i) / sqrt(d)
// dq[c] = Σ_j dl[h,j] K[j,c] ; dK[j,c] (+)= dl[h,j] q[c] ; dk_t = dK[t-1] (total)
// dyc[b,c] = (xcorrᵀ(dq,wq) + xcorrᵀ(dk_t,wk))[c] / HW ; per-b partials of dwq, dwk -> wqk_part[B,2k]
static __global__ void __launch_bounds__(1024) k_base_bwd_attn(const float* __restrict__ sx, const float* __restrict__ q,
                                                               const float* __restrict__ kcache,
                                                               float* __restrict__ dkcache, const float* __restrict__ p,
                                                               const float* __restrict__ dpm, const float* __restrict__ wq,
                                                               const float* __restrict__ wk, float* __restrict__ dyc,
                                                               float* __restrict__ wqk_part, int B, int C, int HW, int d,
                                                               int k, int t, int t_cap, int accumulate) {
  extern __shared__ float sm[];  // y[C] | dq[C] | dk[C] | dl[g*t]
  float* ys = sm;
  float* dq = sm + C;
  float* dk = sm + 2 * C;
  float* dl = sm + 3 * C;
  const int b = blockIdx.x;
  const int g = C / d;
  const int64_t BC = (int64_t)B * C;
  const float inv_hw = 1.f / (float)HW;
  const float norm = rsqrtf((float)d);
  for (int c = threadIdx.x; c < C; c += blockDim.x) ys[c] = sx[(int64_t)b * C + c] * inv_hw;
  for (int h = threadIdx.x; h < g; h += blockDim.x) {
    const float* ph = p + ((int64_t)b * g + h) * t;
    float dot = 0.f;
    for (int j = 0; j < t; ++j) {
      float dp = 0.f;
      for (int i = 0; i < d; ++i) dp += dpm[(int64_t)j * BC + (int64_t)b * C + h * d + i];
      dl[h * t + j] = dp;
      dot = fmaf(ph[j], dp, dot);
    }
    for (int j = 0; j < t; ++j) dl[h * t + j] = ph[j] * (dl[h * t + j] - dot) * norm;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int h = c / d;
    const float qc = q[(int64_t)b * C + c];
    float acc = 0.f;
    for (int j = 0; j < t; ++j) {
      const int64_t ki = ((int64_t)b * t_cap + j) * C + c;
      const float dlj = dl[h * t + j];
      acc = fmaf(dlj, kcache[ki], acc);
      float dkv = dlj * qc;
      if (accumulate) dkv += dkcache[ki];
      dkcache[ki] = dkv;
      if (j == t - 1) dk[c] = dkv;
    }
    dq[c] = acc;
  }
  __syncthreads();
  const int pad = (k - 1) / 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const int cc = c - j + pad;
      if (cc >= 0 && cc < C) acc = fmaf(wq[j], dq[cc], fmaf(wk[j], dk[cc], acc));
    }
    dyc[(int64_t)b * C + c] = acc * inv_hw;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < 2 * k; j += nw) {
    const int jj = (j < k) ? j : j - k;
    const float* src = (j < k) ? dq : dk;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const int cc = c + jj - pad;
      if (cc >= 0 && cc < C) acc = fmaf(ys[cc], src[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) wqk_part[(int64_t)b * 2 * k + j] = acc;
  }
}

// ------------------------------------------------------------------------ B4: dX = res*dY + dwconv^T(dV_t) + dyc ; dWv
template <typename T, int LAYOUT, int CV>
__global__ void __launch_bounds__(512) k_base_dx(const T* __restrict__ dy, const T* __restrict__ x,
                                                 const T* __restrict__ dvt, T* __restrict__ dx,
                                                 const float* __restrict__ wv, const float* __restrict__ dyc,
                                                 float* __restrict__ wv_part, BaseShape s, int64_t bs_dy, int64_t bs_x,
                                                 int64_t bs_dv, int64_t bs_dx, float res) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float w9[9][CV];
  load_wv<CV>(wv, m, w9);
  float dwacc[9 * CV];
#pragma unroll
  for (int i = 0; i < 9 * CV; ++i) dwacc[i] = 0.f;
  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    const T* tb = dvt + (int64_t)b * bs_dv + (int64_t)m.c * m.sC;
    const T* gb = dy + (int64_t)b * bs_dy + (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    T* dxb = dx + (int64_t)b * bs_dx + (int64_t)m.c * m.sC + (int64_t)m.w * m.sW;
    float gap[CV];
#pragma unroll
    for (int i = 0; i < CV; ++i) gap[i] = m.valid ? dyc[(int64_t)b * s.C + m.c + i] : 0.f;
    RowTriple<T, CV> x0, x1, x2, t0, t1, t2;
    load_row<T, CV>(xb, m, -1, s.H, s.W, x0);
    load_row<T, CV>(xb, m, 0, s.H, s.W, x1);
    load_row<T, CV>(tb, m, -1, s.H, s.W, t0);
    load_row<T, CV>(tb, m, 0, s.H, s.W, t1);
    for (int h = 0; h < s.H; ++h) {
      load_row<T, CV>(xb, m, h + 1, s.H, s.W, x2);
      load_row<T, CV>(tb, m, h + 1, s.H, s.W, t2);
      float gv[CV], out[CV];
      ld_vec_pred<T, CV>(gb + (int64_t)h * m.sH, m.valid, gv);
#pragma unroll
      for (int i = 0; i < CV; ++i) {
        // dX[h][w] = Σ_ij wv[i][j] * T[h-i+1][w-j+1]
        float a = fmaf(res, gv[i], gap[i]);
        a = fmaf(w9[0][i], t2.r[i], a);
        a = fmaf(w9[1][i], t2.c[i], a);
        a = fmaf(w9[2][i], t2.l[i], a);
        a = fmaf(w9[3][i], t1.r[i], a);
        a = fmaf(w9[4][i], t1.c[i], a);
        a = fmaf(w9[5][i], t1.l[i], a);
        a = fmaf(w9[6][i], t0.r[i], a);
        a = fmaf(w9[7][i], t0.c[i], a);
        a = fmaf(w9[8][i], t0.l[i], a);
        out[i] = a;
        // dWv[i][j] += T[h][w] * x[h+i-1][w+j-1]
        const float tc = t1.c[i];
        dwacc[0 * CV + i] = fmaf(tc, x0.l[i], dwacc[0 * CV + i]);
        dwacc[1 * CV + i] = fmaf(tc, x0.c[i], dwacc[1 * CV + i]);
        dwacc[2 * CV + i] = fmaf(tc, x0.r[i], dwacc[2 * CV + i]);
        dwacc[3 * CV + i] = fmaf(tc, x1.l[i], dwacc[3 * CV + i]);
        dwacc[4 * CV + i] = fmaf(tc, x1.c[i], dwacc[4 * CV + i]);
        dwacc[5 * CV + i] = fmaf(tc, x1.r[i], dwacc[5 * CV + i]);
        dwacc[6 * CV + i] = fmaf(tc, x2.l[i], dwacc[6 * CV + i]);
        dwacc[7 * CV + i] = fmaf(tc, x2.c[i], dwacc[7 * CV + i]);
        dwacc[8 * CV + i] = fmaf(tc, x2.r[i], dwacc[8 * CV + i]);
      }
      if (m.valid) st_vec<T, CV>(dxb + (int64_t)h * m.sH, out);
      x0 = x1; x1 = x2; t0 = t1; t1 = t2;
    }
  }
  const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
  reduce_over_columns<9 * CV>(dwacc, smem, m, s.W, [&](int slot, int i, float sum) {
    const int tap = i / CV, vi = i - tap * CV;
    const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + vi);
    if (c < s.C) wv_part[((int64_t)blockIdx.y * s.C + c) * 9 + tap] = sum;
  });
}

}  // namespace mrla
