// The four N-sized sweeps of the MRLA-light block tail (SURVEY.md §8a closed form).
//
//   sweep 1  k_light_mom_fwd   reads x,o          -> per-(b,c) moments  Σx ΣV ΣV² ΣVo Σo Σo²
//   sweep 2  k_light_apply_fwd reads x,o          -> y = res*x + A*act(dwconv x) + L*o + D
//   sweep A  k_light_mom_bwd   reads dy,x,o       -> per-(b,c) Σdy Σdy·V Σdy·o
//   sweep B  k_light_apply_bwd reads dy,x,o       -> dx, do, dWv partials
//
// V = dwconv3x3(x) is never materialised: every sweep recomputes it from a 3x3 register
// window while the CTA "marches" down the image rows.  One thread owns one image column of
// CV channels (NCHW: CV = 1, lanes run along w; NHWC: CV channels = one vector, lanes run
// along c), so the new row of the window is three loads (w-1, w, w+1) of which two hit L1.
// Replaces the ATen sequence adaptive_avg_pool2d / conv2d(groups=C) / mul / add /
// batch_norm / add (and its autograd graph) issued by
// resnet/models/modules/mrla_light_module.py:56-72 and resnet_mrla_light.py:42,116.
#pragma once
#include "common.cuh"

namespace mrla {

struct SweepShape {
  int B, C, H, W;
  int slots;  // slots per CTA (planes for NCHW, channel-vector lanes for NHWC)
};

// -------------------------------------------------------------------------------- window
template <typename T, int CV>
struct RowTriple {  // x[h][w-1], x[h][w], x[h][w+1]
  float l[CV], c[CV], r[CV];
};

template <typename T, int CV>
__device__ __forceinline__ void load_row(const T* __restrict__ base, const March& m, int h, int H, int W,
                                         RowTriple<T, CV>& t) {
  const bool row_ok = m.valid && h >= 0 && h < H;
  const T* p = base + (int64_t)h * m.sH + (int64_t)m.w * m.sW;
  ld_vec_pred<T, CV>(p - m.sW, row_ok && m.w > 0, t.l);
  ld_vec_pred<T, CV>(p, row_ok, t.c);
  ld_vec_pred<T, CV>(p + m.sW, row_ok && m.w + 1 < W, t.r);
}

template <int CV>
__device__ __forceinline__ void load_wv(const float* __restrict__ wv, const March& m, float (&w9)[9][CV]) {
#pragma unroll
  for (int v = 0; v < CV; ++v)
#pragma unroll
    for (int i = 0; i < 9; ++i) w9[i][v] = m.valid ? wv[(int64_t)(m.c + v) * 9 + i] : 0.f;
}

template <typename T, int CV>
__device__ __forceinline__ void conv_window(const RowTriple<T, CV>& r0, const RowTriple<T, CV>& r1,
                                            const RowTriple<T, CV>& r2, const float (&w9)[9][CV], float (&u)[CV]) {
#pragma unroll
  for (int v = 0; v < CV; ++v) {
    float s = w9[0][v] * r0.l[v];
    s = fmaf(w9[1][v], r0.c[v], s);
    s = fmaf(w9[2][v], r0.r[v], s);
    s = fmaf(w9[3][v], r1.l[v], s);
    s = fmaf(w9[4][v], r1.c[v], s);
    s = fmaf(w9[5][v], r1.r[v], s);
    s = fmaf(w9[6][v], r2.l[v], s);
    s = fmaf(w9[7][v], r2.c[v], s);
    s = fmaf(w9[8][v], r2.r[v], s);
    u[v] = s;
  }
}

// -------------------------------------------------------------------------------- sweep 1
// FULL = 0: only Σx (GAP) is produced (no BN statistics needed: layer-only / eval-BN / DeiT).
template <typename T, int LAYOUT, int CV, int ACT, bool HAS_O, bool FULL>
__global__ void __launch_bounds__(512) k_light_mom_fwd(const T* __restrict__ x, const T* __restrict__ o,
                                                       const float* __restrict__ wv, float* __restrict__ mom,
                                                       SweepShape s, int64_t bs_x, int64_t bs_o) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float w9[9][CV];
  if (FULL) load_wv<CV>(wv, m, w9);
  const int64_t BC = (int64_t)s.B * s.C;
  constexpr int NM = FULL ? (HAS_O ? 6 : 3) : 1;

  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    const T* ob = HAS_O ? o + (int64_t)b * bs_o + (int64_t)m.c * m.sC : nullptr;
    float acc[NM * CV];
#pragma unroll
    for (int i = 0; i < NM * CV; ++i) acc[i] = 0.f;

    if (FULL) {
      RowTriple<T, CV> r0, r1, r2, pend;
      load_row<T, CV>(xb, m, -1, s.H, s.W, r0);
      load_row<T, CV>(xb, m, 0, s.H, s.W, r1);
      load_row<T, CV>(xb, m, 1, s.H, s.W, r2);
      float oc[CV], on[CV];
      if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)m.w * m.sW, m.valid, oc);
      for (int h = 0; h < s.H; ++h) {
        load_row<T, CV>(xb, m, h + 2, s.H, s.W, pend);  // prefetch two rows ahead
        if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)(h + 1) * m.sH + (int64_t)m.w * m.sW, m.valid && h + 1 < s.H, on);
        float u[CV];
        conv_window<T, CV>(r0, r1, r2, w9, u);
#pragma unroll
        for (int v = 0; v < CV; ++v) {
          const float vv = act_fwd<ACT>(u[v]);
          acc[0 * CV + v] += r1.c[v];
          acc[1 * CV + v] += vv;
          acc[2 * CV + v] = fmaf(vv, vv, acc[2 * CV + v]);
          if (HAS_O) {
            acc[3 * CV + v] = fmaf(vv, oc[v], acc[3 * CV + v]);
            acc[4 * CV + v] += oc[v];
            acc[5 * CV + v] = fmaf(oc[v], oc[v], acc[5 * CV + v]);
          }
        }
        r0 = r1; r1 = r2; r2 = pend;
        if (HAS_O) {
#pragma unroll
          for (int v = 0; v < CV; ++v) oc[v] = on[v];
        }
      }
    } else {
      const T* p = xb + (int64_t)m.w * m.sW;
      for (int h = 0; h < s.H; ++h) {
        float xc[CV];
        ld_vec_pred<T, CV>(p + (int64_t)h * m.sH, m.valid, xc);
#pragma unroll
        for (int v = 0; v < CV; ++v) acc[v] += xc[v];
      }
    }

    const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
    reduce_over_columns<NM * CV>(acc, smem, m, s.W, [&](int slot, int i, float sum) {
      const int mi = i / CV, v = i - mi * CV;
      const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + v);
      if (c < s.C) {
        // moment order in memory: 0 Σx, 1 ΣV, 2 ΣV², 3 ΣVo, 4 Σo, 5 Σo²
        mom[(int64_t)mi * BC + (int64_t)b * s.C + c] = sum;
      }
    });
  }
}

// -------------------------------------------------------------------------------- sweep 2
// y = res*x + A[b,c]*act(dwconv x) + L[b,c]*o + D[b,c];   coef = [3,B,C] (A, L, D)
template <typename T, int LAYOUT, int CV, int ACT, bool HAS_O>
__global__ void __launch_bounds__(512) k_light_apply_fwd(const T* __restrict__ x, const T* __restrict__ o,
                                                         T* __restrict__ y, const float* __restrict__ wv,
                                                         const float* __restrict__ coef, SweepShape s, int64_t bs_x,
                                                         int64_t bs_o, int64_t bs_y, float res) {
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float w9[9][CV];
  load_wv<CV>(wv, m, w9);
  const int64_t BC = (int64_t)s.B * s.C;

  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    const T* ob = HAS_O ? o + (int64_t)b * bs_o + (int64_t)m.c * m.sC : nullptr;
    T* yb = y + (int64_t)b * bs_y + (int64_t)m.c * m.sC;
    float cA[CV], cL[CV], cD[CV];
#pragma unroll
    for (int v = 0; v < CV; ++v) { cA[v] = 0.f; cL[v] = 0.f; cD[v] = 0.f; }
    if (m.valid) {
      const float* cp = coef + (int64_t)b * s.C + m.c;
      ld_f32<CV>(cp, cA);
      if (HAS_O) ld_f32<CV>(cp + BC, cL);
      ld_f32<CV>(cp + 2 * BC, cD);
    }
    RowTriple<T, CV> r0, r1, r2, pend;
    load_row<T, CV>(xb, m, -1, s.H, s.W, r0);
    load_row<T, CV>(xb, m, 0, s.H, s.W, r1);
    load_row<T, CV>(xb, m, 1, s.H, s.W, r2);
    float oc[CV], on[CV];
    if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)m.w * m.sW, m.valid, oc);
    for (int h = 0; h < s.H; ++h) {
      load_row<T, CV>(xb, m, h + 2, s.H, s.W, pend);
      if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)(h + 1) * m.sH + (int64_t)m.w * m.sW, m.valid && h + 1 < s.H, on);
      float u[CV], out[CV];
      conv_window<T, CV>(r0, r1, r2, w9, u);
#pragma unroll
      for (int v = 0; v < CV; ++v) {
        float t = fmaf(cA[v], act_fwd<ACT>(u[v]), cD[v]);
        if (HAS_O) t = fmaf(cL[v], oc[v], t);
        out[v] = fmaf(res, r1.c[v], t);
      }
      if (m.valid) st_vec<T, CV>(yb + (int64_t)h * m.sH + (int64_t)m.w * m.sW, out);
      r0 = r1; r1 = r2; r2 = pend;
      if (HAS_O) {
#pragma unroll
        for (int v = 0; v < CV; ++v) oc[v] = on[v];
      }
    }
  }
}

// -------------------------------------------------------------------------------- sweep A
// gmom = [3,B,C]: Σdy, Σdy·V, Σdy·o  (V = act(dwconv x)); NEED_V = 0 skips the conv (only Σdy, Σdy·o).
template <typename T, int LAYOUT, int CV, int ACT, bool HAS_O>
__global__ void __launch_bounds__(512) k_light_mom_bwd(const T* __restrict__ dy, const T* __restrict__ x,
                                                       const T* __restrict__ o, const float* __restrict__ wv,
                                                       float* __restrict__ gmom, SweepShape s, int64_t bs_dy,
                                                       int64_t bs_x, int64_t bs_o) {
  extern __shared__ float smem[];
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  float w9[9][CV];
  load_wv<CV>(wv, m, w9);
  const int64_t BC = (int64_t)s.B * s.C;
  constexpr int NM = HAS_O ? 3 : 2;

  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    const T* ob = HAS_O ? o + (int64_t)b * bs_o + (int64_t)m.c * m.sC : nullptr;
    const T* gb = dy + (int64_t)b * bs_dy + (int64_t)m.c * m.sC;
    float acc[NM * CV];
#pragma unroll
    for (int i = 0; i < NM * CV; ++i) acc[i] = 0.f;
    RowTriple<T, CV> r0, r1, r2, pend;
    load_row<T, CV>(xb, m, -1, s.H, s.W, r0);
    load_row<T, CV>(xb, m, 0, s.H, s.W, r1);
    load_row<T, CV>(xb, m, 1, s.H, s.W, r2);
    float oc[CV], on[CV], gc[CV], gn[CV];
    if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)m.w * m.sW, m.valid, oc);
    ld_vec_pred<T, CV>(gb + (int64_t)m.w * m.sW, m.valid, gc);
    for (int h = 0; h < s.H; ++h) {
      load_row<T, CV>(xb, m, h + 2, s.H, s.W, pend);
      const int64_t off = (int64_t)(h + 1) * m.sH + (int64_t)m.w * m.sW;
      const bool nok = m.valid && h + 1 < s.H;
      if (HAS_O) ld_vec_pred<T, CV>(ob + off, nok, on);
      ld_vec_pred<T, CV>(gb + off, nok, gn);
      float u[CV];
      conv_window<T, CV>(r0, r1, r2, w9, u);
#pragma unroll
      for (int v = 0; v < CV; ++v) {
        const float vv = act_fwd<ACT>(u[v]);
        acc[0 * CV + v] += gc[v];
        acc[1 * CV + v] = fmaf(gc[v], vv, acc[1 * CV + v]);
        if (HAS_O) acc[2 * CV + v] = fmaf(gc[v], oc[v], acc[2 * CV + v]);
      }
      r0 = r1; r1 = r2; r2 = pend;
#pragma unroll
      for (int v = 0; v < CV; ++v) { gc[v] = gn[v]; if (HAS_O) oc[v] = on[v]; }
    }
    const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
    reduce_over_columns<NM * CV>(acc, smem, m, s.W, [&](int slot, int i, float sum) {
      const int mi = i / CV, v = i - mi * CV;
      const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + v);
      if (c < s.C) gmom[(int64_t)mi * BC + (int64_t)b * s.C + c] = sum;
    });
  }
}

// -------------------------------------------------------------------------------- sweep B
// bcoef = [7,B,C]: Q0,Q1,Q2,Q3 (dS = Q0 + Q1*dy + Q2*V + Q3*o), Ta (T = Ta*dS*act'(U)), dyc (GAP grad / HW)
//   do = lam[c]*dS ;  dx = res*dy + dwconv3x3^T(T) + dyc ;  dWv[c,i,j] = Σ T[h,w]*x[h+i-1,w+j-1]
// T rows are exchanged between the column-threads of a CTA through a 2-slot shared-memory ring.
// wv_part = [gridDim.y, C, 9] per-CTA-row partial sums of dWv (reduced by k_light_finish).
template <typename T, int LAYOUT, int CV, int ACT, bool HAS_O>
__global__ void __launch_bounds__(512) k_light_apply_bwd(const T* __restrict__ dy, const T* __restrict__ x,
                                                         const T* __restrict__ o, T* __restrict__ dx,
                                                         T* __restrict__ dout, const float* __restrict__ wv,
                                                         const float* __restrict__ lam,
                                                         const float* __restrict__ bcoef, float* __restrict__ wv_part,
                                                         SweepShape s, int64_t bs_dy, int64_t bs_x, int64_t bs_o,
                                                         int64_t bs_dx, int64_t bs_do, float res) {
  extern __shared__ float smem[];  // max(2*blockDim*CV, blockDim*9*CV) floats
  const March m = make_march<LAYOUT>(s.C, s.H, s.W, CV, s.slots);
  const int tid = threadIdx.x;
  float w9[9][CV];
  load_wv<CV>(wv, m, w9);
  float lm[CV];
#pragma unroll
  for (int v = 0; v < CV; ++v) lm[v] = (HAS_O && m.valid) ? lam[m.c + v] : 0.f;
  const int64_t BC = (int64_t)s.B * s.C;
  float dwacc[9 * CV];
#pragma unroll
  for (int i = 0; i < 9 * CV; ++i) dwacc[i] = 0.f;
  const bool has_l = m.valid && m.w > 0;
  const bool has_r = m.valid && m.w + 1 < s.W;

  for (int b = blockIdx.y; b < s.B; b += gridDim.y) {
    const T* xb = x + (int64_t)b * bs_x + (int64_t)m.c * m.sC;
    const T* ob = HAS_O ? o + (int64_t)b * bs_o + (int64_t)m.c * m.sC : nullptr;
    const T* gb = dy + (int64_t)b * bs_dy + (int64_t)m.c * m.sC;
    T* dxb = dx + (int64_t)b * bs_dx + (int64_t)m.c * m.sC;
    T* dob = HAS_O ? dout + (int64_t)b * bs_do + (int64_t)m.c * m.sC : nullptr;
    float q0[CV], q1[CV], q2[CV], q3[CV], ta[CV], dyc[CV];
#pragma unroll
    for (int v = 0; v < CV; ++v) { q0[v] = q1[v] = q2[v] = q3[v] = ta[v] = dyc[v] = 0.f; }
    if (m.valid) {
      const float* cp = bcoef + (int64_t)b * s.C + m.c;
      ld_f32<CV>(cp, q0);
      ld_f32<CV>(cp + BC, q1);
      ld_f32<CV>(cp + 2 * BC, q2);
      if (HAS_O) ld_f32<CV>(cp + 3 * BC, q3);
      ld_f32<CV>(cp + 4 * BC, ta);
      ld_f32<CV>(cp + 5 * BC, dyc);
    }
    // x window rows r-1, r, r+1 ; T window rows r-2, r-1, r (columns l,c,r)
    RowTriple<T, CV> r0, r1, r2, pend;
    load_row<T, CV>(xb, m, -1, s.H, s.W, r0);
    load_row<T, CV>(xb, m, 0, s.H, s.W, r1);
    load_row<T, CV>(xb, m, 1, s.H, s.W, r2);
    float t0[3][CV], t1[3][CV], t2[3][CV];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int v = 0; v < CV; ++v) { t0[j][v] = 0.f; t1[j][v] = 0.f; t2[j][v] = 0.f; }
    float gprev[CV], gc[CV], gn[CV], oc[CV], on[CV];
#pragma unroll
    for (int v = 0; v < CV; ++v) { gprev[v] = 0.f; oc[v] = 0.f; on[v] = 0.f; }
    ld_vec_pred<T, CV>(gb + (int64_t)m.w * m.sW, m.valid, gc);
    if (HAS_O) ld_vec_pred<T, CV>(ob + (int64_t)m.w * m.sW, m.valid, oc);

    for (int r = 0; r <= s.H; ++r) {
      float tn[CV];
      if (r < s.H) {
        load_row<T, CV>(xb, m, r + 2, s.H, s.W, pend);
        const int64_t off = (int64_t)(r + 1) * m.sH + (int64_t)m.w * m.sW;
        const bool nok = m.valid && r + 1 < s.H;
        ld_vec_pred<T, CV>(gb + off, nok, gn);
        if (HAS_O) ld_vec_pred<T, CV>(ob + off, nok, on);
        float u[CV], dov[CV];
        conv_window<T, CV>(r0, r1, r2, w9, u);
#pragma unroll
        for (int v = 0; v < CV; ++v) {
          const float vv = act_fwd<ACT>(u[v]);
          float ds = fmaf(q1[v], gc[v], q0[v]);
          ds = fmaf(q2[v], vv, ds);
          if (HAS_O) ds = fmaf(q3[v], oc[v], ds);
          dov[v] = lm[v] * ds;
          tn[v] = ta[v] * ds * act_grad<ACT>(u[v]);
          // dWv[i][j] += T[r][w] * x[r+i-1][w+j-1]
          dwacc[0 * CV + v] = fmaf(tn[v], r0.l[v], dwacc[0 * CV + v]);
          dwacc[1 * CV + v] = fmaf(tn[v], r0.c[v], dwacc[1 * CV + v]);
          dwacc[2 * CV + v] = fmaf(tn[v], r0.r[v], dwacc[2 * CV + v]);
          dwacc[3 * CV + v] = fmaf(tn[v], r1.l[v], dwacc[3 * CV + v]);
          dwacc[4 * CV + v] = fmaf(tn[v], r1.c[v], dwacc[4 * CV + v]);
          dwacc[5 * CV + v] = fmaf(tn[v], r1.r[v], dwacc[5 * CV + v]);
          dwacc[6 * CV + v] = fmaf(tn[v], r2.l[v], dwacc[6 * CV + v]);
          dwacc[7 * CV + v] = fmaf(tn[v], r2.c[v], dwacc[7 * CV + v]);
          dwacc[8 * CV + v] = fmaf(tn[v], r2.r[v], dwacc[8 * CV + v]);
        }
        if (HAS_O && m.valid) st_vec<T, CV>(dob + (int64_t)r * m.sH + (int64_t)m.w * m.sW, dov);
      } else {
#pragma unroll
        for (int v = 0; v < CV; ++v) tn[v] = 0.f;
      }
      // exchange T[r] with the neighbouring columns
      float* ring = smem + (size_t)(r & 1) * blockDim.x * CV;
#pragma unroll
      for (int v = 0; v < CV; ++v) ring[tid * CV + v] = m.valid ? tn[v] : 0.f;
      __syncthreads();
#pragma unroll
      for (int v = 0; v < CV; ++v) {
        t2[0][v] = has_l ? ring[(tid - m.tW) * CV + v] : 0.f;
        t2[1][v] = tn[v];
        t2[2][v] = has_r ? ring[(tid + m.tW) * CV + v] : 0.f;
      }
      if (r >= 1) {
        // dx[r-1][w] = res*dy + dyc + Σ_ij wv[i][j] * T[r-i][w-j+1]
        float out[CV];
#pragma unroll
        for (int v = 0; v < CV; ++v) {
          float a = fmaf(res, gprev[v], dyc[v]);
          a = fmaf(w9[0][v], t2[2][v], a);
          a = fmaf(w9[1][v], t2[1][v], a);
          a = fmaf(w9[2][v], t2[0][v], a);
          a = fmaf(w9[3][v], t1[2][v], a);
          a = fmaf(w9[4][v], t1[1][v], a);
          a = fmaf(w9[5][v], t1[0][v], a);
          a = fmaf(w9[6][v], t0[2][v], a);
          a = fmaf(w9[7][v], t0[1][v], a);
          a = fmaf(w9[8][v], t0[0][v], a);
          out[v] = a;
        }
        if (m.valid) st_vec<T, CV>(dxb + (int64_t)(r - 1) * m.sH + (int64_t)m.w * m.sW, out);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int v = 0; v < CV; ++v) { t0[j][v] = t1[j][v]; t1[j][v] = t2[j][v]; }
      r0 = r1; r1 = r2; r2 = pend;
#pragma unroll
      for (int v = 0; v < CV; ++v) { gprev[v] = gc[v]; gc[v] = gn[v]; if (HAS_O) oc[v] = on[v]; }
    }
    __syncthreads();  // ring slots are reused by the next sample
  }

  const int cbase = (LAYOUT == 0) ? blockIdx.x * s.slots : blockIdx.x * s.slots * CV;
  reduce_over_columns<9 * CV>(dwacc, smem, m, s.W, [&](int slot, int i, float sum) {
    const int tap = i / CV, v = i - tap * CV;
    const int c = cbase + ((LAYOUT == 0) ? slot : slot * CV + v);
    if (c < s.C) wv_part[((int64_t)blockIdx.y * s.C + c) * 9 + tap] = sum;
  });
}

}  // namespace mrla
