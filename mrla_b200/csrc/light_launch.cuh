// Host-side launch logic for the MRLA-light tail (one translation unit per activation dtype).
#pragma once
#include "../../include/mrla_b200.h"
#include "light_mid.cuh"
#include "light_sweeps.cuh"

namespace mrla {

extern thread_local int g_launch_count;

struct LightPlan {
  int slots;      // slots per CTA
  int threads;    // CTA size (multiple of 32)
  int grid_x;     // channel groups
  int grid_y;     // batch chunks (grid-stride over b)
};

constexpr int kNumSMs = 148;
constexpr int kMaxThreads = 512;

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// NCHW: CTA = P planes x W columns.  NHWC: CTA = W columns x LP channel-vector lanes.
inline int make_plan(int layout, int B, int C, int W, int cv, LightPlan* p) {
  if (W > kMaxThreads) return MRLA_ERR_SHAPE;
  if (layout == MRLA_NCHW) {
    int P = kMaxThreads / W;
    if (P > C) P = C;
    p->slots = P;
    p->threads = round_up(P * W, 32);
    p->grid_x = (C + P - 1) / P;
  } else {
    if (C % cv) return MRLA_ERR_ALIGN;
    const int lanes = C / cv;
    int cap = kMaxThreads / W;
    if (cap > 32) cap = 32;
    int lp = 1;
    for (int q = 1; q <= cap; ++q)
      if (lanes % q == 0) lp = q;
    p->slots = lp;
    p->threads = round_up(W * lp, 32);
    p->grid_x = lanes / lp;
  }
  int gy = (kNumSMs * 8 + p->grid_x - 1) / p->grid_x;
  if (gy > B) gy = B;
  if (gy < 1) gy = 1;
  p->grid_y = gy;
  return MRLA_OK;
}

template <typename K>
inline cudaError_t ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

#define MRLA_CHECK_LAUNCH()                       \
  do {                                            \
    cudaError_t e_ = cudaGetLastError();          \
    if (e_ != cudaSuccess) return (int)e_;        \
    ++g_launch_count;                             \
  } while (0)

inline MidShape mid_shape(const MrlaLightArgs& a, bool full) {
  MidShape m;
  m.B = a.B; m.C = a.C; m.HW = a.H * a.W; m.d = a.dim_perhead; m.k = a.k_size;
  m.bn_mode = a.bn_mode; m.has_o = (a.o != nullptr); m.full_mom = full;
  m.update_running = a.update_running; m.eps = a.eps; m.momentum = a.momentum;
  return m;
}

// ------------------------------------------------------------------------------------ forward
template <typename T, int LAYOUT, int CV, int ACT, bool HAS_O>
int light_forward_impl(const MrlaLightArgs& a, cudaStream_t st) {
  LightPlan p;
  int rc = make_plan(LAYOUT, a.B, a.C, a.W, CV, &p);
  if (rc) return rc;
  SweepShape s{a.B, a.C, a.H, a.W, p.slots};
  const bool full = (a.bn_mode == MRLA_BN_TRAIN);
  const dim3 grid(p.grid_x, p.grid_y);
  const T* x = static_cast<const T*>(a.x);
  const T* o = static_cast<const T*>(a.o);
  T* y = static_cast<T*>(a.y);
  // sweep 1
  {
    const int nm = full ? (HAS_O ? 6 : 3) : 1;
    const size_t sm = (size_t)p.threads * nm * CV * sizeof(float);
    if (full) {
      auto k = k_light_mom_fwd<T, LAYOUT, CV, ACT, HAS_O, true>;
      cudaError_t e = ensure_smem(k, sm);
      if (e != cudaSuccess) return (int)e;
      k<<<grid, p.threads, sm, st>>>(x, o, a.wv, a.mom, s, a.bs_x, a.bs_o);
    } else {
      auto k = k_light_mom_fwd<T, LAYOUT, CV, 0, false, false>;
      k<<<grid, p.threads, sm, st>>>(x, nullptr, a.wv, a.mom, s, a.bs_x, 0);
    }
    MRLA_CHECK_LAUNCH();
  }
  // mid
  MidShape ms = mid_shape(a, full);
  {
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    k_light_gate<<<a.B, th, 2 * a.C * sizeof(float), st>>>(a.mom, a.wq, a.wk, a.gate, ms);
    MRLA_CHECK_LAUNCH();
    k_light_bn_coef<<<(a.C + 31) / 32, 1024, 0, st>>>(a.mom, a.gate, a.lam, a.gamma, a.beta, a.running_mean,
                                                       a.running_var, a.drop_scale, a.mean, a.rstd, a.coef, ms);
    MRLA_CHECK_LAUNCH();
  }
  // sweep 2
  {
    auto k = k_light_apply_fwd<T, LAYOUT, CV, ACT, HAS_O>;
    k<<<grid, p.threads, 0, st>>>(x, o, y, a.wv, a.coef, s, a.bs_x, a.bs_o, a.bs_y, a.residual ? 1.f : 0.f);
    MRLA_CHECK_LAUNCH();
  }
  return MRLA_OK;
}

// ------------------------------------------------------------------------------------ backward
template <int LAYOUT>
inline int bwd_cv(int cv_fwd) { return LAYOUT == MRLA_NCHW ? 1 : (cv_fwd > 2 ? 2 : cv_fwd); }

template <typename T, int LAYOUT, int CV, int CVB, int ACT, bool HAS_O>
int light_backward_impl(const MrlaLightArgs& a, cudaStream_t st) {
  LightPlan p, pb;
  int rc = make_plan(LAYOUT, a.B, a.C, a.W, CV, &p);
  if (rc) return rc;
  rc = make_plan(LAYOUT, a.B, a.C, a.W, CVB, &pb);
  if (rc) return rc;
  const size_t need = ((size_t)pb.grid_y * a.C * 9 + (size_t)a.B * 2 * a.k_size) * sizeof(float);
  if (a.scratch == nullptr || a.scratch_bytes < need) return MRLA_ERR_WORKSPACE;
  float* wv_part = a.scratch;
  float* wqk_part = a.scratch + (size_t)pb.grid_y * a.C * 9;
  const bool full = (a.bn_mode == MRLA_BN_TRAIN);
  const T* x = static_cast<const T*>(a.x);
  const T* o = static_cast<const T*>(a.o);
  const T* dy = static_cast<const T*>(a.dy);
  T* dx = static_cast<T*>(a.dx);
  T* dout = static_cast<T*>(a.dout);
  // sweep A
  {
    SweepShape s{a.B, a.C, a.H, a.W, p.slots};
    const int nm = HAS_O ? 3 : 2;
    const size_t sm = (size_t)p.threads * nm * CV * sizeof(float);
    auto k = k_light_mom_bwd<T, LAYOUT, CV, ACT, HAS_O>;
    cudaError_t e = ensure_smem(k, sm);
    if (e != cudaSuccess) return (int)e;
    k<<<dim3(p.grid_x, p.grid_y), p.threads, sm, st>>>(dy, x, o, a.wv, a.gmom, s, a.bs_dy, a.bs_x, a.bs_o);
    MRLA_CHECK_LAUNCH();
  }
  // mid
  MidShape ms = mid_shape(a, full);
  {
    k_light_bwd_chan<<<(a.C + 31) / 32, 1024, 0, st>>>(a.mom, a.gmom, a.gate, a.lam, a.gamma, a.drop_scale, a.mean,
                                                        a.rstd, a.bcoef, a.dlam, a.dgamma, a.dbeta, ms);
    MRLA_CHECK_LAUNCH();
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    const size_t sm = ((size_t)5 * a.C + a.C / a.dim_perhead) * sizeof(float);
    cudaError_t e = ensure_smem(k_light_bwd_gate, sm);
    if (e != cudaSuccess) return (int)e;
    k_light_bwd_gate<<<a.B, th, sm, st>>>(a.mom, a.wq, a.wk, a.gate, a.bcoef, wqk_part, ms);
    MRLA_CHECK_LAUNCH();
  }
  // sweep B
  {
    SweepShape s{a.B, a.C, a.H, a.W, pb.slots};
    const size_t sm = (size_t)pb.threads * 9 * CVB * sizeof(float);
    auto k = k_light_apply_bwd<T, LAYOUT, CVB, ACT, HAS_O>;
    cudaError_t e = ensure_smem(k, sm);
    if (e != cudaSuccess) return (int)e;
    k<<<dim3(pb.grid_x, pb.grid_y), pb.threads, sm, st>>>(dy, x, o, dx, dout, a.wv, a.lam, a.bcoef, wv_part, s,
                                                          a.bs_dy, a.bs_x, a.bs_o, a.bs_dx, a.bs_do,
                                                          a.residual ? 1.f : 0.f);
    MRLA_CHECK_LAUNCH();
  }
  // final reductions
  {
    const int total = a.C * 9 + 2 * a.k_size;
    k_light_finish<<<(total + 255) / 256, 256, 0, st>>>(wv_part, pb.grid_y, wqk_part, a.dwv, a.dwq, a.dwk, a.B, a.C,
                                                         a.k_size);
    MRLA_CHECK_LAUNCH();
  }
  return MRLA_OK;
}

// ------------------------------------------------------------------------------------ dispatch
template <typename T, bool BWD>
int light_dispatch(const MrlaLightArgs& a, cudaStream_t st) {
  const bool has_o = (a.o != nullptr);
#define MRLA_GO(LAYOUT, CV, CVB, ACT, HAS_O)                                           \
  return BWD ? light_backward_impl<T, LAYOUT, CV, CVB, ACT, HAS_O>(a, st)             \
             : light_forward_impl<T, LAYOUT, CV, ACT, HAS_O>(a, st)
  if (a.layout == MRLA_NCHW) {
    if (a.act != MRLA_ACT_NONE) return MRLA_ERR_UNSUPPORTED;  // GELU variant is token-layout only (DeiT)
    if (has_o) { MRLA_GO(0, 1, 1, 0, true); } else { MRLA_GO(0, 1, 1, 0, false); }
  } else {
    if (a.C % 4) return MRLA_ERR_ALIGN;
    if (a.act == MRLA_ACT_GELU) {
      if (has_o) { MRLA_GO(1, 4, 2, 1, true); } else { MRLA_GO(1, 4, 2, 1, false); }
    } else {
      if (has_o) { MRLA_GO(1, 4, 2, 0, true); } else { MRLA_GO(1, 4, 2, 0, false); }
    }
  }
#undef MRLA_GO
}

template <typename T> int light_forward_t(const MrlaLightArgs& a, cudaStream_t st) { return light_dispatch<T, false>(a, st); }
template <typename T> int light_backward_t(const MrlaLightArgs& a, cudaStream_t st) { return light_dispatch<T, true>(a, st); }

// scratch floats needed by backward (dtype independent)
inline size_t light_bwd_scratch_floats(const MrlaLightArgs& a) {
  LightPlan pb;
  const int cvb = a.layout == MRLA_NCHW ? 1 : 2;
  if (make_plan(a.layout, a.B, a.C, a.W, cvb, &pb)) return 0;
  return (size_t)pb.grid_y * a.C * 9 + (size_t)a.B * 2 * a.k_size;
}

}  // namespace mrla
