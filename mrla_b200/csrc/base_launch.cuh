// Host-side launch logic for the MRLA-base tail (one translation unit per activation dtype).
#pragma once
#include "../../include/mrla_b200.h"
#include "base_kernels.cuh"
#include "base_stream.cuh"
#include "light_launch.cuh"

namespace mrla {

constexpr int kBaseMaxParts = 2 * kNumSMs;   // upper bound of dWv partial slots of the TMA dX kernel (grid / ncb)

inline size_t base_bwd_scratch_floats(const MrlaBaseArgs& a) {
  LightPlan p;
  const int cv = a.layout == MRLA_NCHW ? 1 : 4;
  if (make_plan(a.layout, a.B, a.C, a.W, cv, &p)) return 0;
  const int nparts = p.grid_y > kBaseMaxParts ? p.grid_y : kBaseMaxParts;
  // wv partials | wqk partials | bchan[3,C]
  return (size_t)nparts * a.C * 9 + (size_t)a.B * 2 * a.k_size + (size_t)3 * a.C;
}

// ---- NHWC fast path helpers ---------------------------------------------------------------------------
struct StreamPlan { int pl, items, grid; };

inline bool make_stream_plan(const MrlaBaseArgs& a, StreamPlan* p) {
  if (a.layout != MRLA_NHWC || a.C % 64 || a.dim_perhead % 8) return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  const void* ptrs[] = {a.x, a.v, a.s, a.y, a.dy, a.dx, a.dv};
  for (const void* q : ptrs)
    if (q && ((uintptr_t)q % 16)) return false;
  const int64_t strides[] = {a.bs_x, a.bs_y, a.bs_s, a.bs_dy, a.bs_dx, a.bs_v, a.ts_v, a.bs_dv, a.ts_dv};
  for (int64_t s_ : strides)
    if ((s_ * es) % 16) return false;
  const int hw = a.H * a.W;
  p->pl = (hw >= 256 || a.C % 256) ? 32 : 8;     // few pixels: more channel lanes per CTA
  const int cbs = (256 / p->pl) * kSV;
  p->items = a.B * (a.C / cbs);
  p->grid = p->items < kNumSMs * 8 ? p->items : kNumSMs * 8;
  return true;
}

inline MrlaLightArgs light_view_of(const MrlaBaseArgs& a, void* y, int64_t bs_y) {
  MrlaLightArgs l{};
  l.B = a.B; l.C = a.C; l.H = a.H; l.W = a.W; l.dim_perhead = a.dim_perhead; l.k_size = a.k_size;
  l.dtype = a.dtype; l.layout = a.layout; l.act = MRLA_ACT_NONE; l.residual = 0;
  l.x = a.x; l.bs_x = a.bs_x; l.y = y; l.bs_y = bs_y; l.wv = a.wv;
  return l;
}

#define MRLA_STREAM_LAUNCH(KERNEL, SMEM, ...)                                                   \
  do {                                                                                          \
    if (sp.pl == 32) KERNEL<T, 32><<<sp.grid, 256, SMEM, st>>>(__VA_ARGS__);                    \
    else KERNEL<T, 8><<<sp.grid, 256, SMEM, st>>>(__VA_ARGS__);                                 \
    MRLA_CHECK_LAUNCH();                                                                        \
  } while (0)

template <typename T, int LAYOUT, int CV>
int base_forward_impl(const MrlaBaseArgs& a, cudaStream_t st) {
  LightPlan p;
  int rc = make_plan(LAYOUT, a.B, a.C, a.W, CV, &p);
  if (rc) return rc;
  const dim3 grid(p.grid_x, p.grid_y);
  BaseShape s{a.B, a.C, a.H, a.W, p.slots, a.t, a.dim_perhead};
  SweepShape ss{a.B, a.C, a.H, a.W, p.slots};
  const T* x = static_cast<const T*>(a.x);
  T* vslot = static_cast<T*>(a.v) + (int64_t)(a.t - 1) * a.ts_v;
  StreamPlan sp;
  const bool fast = LAYOUT == MRLA_NHWC && make_stream_plan(a, &sp);
  StreamShape ssh{a.B, a.C, a.H * a.W, a.t, a.dim_perhead, fast ? sp.items : 0};
  // F0: GAP sums + v_t into the cache slot
  bool f0_done = false;
  if (fast) {
    MrlaLightArgs la = light_view_of(a, vslot, a.bs_v);
    TmaPlan tp;
    const int es = a.dtype == MRLA_F32 ? 4 : 2;
    if (tma_ptr_ok(a.x, a.bs_x, es) && make_tma_plan(la, 0, 1, &tp)) {
      rc = launch_tma_sweep<T, 0, 3>(la, st, tp, a.x, a.bs_x, nullptr, 0, nullptr, 0, a.sx);
      if (rc) return rc;
      f0_done = true;
    }
  }
  if (!f0_done) {
    const size_t sm = (size_t)p.threads * CV * sizeof(float);
    k_light_mom_fwd<T, LAYOUT, CV, 0, false, false><<<grid, p.threads, sm, st>>>(x, nullptr, a.wv, a.sx, ss, a.bs_x, 0);
    MRLA_CHECK_LAUNCH();
    k_base_conv<T, LAYOUT, CV><<<grid, p.threads, 0, st>>>(x, vslot, a.wv, s, a.bs_x, a.bs_v);
    MRLA_CHECK_LAUNCH();
  }
  // F1: q, k_t, softmax weights
  {
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    k_base_attn<<<a.B, th, 2 * a.C * sizeof(float), st>>>(a.sx, a.wq, a.wk, a.q, a.kcache, a.p, a.B, a.C, a.H * a.W,
                                                          a.dim_perhead, a.k_size, a.t, a.t_cap);
    MRLA_CHECK_LAUNCH();
  }
  // F2: S and its moments
  if (fast) {
    MRLA_STREAM_LAUNCH(k_base_mix_nhwc, 256 * 2 * kSV * sizeof(float), static_cast<const T*>(a.v),
                       static_cast<T*>(a.s), a.p, a.smom, ssh, a.bs_v, a.ts_v, a.bs_s);
  } else {
    const size_t sm = (size_t)p.threads * 2 * CV * sizeof(float);
    k_base_mix<T, LAYOUT, CV><<<grid, p.threads, sm, st>>>(static_cast<const T*>(a.v), static_cast<T*>(a.s), a.p,
                                                           a.smom, s, a.bs_v, a.ts_v, a.bs_s);
    MRLA_CHECK_LAUNCH();
  }
  // F3: BN statistics
  k_base_bn<<<(a.C + 31) / 32, 1024, 0, st>>>(a.smom, a.gamma, a.beta, a.running_mean, a.running_var, a.chan, a.B, a.C,
                                              a.H * a.W, a.bn_mode, a.update_running, a.eps, a.momentum);
  MRLA_CHECK_LAUNCH();
  // F4: output
  if (fast) {
    MRLA_STREAM_LAUNCH(k_base_apply_nhwc, 0, x, static_cast<const T*>(a.s), static_cast<T*>(a.y), a.chan, a.drop_scale,
                       ssh, a.bs_x, a.bs_s, a.bs_y, a.residual ? 1.f : 0.f, a.relu);
  } else {
    k_base_apply<T, LAYOUT, CV><<<grid, p.threads, 0, st>>>(x, static_cast<const T*>(a.s), static_cast<T*>(a.y),
                                                            a.chan, a.drop_scale, s, a.bs_x, a.bs_s, a.bs_y,
                                                            a.residual ? 1.f : 0.f, a.relu);
    MRLA_CHECK_LAUNCH();
  }
  return MRLA_OK;
}

template <typename T, int LAYOUT, int CV>
int base_backward_impl(const MrlaBaseArgs& a, cudaStream_t st) {
  LightPlan p;
  int rc = make_plan(LAYOUT, a.B, a.C, a.W, CV, &p);
  if (rc) return rc;
  const size_t need = base_bwd_scratch_floats(a) * sizeof(float);
  if (a.scratch == nullptr || a.scratch_bytes < need) return MRLA_ERR_WORKSPACE;
  const int nparts_max = p.grid_y > kBaseMaxParts ? p.grid_y : kBaseMaxParts;
  float* wv_part = a.scratch;
  float* wqk_part = wv_part + (size_t)nparts_max * a.C * 9;
  int nparts = p.grid_y;
  float* bchan = wqk_part + (size_t)a.B * 2 * a.k_size;
  const dim3 grid(p.grid_x, p.grid_y);
  BaseShape s{a.B, a.C, a.H, a.W, p.slots, a.t, a.dim_perhead};
  const T* dy = static_cast<const T*>(a.dy);
  const T* sin = static_cast<const T*>(a.s);
  StreamPlan sp;
  const bool fast = LAYOUT == MRLA_NHWC && make_stream_plan(a, &sp);
  StreamShape ssh{a.B, a.C, a.H * a.W, a.t, a.dim_perhead, fast ? sp.items : 0};
  // B0
  if (fast) {
    MRLA_STREAM_LAUNCH(k_base_mom_bwd_nhwc, 256 * 2 * kSV * sizeof(float), dy, sin, a.chan, a.drop_scale, a.gmom, ssh,
                       a.bs_dy, a.bs_s, a.relu);
  } else {
    const size_t sm = (size_t)p.threads * 2 * CV * sizeof(float);
    k_base_mom_bwd<T, LAYOUT, CV><<<grid, p.threads, sm, st>>>(dy, sin, a.chan, a.drop_scale, a.gmom, s, a.bs_dy, a.bs_s,
                                                               a.relu);
    MRLA_CHECK_LAUNCH();
  }
  // B1
  k_base_bwd_chan<<<(a.C + 31) / 32, 1024, 0, st>>>(a.gmom, a.gamma, a.chan, bchan, a.dgamma, a.dbeta, a.B, a.C,
                                                    a.H * a.W, a.bn_mode);
  MRLA_CHECK_LAUNCH();
  // B2 (chunks of kBaseChunk cache slots)
  if (fast) {
    for (int j0 = 0; j0 < a.t; j0 += kBaseChunk)
      MRLA_STREAM_LAUNCH(k_base_scatter_nhwc, 256 * kBaseChunk * sizeof(float), dy, sin, static_cast<const T*>(a.v),
                         static_cast<T*>(a.dv), a.p, a.chan, bchan, a.drop_scale, a.dpm, ssh, j0, a.accumulate, a.bs_dy,
                         a.bs_s, a.bs_v, a.ts_v, a.bs_dv, a.ts_dv, a.relu);
  } else {
    const size_t sm = (size_t)p.threads * kBaseChunk * CV * sizeof(float);
    auto k = k_base_scatter<T, LAYOUT, CV>;
    cudaError_t e = ensure_smem(k, sm);
    if (e != cudaSuccess) return (int)e;
    for (int j0 = 0; j0 < a.t; j0 += kBaseChunk) {
      k<<<grid, p.threads, sm, st>>>(dy, sin, static_cast<const T*>(a.v), static_cast<T*>(a.dv), a.p, a.chan, bchan,
                                     a.drop_scale, a.dpm, s, j0, a.accumulate, a.bs_dy, a.bs_s, a.bs_v, a.ts_v, a.bs_dv,
                                     a.ts_dv, a.relu);
      MRLA_CHECK_LAUNCH();
    }
  }
  // B3
  {
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    const size_t sm = ((size_t)3 * a.C + (size_t)(a.C / a.dim_perhead) * a.t) * sizeof(float);
    cudaError_t e = ensure_smem(k_base_bwd_attn, sm);
    if (e != cudaSuccess) return (int)e;
    k_base_bwd_attn<<<a.B, th, sm, st>>>(a.sx, a.q, a.kcache, a.dkcache, a.p, a.dpm, a.wq, a.wk, a.dyc, wqk_part, a.B,
                                         a.C, a.H * a.W, a.dim_perhead, a.k_size, a.t, a.t_cap, a.accumulate);
    MRLA_CHECK_LAUNCH();
  }
  // B4
  bool b4_done = false;
  if (fast) {
    const T* dvt = static_cast<const T*>(a.dv) + (int64_t)(a.t - 1) * a.ts_dv;
    MrlaLightArgs la = light_view_of(a, a.dx, a.bs_dx);
    la.coef = a.dyc;
    la.residual = a.residual;
    TmaPlan tp;
    const int es = a.dtype == MRLA_F32 ? 4 : 2;
    if (tma_ptr_ok(dvt, a.bs_dv, es) && tma_ptr_ok(a.x, a.bs_x, es) && tma_ptr_ok(a.dy, a.bs_dy, es) &&
        make_tma_plan(la, 2, 9, &tp) && tp.grid % tp.ncb == 0 && tp.grid / tp.ncb <= kBaseMaxParts) {
      // tiles: window tile = dV_t (halo columns), centre tiles = x and dy
      rc = launch_tma_sweep<T, 0, 4>(la, st, tp, dvt, a.bs_dv, a.x, a.bs_x, a.dy, a.bs_dy, wv_part);
      if (rc) return rc;
      nparts = tp.grid / tp.ncb;
      b4_done = true;
    }
  }
  if (!b4_done) {
    const size_t sm = (size_t)p.threads * 9 * CV * sizeof(float);
    auto k = k_base_dx<T, LAYOUT, CV>;
    cudaError_t e = ensure_smem(k, sm);
    if (e != cudaSuccess) return (int)e;
    const T* dvt = static_cast<const T*>(a.dv) + (int64_t)(a.t - 1) * a.ts_dv;
    k<<<grid, p.threads, sm, st>>>(dy, static_cast<const T*>(a.x), dvt, static_cast<T*>(a.dx), a.wv, a.dyc, wv_part, s,
                                   a.bs_dy, a.bs_x, a.bs_dv, a.bs_dx, a.residual ? 1.f : 0.f);
    MRLA_CHECK_LAUNCH();
  }
  // B5
  {
    const int total = a.C * 9 + 2 * a.k_size;
    k_light_finish<<<(a.C * 9 + 255) / 256 + (2 * a.k_size + 7) / 8, 256, 0, st>>>(wv_part, nparts, wqk_part, a.dwv, a.dwq, a.dwk, a.B, a.C,
                                                         a.k_size);
    MRLA_CHECK_LAUNCH();
  }
  return MRLA_OK;
}

template <typename T, bool BWD>
int base_dispatch(const MrlaBaseArgs& a, cudaStream_t st) {
  if (a.layout == MRLA_NCHW)
    return BWD ? base_backward_impl<T, 0, 1>(a, st) : base_forward_impl<T, 0, 1>(a, st);
  if (a.C % 4) return MRLA_ERR_ALIGN;
  return BWD ? base_backward_impl<T, 1, 4>(a, st) : base_forward_impl<T, 1, 4>(a, st);
}
template <typename T> int base_forward_t(const MrlaBaseArgs& a, cudaStream_t st) { return base_dispatch<T, false>(a, st); }
template <typename T> int base_backward_t(const MrlaBaseArgs& a, cudaStream_t st) { return base_dispatch<T, true>(a, st); }

}  // namespace mrla
