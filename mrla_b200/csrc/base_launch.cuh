// Host-side launch logic for the MRLA-base tail (one translation unit per activation dtype).
#pragma once
#include "../../include/mrla_b200.h"
#include "base_kernels.cuh"
#include "light_launch.cuh"

namespace mrla {

inline size_t base_bwd_scratch_floats(const MrlaBaseArgs& a) {
  LightPlan p;
  const int cv = a.layout == MRLA_NCHW ? 1 : 4;
  if (make_plan(a.layout, a.B, a.C, a.W, cv, &p)) return 0;
  // wv partials | wqk partials | bchan[3,C]
  return (size_t)p.grid_y * a.C * 9 + (size_t)a.B * 2 * a.k_size + (size_t)3 * a.C;
}

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
  // F0: GAP sums + v_t into the cache slot
  {
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
  {
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
  k_base_apply<T, LAYOUT, CV><<<grid, p.threads, 0, st>>>(x, static_cast<const T*>(a.s), static_cast<T*>(a.y), a.chan,
                                                          a.drop_scale, s, a.bs_x, a.bs_s, a.bs_y,
                                                          a.residual ? 1.f : 0.f, a.relu);
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

template <typename T, int LAYOUT, int CV>
int base_backward_impl(const MrlaBaseArgs& a, cudaStream_t st) {
  LightPlan p;
  int rc = make_plan(LAYOUT, a.B, a.C, a.W, CV, &p);
  if (rc) return rc;
  const size_t need = base_bwd_scratch_floats(a) * sizeof(float);
  if (a.scratch == nullptr || a.scratch_bytes < need) return MRLA_ERR_WORKSPACE;
  float* wv_part = a.scratch;
  float* wqk_part = wv_part + (size_t)p.grid_y * a.C * 9;
  float* bchan = wqk_part + (size_t)a.B * 2 * a.k_size;
  const dim3 grid(p.grid_x, p.grid_y);
  BaseShape s{a.B, a.C, a.H, a.W, p.slots, a.t, a.dim_perhead};
  const T* dy = static_cast<const T*>(a.dy);
  const T* sin = static_cast<const T*>(a.s);
  // B0
  {
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
  {
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
  {
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
    k_light_finish<<<(a.C * 9 + 255) / 256 + (2 * a.k_size + 7) / 8, 256, 0, st>>>(wv_part, p.grid_y, wqk_part, a.dwv, a.dwq, a.dwk, a.B, a.C,
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
