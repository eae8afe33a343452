// Host-side launch logic for the MRLA-light tail (one translation unit per activation dtype).
#pragma once
#include "../../include/mrla_b200.h"
#include "light_mid.cuh"
#include "light_mid_cluster.cuh"
#include "light_sweeps.cuh"
#include "light_nhwc_tma.cuh"
#include "light_nhwc_ring.cuh"
#include "layout_kernels.cuh"
#include "light_v7_launch.cuh"

namespace mrla {

extern thread_local int g_launch_count;

// the v7 kernels are compiled in their own translation units (v7_<dtype>.cu)
#define MRLA_V7_EXTERN(T)                                                                                                  \
  extern template int v7_launch_fwd<T, 0>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t,  \
                                          float*, int, int);                                                               \
  extern template int v7_launch_fwd<T, 1>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t,  \
                                          float*, int, int);                                                               \
  extern template int v7_launch_fwd<T, 2>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, const void*, int64_t,  \
                                          float*, int, int);                                                               \
  extern template int v7_launch_bwd<T>(const MrlaLightArgs&, cudaStream_t, const V7Plan&, bool, bool, const void*, int64_t, \
                                       float*, float*, float*);
MRLA_V7_EXTERN(float)
MRLA_V7_EXTERN(__nv_bfloat16)
MRLA_V7_EXTERN(__half)
#undef MRLA_V7_EXTERN

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
  m.da_summed = 0;
  return m;
}

// ------------------------------------------------------------------------------------ v2 (TMA) planning
struct TmaPlan {
  int CB, NQ, G, S, ncb, items, cons_threads, grid, big;
  uint32_t x_bytes, o_bytes, stage_bytes;
  size_t smem;
};

// ntiles = number of [CB,W,G] tiles per stage besides the x tile (o and/or dy); nacc = float2 accumulators/thread
inline bool make_tma_plan(const MrlaLightArgs& a, int ntiles, int nacc, TmaPlan* p, bool ohalo = false) {
  if (a.layout != MRLA_NHWC || a.C % 8 || a.W > 56) return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  const int NQ = (a.W + kCols - 1) / kCols;
  int CB = 0;
  for (int cb : {256, 128, 64})
    if (NQ * cb / 2 <= 448 && (a.C % cb == 0 || (cb == 64 && a.C > 64))) { CB = cb; break; }
  if (CB == 0) {
    if (NQ * 32 <= 448) CB = 64; else return false;
  }
  if (a.act == MRLA_ACT_GELU && CB == 256) CB = 128;
  const uint32_t xrow = (uint32_t)(NQ * kCols + 2) * CB * es;
  const uint32_t orow = ohalo ? xrow : (uint32_t)(NQ * kCols) * CB * es;   // MODE 5: the o tile carries halo columns
  p->cons_threads = NQ * (CB / 2);
  p->big = p->cons_threads > 256;
  const size_t budget = (p->big ? 200 : 100) * 1024;   // !big: two CTAs share one SM
  // rows per TMA group: about a fifth of the ring per stage (>= 4 stages in flight)
  const uint32_t rowtot = xrow + (uint32_t)ntiles * orow;
  int G = (int)((budget / 5) / rowtot);
  if (G < 1) G = 1;
  if (G > a.H) G = a.H;
  p->CB = CB; p->NQ = NQ; p->G = G;
  p->x_bytes = (uint32_t)G * xrow;
  p->o_bytes = (uint32_t)G * orow;
  p->stage_bytes = p->x_bytes + (uint32_t)ntiles * p->o_bytes;
  const size_t red = (size_t)2 * NQ * nacc * (CB / 2) * sizeof(float2);   // double buffered
  int S = (int)((budget - red) / p->stage_bytes);
  if (S > 8) S = 8;
  if (S < 2) return false;
  p->S = S;
  p->ncb = (a.C + CB - 1) / CB;
  p->items = a.B * p->ncb;
  const int slots = p->big ? kNumSMs : 2 * kNumSMs;
  p->grid = p->items < slots ? p->items : slots;
  if (p->grid >= p->ncb) p->grid -= p->grid % p->ncb;   // a CTA then always sees the same channel block
  p->smem = 256 + (size_t)S * p->stage_bytes + red;
  return true;
}

inline bool tma_ptr_ok(const void* ptr, int64_t bs, int es) {
  return ptr != nullptr && ((uintptr_t)ptr % 16 == 0) && ((bs * es) % 16 == 0);
}

template <typename T, int ACT, int MODE>
int launch_tma_sweep(const MrlaLightArgs& a, cudaStream_t st, const TmaPlan& p, const void* xptr, int64_t bs_x,
                     const void* optr, int64_t bs_o, const void* dyptr, int64_t bs_dy, float* mom) {
  CUtensorMap tx, to, tdy;
  if (make_nhwc_tmap(&tx, xptr, a.dtype, a.B, a.C, a.H, a.W, bs_x, p.CB, p.NQ * kCols + 2, p.G)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (MODE == 3) to = tx;
  else if (make_nhwc_tmap(&to, optr, a.dtype, a.B, a.C, a.H, a.W, bs_o, p.CB, p.NQ * kCols + ((MODE == 5 || MODE == 6) ? 2 : 0), p.G))
    return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  tdy = to;
  if ((MODE == 2 || MODE == 4) && make_nhwc_tmap(&tdy, dyptr, a.dtype, a.B, a.C, a.H, a.W, bs_dy, p.CB, p.NQ * kCols, p.G))
    return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  TmaSweepParams P;
  P.B = a.B; P.C = a.C; P.H = a.H; P.W = a.W;
  P.G = p.G; P.S = p.S; P.NQ = p.NQ; P.ncb = p.ncb; P.items = p.items; P.cons_threads = p.cons_threads;
  P.rev = (MODE == 1) ? 1 : 0;   // sweep 2 meets the end of the batch sweep 1 just left in L2
  P.x_bytes = p.x_bytes; P.o_bytes = p.o_bytes; P.stage_bytes = p.stage_bytes;
  P.wv = a.wv; P.mom = mom; P.coef = a.coef; P.y = a.y; P.bs_y = a.bs_y; P.res = a.residual ? 1.f : 0.f;
  P.wv_part = (MODE == 4) ? mom : nullptr;   // MODE 4 passes the partial buffer through `mom`
  P.zcoef = (MODE == 6) ? a.z_coef : nullptr;
  const int threads = 32 + p.cons_threads;
#define MRLA_TMA_LAUNCH1(CBV, BIGV)                                                                       \
  {                                                                                                       \
    auto k = k_light_nhwc_tma<T, CBV, ACT, (MODE != 3), MODE, BIGV>;                                      \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);    \
    if (e != cudaSuccess) return (int)e;                                                                  \
    k<<<p.grid, threads, p.smem, st>>>(tx, to, tdy, P);                                                   \
  }
#define MRLA_TMA_LAUNCH(CBV)                                                                              \
  {                                                                                                       \
    if (p.big) MRLA_TMA_LAUNCH1(CBV, true) else MRLA_TMA_LAUNCH1(CBV, false)                              \
  }
  if (p.CB == 64) MRLA_TMA_LAUNCH(64)
  else if (p.CB == 128) MRLA_TMA_LAUNCH(128)
  else {
    if (ACT == 1) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
    MRLA_TMA_LAUNCH(ACT == 1 ? 128 : 256)
  }
#undef MRLA_TMA_LAUNCH
#undef MRLA_TMA_LAUNCH1
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

struct TmaBwdPlan {
  int CB, NQ, NT, WT, G, S, ncb, items, ipc, grid, cons_threads, maxslots, big;
  uint32_t x_bytes, t_bytes, stage_bytes;
  size_t smem;
};

inline bool make_tma_bwd_plan(const MrlaLightArgs& a, TmaBwdPlan* p) {
  if (a.layout != MRLA_NHWC || a.C % 8 || a.W > 224) return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  const int NT = (a.W + 27) / 28;
  int WT = (a.W + NT - 1) / NT;
  if (NT > 1) WT = (WT + 3) / 4 * 4;
  const int NQ = (WT + kCols - 1) / kCols;
  int CB = 0;
  for (int cb : {256, 128, 64})
    if (NQ * cb / 2 <= 224 && a.C % cb == 0) { CB = cb; break; }
  if (CB == 0) CB = 64;
  if (NQ * CB / 2 > 224) return false;
  if (a.act == MRLA_ACT_GELU && CB == 256) CB = 128;
  const uint32_t xrow = (uint32_t)(NQ * kCols + 4) * CB * es, trow = (uint32_t)(NQ * kCols + 2) * CB * es;
  const uint32_t rowtot = xrow + 2 * trow;
  p->big = NQ * (CB / 2) > 128;
  int G = (int)((((size_t)(p->big ? 200 : 100) * 1024) / 5) / rowtot);
  if (G < 1) G = 1;
  if (G > a.H) G = a.H;
  p->CB = CB; p->NQ = NQ; p->NT = NT; p->WT = WT; p->G = G;
  p->x_bytes = (uint32_t)G * xrow;
  p->t_bytes = (uint32_t)G * trow;
  p->stage_bytes = p->x_bytes + 2 * p->t_bytes;
  const size_t red = (size_t)NQ * 9 * (CB / 2) * sizeof(float2);
  p->big = NQ * (CB / 2) > 128;
  int S = (int)(((p->big ? 200 : 100) * 1024 - red) / p->stage_bytes);
  if (S > 8) S = 8;
  if (S < 2) return false;
  p->S = S;
  p->ncb = (a.C + CB - 1) / CB;
  p->items = p->ncb * a.B * NT;
  const int slots = p->big ? kNumSMs : 2 * kNumSMs;
  int grid = p->items < slots ? p->items : slots;
  p->ipc = (p->items + grid - 1) / grid;
  p->grid = (p->items + p->ipc - 1) / p->ipc;
  const int ipcb = a.B * NT;
  p->maxslots = (ipcb + p->ipc - 1) / p->ipc + 1;
  p->cons_threads = NQ * (CB / 2);
  p->smem = 256 + (size_t)S * p->stage_bytes + red;
  return true;
}

template <typename T, int ACT>
int launch_tma_bwd(const MrlaLightArgs& a, cudaStream_t st, const TmaBwdPlan& p, float* wv_part) {
  CUtensorMap tx, tdy, to;
  if (make_nhwc_tmap(&tx, a.x, a.dtype, a.B, a.C, a.H, a.W, a.bs_x, p.CB, p.NQ * kCols + 4, p.G)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&tdy, a.dy, a.dtype, a.B, a.C, a.H, a.W, a.bs_dy, p.CB, p.NQ * kCols + 2, p.G)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&to, a.o, a.dtype, a.B, a.C, a.H, a.W, a.bs_o, p.CB, p.NQ * kCols + 2, p.G)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  cudaError_t e = cudaMemsetAsync(wv_part, 0, (size_t)p.maxslots * a.C * 9 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  TmaBwdParams P;
  P.B = a.B; P.C = a.C; P.H = a.H; P.W = a.W;
  P.G = p.G; P.S = p.S; P.NQ = p.NQ; P.ncb = p.ncb; P.NT = p.NT; P.WT = p.WT; P.items = p.items; P.ipc = p.ipc;
  P.cons_threads = p.cons_threads; P.maxslots = p.maxslots;
  P.x_bytes = p.x_bytes; P.t_bytes = p.t_bytes; P.stage_bytes = p.stage_bytes;
  P.wv = a.wv; P.lam = a.lam; P.bcoef = a.bcoef; P.dx = a.dx; P.dout = a.dout; P.bs_dx = a.bs_dx; P.bs_do = a.bs_do;
  P.res = a.residual ? 1.f : 0.f; P.wv_part = wv_part;
  const int threads = 32 + p.cons_threads;
#define MRLA_TMA_LAUNCH1(CBV, BIGV)                                                                       \
  {                                                                                                       \
    auto k = k_light_nhwc_tma_bwd<T, CBV, ACT, BIGV>;                                                     \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);                \
    if (e != cudaSuccess) return (int)e;                                                                  \
    k<<<p.grid, threads, p.smem, st>>>(tx, tdy, to, P);                                                   \
  }
#define MRLA_TMA_LAUNCH(CBV)                                                                              \
  {                                                                                                       \
    if (p.big) MRLA_TMA_LAUNCH1(CBV, true) else MRLA_TMA_LAUNCH1(CBV, false)                              \
  }
  if (p.CB == 64) MRLA_TMA_LAUNCH(64)
  else if (p.CB == 128) MRLA_TMA_LAUNCH(128)
  else {
    if (ACT == 1) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
    MRLA_TMA_LAUNCH(ACT == 1 ? 128 : 256)
  }
#undef MRLA_TMA_LAUNCH
#undef MRLA_TMA_LAUNCH1
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

// v4 sweep B (T ring, light_nhwc_ring.cuh): full-width rows, one image row per stage, one channel block per CTA.
// TmaBwdPlan::big doubles as the column count per thread: big (KC = 8, one CTA per SM) or small (KC = 4, two per SM).
inline bool ring_plan_try(const MrlaLightArgs& a, int KC, int CB, TmaBwdPlan* p) {
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  const int NQ = (a.W + KC - 1) / KC;
  p->cons_threads = NQ * (CB / 2);
  p->big = (KC == 8);
  const uint32_t xrow = (uint32_t)(NQ * KC + 2) * CB * es, trow = (uint32_t)(NQ * KC) * CB * es;
  // 4 T-ring slots of 128-bit column pairs + 3 staging buffers of one dX row and one dO row for the TMA stores
  const size_t ring = (size_t)4 * (KC / 2 * NQ + 2) * (CB / 2) * 16 + (size_t)3 * 2 * trow;
  const size_t budget = (size_t)(p->big ? 226 : 100) * 1024;   // sm_100a: 227 KB dynamic shared memory per CTA
  if (ring + 8192 > budget) return false;
  p->CB = CB; p->NQ = NQ; p->NT = 1; p->WT = a.W; p->G = 1;
  p->x_bytes = xrow;
  p->t_bytes = trow;
  p->stage_bytes = xrow + 2 * trow;
  int S = (int)((budget - ring) / p->stage_bytes);
  if (S > 8) S = 8;
  if (S < 3) return false;
  p->S = S;
  p->ncb = (a.C + CB - 1) / CB;
  p->items = p->ncb * a.B;
  const int slots = p->big ? kNumSMs : 2 * kNumSMs;
  int cpc = slots / p->ncb;
  if (cpc < 1) cpc = 1;
  if (cpc > a.B) cpc = a.B;
  p->ipc = (a.B + cpc - 1) / cpc;
  p->grid = p->ncb * cpc;
  p->maxslots = cpc;
  p->smem = 256 + (size_t)S * p->stage_bytes + ring;
  return true;
}

inline bool make_tma_ring_plan(const MrlaLightArgs& a, TmaBwdPlan* p) {
  if (a.layout != MRLA_NHWC || a.C % 8 || a.W > 64 || a.H < 3) return false;
  // big: 8 columns per thread, 129..256 threads, one CTA per SM; small: 4 columns per thread, up to 128 threads, two
  // CTAs per SM.  First candidate whose ring + staging + >= 4 pipeline stages fit in shared memory.
  for (int cb : {256, 128, 64}) {
    const int n = ((a.W + 7) / 8) * cb / 2;
    if (a.act == MRLA_ACT_GELU && cb == 256) continue;
    if (n <= 256 && n > 128 && (a.C % cb == 0 || (cb == 64 && a.C > 64)) && ring_plan_try(a, 8, cb, p)) return true;
  }
  for (int cb : {256, 128, 64}) {
    const int n = ((a.W + 3) / 4) * cb / 2;
    if (a.act == MRLA_ACT_GELU && cb == 256) continue;
    if (n <= 128 && (a.C % cb == 0 || cb == 64) && ring_plan_try(a, 4, cb, p)) return true;
  }
  return false;
}

template <typename T, int ACT>
int launch_tma_bwd_ring(const MrlaLightArgs& a, cudaStream_t st, const TmaBwdPlan& p, float* wv_part) {
  const int KC = p.big ? 8 : 4;
  CUtensorMap tx, tdy, to, tdx, tdo;
  if (make_nhwc_tmap(&tx, a.x, a.dtype, a.B, a.C, a.H, a.W, a.bs_x, p.CB, p.NQ * KC + 2, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&tdy, a.dy, a.dtype, a.B, a.C, a.H, a.W, a.bs_dy, p.CB, p.NQ * KC, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&to, a.o, a.dtype, a.B, a.C, a.H, a.W, a.bs_o, p.CB, p.NQ * KC, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&tdx, a.dx, a.dtype, a.B, a.C, a.H, a.W, a.bs_dx, p.CB, p.NQ * KC, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (make_nhwc_tmap(&tdo, a.dout, a.dtype, a.B, a.C, a.H, a.W, a.bs_do, p.CB, p.NQ * KC, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  cudaError_t e = cudaSuccess;
  TmaBwdParams P;
  P.B = a.B; P.C = a.C; P.H = a.H; P.W = a.W;
  P.G = 1; P.S = p.S; P.NQ = p.NQ; P.ncb = p.ncb; P.NT = 1; P.WT = a.W; P.items = p.items; P.ipc = p.ipc;
  P.cons_threads = p.cons_threads; P.maxslots = p.maxslots; P.cpc = p.maxslots;
  P.x_bytes = p.x_bytes; P.t_bytes = p.t_bytes; P.stage_bytes = p.stage_bytes;
  P.wv = a.wv; P.lam = a.lam; P.bcoef = a.bcoef; P.dx = a.dx; P.dout = a.dout; P.bs_dx = a.bs_dx; P.bs_do = a.bs_do;
  P.res = a.residual ? 1.f : 0.f; P.wv_part = wv_part;   // every (slot, channel) of wv_part is written by its CTA
  const int threads = p.cons_threads;   // no producer warp: thread 0 issues the TMA traffic
  const bool ragged = (a.W % KC) != 0;
#define MRLA_TMA_LAUNCH3(CBV, KCV, FUSEV, RAGV)                                                           \
  {                                                                                                       \
    auto k = k_light_nhwc_ring<T, CBV, ACT, FUSEV, RAGV, KCV>;                                            \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);                \
    if (e != cudaSuccess) return (int)e;                                                                  \
    k<<<p.grid, threads, p.smem, st>>>(tx, tdy, to, tdx, tdo, P);                                         \
  }
#define MRLA_TMA_LAUNCH2(CBV, KCV, FUSEV)                                                                 \
  {                                                                                                       \
    if (ragged) MRLA_TMA_LAUNCH3(CBV, KCV, FUSEV, true) else MRLA_TMA_LAUNCH3(CBV, KCV, FUSEV, false)     \
  }
#define MRLA_TMA_LAUNCH1(CBV, KCV)                                                                        \
  {                                                                                                       \
    if (a.fuse_relu_bwd) MRLA_TMA_LAUNCH2(CBV, KCV, true) else MRLA_TMA_LAUNCH2(CBV, KCV, false)          \
  }
#define MRLA_TMA_LAUNCH(CBV)                                                                              \
  {                                                                                                       \
    if (p.big) MRLA_TMA_LAUNCH1(CBV, 8) else MRLA_TMA_LAUNCH1(CBV, 4)                                     \
  }
  if (p.CB == 64) MRLA_TMA_LAUNCH(64)
  else if (p.CB == 128) MRLA_TMA_LAUNCH(128)
  else {
    if (ACT == 1) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
    MRLA_TMA_LAUNCH(ACT == 1 ? 128 : 256)
  }
#undef MRLA_TMA_LAUNCH
#undef MRLA_TMA_LAUNCH1
#undef MRLA_TMA_LAUNCH2
#undef MRLA_TMA_LAUNCH3
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

// does the backward of these arguments run the kernel that implements fuse_relu_bwd?
inline bool light_bwd_can_fuse_relu(const MrlaLightArgs& a) {
  if (a.layout != MRLA_NHWC || a.o == nullptr || a.act != MRLA_ACT_NONE) return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  TmaBwdPlan tpb, tpr;
  return tma_ptr_ok(a.x, a.bs_x, es) && tma_ptr_ok(a.o, a.bs_o, es) && tma_ptr_ok(a.dy, a.bs_dy, es) &&
         tma_ptr_ok(a.dx, a.bs_dx, es) && tma_ptr_ok(a.dout, a.bs_do, es) && make_tma_bwd_plan(a, &tpb) &&
         make_tma_ring_plan(a, &tpr);
}

// does the forward of these arguments run the sweep-1 variant that folds the bn3 affine on z (MODE 6)?
inline bool light_fwd_can_fold_bn(const MrlaLightArgs& a) {
  if (a.layout != MRLA_NHWC || a.o == nullptr || a.z == nullptr || a.act != MRLA_ACT_NONE || a.bn_mode != MRLA_BN_TRAIN)
    return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  TmaPlan tp1, tp2, tp5;
  return tma_ptr_ok(a.x, a.bs_x, es) && tma_ptr_ok(a.o, a.bs_o, es) && tma_ptr_ok(a.z, a.bs_z, es) &&
         (a.bs_y * es) % 4 == 0 && make_tma_plan(a, 1, 6, &tp1) && make_tma_plan(a, 1, 0, &tp2) &&
         make_tma_plan(a, 1, 6, &tp5, true);
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
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  TmaPlan tp1, tp2, tp5;
  // ---- v7 sweeps (light_v7.cuh): x re-formed on the fly (x_virtual) or a plain materialised x
  V7Plan v1, v2;
  const bool v7_base = LAYOUT == MRLA_NHWC && HAS_O && ACT == 0 && v7_ptr_ok(a.o, a.bs_o, es) && v7_ptr_ok(a.y, a.bs_y, es);
  if (a.x_virtual) {
    if (!(v7_base && full && a.z && a.z_coef && v7_ptr_ok(a.z, a.bs_z, es) && v7_plan(a, V7_S1, true, &v1) &&
          v7_plan(a, V7_S2, true, &v2)))
      return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);   // callers ask mrla_light_virtual_x() first
  }
  const bool v7x = a.x_virtual != 0;
  const bool v7p = !v7x && v7_base && a.z == nullptr && v7_ptr_ok(a.x, a.bs_x, es) && v7_plan(a, V7_S1, false, &v1) &&
                   v7_plan(a, V7_S2, false, &v2);
  const bool tma_ok = !v7x && LAYOUT == MRLA_NHWC && HAS_O && tma_ptr_ok(a.x, a.bs_x, es) && tma_ptr_ok(a.o, a.bs_o, es) &&
                      (a.bs_y * es) % 4 == 0 && make_tma_plan(a, 1, 6, &tp1) && make_tma_plan(a, 1, 0, &tp2);
  // optional producer fold: x = relu(z + o)
  bool x_ready = (a.z == nullptr) || v7x;
  if (!x_ready && !HAS_O) return MRLA_ERR_NULL;
  if (v7x) {
    rc = v7_launch_fwd<T, 0>(a, st, v1, true, a.z, a.bs_z, a.mom, 0, 0);
    if (rc) return rc;
  } else
  if (!x_ready && tma_ok && full && ACT == 0 && tma_ptr_ok(a.z, a.bs_z, es) && make_tma_plan(a, 1, 6, &tp5, true)) {
    // sweep 1 forms and stores x itself (MODE 5): no separate add+relu pass
    MrlaLightArgs a5 = a;
    a5.y = const_cast<void*>(a.x);
    a5.bs_y = a.bs_x;
    rc = a.z_coef ? launch_tma_sweep<T, 0, 6>(a5, st, tp5, a.z, a.bs_z, a.o, a.bs_o, nullptr, 0, a.mom)
                  : launch_tma_sweep<T, 0, 5>(a5, st, tp5, a.z, a.bs_z, a.o, a.bs_o, nullptr, 0, a.mom);
    if (rc) return rc;
    x_ready = true;
  } else {
    if (a.z_coef) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);   // callers ask mrla_light_fwd_folds_bn() first
    if (!x_ready) {
      const int64_t n = (int64_t)a.C * a.H * a.W;
      const int64_t v = 16 / es;
      if (a.bs_z != n || a.bs_o != n || a.bs_x != n || (a.B * n) % v || ((uintptr_t)a.z % 16) || ((uintptr_t)a.o % 16) ||
          ((uintptr_t)a.x % 16))
        return MRLA_ERR_ALIGN;
      const int64_t nv = a.B * n / v;
      int64_t blocks = (nv + 255) / 256;
      if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
      k_add_relu<T><<<(int)blocks, 256, 0, st>>>(static_cast<const T*>(a.z), o, const_cast<T*>(x), nv);
      MRLA_CHECK_LAUNCH();
    }
  // sweep 1
  if (v7p && full) {
    rc = v7_launch_fwd<T, 0>(a, st, v1, false, a.x, a.bs_x, a.mom, 0, 0);
    if (rc) return rc;
  } else if (tma_ok && full) {
    rc = launch_tma_sweep<T, ACT, 0>(a, st, tp1, a.x, a.bs_x, a.o, a.bs_o, nullptr, 0, a.mom);
    if (rc) return rc;
  } else {
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
  }  // sweep 1 (skipped when MODE 5 produced the moments together with x)
  // mid
  MidShape ms = mid_shape(a, full);
  if (mid_cluster_ok(a.C, a.dim_perhead, a.k_size)) {
    // one cluster launch: gate + BN statistics + coefficients
    cudaError_t e = launch_mid_cluster(k_light_mid_fwd, a.C, mid_cluster_ns(a.B), st, (const float*)a.mom, a.wq, a.wk, a.lam,
                                       a.gamma, a.beta, a.running_mean, a.running_var, a.drop_scale, a.gate, a.mean, a.rstd,
                                       a.coef, ms);
    if (e != cudaSuccess) return (int)e;
    MRLA_CHECK_LAUNCH();
  } else {
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    k_light_gate<<<a.B, th, 2 * a.C * sizeof(float), st>>>(a.mom, a.wq, a.wk, a.gate, ms);
    MRLA_CHECK_LAUNCH();
    k_light_bn_coef<<<(a.C + kMidCPB - 1) / kMidCPB, kMidCPB * kMidBL, 0, st>>>(a.mom, a.gate, a.lam, a.gamma, a.beta, a.running_mean,
                                                       a.running_var, a.drop_scale, a.mean, a.rstd, a.coef, ms);
    MRLA_CHECK_LAUNCH();
  }
  // sweep 2
  if (v7x || v7p) {
    // walks the batch downwards: the samples sweep 1 read last are the ones still in L2
    rc = v7x ? v7_launch_fwd<T, 1>(a, st, v2, true, a.z, a.bs_z, nullptr, 1, 1)
             : v7_launch_fwd<T, 1>(a, st, v2, false, a.x, a.bs_x, nullptr, 1, 0);
    if (rc) return rc;
  } else if (tma_ok) {
    rc = launch_tma_sweep<T, ACT, 1>(a, st, tp2, a.x, a.bs_x, a.o, a.bs_o, nullptr, 0, nullptr);
    if (rc) return rc;
  } else {
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
  const int es_ = a.dtype == MRLA_F32 ? 4 : 2;
  TmaBwdPlan tpb;
  const bool tma_b = LAYOUT == MRLA_NHWC && HAS_O && tma_ptr_ok(a.x, a.bs_x, es_) && tma_ptr_ok(a.o, a.bs_o, es_) &&
                     tma_ptr_ok(a.dy, a.bs_dy, es_) && (a.bs_dx * es_) % 4 == 0 && (a.bs_do * es_) % 4 == 0 &&
                     make_tma_bwd_plan(a, &tpb);
  TmaBwdPlan tpr;
  // ---- v7 sweeps
  V7Plan va, vb;
  const bool v7_base = LAYOUT == MRLA_NHWC && HAS_O && ACT == 0 && v7_ptr_ok(a.o, a.bs_o, es_) && v7_ptr_ok(a.dy, a.bs_dy, es_) &&
                       v7_ptr_ok(a.dx, a.bs_dx, es_) && v7_ptr_ok(a.dout, a.bs_do, es_);
  const bool v7x = a.x_virtual != 0;
  if (v7x) {
    if (!(v7_base && a.fuse_relu_bwd && a.bn_mode == MRLA_BN_TRAIN && a.z && a.z_coef && v7_ptr_ok(a.z, a.bs_z, es_) &&
          v7_plan(a, V7_SA, true, &va) && v7_plan(a, V7_SB, true, &vb)))
      return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);   // callers ask mrla_light_virtual_x() first
  }
  const bool v7p = !v7x && v7_base && !a.fuse_relu_bwd && v7_ptr_ok(a.x, a.bs_x, es_) && v7_plan(a, V7_SA, false, &va) &&
                   v7_plan(a, V7_SB, false, &vb);
  const bool v7 = v7x || v7p;
  const bool tma_r = !v7 && tma_b && tma_ptr_ok(a.dx, a.bs_dx, es_) && tma_ptr_ok(a.dout, a.bs_do, es_) && make_tma_ring_plan(a, &tpr);
  if (a.fuse_relu_bwd && !tma_r && !v7x) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);   // callers ask mrla_light_bwd_fuses_relu() first
  const int nparts = v7 ? vb.cpc : (tma_r ? tpr.maxslots : (tma_b ? tpb.maxslots : pb.grid_y));
  const size_t nfl = (size_t)nparts * a.C * 9 + (size_t)a.B * 2 * a.k_size + (v7x ? (size_t)vb.cpc * 2 * a.C : 0);
  const size_t need = nfl * sizeof(float);
  if (a.scratch == nullptr || a.scratch_bytes < need) return MRLA_ERR_WORKSPACE;
  float* wv_part = a.scratch;
  float* wqk_part = a.scratch + (size_t)nparts * a.C * 9;
  float* dz_part = wqk_part + (size_t)a.B * 2 * a.k_size;
  const bool full = (a.bn_mode == MRLA_BN_TRAIN);
  const T* x = static_cast<const T*>(a.x);
  const T* o = static_cast<const T*>(a.o);
  const T* dy = static_cast<const T*>(a.dy);
  T* dx = static_cast<T*>(a.dx);
  T* dout = static_cast<T*>(a.dout);
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  TmaPlan tpa;
  const bool tma_ok = !v7 && LAYOUT == MRLA_NHWC && HAS_O && tma_ptr_ok(a.x, a.bs_x, es) && tma_ptr_ok(a.o, a.bs_o, es) &&
                      tma_ptr_ok(a.dy, a.bs_dy, es) && make_tma_plan(a, 2, 3, &tpa);
  // sweep A
  if (v7) {
    rc = v7x ? v7_launch_fwd<T, 2>(a, st, va, true, a.z, a.bs_z, a.gmom, 0, 0)
             : v7_launch_fwd<T, 2>(a, st, va, false, a.x, a.bs_x, a.gmom, 0, 0);
    if (rc) return rc;
  } else if (tma_ok) {
    rc = launch_tma_sweep<T, ACT, 2>(a, st, tpa, a.x, a.bs_x, a.o, a.bs_o, a.dy, a.bs_dy, a.gmom);
    if (rc) return rc;
  } else {
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
    if (mid_cluster_ok(a.C, a.dim_perhead, a.k_size)) {
      ms.da_summed = 1;
      cudaError_t e = launch_mid_cluster(k_light_mid_bwd, a.C, mid_cluster_ns(a.B), st, (const float*)a.mom, (const float*)a.gmom,
                                         (const float*)a.gate, a.lam, a.gamma, a.drop_scale, (const float*)a.mean,
                                         (const float*)a.rstd, a.bcoef, a.dlam, a.dgamma, a.dbeta, ms);
      if (e != cudaSuccess) return (int)e;
    } else {
      k_light_bwd_chan<<<(a.C + kMidCPB - 1) / kMidCPB, kMidCPB * kMidBL, 0, st>>>(a.mom, a.gmom, a.gate, a.lam, a.gamma, a.drop_scale, a.mean,
                                                          a.rstd, a.bcoef, a.dlam, a.dgamma, a.dbeta, ms);
    }
    MRLA_CHECK_LAUNCH();
    int th = round_up(a.C < 1024 ? a.C : 1024, 32);
    const size_t sm = ((size_t)5 * a.C + a.C / a.dim_perhead) * sizeof(float);
    cudaError_t e = ensure_smem(k_light_bwd_gate, sm);
    if (e != cudaSuccess) return (int)e;
    k_light_bwd_gate<<<a.B, th, sm, st>>>(a.mom, a.wq, a.wk, a.gate, a.bcoef, wqk_part, ms);
    MRLA_CHECK_LAUNCH();
  }
  // sweep B
  if (v7) {
    rc = v7x ? v7_launch_bwd<T>(a, st, vb, true, true, a.z, a.bs_z, wv_part, dz_part, a.dz_sums)
             : v7_launch_bwd<T>(a, st, vb, false, false, a.x, a.bs_x, wv_part, nullptr, nullptr);
    if (rc) return rc;
  } else if (tma_r) {
    rc = launch_tma_bwd_ring<T, ACT>(a, st, tpr, wv_part);
    if (rc) return rc;
  } else if (tma_b) {
    rc = launch_tma_bwd<T, ACT>(a, st, tpb, wv_part);
    if (rc) return rc;
  } else {
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
    k_light_finish<<<(a.C * 9 + 255) / 256 + (2 * a.k_size + 7) / 8, 256, 0, st>>>(wv_part, nparts, wqk_part, a.dwv, a.dwq, a.dwk, a.B, a.C,
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
    if (a.act != MRLA_ACT_NONE) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);  // GELU variant is token-layout only (DeiT)
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
  int nparts = pb.grid_y;
  TmaBwdPlan tpb;
  if (make_tma_bwd_plan(a, &tpb) && tpb.maxslots > nparts) nparts = tpb.maxslots;
  if (make_tma_ring_plan(a, &tpb) && tpb.maxslots > nparts) nparts = tpb.maxslots;
  size_t extra = 0;
  V7Plan vb;
  for (int xf = 0; xf < 2; ++xf)
    if (a.o != nullptr && a.act == MRLA_ACT_NONE && v7_plan(a, V7_SB, xf != 0, &vb)) {
      if (vb.cpc > nparts) nparts = vb.cpc;
      extra = (size_t)vb.cpc * 2 * a.C;
    }
  return (size_t)nparts * a.C * 9 + (size_t)a.B * 2 * a.k_size + extra;
}

}  // namespace mrla
