// Host side of the v7 sweeps (light_v7.cuh): planning, cached tensor maps, launches.
#pragma once
#include <cstdio>
#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "../../include/mrla_b200.h"
#include "light_v7.cuh"

namespace mrla {

extern thread_local int g_launch_count;
extern thread_local char g_err_detail[192];
constexpr int kV7SMs = 148;

// ------------------------------------------------------------------------------------ tensor-map cache
// cuTensorMapEncodeTiled costs ~1.5 us and a call of the tail needs up to 14 maps; PyTorch's caching allocator hands the
// same addresses back step after step, so the encoded descriptors are kept per (pointer, geometry, box).
struct TmapKey {
  const void* base;
  int64_t bs;
  int32_t dtype, B, C, H, W, bc, bw, bh;
  bool operator==(const TmapKey& o) const {
    return base == o.base && bs == o.bs && dtype == o.dtype && B == o.B && C == o.C && H == o.H && W == o.W && bc == o.bc &&
           bw == o.bw && bh == o.bh;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    auto mix = [&h](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix((uint64_t)k.bs); mix((uint64_t)k.dtype | ((uint64_t)k.B << 8) | ((uint64_t)k.C << 32));
    mix((uint64_t)k.H | ((uint64_t)k.W << 20) | ((uint64_t)k.bc << 40));
    mix((uint64_t)k.bw | ((uint64_t)k.bh << 20));
    return (size_t)h;
  }
};
inline int cached_nhwc_tmap(CUtensorMap* out, const void* base, int dtype, int B, int C, int H, int W, int64_t bs, int box_c,
                            int box_w, int box_h) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, bs, dtype, B, C, H, W, box_c, box_w, box_h};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  const int rc = make_nhwc_tmap(out, base, dtype, B, C, H, W, bs, box_c, box_w, box_h);
  if (rc != 0) {
    snprintf(g_err_detail, sizeof(g_err_detail), "tensor map rc=%d base=%p dtype=%d B=%d C=%d H=%d W=%d bs=%lld box=(%d,%d,%d)", rc,
             base, dtype, B, C, H, W, (long long)bs, box_c, box_w, box_h);
  }
  if (rc == 0) {
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return rc;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel function, device).  Keyed by the function POINTER:
// every instantiation of a kernel template has the same C++ type, so a static per template type would be shared.
inline cudaError_t ensure_smem_once_ptr(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> done[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  std::lock_guard<std::mutex> lk(mu);
  auto it = done[dev].find(kernel);
  if (it != done[dev].end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done[dev][kernel] = bytes;
  return e;
}
template <typename Kern>
inline cudaError_t ensure_smem_once(Kern kernel, size_t bytes) {
  return ensure_smem_once_ptr(reinterpret_cast<const void*>(kernel), bytes);
}

// ------------------------------------------------------------------------------------ planning
enum { V7_S1 = 0, V7_S2 = 1, V7_SA = 2, V7_SB = 3 };

struct V7Plan {
  int CB, NQ, NT, U, TPU, S, ncb, cpc, grid, ncw, threads, ctas;
  int xcols, ocols, dycols;
  uint32_t x_bytes, o_bytes, dy_bytes, stage_bytes;
  size_t smem;
  bool ragged;
};

inline bool v7_ptr_ok(const void* ptr, int64_t bs, int es) {
  return ptr != nullptr && ((uintptr_t)ptr % 16 == 0) && ((bs * es) % 16 == 0);
}

// shape-level eligibility + resources of one sweep; xf = x is re-formed from (z, z_coef, o)
inline bool v7_plan(const MrlaLightArgs& a, int kind, bool xf, V7Plan* p) {
  if (a.layout != MRLA_NHWC || a.o == nullptr || a.act != MRLA_ACT_NONE) return false;
  if (a.C % 64 || a.H < 3 || a.W < 1 || a.W > 512) return false;   // 512: the width limit of the generic plan
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  // W > 56 (mmdet feature maps): column tiles of 8 groups of 7 columns; halo columns of a tile are real data of its
  // neighbours (the TMA box simply starts one / two columns to the left), only image borders are masked
  int NQ = a.W > 8 * kV7 ? 8 : (a.W + kV7 - 1) / kV7;
  // small batches of wide maps (detection: 2 images per GPU): one work unit per column tile instead of per image, and
  // narrower tiles, until the units cover the SMs (identical for all four sweeps: a function of B, C, W only)
  bool split = false;
  if (a.W > 8 * kV7 && (int64_t)a.B * (a.C / 64) < kV7SMs) {
    split = true;
    if ((int64_t)a.B * ((a.W + 8 * kV7 - 1) / (8 * kV7)) * (a.C / 64) < kV7SMs) NQ = 4;
  }
  p->NT = (a.W + NQ * kV7 - 1) / (NQ * kV7);
  p->TPU = split ? 1 : p->NT;
  p->U = split ? a.B * p->NT : a.B;
  int CB = 0;
  for (int cb : {256, 128, 64})
    if (a.C % cb == 0 && NQ * cb / 2 <= 256 && !(split && cb > 64)) { CB = cb; break; }
  if (CB == 0) return false;
  p->CB = CB; p->NQ = NQ;
  p->ncw = NQ * CB / 64;
  // sweeps 1 / 2 / A: one producer warp + the consumers; sweep B has no producer warp (thread 0 issues the loads) and
  // keeps ~220-240 registers per thread: one CTA per SM unless it is small
  p->threads = (kind == V7_SB ? 0 : 32) + 32 * p->ncw;
  if (kind == V7_SB) p->ctas = p->threads <= 128 ? 2 : 1;
  else p->ctas = p->threads <= 96 ? 4 : (p->threads <= 160 ? 2 : 1);
  const int halo = (kind == V7_SB) ? 4 : 2;
  p->xcols = NQ * kV7 + halo;
  p->ocols = xf ? p->xcols : NQ * kV7 + (kind == V7_SB ? 2 : 0);
  p->dycols = (kind == V7_SB) ? NQ * kV7 + 2 : (kind == V7_SA ? NQ * kV7 : 0);
  if (p->xcols > 256) return false;
  p->x_bytes = (uint32_t)p->xcols * CB * es;
  p->o_bytes = (uint32_t)p->ocols * CB * es;
  p->dy_bytes = (uint32_t)p->dycols * CB * es;
  p->stage_bytes = p->x_bytes + p->o_bytes + p->dy_bytes;
  size_t tail = 0;
  const int NP = CB / 2;
  if (kind == V7_S2) tail = (size_t)p->ncw * 2 * kV7 * 64 * es;
  else if (kind == V7_S1) tail = (size_t)2 * NQ * 6 * NP * sizeof(float2);
  else if (kind == V7_SA) tail = (size_t)2 * NQ * 3 * NP * sizeof(float2);
  else tail = (size_t)p->ncw * 3 * 2 * kV7 * 64 * es;
  const size_t budget = (size_t)227 * 1024 / p->ctas - 1024;
  if (kV7Hdr + tail + 4 * (size_t)p->stage_bytes > budget) return false;
  int S = (int)((budget - kV7Hdr - tail) / p->stage_bytes);
  if (S > 12) S = 12;
  const int smin = (kind == V7_SB) ? 5 : 4;
  if (S < smin) return false;
  if (kind == V7_SB && (size_t)NQ * 11 * NP * sizeof(float2) > (size_t)S * p->stage_bytes) return false;
  p->S = S;
  p->smem = kV7Hdr + (size_t)S * p->stage_bytes + tail;
  p->ncb = a.C / CB;
  int cpc = (kV7SMs * p->ctas) / p->ncb;
  if (cpc < 1) cpc = 1;
  if (cpc > p->U) cpc = p->U;
  p->cpc = cpc;
  p->grid = p->ncb * cpc;
  {  // sweep B finds (unit, tile, row) of a pipeline row with multiply-high reciprocals: exact while rows * d < 2^32
    const uint64_t per_unit = (uint64_t)p->TPU * a.H, n_my = ((uint64_t)p->U + cpc - 1) / cpc;
    if (n_my * per_unit * per_unit >= ((uint64_t)1 << 32)) return false;
  }
  p->ragged = (a.W % kV7) != 0 || p->NT > 1;
  return true;
}

// can forward AND backward of these arguments run without a materialised x (x_virtual)?
inline bool v7_virtual_x_ok(const MrlaLightArgs& a) {
  if (a.bn_mode != MRLA_BN_TRAIN || a.z == nullptr) return false;
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  V7Plan p;
  return v7_ptr_ok(a.z, a.bs_z, es) && v7_ptr_ok(a.o, a.bs_o, es) && v7_plan(a, V7_S1, true, &p) &&
         v7_plan(a, V7_S2, true, &p) && v7_plan(a, V7_SA, true, &p) && v7_plan(a, V7_SB, true, &p);
}

inline void v7_fill(V7Params* P, const MrlaLightArgs& a, const V7Plan& p) {
  P->B = a.B; P->C = a.C; P->H = a.H; P->W = a.W;
  auto magic = [](int d) { return (uint32_t)((((uint64_t)1 << 32) + (uint64_t)d - 1) / (uint64_t)d); };   // d >= 2
  const int ns = p.NT / p.TPU;
  P->mH = magic(a.H); P->mPU = magic(p.TPU * a.H); P->mNS = ns > 1 ? magic(ns) : 0;
  P->NQ = p.NQ; P->NT = p.NT; P->U = p.U; P->TPU = p.TPU; P->ncb = p.ncb; P->S = p.S; P->cpc = p.cpc; P->rev = 0; P->ncw = p.ncw; P->hint = 0;
  P->x_bytes = p.x_bytes; P->o_bytes = p.o_bytes; P->dy_bytes = p.dy_bytes; P->stage_bytes = p.stage_bytes;
  P->xo_cols = 0;
  P->wv = a.wv; P->zcoef = a.z_coef; P->coef = a.coef; P->mom = nullptr; P->res = a.residual ? 1.f : 0.f;
  P->lam = a.lam; P->bcoef = a.bcoef; P->wv_part = nullptr; P->dz_part = nullptr;
}

#define MRLA_V7_CHECK()                        \
  do {                                         \
    cudaError_t e_ = cudaGetLastError();       \
    if (e_ != cudaSuccess) return (int)e_;     \
    ++g_launch_count;                          \
  } while (0)

// kind: V7_S1 / V7_S2 / V7_SA.  xsrc = x (or the raw conv3 output when xf)
template <typename T, int MODE>
int v7_launch_fwd(const MrlaLightArgs& a, cudaStream_t st, const V7Plan& p, bool xf, const void* xsrc, int64_t bs_x,
                  float* mom, int rev, int hint) {
  CUtensorMap tx, to, tdy, ty;
  if (cached_nhwc_tmap(&tx, xsrc, a.dtype, a.B, a.C, a.H, a.W, bs_x, p.CB, p.xcols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (cached_nhwc_tmap(&to, a.o, a.dtype, a.B, a.C, a.H, a.W, a.bs_o, p.CB, p.ocols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  tdy = to;
  ty = to;
  if (MODE == 2 && cached_nhwc_tmap(&tdy, a.dy, a.dtype, a.B, a.C, a.H, a.W, a.bs_dy, p.CB, p.dycols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (MODE == 1 && cached_nhwc_tmap(&ty, a.y, a.dtype, a.B, a.C, a.H, a.W, a.bs_y, 64, kV7, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  V7Params P;
  v7_fill(&P, a, p);
  P.mom = mom;
  P.rev = rev;
  P.hint = hint;
  if (p.TPU != p.NT && MODE != 1 && mom != nullptr) {   // per-tile units add their moments atomically
    const size_t nacc = MODE == 0 ? 6 : 3;
    cudaError_t em = cudaMemsetAsync(mom, 0, nacc * (size_t)a.B * a.C * sizeof(float), st);
    if (em != cudaSuccess) return (int)em;
  }
  constexpr bool LEAN = std::is_same<T, __nv_bfloat16>::value;   // only bf16 instantiates the non-ragged variant
  const bool ragged = p.ragged || !LEAN;
  cudaError_t e = cudaSuccess;
#define MRLA_V7_L3(CBV, XFV, RAGV)                                       \
  {                                                                      \
    auto k = k_v7_fwd<T, CBV, XFV, MODE, RAGV>;                          \
    e = ensure_smem_once(k, p.smem);                                     \
    if (e != cudaSuccess) return (int)e;                                 \
    k<<<p.grid, p.threads, p.smem, st>>>(tx, to, tdy, ty, P);            \
  }
#define MRLA_V7_L2(CBV, XFV)                                             \
  {                                                                      \
    if (ragged) MRLA_V7_L3(CBV, XFV, true)                               \
    else MRLA_V7_L3(CBV, XFV, !LEAN)                                    \
  }
#define MRLA_V7_L1(CBV)                                                  \
  {                                                                      \
    if (xf) MRLA_V7_L2(CBV, true) else MRLA_V7_L2(CBV, false)            \
  }
  if (p.CB == 64) MRLA_V7_L1(64)
  else if (p.CB == 128) MRLA_V7_L1(128)
  else MRLA_V7_L1(256)
#undef MRLA_V7_L1
#undef MRLA_V7_L2
#undef MRLA_V7_L3
  MRLA_V7_CHECK();
  return MRLA_OK;
}

// sweep B.  dz_sums != nullptr (xf + fuse): [2,C] sum dz, sum dz*c3 are produced as well.
template <typename T>
int v7_launch_bwd(const MrlaLightArgs& a, cudaStream_t st, const V7Plan& p, bool xf, bool fuse, const void* xsrc,
                  int64_t bs_x, float* wv_part, float* dz_part, float* dz_sums) {
  if (xf != fuse) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  CUtensorMap tx, to, tdy, tdx, tdo;
  if (cached_nhwc_tmap(&tx, xsrc, a.dtype, a.B, a.C, a.H, a.W, bs_x, p.CB, p.xcols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (cached_nhwc_tmap(&to, a.o, a.dtype, a.B, a.C, a.H, a.W, a.bs_o, p.CB, p.ocols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (cached_nhwc_tmap(&tdy, a.dy, a.dtype, a.B, a.C, a.H, a.W, a.bs_dy, p.CB, p.dycols, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (cached_nhwc_tmap(&tdx, a.dx, a.dtype, a.B, a.C, a.H, a.W, a.bs_dx, 64, kV7, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  if (cached_nhwc_tmap(&tdo, a.dout, a.dtype, a.B, a.C, a.H, a.W, a.bs_do, 64, kV7, 1)) return MRLA_FAIL(MRLA_ERR_UNSUPPORTED);
  V7Params P;
  v7_fill(&P, a, p);
  P.rev = 1;    // sweep A walked the batch upwards
  P.hint = 1;   // nothing sweep B reads is needed again soon
  const int es = a.dtype == MRLA_F32 ? 4 : 2;
  P.xo_cols = xf ? (uint32_t)p.CB * es : 0;
  P.wv_part = wv_part;
  P.dz_part = dz_part;
  constexpr bool LEAN = std::is_same<T, __nv_bfloat16>::value;
  const bool ragged = p.ragged || !LEAN;
  cudaError_t e = cudaSuccess;
#define MRLA_V7_B3(CBV, XFV, RAGV)                                       \
  {                                                                      \
    auto k = k_v7_bwd<T, CBV, XFV, XFV, RAGV>;                           \
    e = ensure_smem_once(k, p.smem);                                     \
    if (e != cudaSuccess) return (int)e;                                 \
    k<<<p.grid, p.threads, p.smem, st>>>(tx, to, tdy, tdx, tdo, P);      \
  }
#define MRLA_V7_B2(CBV, XFV)                                             \
  {                                                                      \
    if (ragged) MRLA_V7_B3(CBV, XFV, true)                               \
    else MRLA_V7_B3(CBV, XFV, !LEAN)                                    \
  }
#define MRLA_V7_B1(CBV)                                                  \
  {                                                                      \
    if (xf) MRLA_V7_B2(CBV, true) else MRLA_V7_B2(CBV, false)            \
  }
  if (p.CB == 64) MRLA_V7_B1(64)
  else if (p.CB == 128) MRLA_V7_B1(128)
  else MRLA_V7_B1(256)
#undef MRLA_V7_B1
#undef MRLA_V7_B2
#undef MRLA_V7_B3
  MRLA_V7_CHECK();
  if (xf && fuse && dz_sums != nullptr) {
    k_v7_dz_finish<<<(2 * a.C + 255) / 256, 256, 0, st>>>(dz_part, p.cpc, a.C, dz_sums);
    MRLA_V7_CHECK();
  }
  return MRLA_OK;
}

}  // namespace mrla
