// v7 NHWC sweeps of the MRLA-light tail (round 2): seven columns per thread, x never materialised.
//
// What changed against the v2..v6 kernels (light_nhwc_tma.cuh / light_nhwc_ring.cuh), and why:
//   * XF ("x formed"): the bottleneck's  x = relu(bn3(c3) + identity)  (resnet_mrla_light.py:101-102,113-114) is
//     re-formed from the RAW conv3 output and the identity inside EVERY sweep, in storage precision, exactly as
//     sweep 1 MODE 6 used to do before storing it.  x is no longer written or read: forward reads c3,id (sweep 1),
//     c3,id -> y (sweep 2); backward reads dy,c3,id (sweep A), dy,c3,id -> dz,d_id (sweep B) = 13 N instead of 14 N
//     of DRAM traffic, and one activation less is kept alive per block.
//   * seven columns per thread (56 = 8*7, 28 = 4*7, 14 = 2*7, 7 = 1*7): no idle column lanes at any ResNet width
//     and 9/7 instead of 6/4 window loads per output column.
//   * one image row per pipeline stage, row loop peeled (first rows / steady state / last rows are separate template
//     instances): the steady-state step has no per-row conditionals, running pointers or select chains.  The v6
//     sweeps spent about half of their issue slots on those (43 instructions per output pair, 21 of them useful).
//   * outputs leave through per-warp shared-memory staging rows and one TMA tensor store per warp and row: no
//     per-thread global addresses or store predicates; ragged edges are clipped by the TMA unit.
//   * sweep B recomputes the T halo column (x window of 11 columns) instead of exchanging T rows between warps: no
//     CTA-wide barrier per row (the v4 ring kernel ran at 11 % warps-active because of it), warps only meet at the
//     end of the kernel.  dX rows are built in scatter form (three rotating row accumulators).
//   * sweep B also accumulates bn3's backward reductions  sum dz, sum dz*c3  per channel, so the
//     separate k_bn_bwd_reduce pass over dz and c3 (2 N) disappears.
// Thread map (as before): lanes of a warp = 32 consecutive channel pairs of one pixel (conflict-free LDS.32), warp =
// (column group q, 64-channel sub-block); NQ*CB/64 consumer warps per CTA (+ 1 producer warp in sweeps 1 / 2 / A).
// Replaces (paths relative to /root/reference) mrla_light_module.py:56-72, resnet_mrla_light.py:42,101-102,113-116.
#pragma once
#include "light_nhwc_ring.cuh"

namespace mrla {

constexpr int kV7 = 7;   // output columns per consumer thread

struct V7Params {
  int B, C, H, W;
  int NQ, ncb, S, cpc;     // column groups, channel blocks, pipeline stages, CTAs per channel block (grid = ncb*cpc)
  int NT;                  // column tiles of NQ*7 columns per image (W > 56: mmdet feature maps); 1 otherwise
  uint32_t mH, mPU, mNS;   // ceil(2^32 / d) for d = H, TPU*H, NT/TPU (v7_fdiv; exact for a*d < 2^32)
  int U, TPU;              // work units and column tiles per unit: (B, NT) = one unit per image, or (B*NT, 1) = one unit per
                           // tile for small batches (the per-image moments are then accumulated with atomics)
  int rev;                 // 1: walk the batch from the last sample down (the previous sweep left that end in L2)
  int ncw;                 // consumer warps per CTA
  int hint;                // L2 policy of the tile loads: 0 default, 1 evict_first
  uint32_t x_bytes, o_bytes, dy_bytes, stage_bytes;   // per stage (one image row of each tile)
  uint32_t xo_cols;        // byte offset of a thread's first OWN column inside its o-tile row (CS if the o tile has halo)
  const float* wv;         // [C,9]
  const float* zcoef;      // XF: [2,C] bn3 coefficients (a_c, b_c)
  const float* coef;       // sweep 2: [3,B,C]
  float* mom;              // sweep 1: [6,B,C] ; sweep A: [3,B,C]
  float res;
  // sweep B
  const float* lam;        // [C]
  const float* bcoef;      // [7,B,C]
  float* wv_part;          // [cpc, C, 9]
  float* dz_part;          // [cpc, 2, C]   sum dz, sum dz*c3   (XF + FUSE)
};

// ---------------------------------------------------------------------------- small helpers
// a / d for 0 <= a, a*d < 2^32, with m = ceil(2^32 / d) from the host (d == 1: m is unused)
__device__ __forceinline__ int v7_fdiv(int a, int d, uint32_t m) { return d == 1 ? a : (int)__umulhi((uint32_t)a, m); }
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_4d_hint(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar_s, int c0, int c1,
                                                 int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_dst), "l"(tmap), "r"(bar_s), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_s, uint32_t parity) {
  while (!mbar_try_wait_s(bar_s, parity)) {
  }
}
__device__ __forceinline__ bool mbar_test_wait_s(uint32_t bar_s, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_s), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
template <typename T> __device__ __forceinline__ typename RawPair<T>::type raw_zero();
template <> __device__ __forceinline__ uint32_t raw_zero<__nv_bfloat16>() { return 0u; }
template <> __device__ __forceinline__ uint32_t raw_zero<__half>() { return 0u; }
template <> __device__ __forceinline__ float2 raw_zero<float>() { return make_float2(0.f, 0.f); }

// x = relu(round(a*z + b) + identity) in storage precision (the value the reference's bf16 graph holds after
// `out = bn3(out); out += identity; out = relu(out)`), returned unpacked
template <typename T>
__device__ __forceinline__ float2 form_x(typename RawPair<T>::type zr, typename RawPair<T>::type ir, float2 za, float2 zb,
                                         bool valid) {
  typename RawPair<T>::type z2 = pack_pair<T>(ffma2(za, unpack_pair<T>(zr), zb));
  typename RawPair<T>::type xr = raw_relu<T>(raw_add<T>(z2, ir));
  if (!valid) xr = raw_zero<T>();   // one select on the packed pair
  return unpack_pair<T>(xr);
}
// same with the column's validity as DATA (all ones / zero): one AND, no predicate logic in the row loop
__device__ __forceinline__ uint32_t raw_and(uint32_t a, uint32_t m) { return a & m; }
__device__ __forceinline__ float2 raw_and(float2 a, uint32_t m) {
  return make_float2(__uint_as_float(__float_as_uint(a.x) & m), __uint_as_float(__float_as_uint(a.y) & m));
}
template <typename T>
__device__ __forceinline__ float2 form_x_m(typename RawPair<T>::type zr, typename RawPair<T>::type ir, float2 za, float2 zb,
                                           uint32_t m) {
  typename RawPair<T>::type z2 = pack_pair<T>(ffma2(za, unpack_pair<T>(zr), zb));
  return unpack_pair<T>(raw_and(raw_relu<T>(raw_add<T>(z2, ir)), m));
}

// shared-memory header: full barriers at +0, empty barriers at +128 (up to 16 stages each)
constexpr int kV7Hdr = 256;

// producer warp: lane 0 streams (sample, channel block) images one row per stage
template <int NT>
__device__ __forceinline__ void v7_producer(const CUtensorMap* tm0, const CUtensorMap* tm1, const CUtensorMap* tm2,
                                            const V7Params& P, uint32_t bar_s, uint32_t stages_s, int cb, int m, int CB,
                                            int w0, int w1, int w2) {
  tma_prefetch_desc(tm0);
  if (NT > 1) tma_prefetch_desc(tm1);
  if (NT > 2) tma_prefetch_desc(tm2);
  const uint64_t pol = l2_policy_evict_first();
  int st = 0;
  uint32_t ph = 1;   // first pass over the ring: slots are free
  const int NS = P.NT / P.TPU;   // units per image
  for (int u = m; u < P.U; u += P.cpc) {
    const int bi = v7_fdiv(u, NS, P.mNS), tile0 = (u - bi * NS) * P.TPU;
    const int b = P.rev ? P.B - 1 - bi : bi;
    for (int tile = tile0; tile < tile0 + P.TPU; ++tile) {
      const int t0 = tile * P.NQ * kV7;   // first image column of the tile
      for (int r = 0; r < P.H; ++r) {
        mbar_wait_s(bar_s + 128 + st * 8, ph);
        const uint32_t fb = bar_s + st * 8;
        const uint32_t dst = stages_s + (uint32_t)st * P.stage_bytes;
        mbar_expect_tx_s(fb, P.stage_bytes);
        if (P.hint) {
          tma_load_4d_hint(dst, tm0, fb, cb * CB, t0 + w0, r, b, pol);
          if (NT > 1) tma_load_4d_hint(dst + P.x_bytes, tm1, fb, cb * CB, t0 + w1, r, b, pol);
          if (NT > 2) tma_load_4d_hint(dst + P.x_bytes + P.o_bytes, tm2, fb, cb * CB, t0 + w2, r, b, pol);
        } else {
          tma_load_4d_s(dst, tm0, fb, cb * CB, t0 + w0, r, b);
          if (NT > 1) tma_load_4d_s(dst + P.x_bytes, tm1, fb, cb * CB, t0 + w1, r, b);
          if (NT > 2) tma_load_4d_s(dst + P.x_bytes + P.o_bytes, tm2, fb, cb * CB, t0 + w2, r, b);
        }
        if (++st == P.S) { st = 0; ph ^= 1; }
      }
    }
  }
}

// =====================================================================================================
// forward family: MODE 0 sweep 1 (moments), MODE 1 sweep 2 (y), MODE 2 sweep A (backward moments)
// =====================================================================================================
template <typename T, int CB, bool XF, int MODE, bool RAGGED>
struct V7Fwd {
  static constexpr int K = kV7;
  static constexpr int KW = K + 2;
  static constexpr int NP = CB / 2;
  static constexpr int ES = sizeof(T);
  static constexpr uint32_t CS = CB * ES;        // bytes between adjacent tile columns
  static constexpr uint32_t OW = 64 * ES;        // bytes between adjacent columns of a warp's staging row
  static constexpr int NACC = MODE == 0 ? 6 : (MODE == 2 ? 3 : 0);
  typedef typename RawPair<T>::type Raw;

  const V7Params& P;
  const CUtensorMap* tm_y;
  uint32_t bar_s, stages_s, tbase;
  uint32_t obuf;             // this warp's staging rows (2 buffers), lane offset included
  int lane, ycol, ychan;     // TMA store coordinates of the warp (first own column, first channel)
  uint32_t vmask;            // bit j: window column j lies inside the image
  uint32_t em0, em1;         // validity of the two halo columns as data (all ones / zero)
  float2 w9[9], za, zb;
  float2 cA, cL, cD;
  float2 acc[NACC > 0 ? NACC : 1], accb[NACC > 0 ? NACC : 1];
  int st, ob, b, r;
  uint32_t ph;
  float2 win[3][KW];
  Raw oc[3][K];              // own-column o (identity) of the rows in flight, storage precision
  Raw gc[3][MODE == 2 ? K : 1];

  __device__ __forceinline__ V7Fwd(const V7Params& P_) : P(P_) {}

  __device__ __forceinline__ bool col_ok(int j) const { return (vmask >> j) & 1u; }
  // column tile starting at image column t0: validity of this thread's window columns, TMA store column
  __device__ __forceinline__ void set_tile(int t0, int q) {
    vmask = 0;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      const int col = t0 + q * K - 1 + j;
      if (col >= 0 && col < P.W) vmask |= 1u << j;
    }
    em0 = col_ok(0) ? 0xffffffffu : 0u;
    em1 = col_ok(KW - 1) ? 0xffffffffu : 0u;
    ycol = t0 + q * K;
  }

  template <int I, bool FETCH, bool OUT>
  __device__ __forceinline__ void step() {
    if (FETCH) {
      mbar_wait_s(bar_s + st * 8, ph);
      const uint32_t xa = stages_s + (uint32_t)st * P.stage_bytes + tbase;
      const uint32_t oa = xa + P.x_bytes;
      if (XF) {
#pragma unroll
        for (int j = 0; j < KW; ++j) {
          const Raw zr = lds_raw<T>(xa + j * CS);
          const Raw ir = lds_raw<T>(oa + j * CS);
          if (RAGGED) win[I][j] = form_x<T>(zr, ir, za, zb, col_ok(j));
          else if (j == 0) win[I][j] = form_x_m<T>(zr, ir, za, zb, em0);
          else if (j == KW - 1) win[I][j] = form_x_m<T>(zr, ir, za, zb, em1);
          else win[I][j] = form_x<T>(zr, ir, za, zb, true);
          if (j >= 1 && j <= K) oc[I][j - 1] = ir;
        }
      } else {
#pragma unroll
        for (int j = 0; j < KW; ++j) win[I][j] = lds_pair<T>(xa + j * CS);
#pragma unroll
        for (int j = 0; j < K; ++j) oc[I][j] = lds_raw<T>(oa + j * CS);
      }
      if (MODE == 2) {
        const uint32_t ga = oa + P.o_bytes;
#pragma unroll
        for (int j = 0; j < K; ++j) gc[I][j] = lds_raw<T>(ga + j * CS);
      }
      // everything this row contributes is in registers: hand the stage back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bar_s + 128 + st * 8);
      if (++st == P.S) { st = 0; ph ^= 1; }
    } else {
#pragma unroll
      for (int j = 0; j < KW; ++j) win[I][j] = f2(0.f, 0.f);
    }
    if (OUT) {
      // output row r-1: top = row r-2, mid = row r-1, bot = row r
      const float2(&top)[KW] = win[(I + 1) % 3];
      const float2(&mid)[KW] = win[(I + 2) % 3];
      const float2(&bot)[KW] = win[I];
      float2 u[K];
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = fmul2(w9[0], top[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[1], top[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[2], top[j + 2], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[3], mid[j], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[4], mid[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[5], mid[j + 2], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[6], bot[j], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[7], bot[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) u[j] = ffma2(w9[8], bot[j + 2], u[j]);
      const Raw(&orow)[K] = oc[(I + 2) % 3];
      if (MODE == 1) {
        if (lane == 0) bulk_wait_read<1>();   // the staging row written two rows ago has left
        __syncwarp();
        const uint32_t dst = obuf + (uint32_t)ob * (K * OW);
        const float2 res2 = f2(P.res, P.res);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float2 t = ffma2(cA, u[j], cD);
          t = ffma2(cL, unpack_pair<T>(orow[j]), t);
          t = ffma2(res2, mid[j + 1], t);
          sts_raw(dst + j * OW, pack_pair<T>(t));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(tm_y, dst - (uint32_t)lane * 2 * ES, ychan, ycol, r - 1, b);
          bulk_commit();
        }
        ob ^= 1;
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float2 v = u[j];
          if (RAGGED && !col_ok(j + 1)) v = f2(0.f, 0.f);
          const float2 ov = unpack_pair<T>(orow[j]);
          float2(&A)[NACC > 0 ? NACC : 1] = (j & 1) ? accb : acc;
          if (MODE == 0) {
            A[0] = fadd2(A[0], mid[j + 1]);
            A[1] = fadd2(A[1], v);
            A[2] = ffma2(v, v, A[2]);
            A[3] = ffma2(v, ov, A[3]);
            A[4] = fadd2(A[4], ov);
            A[5] = ffma2(ov, ov, A[5]);
          } else {
            const float2 gv = unpack_pair<T>(gc[(I + 2) % 3][j]);
            A[0] = fadd2(A[0], gv);
            A[1] = ffma2(gv, v, A[1]);
            A[2] = ffma2(gv, ov, A[2]);
          }
        }
      }
    }
    ++r;
  }

  // one image of H >= 3 rows
  __device__ __forceinline__ void image() {
    const int H = P.H;
    r = 0;
#pragma unroll
    for (int j = 0; j < KW; ++j) win[2][j] = f2(0.f, 0.f);   // row -1
    step<0, true, false>();
    step<1, true, true>();
    step<2, true, true>();
    while (r + 2 < H) {
      step<0, true, true>();
      step<1, true, true>();
      step<2, true, true>();
    }
    // r % 3 == 0; up to two more rows, then the zero row H
    if (r < H) {
      step<0, true, true>();
      if (r < H) {
        step<1, true, true>();
        step<2, false, true>();
      } else {
        step<1, false, true>();
      }
    } else {
      step<0, false, true>();
    }
  }
};

template <typename T, int CB, bool XF, int MODE, bool RAGGED>
__global__ void __launch_bounds__(288, 1)
k_v7_fwd(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_o,
         const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_y,
         const __grid_constant__ V7Params P) {
  typedef V7Fwd<T, CB, XF, MODE, RAGGED> F;
  constexpr int NP = CB / 2;
  constexpr int ES = sizeof(T);
  constexpr int K = kV7;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 16;
  unsigned char* stages = smem_raw + kV7Hdr;
  unsigned char* tail = stages + (size_t)P.S * P.stage_bytes;   // MODE 1: staging rows ; MODE 0/2: reduction scratch
  const int cb = blockIdx.x / P.cpc, m = blockIdx.x - cb * P.cpc;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], P.ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t bar_s = smem_u32(smem_raw), stages_s = smem_u32(stages);
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      v7_producer<(MODE == 2 ? 3 : 2)>(&tm_x, &tm_o, &tm_dy, P, bar_s, stages_s, cb, m, CB, -1, XF ? -1 : 0, 0);
    return;
  }
  const int ct = threadIdx.x - 32;
  const int p = ct % NP, q = ct / NP;
  const int warp = ct >> 5;
  F S(P);
  S.tm_y = &tm_y;
  S.bar_s = bar_s;
  S.stages_s = stages_s;
  S.lane = threadIdx.x & 31;
  S.tbase = (uint32_t)(q * K) * F::CS + (uint32_t)p * 2 * ES;
  S.obuf = smem_u32(tail) + (uint32_t)warp * (2 * K * F::OW) + (uint32_t)S.lane * 2 * ES;
  S.ychan = cb * CB + ((warp * 32) % NP) * 2;
  S.set_tile(0, q);
  S.st = 0;
  S.ph = 0;
  S.ob = 0;
  const int c = cb * CB + 2 * p;   // C % 64 == 0 and CB | C : always a valid channel pair
#pragma unroll
  for (int i = 0; i < 9; ++i) S.w9[i] = f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]);
  S.za = S.zb = f2(0.f, 0.f);
  if (XF) {
    S.za = *reinterpret_cast<const float2*>(P.zcoef + c);
    S.zb = *reinterpret_cast<const float2*>(P.zcoef + P.C + c);
  }
  const int64_t BC = (int64_t)P.B * P.C;
  float2* red_base = reinterpret_cast<float2*>(tail);   // [2][NQ][NACC][NP]
  int red_sel = 0;
  const int NS = P.NT / P.TPU;   // units per image
  for (int u = m; u < P.U; u += P.cpc) {
    const int bi = v7_fdiv(u, NS, P.mNS), tile0 = (u - bi * NS) * P.TPU;
    const int b = P.rev ? P.B - 1 - bi : bi;
    S.b = b;
    if (MODE == 1) {
      const float* cp = P.coef + (int64_t)b * P.C + c;
      S.cA = *reinterpret_cast<const float2*>(cp);
      S.cL = *reinterpret_cast<const float2*>(cp + BC);
      S.cD = *reinterpret_cast<const float2*>(cp + 2 * BC);
    }
#pragma unroll
    for (int i = 0; i < (F::NACC > 0 ? F::NACC : 1); ++i) { S.acc[i] = f2(0.f, 0.f); S.accb[i] = f2(0.f, 0.f); }
    for (int tile = tile0; tile < tile0 + P.TPU; ++tile) {   // W > 56: column tiles; the moments run over all of them
      if (P.NT > 1) S.set_tile(tile * P.NQ * K, q);
      S.image();                                              // (one call site: the body is ~10k instructions)
    }
    if (F::NACC > 0) {
      if (P.NQ == 1) {
#pragma unroll
        for (int i = 0; i < F::NACC; ++i)
          *reinterpret_cast<float2*>(P.mom + (int64_t)i * BC + (int64_t)b * P.C + c) = fadd2(S.acc[i], S.accb[i]);
      } else {
        // deterministic reduction over the NQ column groups (double-buffered scratch: one consumer barrier per image)
        float2* red = red_base + (size_t)red_sel * P.NQ * F::NACC * NP;
        red_sel ^= 1;
#pragma unroll
        for (int i = 0; i < F::NACC; ++i) red[((size_t)q * F::NACC + i) * NP + p] = fadd2(S.acc[i], S.accb[i]);
        named_bar_sync(1, P.ncw * 32);
        for (int idx = ct; idx < F::NACC * NP; idx += P.ncw * 32) {
          const int mi = idx / NP, pp = idx - mi * NP;
          float2 s = f2(0.f, 0.f);
          for (int qq = 0; qq < P.NQ; ++qq) {
            const float2 v = red[((size_t)qq * F::NACC + mi) * NP + pp];
            s.x += v.x;
            s.y += v.y;
          }
          float* dst = P.mom + (int64_t)mi * BC + (int64_t)b * P.C + cb * CB + 2 * pp;
          if (NS == 1) {
            *reinterpret_cast<float2*>(dst) = s;
          } else {   // the image's tiles are spread over CTAs: the launcher zeroed `mom`
            atomicAdd(dst, s.x);
            atomicAdd(dst + 1, s.y);
          }
        }
      }
    }
  }
  if (MODE == 1 && S.lane == 0) bulk_wait_all<0>();
}

// =====================================================================================================
// sweep B:  dS = Q0 + Q1*dy + Q2*V + Q3*o ;  T = a*dS ;  do = lam*dS ;  dx = res*dy + dyc + dwconv3x3^T(T)
//           dWv[c,i,j] = sum T[h,w]*x[h+i-1,w+j-1]
//   FUSE: x = relu(z' + o) was formed in front of the tail -> `dx` receives dz = dx*[x>0], `dout` the total identity
//         gradient lam*dS + dz; with XF also sum dz, sum dz*c3 per channel (bn3's backward reductions).
// Step s (fetches row s):  x row s -> window;  T row s-1 on K+2 columns (halo recomputed) from the dy / o rows of the
// stage fetched one step ago;  T scattered into the dX rows s-2, s-1, s;  dX row s-2 finished, staged and stored.
// =====================================================================================================
template <typename T, int CB, bool XF, bool FUSE, bool RAGGED>
struct V7Bwd {
  static constexpr int K = kV7;
  static constexpr int KT = K + 2;    // T columns (one halo column each side)
  static constexpr int KX = K + 4;    // x window columns
  static constexpr int NP = CB / 2;
  static constexpr int ES = sizeof(T);
  static constexpr uint32_t CS = CB * ES;
  static constexpr uint32_t OW = 64 * ES;
  static constexpr uint32_t OROW = K * OW;             // one staging row of a warp
  static constexpr bool HOLD2 = XF && FUSE;            // the c3 row is read again two steps after its fetch
  typedef typename RawPair<T>::type Raw;

  const V7Params& P;
  const CUtensorMap* tm_dx;
  const CUtensorMap* tm_do;
  uint32_t bar_s, stages_s, tbase;
  uint32_t obuf;             // this warp's staging: [3 rows][dx row | do row], lane offset included
  int lane, ycol, ychan;
  uint32_t vmask;            // bit j: x-window column j (image column q*K-2+j) lies inside the image
  uint32_t em[4];            // the same as data for the four window columns that can fall outside (0, 1, KX-2, KX-1)
  float2 tm0, tm1;           // 1 / 0 for the two T halo columns
  float2 w9[9], za, zb, lm;
  float2 q0, q1, q2, q3, ta, dyc;
  float2 dw[9];
  float2 sdz, sdzc;
  int st, b, r;
  uint32_t ph;
  int sidx[3];               // pipeline stage that holds row s, per window slot
  float2 xw[3][KX];
  float2 da[3][K];           // dX row accumulators, row h lives in slot h % 3
  // No producer warp (8 consumer warps + 1 would be allocated like 12 warps — registers are granted per 4 warps — which
  // caps a thread at 168 registers; at 256 threads this kernel gets its ~220).  Stages are refilled by whichever warp
  // releases them LAST: every warp bumps a per-stage shared counter when it is done with a stage, and the lane whose
  // increment completes the round requests the row that goes into that stage next (global row + S) on the spot.
  const CUtensorMap* tm_x;
  const CUtensorMap* tm_o;
  const CUtensorMap* tm_dy;
  int cb, m, rows_total;     // rows_total = rows of all images of this CTA
  int grow;                  // global index (over the CTA's images) of the row being fetched
  int gidx[3];               // global row index per window slot
  uint32_t cnt_s;            // shared address of the per-stage release counters
  uint64_t pol;

  __device__ __forceinline__ V7Bwd(const V7Params& P_) : P(P_) {}
  __device__ __forceinline__ bool col_ok(int j) const { return (vmask >> j) & 1u; }
  // request global row g (of this CTA's row sequence) into stage sx
  __device__ __forceinline__ void issue_row(int g, int sx) {
    const int per_unit = P.TPU * P.H, NS = P.NT / P.TPU;
    const int ui = v7_fdiv(g, per_unit, P.mPU), rem = g - ui * per_unit;
    const int t = v7_fdiv(rem, P.H, P.mH), row = rem - t * P.H;
    const int u = m + ui * P.cpc;
    const int bi = v7_fdiv(u, NS, P.mNS);
    const int t0 = ((u - bi * NS) * P.TPU + t) * P.NQ * K;
    const int bb = P.rev ? P.B - 1 - bi : bi;
    const uint32_t fb = bar_s + sx * 8;
    const uint32_t dst = stages_s + (uint32_t)sx * P.stage_bytes;
    mbar_expect_tx_s(fb, P.stage_bytes);
    tma_load_4d_hint(dst, tm_x, fb, cb * CB, t0 - 2, row, bb, pol);
    tma_load_4d_hint(dst + P.x_bytes, tm_o, fb, cb * CB, t0 + (XF ? -2 : -1), row, bb, pol);
    tma_load_4d_hint(dst + P.x_bytes + P.o_bytes, tm_dy, fb, cb * CB, t0 - 1, row, bb, pol);
  }
  // column tile starting at image column t0: validity of the x-window / T columns, TMA store column
  __device__ __forceinline__ void set_tile(int t0, int q) {
    vmask = 0;
#pragma unroll
    for (int j = 0; j < KX; ++j) {
      const int col = t0 + q * K - 2 + j;
      if (col >= 0 && col < P.W) vmask |= 1u << j;
    }
    em[0] = col_ok(0) ? 0xffffffffu : 0u;
    em[1] = col_ok(1) ? 0xffffffffu : 0u;
    em[2] = col_ok(KX - 2) ? 0xffffffffu : 0u;
    em[3] = col_ok(KX - 1) ? 0xffffffffu : 0u;
    tm0 = col_ok(1) ? f2(1.f, 1.f) : f2(0.f, 0.f);
    tm1 = col_ok(KX - 2) ? f2(1.f, 1.f) : f2(0.f, 0.f);
    ycol = t0 + q * K;
  }
  // this warp is done with stage sx, which held global row g
  __device__ __forceinline__ void release(int sx, int g) {
    __syncwarp();
    if (lane == 0) {
      uint32_t old;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt_s + sx * 4) : "memory");
      if (old + 1 == (uint32_t)P.ncw) {
        // last warp of the round: rearm the counter (nobody touches it before the refilled stage's barrier completes, and
        // the expect_tx arrive below releases this store), then refill the stage with global row g + S
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt_s + sx * 4), "r"(0u) : "memory");
        if (g + P.S < rows_total) issue_row(g + P.S, sx);
      }
    }
  }

  // I = s % 3.  FETCH: s < H.  TROW: 1 <= s <= H (T row s-1).  TFIRST: s == 1.  TLAST: s == H.  FIN: s >= 2.
  template <int I, bool FETCH, bool TROW, bool TFIRST, bool TLAST, bool FIN>
  __device__ __forceinline__ void step() {
    constexpr int IM1 = (I + 2) % 3;   // slot of row s-1
    constexpr int IM2 = (I + 1) % 3;   // slot of row s-2
    if (FETCH) {
      mbar_wait_s(bar_s + st * 8, ph);
      sidx[I] = st;
      gidx[I] = grow++;
      const uint32_t xa = stages_s + (uint32_t)st * P.stage_bytes + tbase;
      if (XF) {
        const uint32_t oa = xa + P.x_bytes;
#pragma unroll
        for (int j = 0; j < KX; ++j) {
          const Raw zr = lds_raw<T>(xa + j * CS);
          const Raw ir = lds_raw<T>(oa + j * CS);
          if (RAGGED) xw[I][j] = form_x<T>(zr, ir, za, zb, col_ok(j));
          else if (j < 2) xw[I][j] = form_x_m<T>(zr, ir, za, zb, j == 0 ? em[0] : em[1]);
          else if (j >= KX - 2) xw[I][j] = form_x_m<T>(zr, ir, za, zb, j == KX - 2 ? em[2] : em[3]);
          else xw[I][j] = form_x<T>(zr, ir, za, zb, true);
        }
      } else {
#pragma unroll
        for (int j = 0; j < KX; ++j) xw[I][j] = lds_pair<T>(xa + j * CS);
      }
      if (++st == P.S) { st = 0; ph ^= 1; }
    } else if (TROW) {
#pragma unroll
      for (int j = 0; j < KX; ++j) xw[I][j] = f2(0.f, 0.f);   // row H: zero padding
    }
    if (TROW) {
      const float2(&top)[KX] = xw[IM2];
      const float2(&mid)[KX] = xw[IM1];
      const float2(&bot)[KX] = xw[I];
      // T column jt = image column q*K-1+jt ; its 3x3 window starts at x-window column jt
      float2 u[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = fmul2(w9[0], top[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[1], top[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[2], top[j + 2], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[3], mid[j], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[4], mid[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[5], mid[j + 2], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[6], bot[j], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[7], bot[j + 1], u[j]);
#pragma unroll
      for (int j = 0; j < KT; ++j) u[j] = ffma2(w9[8], bot[j + 2], u[j]);
      // dy / o of row s-1 come from the stage fetched one step ago (still held)
      const uint32_t sb = stages_s + (uint32_t)sidx[IM1] * P.stage_bytes + tbase;
      const uint32_t oa = sb + P.x_bytes + P.xo_cols;     // o at T column 0
      const uint32_t ga = sb + P.x_bytes + P.o_bytes;     // dy tile starts at image column -1
      // staging slot of row s-1 receives lam*dS now and is completed one step later
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      const uint32_t dst = obuf + (uint32_t)IM1 * (2 * OROW);
      const float2 res2 = f2(P.res, P.res);
      float2 tt[KT];
#pragma unroll
      for (int j = 0; j < KT; ++j) {
        const float2 gy = lds_pair<T>(ga + j * CS);
        const float2 ov = lds_pair<T>(oa + j * CS);
        float2 ds = ffma2(q1, gy, q0);
        ds = ffma2(q2, u[j], ds);
        ds = ffma2(q3, ov, ds);
        float2 t = fmul2(ta, ds);
        if (RAGGED) {
          if (!col_ok(j + 1)) t = f2(0.f, 0.f);
        } else if (j == 0) {
          t = fmul2(t, tm0);
        } else if (j == KT - 1) {
          t = fmul2(t, tm1);
        }
        tt[j] = t;
        if (j >= 1 && j <= K) {
          const int jo = j - 1;
          sts_raw(dst + OROW + jo * OW, pack_pair<T>(fmul2(lm, ds)));
          // residual / GAP term of dX row s-1
          if (TFIRST) da[IM1][jo] = ffma2(res2, gy, dyc);
          else da[IM1][jo] = fadd2(da[IM1][jo], ffma2(res2, gy, dyc));
          // dWv[i][dj] += T[t][w] * x[t+i-1][w+dj-1]
          dw[0] = ffma2(t, top[j], dw[0]);
          dw[1] = ffma2(t, top[j + 1], dw[1]);
          dw[2] = ffma2(t, top[j + 2], dw[2]);
          dw[3] = ffma2(t, mid[j], dw[3]);
          dw[4] = ffma2(t, mid[j + 1], dw[4]);
          dw[5] = ffma2(t, mid[j + 2], dw[5]);
          dw[6] = ffma2(t, bot[j], dw[6]);
          dw[7] = ffma2(t, bot[j + 1], dw[7]);
          dw[8] = ffma2(t, bot[j + 2], dw[8]);
        }
      }
      if (!HOLD2) release(sidx[IM1], gidx[IM1]);
      // scatter T row t = s-1:  dX[h][w] += wv[i][dj] * T[h-i+1][w-dj+1]
#pragma unroll
      for (int jo = 0; jo < K; ++jo) {
        if (!TFIRST) {
          float2 a = da[IM2][jo];
          a = ffma2(w9[0], tt[jo + 2], a);
          a = ffma2(w9[1], tt[jo + 1], a);
          a = ffma2(w9[2], tt[jo], a);
          da[IM2][jo] = a;
        }
        {
          float2 a = da[IM1][jo];
          a = ffma2(w9[3], tt[jo + 2], a);
          a = ffma2(w9[4], tt[jo + 1], a);
          a = ffma2(w9[5], tt[jo], a);
          da[IM1][jo] = a;
        }
        if (!TLAST) {
          float2 n = fmul2(w9[6], tt[jo + 2]);
          n = ffma2(w9[7], tt[jo + 1], n);
          da[I][jo] = ffma2(w9[8], tt[jo], n);
        }
      }
    }
    if (FIN) {
      // dX row h = s-2 is complete
      const uint32_t dst = obuf + (uint32_t)IM2 * (2 * OROW);
      const float2(&xrow)[KX] = xw[IM2];
      uint32_t ca = 0;
      if (HOLD2) ca = stages_s + (uint32_t)sidx[IM2] * P.stage_bytes + tbase + 2 * CS;   // c3 at own column 0
#pragma unroll
      for (int jo = 0; jo < K; ++jo) {
        float2 tot = da[IM2][jo];
        if (FUSE) {
          const float2 xc = xrow[jo + 2];
          tot = f2(xc.x > 0.f ? tot.x : 0.f, xc.y > 0.f ? tot.y : 0.f);
          const Raw dz = pack_pair<T>(tot);
          sts_raw(dst + jo * OW, dz);
          const Raw ld = lds_raw<T>(dst + OROW + jo * OW);
          sts_raw(dst + OROW + jo * OW, raw_add<T>(ld, dz));   // total identity gradient
          if (XF) {
            const float2 c3 = lds_pair<T>(ca + jo * CS);
            sdz = fadd2(sdz, tot);
            sdzc = ffma2(tot, c3, sdzc);
          }
        } else {
          sts_raw(dst + jo * OW, pack_pair<T>(tot));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t src = dst - (uint32_t)lane * 2 * ES;
        tma_store_4d(tm_dx, src, ychan, ycol, r - 2, b);
        tma_store_4d(tm_do, src + OROW, ychan, ycol, r - 2, b);
        bulk_commit();
      }
      if (HOLD2) release(sidx[IM2], gidx[IM2]);
    }
    ++r;
  }

  template <int I> __device__ __forceinline__ void tail_steps() {
    step<I, false, true, false, true, true>();                    // s = H
    step<(I + 1) % 3, false, false, false, false, true>();        // s = H + 1
  }

  // one image of H >= 3 rows
  __device__ __forceinline__ void image() {
    const int H = P.H;
    r = 0;
#pragma unroll
    for (int j = 0; j < KX; ++j) xw[2][j] = f2(0.f, 0.f);   // row -1
    step<0, true, false, false, false, false>();
    step<1, true, true, true, false, false>();
    step<2, true, true, false, false, true>();
    while (r + 2 < H) {
      step<0, true, true, false, false, true>();
      step<1, true, true, false, false, true>();
      step<2, true, true, false, false, true>();
    }
    if (r < H) {
      step<0, true, true, false, false, true>();
      if (r < H) {
        step<1, true, true, false, false, true>();
        tail_steps<2>();
      } else {
        tail_steps<1>();
      }
    } else {
      tail_steps<0>();
    }
  }
};

template <typename T, int CB, bool XF, bool FUSE, bool RAGGED>
__global__ void __launch_bounds__(256, 1)
k_v7_bwd(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_o,
         const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_dx,
         const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ V7Params P) {
  typedef V7Bwd<T, CB, XF, FUSE, RAGGED> Bk;
  constexpr int NP = CB / 2;
  constexpr int ES = sizeof(T);
  constexpr int K = kV7;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 16;
  unsigned char* stages = smem_raw + kV7Hdr;
  unsigned char* tail = stages + (size_t)P.S * P.stage_bytes;   // staging rows of all warps
  const int cb = blockIdx.x / P.cpc, m = blockIdx.x - cb * P.cpc;
  const int n_my = (m < P.U) ? (P.U - 1 - m) / P.cpc + 1 : 0;
  const uint32_t bar_s = smem_u32(smem_raw), stages_s = smem_u32(stages);
  const int ct = threadIdx.x;
  const int p = ct % NP, q = ct / NP;
  const int warp = ct >> 5;
  Bk S(P);
  S.tm_x = &tm_x;
  S.tm_o = &tm_o;
  S.tm_dy = &tm_dy;
  S.tm_dx = &tm_dx;
  S.tm_do = &tm_do;
  S.bar_s = bar_s;
  S.stages_s = stages_s;
  S.lane = threadIdx.x & 31;
  S.cb = cb;
  S.m = m;
  S.rows_total = n_my * P.TPU * P.H;
  S.grow = 0;
  S.gidx[0] = S.gidx[1] = S.gidx[2] = 0;
  S.cnt_s = bar_s + 128;   // the release counters live where the forward kernels keep their empty barriers
  S.pol = l2_policy_evict_first();
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.S; ++s) {
      mbar_init(&full[s], 1);
      reinterpret_cast<uint32_t*>(empty)[s] = 0u;
    }
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_o);
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_dx);
    tma_prefetch_desc(&tm_do);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int s = 0; s < P.S && s < S.rows_total; ++s) S.issue_row(s, s);   // fill the pipeline
  S.tbase = (uint32_t)(q * K) * Bk::CS + (uint32_t)p * 2 * ES;
  S.obuf = smem_u32(tail) + (uint32_t)warp * (3 * 2 * Bk::OROW) + (uint32_t)S.lane * 2 * ES;
  S.ychan = cb * CB + ((warp * 32) % NP) * 2;
  S.set_tile(0, q);
  S.st = 0;
  S.ph = 0;
  S.sidx[0] = S.sidx[1] = S.sidx[2] = 0;
  const int c = cb * CB + 2 * p;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    S.w9[i] = f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]);
    S.dw[i] = f2(0.f, 0.f);
  }
  S.lm = P.lam ? f2(P.lam[c], P.lam[c + 1]) : f2(0.f, 0.f);
  S.za = S.zb = f2(0.f, 0.f);
  if (XF) {
    S.za = *reinterpret_cast<const float2*>(P.zcoef + c);
    S.zb = *reinterpret_cast<const float2*>(P.zcoef + P.C + c);
  }
  S.sdz = S.sdzc = f2(0.f, 0.f);
  const int64_t BC = (int64_t)P.B * P.C;
  const int NS = P.NT / P.TPU;   // units per image
  for (int u = m; u < P.U; u += P.cpc) {
    const int bi = v7_fdiv(u, NS, P.mNS), tile0 = (u - bi * NS) * P.TPU;
    const int b = P.rev ? P.B - 1 - bi : bi;
    S.b = b;
    const float* cp = P.bcoef + (int64_t)b * P.C + c;
    S.q0 = *reinterpret_cast<const float2*>(cp);
    S.q1 = *reinterpret_cast<const float2*>(cp + BC);
    S.q2 = *reinterpret_cast<const float2*>(cp + 2 * BC);
    S.q3 = *reinterpret_cast<const float2*>(cp + 3 * BC);
    S.ta = *reinterpret_cast<const float2*>(cp + 4 * BC);
    S.dyc = *reinterpret_cast<const float2*>(cp + 5 * BC);
    for (int tile = tile0; tile < tile0 + P.TPU; ++tile) {
      if (P.NT > 1) S.set_tile(tile * P.NQ * K, q);
      S.image();
    }
  }
  if (S.lane == 0) bulk_wait_all<0>();
  // per-CTA partials of dWv (and the bn3 sums): reduce over the NQ column groups; the scratch aliases the pipeline
  // stages, which every warp has finished reading once it arrives at the barrier
  __syncthreads();
  constexpr int NR = 11;
  float2* red = reinterpret_cast<float2*>(stages);   // [NQ][NR][NP]
#pragma unroll
  for (int i = 0; i < 9; ++i) red[((size_t)q * NR + i) * NP + p] = S.dw[i];
  red[((size_t)q * NR + 9) * NP + p] = S.sdz;
  red[((size_t)q * NR + 10) * NP + p] = S.sdzc;
  __syncthreads();
  for (int idx = ct; idx < NR * NP; idx += blockDim.x) {
    const int k = idx / NP, pp = idx - k * NP;
    float2 s = f2(0.f, 0.f);
    for (int qq = 0; qq < P.NQ; ++qq) {
      const float2 v = red[((size_t)qq * NR + k) * NP + pp];
      s.x += v.x;
      s.y += v.y;
    }
    const int cc = cb * CB + 2 * pp;
    if (k < 9) {
      float* dst = P.wv_part + ((int64_t)m * P.C + cc) * 9 + k;
      dst[0] = s.x;
      dst[9] = s.y;
    } else if (XF && FUSE && P.dz_part != nullptr) {
      *reinterpret_cast<float2*>(P.dz_part + ((int64_t)m * 2 + (k - 9)) * P.C + cc) = s;
    }
  }
}

// partial [nparts, 2, C] -> sums [2, C] (deterministic, fp64 accumulate)
static __global__ void k_v7_dz_finish(const float* __restrict__ part, int nparts, int C, float* __restrict__ sums) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * C) return;
  double a = 0.0;
  for (int p = 0; p < nparts; ++p) a += (double)part[(int64_t)p * 2 * C + idx];
  sums[idx] = (float)a;
}

}  // namespace mrla
