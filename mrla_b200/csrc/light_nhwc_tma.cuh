// v2 NHWC sweeps of the MRLA-light tail: TMA-fed shared-memory row ring + packed fp32x2 math.
//
// One CTA = 1 producer warp + NQ*(CB/2) consumer threads, persistent over work items (b, channel block).
//   producer : one elected lane streams the image in groups of G rows with cp.async.bulk.tensor (TMA):
//              x box = [CB ch, W+2 cols (from w=-1), G rows]  — out-of-bounds columns / rows / channels are
//              zero-filled by the TMA unit, which IS the conv zero padding;  o / dy box = [CB, W, G].
//              S stages, full/empty mbarriers, up to ~190 KB in flight per SM.
//   consumers: thread = (channel pair p, column group q of 4 columns); lanes of a warp are 32 consecutive
//              channel pairs (128 B of one pixel -> conflict-free LDS.32).  Each thread marches down the rows
//              with a 3-row x 6-column register window of fp32x2 pairs; every x row is read from shared
//              memory exactly once per thread and the depthwise 3x3 is 9 FFMA2 per output pair.
// Replaces (paths relative to /root/reference) mrla_light_module.py:56-72, resnet_mrla_light.py:42,116.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace mrla {

constexpr int kCols = 4;          // output columns per consumer thread
constexpr int kWin = kCols + 2;   // window columns (one halo column each side)

struct TmaSweepParams {
  int B, C, H, W;
  int G, S, NQ, ncb, items;
  int cons_threads;
  int rev;              // 1: walk the batch from the last sample down (the previous sweep left that end in L2)
  uint32_t x_bytes, o_bytes, stage_bytes;   // per stage (x tile, o/dy tile, total)
  const float* wv;      // [C,9]
  float* mom;           // MODE 0: [6,B,C] ; MODE 2: gmom [3,B,C]
  const float* coef;    // MODE 1: [3,B,C]
  void* y;              // MODE 1 / 3 / 4 output
  int64_t bs_y;
  float res;
  float* wv_part;       // MODE 4: [grid/ncb, C, 9] per-CTA dWv partials
  const float* zcoef;   // MODE 6: [2,C] affine applied to z in front of the fold (the bottleneck's bn3 apply)
};

// ---------------------------------------------------------------------------- packed helpers
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }

template <typename T> __device__ __forceinline__ float2 lds_pair(uint32_t saddr);
template <> __device__ __forceinline__ float2 lds_pair<__nv_bfloat16>(uint32_t saddr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 lds_pair<__half>(uint32_t saddr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <> __device__ __forceinline__ float2 lds_pair<float>(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
template <typename T> __device__ __forceinline__ void stg_pair(T* p, float2 v);
template <> __device__ __forceinline__ void stg_pair<__nv_bfloat16>(__nv_bfloat16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}
template <> __device__ __forceinline__ void stg_pair<__half>(__half* p, float2 v) {
  *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y);
}
template <> __device__ __forceinline__ void stg_pair<float>(float* p, float2 v) {
  *reinterpret_cast<float2*>(p) = v;
}

template <typename T> __device__ __forceinline__ float2 ldg_pair(const T* p);
template <> __device__ __forceinline__ float2 ldg_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const volatile uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 ldg_pair<__half>(const __half* p) {
  uint32_t u = *reinterpret_cast<const volatile uint32_t*>(p);
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <> __device__ __forceinline__ float2 ldg_pair<float>(const float* p) {
  const volatile float* q = p;
  return make_float2(q[0], q[1]);
}

// an output pair in the precision it will be stored with (one 32-bit register for bf16 / fp16)
template <typename T> struct RawPair { typedef uint32_t type; };
template <> struct RawPair<float> { typedef float2 type; };
template <typename T> __device__ __forceinline__ typename RawPair<T>::type pack_pair(float2 v);
template <> __device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_pair<__half>(float2 v) {
  const __half2 h = __floats2half2_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ float2 pack_pair<float>(float2 v) { return v; }
template <typename T> __device__ __forceinline__ float2 unpack_pair(typename RawPair<T>::type r);
template <> __device__ __forceinline__ float2 unpack_pair<__nv_bfloat16>(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack_pair<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <> __device__ __forceinline__ float2 unpack_pair<float>(float2 v) { return v; }

template <typename T> __device__ __forceinline__ typename RawPair<T>::type lds_raw(uint32_t saddr);
template <> __device__ __forceinline__ uint32_t lds_raw<__nv_bfloat16>(uint32_t saddr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
  return u;
}
template <> __device__ __forceinline__ uint32_t lds_raw<__half>(uint32_t saddr) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
  return u;
}
template <> __device__ __forceinline__ float2 lds_raw<float>(uint32_t saddr) { return lds_pair<float>(saddr); }

template <typename T> __device__ __forceinline__ typename RawPair<T>::type raw_add(typename RawPair<T>::type a,
                                                                                  typename RawPair<T>::type b);
template <> __device__ __forceinline__ uint32_t raw_add<__nv_bfloat16>(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t raw_add<__half>(uint32_t a, uint32_t b) {
  const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ float2 raw_add<float>(float2 a, float2 b) { return fadd2(a, b); }

template <typename T> __device__ __forceinline__ void stg_raw(T* p, typename RawPair<T>::type r) {
  *reinterpret_cast<typename RawPair<T>::type*>(p) = r;
}
// relu on a pair in storage precision
template <typename T> __device__ __forceinline__ typename RawPair<T>::type raw_relu(typename RawPair<T>::type a);
template <> __device__ __forceinline__ uint32_t raw_relu<__nv_bfloat16>(uint32_t a) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), z);
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t raw_relu<__half>(uint32_t a) {
  const __half2 z = __floats2half2_rn(0.f, 0.f);
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), z);
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ float2 raw_relu<float>(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }

template <int ACT> __device__ __forceinline__ float2 act2(float2 u) {
  if (ACT == 1) return make_float2(act_fwd<1>(u.x), act_fwd<1>(u.y));
  return u;
}

// 3x3 depthwise taps on one output column `j` of the window (w9[i*3+dj] are channel-pair weights)
__device__ __forceinline__ float2 conv9(const float2 (&top)[kWin], const float2 (&mid)[kWin], const float2 (&bot)[kWin],
                                        const float2 (&w9)[9], int j) {
  float2 s = fmul2(w9[0], top[j]);
  s = ffma2(w9[1], top[j + 1], s);
  s = ffma2(w9[2], top[j + 2], s);
  s = ffma2(w9[3], mid[j], s);
  s = ffma2(w9[4], mid[j + 1], s);
  s = ffma2(w9[5], mid[j + 2], s);
  s = ffma2(w9[6], bot[j], s);
  s = ffma2(w9[7], bot[j + 1], s);
  s = ffma2(w9[8], bot[j + 2], s);
  return s;
}

// all kCols outputs of one row, tap-major so the kCols FFMA2 chains are independent and interleaved
__device__ __forceinline__ void conv9x4(const float2 (&top)[kWin], const float2 (&mid)[kWin], const float2 (&bot)[kWin],
                                        const float2 (&w9)[9], float2 (&u)[kCols]) {
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = fmul2(w9[0], top[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[1], top[j + 1], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[2], top[j + 2], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[3], mid[j], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[4], mid[j + 1], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[5], mid[j + 2], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[6], bot[j], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[7], bot[j + 1], u[j]);
#pragma unroll
  for (int j = 0; j < kCols; ++j) u[j] = ffma2(w9[8], bot[j + 2], u[j]);
}

// ---------------------------------------------------------------------------- the kernel
// MODE 0: forward moments (Σx ΣV ΣV² [ΣVo Σo Σo²]) ; MODE 1: forward apply (y) ; MODE 2: backward moments (Σdy ΣdyV [Σdyo])
// MODE 3: MRLA-base F0 — y = dwconv3x3(x) stored into the V-cache slot, plus Σx (no o tile)
// MODE 4: MRLA-base B4 — window tile = dV_t: dX and dWv ; MODE 5: MODE 0 with x = relu(z + o) formed (and stored) here
// MODE 6: MODE 5 with z' = round(a_c z + b_c) first — z is the raw conv3 output and (a, b) the bn3 coefficients, so the
//         bottleneck's bn3 apply pass (resnet_mrla_light.py:101-102) disappears as well
// BIG: one 480-thread CTA per SM (W = 56 / 28); !BIG: <= 288 threads, two CTAs per SM (small images)
template <typename T, int CB, int ACT, bool HAS_O, int MODE, bool BIG>
__global__ void __launch_bounds__(BIG ? 480 : 288, BIG ? 1 : 2)
k_light_nhwc_tma(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_o,
                 const __grid_constant__ CUtensorMap tm_dy, TmaSweepParams P) {
  constexpr int NP = CB / 2;                                  // channel pairs per block
  constexpr bool MOM = (MODE == 0 || MODE == 5 || MODE == 6);   // forward moments
  constexpr int NACC = MOM ? (HAS_O ? 6 : 3) : (MODE == 2 ? (HAS_O ? 3 : 2) : (MODE == 3 ? 1 : 0));
  constexpr bool XFOLD = (MODE == 5 || MODE == 6);   // x = relu(z + o) is formed here: the o tile carries halo columns too
  constexpr bool ZAFF = (MODE == 6);
  constexpr bool HAS_DY = (MODE == 2 || MODE == 4);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 16;
  unsigned char* stages = smem_raw + 256;
  float2* red_base = reinterpret_cast<float2*>(stages + (size_t)P.S * P.stage_bytes);  // [2][NQ][NACC][NP]
  const int ncw = P.cons_threads / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int ngroups = (P.H + P.G - 1) / P.G;

  if (threadIdx.x < 32) {
    // ================================ producer warp ================================
    if (threadIdx.x == 0) {
      tma_prefetch_desc(&tm_x);
      if (HAS_O) tma_prefetch_desc(&tm_o);
      if (HAS_DY) tma_prefetch_desc(&tm_dy);
      int st = 0;
      uint32_t ph = 1;  // first pass over the ring: slots are free
      for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
        const int b0 = item / P.ncb, cb = item - b0 * P.ncb;
        const int b = P.rev ? P.B - 1 - b0 : b0;
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(&empty[st], ph);
          unsigned char* sx = stages + (size_t)st * P.stage_bytes;
          mbar_arrive_expect_tx(&full[st], P.stage_bytes);
          tma_load_4d(sx, &tm_x, &full[st], cb * CB, -1, g * P.G, b);
          if (HAS_O) tma_load_4d(sx + P.x_bytes, &tm_o, &full[st], cb * CB, XFOLD ? -1 : 0, g * P.G, b);
          if (HAS_DY) tma_load_4d(sx + P.x_bytes + (HAS_O ? P.o_bytes : 0), &tm_dy, &full[st], cb * CB, 0, g * P.G, b);
          if (++st == P.S) { st = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // ================================== consumers ==================================
  const int ct = threadIdx.x - 32;
  const int p = ct % NP, q = ct / NP;
  const int lane = threadIdx.x & 31;
  constexpr int ES = sizeof(T);
  const uint32_t stages_s = smem_u32(stages);
  // Tile rows are 4*NQ (+2 halo) columns wide: columns past the image are zero-filled by TMA, so every
  // thread reads its window at compile-time offsets from one per-thread base (no clamping, no o/dy masks).
  constexpr uint32_t CS = CB * ES;                      // bytes between adjacent columns
  const uint32_t xrow_bytes = (uint32_t)(P.NQ * kCols + 2) * CS;
  const uint32_t orow_bytes = XFOLD ? xrow_bytes : (uint32_t)(P.NQ * kCols) * CS;
  constexpr uint32_t OC = XFOLD ? CS : 0;   // byte offset of a thread's first own column inside its o row
  const uint32_t tbase = (uint32_t)(q * kCols) * CS + (uint32_t)p * 2 * ES;
  const bool ragged = (P.W % kCols) != 0;               // last column group is partial
  bool cvalid[kCols];
#pragma unroll
  for (int j = 0; j < kCols; ++j) cvalid[j] = (q * kCols + j) < P.W;

  // ring cursor: shared-memory address of the next x row / its o(dy) row, advanced incrementally so the row
  // loop carries no multiplies or divisions
  int st_cur = 0;
  uint32_t ph_cur = 0;
  int rr = 0;
  uint32_t xa = stages_s + tbase;                 // address of x row r (window column 0 of this thread)
  uint32_t oa = stages_s + P.x_bytes + tbase;     // address of o row r
  int cur_cb = -1;   // the grid is a multiple of ncb, so a CTA keeps its channel block: weights are loaded once
  float2 w9[9];
  float2 za = f2(0.f, 0.f), zb = f2(0.f, 0.f);   // MODE 6: bn3 coefficients of this thread's channel pair
  // MODE 6: window columns outside the image must stay exactly 0 (conv zero padding): the affine would turn the
  // TMA zero fill into b_c
  bool wvalid[kWin];
#pragma unroll
  for (int j = 0; j < kWin; ++j) wvalid[j] = (q * kCols - 1 + j) >= 0 && (q * kCols - 1 + j) < P.W;
  float2 dwf[MODE == 4 ? 9 : 1];   // MODE 4: dWv accumulators (flipped tap order), kept across all items of the CTA
#pragma unroll
  for (int i = 0; i < (MODE == 4 ? 9 : 1); ++i) dwf[i] = f2(0.f, 0.f);
  int red_sel = 0;   // double-buffered reduction scratch: one consumer barrier per image

  for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
    const int b0 = item / P.ncb, cb = item - b0 * P.ncb;
    const int b = P.rev ? P.B - 1 - b0 : b0;
    const int c = cb * CB + 2 * p;
    const bool chan_ok = c < P.C;
    if (cb != cur_cb) {
      cur_cb = cb;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int wi = (MODE == 4) ? 8 - i : i;   // MODE 4 correlates with the flipped kernel (transposed conv)
        w9[i] = chan_ok ? f2(P.wv[(int64_t)c * 9 + wi], P.wv[(int64_t)(c + 1) * 9 + wi]) : f2(0.f, 0.f);
      }
      if (ZAFF && chan_ok) {
        za = *reinterpret_cast<const float2*>(P.zcoef + c);
        zb = *reinterpret_cast<const float2*>(P.zcoef + P.C + c);
      }
    }
    float2 cA = f2(0.f, 0.f), cL = f2(0.f, 0.f), cD = f2(0.f, 0.f);
    if (MODE == 1 && chan_ok) {
      const int64_t BC = (int64_t)P.B * P.C;
      const float* cp = P.coef + (int64_t)b * P.C + c;
      cA = *reinterpret_cast<const float2*>(cp);
      if (HAS_O) cL = *reinterpret_cast<const float2*>(cp + BC);
      cD = *reinterpret_cast<const float2*>(cp + 2 * BC);
    }
    if (MODE == 4 && chan_ok) cD = *reinterpret_cast<const float2*>(P.coef + (int64_t)b * P.C + c);  // GAP grad / HW
    const float2 res2 = f2(P.res, P.res);
    float2 acc[NACC > 0 ? NACC : 1], accb[NACC > 0 ? NACC : 1];  // even / odd columns: two independent chains
#pragma unroll
    for (int i = 0; i < (NACC > 0 ? NACC : 1); ++i) { acc[i] = f2(0.f, 0.f); accb[i] = f2(0.f, 0.f); }
    // running pointer to this thread's first output column of the row being produced
    // MODE 5 writes x (row r, at fetch time); the others write the output row r-1
    T* yrow = (MODE == 1 || MODE == 3 || MODE == 4 || XFOLD) ? static_cast<T*>(P.y) + (int64_t)b * P.bs_y + (int64_t)(q * kCols) * P.C + c : nullptr;
    const int64_t y_row_stride = (int64_t)P.W * P.C;
    bool sv[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) sv[j] = chan_ok && cvalid[j];

    float2 win[3][kWin];
#pragma unroll
    for (int j = 0; j < kWin; ++j) win[2][j] = f2(0.f, 0.f);  // row -1 lives in slot 2

    uint32_t o_prev = 0;      // address of the o(dy) row of the previous x row
    int rel_stage = -1;       // stage to release once the previous row has been consumed (-1: none)
    // r = x row being fetched into the window; output row r-1 is produced in the same step
    for (int r0 = 0; r0 <= P.H; r0 += 3) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int r = r0 + i;
        if (r <= P.H) {
          uint32_t o_this = 0;
          int rel_next = -1;
          // ---- fetch x row r into window slot i ----
          if (r < P.H) {
            if (rr == 0) mbar_wait(&full[st_cur], ph_cur);
            if (XFOLD) {
              // x = relu(z + identity) on the whole window (halo columns are 0 + 0), formed in storage precision like
              // the reference's in-place `out += identity; relu_` (resnet_mrla_light.py:113-114) — the conv below then
              // sees exactly the x that is stored — and written out for the own columns
#pragma unroll
              for (int j = 0; j < kWin; ++j) {
                typename RawPair<T>::type zr = lds_raw<T>(xa + j * CS);
                if (ZAFF) zr = pack_pair<T>(ffma2(za, unpack_pair<T>(zr), zb));   // bn3 output in storage precision
                typename RawPair<T>::type xr = raw_relu<T>(raw_add<T>(zr, lds_raw<T>(oa + j * CS)));
                if (ZAFF && !wvalid[j]) xr = pack_pair<T>(f2(0.f, 0.f));
                win[i][j] = unpack_pair<T>(xr);
                if (j >= 1 && j <= kCols)
                  if (sv[j - 1]) stg_raw<T>(yrow + (j - 1) * P.C, xr);
              }
              yrow += y_row_stride;
            } else {
#pragma unroll
              for (int j = 0; j < kWin; ++j) win[i][j] = lds_pair<T>(xa + j * CS);
            }
            o_this = oa;
            xa += xrow_bytes;
            oa += orow_bytes;
            if (++rr == P.G || r == P.H - 1) {   // last row of its group: move the cursor to the next stage
              rel_next = st_cur;
              rr = 0;
              if (++st_cur == P.S) { st_cur = 0; ph_cur ^= 1; }
              xa = stages_s + (uint32_t)st_cur * P.stage_bytes + tbase;
              oa = xa + P.x_bytes;
            }
          } else {
#pragma unroll
            for (int j = 0; j < kWin; ++j) win[i][j] = f2(0.f, 0.f);
          }
          // ---- produce output row r-1 ----
          if (r >= 1) {
            const uint32_t ob = o_prev;
            const float2(&top)[kWin] = win[(i + 1) % 3];
            const float2(&mid)[kWin] = win[(i + 2) % 3];
            const float2(&bot)[kWin] = win[i];
            float2 u4[kCols];
            conv9x4(top, mid, bot, w9, u4);
#pragma unroll
            for (int j = 0; j < kCols; ++j) {
              float2 v = act2<ACT>(u4[j]);
              float2 ov = f2(0.f, 0.f), gv = f2(0.f, 0.f);
              if (HAS_O) ov = lds_pair<T>(ob + OC + j * CS);
              if (HAS_DY) gv = lds_pair<T>(ob + (HAS_O ? P.o_bytes : 0) + j * CS);
              const float2 xc = mid[j + 1];
              float2(&A)[NACC > 0 ? NACC : 1] = (j & 1) ? accb : acc;
              if (MOM) {
                if (ragged && !cvalid[j]) v = f2(0.f, 0.f);
                A[0] = fadd2(A[0], xc);
                A[1] = fadd2(A[1], v);
                A[2] = ffma2(v, v, A[2]);
                if (HAS_O) {
                  A[3] = ffma2(v, ov, A[3]);
                  A[4] = fadd2(A[4], ov);
                  A[5] = ffma2(ov, ov, A[5]);
                }
              } else if (MODE == 2) {
                A[0] = fadd2(A[0], gv);
                A[1] = ffma2(gv, v, A[1]);
                if (HAS_O) A[2] = ffma2(gv, ov, A[2]);
              } else if (MODE == 3) {
                A[0] = fadd2(A[0], xc);
                if (sv[j]) stg_pair<T>(yrow + j * P.C, v);
              } else if (MODE == 4) {
                // window = dV_t, ov = x centre, gv = dy:  dX = res*dy + dyc + conv(dV_t, flipped wv) ;
                // dW'[k] += x * window_k  (k over the flipped tap order; un-flipped when flushed)
                if (sv[j]) stg_pair<T>(yrow + j * P.C, fadd2(ffma2(res2, gv, cD), v));
                dwf[0] = ffma2(ov, top[j], dwf[0]);
                dwf[1] = ffma2(ov, top[j + 1], dwf[1]);
                dwf[2] = ffma2(ov, top[j + 2], dwf[2]);
                dwf[3] = ffma2(ov, mid[j], dwf[3]);
                dwf[4] = ffma2(ov, mid[j + 1], dwf[4]);
                dwf[5] = ffma2(ov, mid[j + 2], dwf[5]);
                dwf[6] = ffma2(ov, bot[j], dwf[6]);
                dwf[7] = ffma2(ov, bot[j + 1], dwf[7]);
                dwf[8] = ffma2(ov, bot[j + 2], dwf[8]);
              } else {
                float2 t = ffma2(cA, v, cD);
                if (HAS_O) t = ffma2(cL, ov, t);
                t = ffma2(res2, xc, t);
                if (sv[j]) stg_pair<T>(yrow + j * P.C, t);
              }
            }
            if (MODE == 1 || MODE == 3 || MODE == 4) yrow += y_row_stride;
            // release the group whose last row was just consumed
            if (rel_stage >= 0) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty[rel_stage]);
            }
          }
          o_prev = o_this;
          rel_stage = rel_next;
        }
      }
    }

    if (NACC > 0) {
      // deterministic reduction over the NQ column groups of every channel pair.  The scratch is double
      // buffered: a thread can only reach the write of image i+2 after the barrier of image i+1, which every
      // thread passes after finishing its reads of image i.
      float2* red = red_base + (size_t)red_sel * P.NQ * NACC * NP;
      red_sel ^= 1;
#pragma unroll
      for (int i = 0; i < NACC; ++i) red[((size_t)q * NACC + i) * NP + p] = fadd2(acc[i], accb[i]);
      named_bar_sync(1, P.cons_threads);
      const int64_t BC = (int64_t)P.B * P.C;
      for (int idx = ct; idx < NACC * NP; idx += P.cons_threads) {
        const int mi = idx / NP, pp = idx - mi * NP;
        float2 s = f2(0.f, 0.f);
        for (int qq = 0; qq < P.NQ; ++qq) {
          const float2 v = red[((size_t)qq * NACC + mi) * NP + pp];
          s.x += v.x;
          s.y += v.y;
        }
        const int cc = cb * CB + 2 * pp;
        if (cc < P.C) *reinterpret_cast<float2*>(P.mom + (int64_t)mi * BC + (int64_t)b * P.C + cc) = s;
      }
    }
  }
  if (MODE == 4) {
    // one partial per CTA: CTAs with blockIdx = cb (mod ncb) share a channel block -> slot = blockIdx / ncb
    float2* red = red_base;
#pragma unroll
    for (int i = 0; i < 9; ++i) red[((size_t)q * 9 + i) * NP + p] = dwf[i];
    named_bar_sync(1, P.cons_threads);
    const int slot = blockIdx.x / P.ncb;
    for (int idx = ct; idx < 9 * NP; idx += P.cons_threads) {
      const int kf = idx / NP, pp = idx - kf * NP;
      float2 s = f2(0.f, 0.f);
      for (int qq = 0; qq < P.NQ; ++qq) {
        const float2 v = red[((size_t)qq * 9 + kf) * NP + pp];
        s.x += v.x;
        s.y += v.y;
      }
      const int cc = cur_cb * CB + 2 * pp;
      if (cur_cb >= 0 && cc < P.C) {
        float* dst = P.wv_part + ((int64_t)slot * P.C + cc) * 9 + (8 - kf);
        dst[0] = s.x;
        dst[9] = s.y;
      }
    }
  }
}


// =====================================================================================================
// v2 sweep B (backward apply), NHWC:   dS = Q0 + Q1*dy + Q2*V + Q3*o ;  T = Ta*dS*act'(U)
//     do = lam*dS ;  dx = res*dy + dwconv3x3^T(T) + dyc ;  dWv[c,i,j] = Σ T[h,w]*x[h+i-1,w+j-1]
// Work item = (channel block, sample, column tile of WT<=28 columns).  A consumer thread owns 4 output
// columns of one channel pair and recomputes T on one halo column each side (x window = 3 rows x 8 cols),
// so the transposed conv needs no exchange between threads; dX rows are built in scatter form with three
// rotating row accumulators.  CTAs take CONTIGUOUS item ranges ordered (cb, b, tile) so the dWv partial
// sums stay in registers across samples and are flushed once per channel block into
// wv_part[slot][C][9] (slot = CTA index relative to the first CTA touching that channel block).
struct TmaBwdParams {
  int B, C, H, W;
  int G, S, NQ, ncb, NT, WT, items, ipc, cons_threads, maxslots;
  int cpc;             // ring kernel: CTAs per channel block (grid = ncb * cpc)
  uint32_t x_bytes, t_bytes, stage_bytes;
  const float* wv;
  const float* lam;
  const float* bcoef;  // [7,B,C]
  void* dx;
  void* dout;
  int64_t bs_dx, bs_do;
  float res;
  float* wv_part;      // [maxslots, C, 9]
};

// BIG: <= 256 threads, one CTA per SM; !BIG: <= 160 threads, two CTAs per SM (small images)
template <typename T, int CB, int ACT, bool BIG>
__global__ void __launch_bounds__(BIG ? 256 : 160, BIG ? 1 : 2)
k_light_nhwc_tma_bwd(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                     const __grid_constant__ CUtensorMap tm_o, TmaBwdParams P) {
  constexpr int NP = CB / 2;
  constexpr int XW = kCols + 4;  // x window columns
  constexpr int TW = kCols + 2;  // T columns
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 16;
  unsigned char* stages = smem_raw + 256;
  float2* red = reinterpret_cast<float2*>(stages + (size_t)P.S * P.stage_bytes);  // [NQ][9][NP]
  const int ncw = P.cons_threads / 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int ngroups = (P.H + P.G - 1) / P.G;
  const int item_beg = blockIdx.x * P.ipc;
  const int item_end = min(item_beg + P.ipc, P.items);
  const int ipcb = P.B * P.NT;  // items per channel block

  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      tma_prefetch_desc(&tm_x);
      tma_prefetch_desc(&tm_dy);
      tma_prefetch_desc(&tm_o);
      int st = 0;
      uint32_t ph = 1;
      for (int item = item_beg; item < item_end; ++item) {
        const int cb = item / ipcb;
        const int rem = item - cb * ipcb;
        const int b = rem / P.NT, tile = rem - b * P.NT;
        const int w0 = tile * P.WT;
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(&empty[st], ph);
          unsigned char* sx = stages + (size_t)st * P.stage_bytes;
          mbar_arrive_expect_tx(&full[st], P.stage_bytes);
          tma_load_4d(sx, &tm_x, &full[st], cb * CB, w0 - 2, g * P.G, b);
          tma_load_4d(sx + P.x_bytes, &tm_dy, &full[st], cb * CB, w0 - 1, g * P.G, b);
          tma_load_4d(sx + P.x_bytes + P.t_bytes, &tm_o, &full[st], cb * CB, w0 - 1, g * P.G, b);
          if (++st == P.S) { st = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  const int ct = threadIdx.x - 32;
  const int p = ct % NP, q = ct / NP;
  const int lane = threadIdx.x & 31;
  constexpr int ES = sizeof(T);
  const uint32_t stages_s = smem_u32(stages);
  // tile rows are 4*NQ+4 (x) / 4*NQ+2 (dy, o) columns wide; columns outside the image are TMA zero fill
  constexpr uint32_t CS = CB * ES;
  const uint32_t xrow_bytes = (uint32_t)(P.NQ * kCols + 4) * CS;
  const uint32_t trow_bytes = (uint32_t)(P.NQ * kCols + 2) * CS;
  const uint32_t tbase = (uint32_t)(q * kCols) * CS + (uint32_t)p * 2 * ES;

  int st_cur = 0, st_prev = 0;
  uint32_t ph_cur = 0;
  int cur_cb = -1;
  float2 dw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) dw[i] = f2(0.f, 0.f);
  float2 w9[9];
  float2 lm = f2(0.f, 0.f);
  const int64_t BC = (int64_t)P.B * P.C;

  auto flush_dw = [&](int cb) {
    // reduce dw over the NQ column groups, write this CTA's partial for channel block `cb`
#pragma unroll
    for (int i = 0; i < 9; ++i) red[((size_t)q * 9 + i) * NP + p] = dw[i];
    named_bar_sync(1, P.cons_threads);
    const int first_cta = (cb * ipcb) / P.ipc;
    const int slot = blockIdx.x - first_cta;
    for (int idx = ct; idx < 9 * NP; idx += P.cons_threads) {
      const int tap = idx / NP, pp = idx - tap * NP;
      float2 s = f2(0.f, 0.f);
      for (int qq = 0; qq < P.NQ; ++qq) {
        const float2 v = red[((size_t)qq * 9 + tap) * NP + pp];
        s.x += v.x;
        s.y += v.y;
      }
      const int cc = cb * CB + 2 * pp;
      if (cc < P.C) {
        float* dst = P.wv_part + ((int64_t)slot * P.C + cc) * 9 + tap;
        dst[0] = s.x;    // each (slot, channel block) is written by exactly one CTA, exactly once;
        dst[9] = s.y;    // slots no CTA maps to stay at their memset zero
      }
    }
    named_bar_sync(1, P.cons_threads);
#pragma unroll
    for (int i = 0; i < 9; ++i) dw[i] = f2(0.f, 0.f);
  };

  for (int item = item_beg; item < item_end; ++item) {
    const int cb = item / ipcb;
    const int rem = item - cb * ipcb;
    const int b = rem / P.NT, tile = rem - b * P.NT;
    const int w0 = tile * P.WT;
    const int wt = min(P.WT, P.W - w0);
    const int c = cb * CB + 2 * p;
    const bool chan_ok = c < P.C;
    if (cb != cur_cb) {
      if (cur_cb >= 0) flush_dw(cur_cb);
      cur_cb = cb;
#pragma unroll
      for (int i = 0; i < 9; ++i)
        w9[i] = chan_ok ? f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]) : f2(0.f, 0.f);
      lm = (chan_ok && P.lam) ? f2(P.lam[c], P.lam[c + 1]) : f2(0.f, 0.f);
    }
    float2 q0 = f2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0, ta = q0, dyc = q0;
    if (chan_ok) {
      const float* cp = P.bcoef + (int64_t)b * P.C + c;
      q0 = *reinterpret_cast<const float2*>(cp);
      q1 = *reinterpret_cast<const float2*>(cp + BC);
      q2 = *reinterpret_cast<const float2*>(cp + 2 * BC);
      q3 = *reinterpret_cast<const float2*>(cp + 3 * BC);
      ta = *reinterpret_cast<const float2*>(cp + 4 * BC);
      dyc = *reinterpret_cast<const float2*>(cp + 5 * BC);
    }
    const float2 res2 = f2(P.res, P.res);
    bool tvalid[TW], ovalid[kCols];
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      const int col = w0 + q * kCols - 1 + j;
      tvalid[j] = col >= 0 && col < P.W;
    }
#pragma unroll
    for (int j = 0; j < kCols; ++j) ovalid[j] = chan_ok && (q * kCols + j < wt);
    const int64_t row_stride = (int64_t)P.W * P.C;
    // running row pointers: dxp -> row t-1 (the dX row emitted in a step), dop -> row t
    T* dxp = static_cast<T*>(P.dx) + (int64_t)b * P.bs_dx + (int64_t)(w0 + q * kCols) * P.C + c - row_stride;
    T* dop = static_cast<T*>(P.dout) + (int64_t)b * P.bs_do + (int64_t)(w0 + q * kCols) * P.C + c;

    float2 xw[3][XW];
    float2 da[3][kCols];   // rotating dX row accumulators: row t lives in slot t % 3
    float2 dyprev[kCols];
#pragma unroll
    for (int k = 0; k < XW; ++k) xw[2][k] = f2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kCols; ++j) {
      da[0][j] = f2(0.f, 0.f); da[1][j] = f2(0.f, 0.f); da[2][j] = f2(0.f, 0.f);
      dyprev[j] = f2(0.f, 0.f);
    }
    int rr = 0;
    for (int r0 = 0; r0 <= P.H; r0 += 3) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int r = r0 + i;
        if (r <= P.H) {
          if (r < P.H) {
            if (rr == 0) mbar_wait(&full[st_cur], ph_cur);
            const uint32_t rowb = stages_s + (uint32_t)st_cur * P.stage_bytes + (uint32_t)rr * xrow_bytes + tbase;
#pragma unroll
            for (int k = 0; k < XW; ++k) xw[i][k] = lds_pair<T>(rowb + k * CS);
          } else {
#pragma unroll
            for (int k = 0; k < XW; ++k) xw[i][k] = f2(0.f, 0.f);
          }
          if (r >= 1) {
            const int t = r - 1;                                     // T row produced in this step
            const int rro = (r == P.H) ? ((P.H - 1) % P.G) : ((rr == 0) ? (P.G - 1) : rr - 1);
            const int sto = (r == P.H || rr == 0) ? st_prev : st_cur;
            const uint32_t tb = stages_s + (uint32_t)sto * P.stage_bytes + P.x_bytes + (uint32_t)rro * trow_bytes + tbase;
            const float2(&top)[XW] = xw[(i + 1) % 3];
            const float2(&mid)[XW] = xw[(i + 2) % 3];
            const float2(&bot)[XW] = xw[i];
            float2 tn[TW];
            float2 dyc_row[kCols];
#pragma unroll
            for (int j = 0; j < TW; ++j) {
              float2 u = fmul2(w9[0], top[j]);
              u = ffma2(w9[1], top[j + 1], u);
              u = ffma2(w9[2], top[j + 2], u);
              u = ffma2(w9[3], mid[j], u);
              u = ffma2(w9[4], mid[j + 1], u);
              u = ffma2(w9[5], mid[j + 2], u);
              u = ffma2(w9[6], bot[j], u);
              u = ffma2(w9[7], bot[j + 1], u);
              u = ffma2(w9[8], bot[j + 2], u);
              const float2 v = act2<ACT>(u);
              const float2 gy = lds_pair<T>(tb + j * CS);
              const float2 ov = lds_pair<T>(tb + P.t_bytes + j * CS);
              float2 ds = ffma2(q1, gy, q0);
              ds = ffma2(q2, v, ds);
              ds = ffma2(q3, ov, ds);
              float2 tt = fmul2(ta, ds);
              if (ACT == 1) tt = fmul2(tt, f2(act_grad<1>(u.x), act_grad<1>(u.y)));
              tn[j] = tvalid[j] ? tt : f2(0.f, 0.f);
              if (j >= 1 && j <= kCols) {
                const int jo = j - 1;
                dyc_row[jo] = gy;
                if (ovalid[jo]) stg_pair<T>(dop + jo * P.C, fmul2(lm, ds));
                // dWv[i][dj] += T[t][w] * x[t+i-1][w+dj-1]  (x window index = j + dj)
                dw[0] = ffma2(tn[j], top[j], dw[0]);
                dw[1] = ffma2(tn[j], top[j + 1], dw[1]);
                dw[2] = ffma2(tn[j], top[j + 2], dw[2]);
                dw[3] = ffma2(tn[j], mid[j], dw[3]);
                dw[4] = ffma2(tn[j], mid[j + 1], dw[4]);
                dw[5] = ffma2(tn[j], mid[j + 2], dw[5]);
                dw[6] = ffma2(tn[j], bot[j], dw[6]);
                dw[7] = ffma2(tn[j], bot[j + 1], dw[7]);
                dw[8] = ffma2(tn[j], bot[j + 2], dw[8]);
              }
            }
            // scatter T[t] into dX rows t-1 (taps i=0), t (i=1), t+1 (i=2, first contribution)
            // dX[h][w] += wv[i][jj] * T[h-i+1][w-jj+1]  ->  T index for own column jo: jo + 2 - jj
            float2(&a_m1)[kCols] = da[(i + 1) % 3];   // row t-1 = r-2
            float2(&a_0)[kCols] = da[(i + 2) % 3];    // row t   = r-1
            float2(&a_p1)[kCols] = da[i];             // row t+1 = r
#pragma unroll
            for (int jo = 0; jo < kCols; ++jo) {
              a_m1[jo] = ffma2(w9[0], tn[jo + 2], a_m1[jo]);
              a_m1[jo] = ffma2(w9[1], tn[jo + 1], a_m1[jo]);
              a_m1[jo] = ffma2(w9[2], tn[jo], a_m1[jo]);
              a_0[jo] = ffma2(w9[3], tn[jo + 2], a_0[jo]);
              a_0[jo] = ffma2(w9[4], tn[jo + 1], a_0[jo]);
              a_0[jo] = ffma2(w9[5], tn[jo], a_0[jo]);
              float2 n = fmul2(w9[6], tn[jo + 2]);
              n = ffma2(w9[7], tn[jo + 1], n);
              a_p1[jo] = ffma2(w9[8], tn[jo], n);
            }
            if (t >= 1) {
#pragma unroll
              for (int jo = 0; jo < kCols; ++jo) {
                const float2 out = fadd2(ffma2(res2, dyprev[jo], dyc), a_m1[jo]);
                if (ovalid[jo]) stg_pair<T>(dxp + jo * P.C, out);
              }
            }
            if (t == P.H - 1) {
#pragma unroll
              for (int jo = 0; jo < kCols; ++jo) {
                const float2 out = fadd2(ffma2(res2, dyc_row[jo], dyc), a_0[jo]);
                if (ovalid[jo]) stg_pair<T>(dxp + row_stride + jo * P.C, out);
              }
            }
            dxp += row_stride;
            dop += row_stride;
#pragma unroll
            for (int jo = 0; jo < kCols; ++jo) dyprev[jo] = dyc_row[jo];
            if (rro == P.G - 1 || t == P.H - 1) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&empty[sto]);
            }
          }
          if (r < P.H) {
            if (++rr == P.G || r == P.H - 1) {
              rr = 0;
              st_prev = st_cur;
              if (++st_cur == P.S) { st_cur = 0; ph_cur ^= 1; }
            }
          }
        }
      }
    }
  }
  if (cur_cb >= 0) flush_dw(cur_cb);
}

}  // namespace mrla
