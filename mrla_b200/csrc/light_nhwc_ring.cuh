// v4 sweep B (backward apply) of the MRLA-light tail, NHWC:  dS = Q0 + Q1*dy + Q2*V + Q3*o ;  T = Ta*dS*act'(U)
//     do = lam*dS ;  dx = res*dy + dwconv3x3^T(T) + dyc ;  dWv[c,i,j] = sum T[h,w]*x[h+i-1,w+j-1]
//     FUSE: x = relu(z + o) was formed in front of the tail -> dx receives dz = dx*[x>0], dout the total identity
//           gradient lam*dS + dz (replaces threshold_backward + the gradient-accumulation add of the reference graph).
// Replaces (paths relative to /root/reference) the autograd backward of mrla_light_module.py:56-72,
// resnet_mrla_light.py:42,113-116.
//
// Same tiles as the forward sweeps (light_nhwc_tma.cuh), one image row per pipeline stage; thread = (channel pair, KC
// columns).  KC = 8 for large images (7 warps at W = 56, 8 at W = 28: at most two warps per scheduler, so every
// thread may use ~250 registers — with 4 columns and 14 warps the 128-register cap forced spills and index
// re-derivation into every row step), KC = 4 with two CTAs per SM for small images.  There is no producer warp: the
// CTA-wide barrier of every row step already proves that the stage consumed one step earlier is free, so thread 0
// issues the next TMA loads (and the row stores) right after it.  T rows travel between column groups through a 4-slot
// shared-memory ring laid out in 128-bit column PAIRS (2 STS.128 per row step, 4 LDS.128 per tap row).
// What differs from the v3 kernel it replaces:
//   * the row loop is peeled (prologue rows 0..2, steady state, rows H and H+1), so the steady-state step carries no
//     per-row conditionals, no select chains and no register shuffles — v3 spent 2/3 of its issue slots on them;
//   * a CTA owns ONE channel block and walks the batch from the last sample down: dWv lives in registers for the whole
//     kernel, and the first samples it touches are the ones sweep A left in L2;
//   * dX / dO rows leave through a shared-memory staging row and one TMA tensor store per row (three rotating
//     buffers, bulk async groups): no per-thread global pointers, strides or store predicates — the TMA unit clips
//     ragged columns and partial channel blocks;
//   * the FUSE epilogue works on packed pairs (bf16x2 add for the identity gradient).
#pragma once
#include "light_nhwc_tma.cuh"

namespace mrla {

__device__ __forceinline__ void sts_v4(uint32_t saddr, float2 a, float2 b) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}
__device__ __forceinline__ void lds_v4(uint32_t saddr, float2& a, float2& b) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y) : "r"(saddr));
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar_s, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_s), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}

enum { DX_NONE = 0, DX_FULL = 1, DX_FIRST = 2, DX_LAST = 3 };

__device__ __forceinline__ void sts_raw(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_raw(uint32_t saddr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory");
}

// Per-thread state of a consumer; every array is indexed with compile-time constants after unrolling, so it all
// lives in registers.  CTA-uniform values are read from the kernel parameters where they are used.
template <typename T, int CB, int ACT, bool FUSE, bool RAGGED, int KC>
struct RingConsumer {
  static constexpr int NP = CB / 2;
  static constexpr int ES = sizeof(T);
  static constexpr int KW = KC + 2;         // x window columns
  static constexpr int HP = KC / 2;         // own ring column pairs
  static constexpr uint32_t CS = CB * ES;   // bytes between adjacent tile columns
  static constexpr uint32_t PS = NP * 16;   // bytes between adjacent ring column pairs
  typedef typename RawPair<T>::type Raw;

  const TmaBwdParams& P;
  const CUtensorMap* tm_dx;
  const CUtensorMap* tm_do;
  uint32_t full_s, stages_s, obuf_s;
  uint32_t tbase;            // byte offset of this thread's window column 0 inside a tile row
  uint32_t ring_rd;          // ring address of the first pair this thread reads (pair HP*q), slot 0
  uint32_t ring_slot_bytes;
  bool issuer;               // the one consumer thread that issues the TMA stores
  bool cv[KC];
  float2 w9[9], lm;
  float2 dw[9];
  // pipeline cursors
  int st;
  uint32_t ph;
  uint32_t t_prev;           // shared-memory address of the dy row fetched one step ago (o row = + t_bytes)
  int ob;                    // staging buffer the next dX row goes to (0..2)
  int pend_h, pend_ob;       // staged row waiting for its TMA store (pend_h < 0: none)
  int b, cb;
  // TMA load cursor (meaningful in the issuer thread): next (sample, row) to request and the stage it goes to
  const CUtensorMap* tm_x;
  const CUtensorMap* tm_dy;
  const CUtensorMap* tm_o;
  int ld_b, ld_row, ld_st, ld_left;   // ld_left = rows still to be requested
  // per sample
  float2 q0, q1, q2, q3, ta, dyc;
  int r;
  // rows in flight
  float2 xw[3][KW];
  Raw dyr[3][KC];
  Raw dor[3][KC];

  __device__ __forceinline__ RingConsumer(const TmaBwdParams& P_) : P(P_) {}

  // issuer only: request the next image row (x with halo columns, dy, o) into stage ld_st
  __device__ __forceinline__ void issue_load() {
    const uint32_t fb = full_s + (uint32_t)ld_st * 8;
    const uint32_t dst = stages_s + (uint32_t)ld_st * P.stage_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(P.stage_bytes) : "memory");
    tma_load_4d_s(dst, tm_x, fb, cb * CB, -1, ld_row, ld_b);
    tma_load_4d_s(dst + P.x_bytes, tm_dy, fb, cb * CB, 0, ld_row, ld_b);
    tma_load_4d_s(dst + P.x_bytes + P.t_bytes, tm_o, fb, cb * CB, 0, ld_row, ld_b);
    if (++ld_st == P.S) ld_st = 0;
    if (++ld_row == P.H) { ld_row = 0; ld_b -= P.cpc; }
    --ld_left;
  }

  // CTA-wide barrier of a row step; around it the issuer retires / launches the row stores and refills the stage
  // whose dy / o row every warp has just consumed (TROW)
  template <bool TROW>
  __device__ __forceinline__ void barrier_and_io(int b_of_pending) {
    if (issuer) bulk_wait_read<1>();   // the buffer staged two rows ago has been read: it is written again below
    __syncthreads();
    if (issuer) {
      if (TROW && ld_left > 0) issue_load();
      if (pend_h >= 0) {
        const uint32_t src = obuf_s + (uint32_t)pend_ob * 2 * P.t_bytes;
        tma_store_4d(tm_dx, src, cb * CB, 0, pend_h, b_of_pending);
        tma_store_4d(tm_do, src + P.t_bytes, cb * CB, 0, pend_h, b_of_pending);
        bulk_commit();
      }
    }
    pend_h = -1;
  }

  // One row step: fetch x row r into window slot I; T row r-1 -> ring; barrier; dX row r-2 from the ring.
  template <int I, bool FETCH, bool TROW, int DX>
  __device__ __forceinline__ void step(int b_of_pending) {
    uint32_t t_this = 0;
    if (FETCH) {
      const uint32_t fb = full_s + (uint32_t)st * 8;
      while (!mbar_try_wait_s(fb, ph)) {
      }
      const uint32_t xa = stages_s + (uint32_t)st * P.stage_bytes + tbase;
#pragma unroll
      for (int k = 0; k < KW; ++k) xw[I][k] = lds_pair<T>(xa + k * CS);
      t_this = xa + P.x_bytes;
      if (++st == P.S) { st = 0; ph ^= 1; }
    } else if (TROW) {
#pragma unroll
      for (int k = 0; k < KW; ++k) xw[I][k] = f2(0.f, 0.f);
    }
    if (TROW) {
      const float2(&top)[KW] = xw[(I + 1) % 3];
      const float2(&mid)[KW] = xw[(I + 2) % 3];
      const float2(&bot)[KW] = xw[I];
      const uint32_t ws = ring_rd + PS + (uint32_t)((r - 1) & 3) * ring_slot_bytes;   // own pairs HP*q+1 ..
      // four columns at a time: tap-major conv (independent FFMA2 chains), then dS / T / dWv, then one STS.128 pair
#pragma unroll
      for (int j0 = 0; j0 < KC; j0 += 4) {
        float2 u[4], tt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = fmul2(w9[0], top[j0 + j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[1], top[j0 + j + 1], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[2], top[j0 + j + 2], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[3], mid[j0 + j], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[4], mid[j0 + j + 1], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[5], mid[j0 + j + 2], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[6], bot[j0 + j], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[7], bot[j0 + j + 1], u[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ffma2(w9[8], bot[j0 + j + 2], u[j]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = j0 + jj;
          const float2 v = act2<ACT>(u[jj]);
          const Raw graw = lds_raw<T>(t_prev + j * CS);
          const float2 gy = unpack_pair<T>(graw);
          const float2 ov = lds_pair<T>(t_prev + P.t_bytes + j * CS);
          float2 ds = ffma2(q1, gy, q0);
          ds = ffma2(q2, v, ds);
          ds = ffma2(q3, ov, ds);
          float2 t = fmul2(ta, ds);
          if (ACT == 1) t = fmul2(t, f2(act_grad<1>(u[jj].x), act_grad<1>(u[jj].y)));
          if (RAGGED && !cv[j]) t = f2(0.f, 0.f);
          tt[jj] = t;
          dyr[I][j] = graw;
          dor[I][j] = pack_pair<T>(fmul2(lm, ds));
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
        sts_v4(ws + (j0 / 2) * PS, tt[0], tt[1]);
        sts_v4(ws + (j0 / 2 + 1) * PS, tt[2], tt[3]);
      }
    }
    t_prev = t_this;
    // after this barrier the stage holding dy / o row r-1 (x row r-1 is in registers) is free again
    barrier_and_io<TROW>(b_of_pending);
    if (DX != DX_NONE) {
      float2 acc[KC];
#pragma unroll
      for (int j = 0; j < KC; ++j) acc[j] = ffma2(f2(P.res, P.res), unpack_pair<T>(dyr[(I + 2) % 3][j]), dyc);
#pragma unroll
      for (int ii = 0; ii < 3; ++ii) {
        if (DX == DX_FIRST && ii == 2) continue;   // T row -1
        if (DX == DX_LAST && ii == 0) continue;    // T row H
        // tap row ii reads T row r-1-ii;  tw[k] = image column KC*q-2+k
        const uint32_t rs = ring_rd + (uint32_t)((r - 1 - ii) & 3) * ring_slot_bytes;
        float2 tw[KC + 4];
#pragma unroll
        for (int k = 0; k < HP + 2; ++k) lds_v4(rs + k * PS, tw[2 * k], tw[2 * k + 1]);
        // dX[h][w] += wv[ii][dj] * T[h-ii+1][w-dj+1]
#pragma unroll
        for (int j = 0; j < KC; ++j) acc[j] = ffma2(w9[ii * 3 + 0], tw[j + 3], acc[j]);
#pragma unroll
        for (int j = 0; j < KC; ++j) acc[j] = ffma2(w9[ii * 3 + 1], tw[j + 2], acc[j]);
#pragma unroll
        for (int j = 0; j < KC; ++j) acc[j] = ffma2(w9[ii * 3 + 2], tw[j + 1], acc[j]);
      }
      // stage the dX / dO row (tile layout [column][CB channels], like the input tiles) for the TMA store
      const uint32_t oa = obuf_s + (uint32_t)ob * 2 * P.t_bytes + tbase;
      const float2(&xrow)[KW] = xw[(I + 1) % 3];   // x row h = r-2
#pragma unroll
      for (int j = 0; j < KC; ++j) {
        if (FUSE) {
          const float2 xc = xrow[j + 1];
          const Raw dz = pack_pair<T>(f2(xc.x > 0.f ? acc[j].x : 0.f, xc.y > 0.f ? acc[j].y : 0.f));
          sts_raw(oa + j * CS, dz);
          sts_raw(oa + P.t_bytes + j * CS, raw_add<T>(dor[(I + 2) % 3][j], dz));   // total identity gradient
        } else {
          sts_raw(oa + j * CS, pack_pair<T>(acc[j]));
          sts_raw(oa + P.t_bytes + j * CS, dor[(I + 2) % 3][j]);
        }
      }
      fence_proxy_async_smem();
      pend_h = r - 2;
      pend_ob = ob;
      ob = (ob == 2) ? 0 : ob + 1;
    }
    ++r;
  }

  // bring window slot 1 to slot 0 (with the per-row side registers)
  __device__ __forceinline__ void rotate_left1() {
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      const float2 t0 = xw[0][k];
      xw[0][k] = xw[1][k]; xw[1][k] = xw[2][k]; xw[2][k] = t0;
    }
#pragma unroll
    for (int j = 0; j < KC; ++j) {
      const Raw t0 = dyr[0][j];
      dyr[0][j] = dyr[1][j]; dyr[1][j] = dyr[2][j]; dyr[2][j] = t0;
      const Raw t1 = dor[0][j];
      dor[0][j] = dor[1][j]; dor[1][j] = dor[2][j]; dor[2][j] = t1;
    }
  }

  // one image of H >= 3 rows; b_prev = sample of the row store still pending from the previous image
  __device__ __forceinline__ void image(int b_prev) {
    const int H = P.H;
    r = 0;
#pragma unroll
    for (int k = 0; k < KW; ++k) xw[2][k] = f2(0.f, 0.f);   // x row -1
    // the barrier of step 0 separates the previous image's last ring reads from this image's first ring write
    step<0, true, false, DX_NONE>(b_prev);
    step<1, true, true, DX_NONE>(b);
    step<2, true, true, DX_FIRST>(b);
    while (r + 2 < H) {
      step<0, true, true, DX_FULL>(b);
      step<1, true, true, DX_FULL>(b);
      step<2, true, true, DX_FULL>(b);
    }
    // r % 3 == 0 here; up to two more steady rows, then bring the slot of row H to index 0
    if (r < H) {
      step<0, true, true, DX_FULL>(b);
      if (r < H) {
        step<1, true, true, DX_FULL>(b);
        rotate_left1();
        rotate_left1();
      } else {
        rotate_left1();
      }
    }
    step<0, false, true, DX_FULL>(b);    // r = H: x row H is zero padding
    step<1, false, false, DX_LAST>(b);   // r = H+1
  }
};

// KC = 8: one CTA of up to 256 threads per SM (W = 56 / 28); KC = 4: up to 128 threads, two CTAs per SM
template <typename T, int CB, int ACT, bool FUSE, bool RAGGED, int KC>
__global__ void __launch_bounds__(KC == 8 ? 256 : 128, KC == 8 ? 1 : 2)
k_light_nhwc_ring(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                  const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_dx,
                  const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ TmaBwdParams P) {
  typedef RingConsumer<T, CB, ACT, FUSE, RAGGED, KC> Cons;
  constexpr int NP = CB / 2;
  constexpr int ES = sizeof(T);
  constexpr int HP = KC / 2;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned char* stages = smem_raw + 256;
  unsigned char* ring = stages + (size_t)P.S * P.stage_bytes;     // [4][HP*NQ+2 column pairs][NP] float4
  const int npairs = HP * P.NQ + 2;
  const uint32_t ring_slot_bytes = (uint32_t)npairs * Cons::PS;
  unsigned char* obuf = ring + (size_t)4 * ring_slot_bytes;       // [3][dX row, dO row] staging for the TMA stores
  // CTA -> (channel block, lane m of that block's CTAs); samples b = B-1-m, B-1-m-cpc, ...
  const int cb = blockIdx.x / P.cpc, m = blockIdx.x - cb * P.cpc;
  const int n_my = (m < P.B) ? (P.B - 1 - m) / P.cpc + 1 : 0;

  const int ct = threadIdx.x;
  const int p = ct % NP, q = ct / NP;
  const uint32_t ring_s = smem_u32(ring);
  Cons K(P);
  K.tm_x = &tm_x;
  K.tm_dy = &tm_dy;
  K.tm_o = &tm_o;
  K.tm_dx = &tm_dx;
  K.tm_do = &tm_do;
  K.full_s = smem_u32(full);
  K.stages_s = smem_u32(stages);
  K.obuf_s = smem_u32(obuf);
  K.tbase = (uint32_t)(q * KC) * Cons::CS + (uint32_t)p * 2 * ES;
  K.ring_rd = ring_s + (uint32_t)(HP * q) * Cons::PS + (uint32_t)p * 16;
  K.ring_slot_bytes = ring_slot_bytes;
  K.issuer = (ct == 0);
  K.st = 0;
  K.ph = 0;
  K.t_prev = 0;
  K.ob = 0;
  K.pend_h = -1;
  K.pend_ob = 0;
  K.cb = cb;
  K.ld_b = P.B - 1 - m;
  K.ld_row = 0;
  K.ld_st = 0;
  K.ld_left = n_my * P.H;
  if (K.issuer) {
    for (int s = 0; s < P.S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_o);
    tma_prefetch_desc(&tm_dx);
    tma_prefetch_desc(&tm_do);
    for (int s = 0; s < P.S && K.ld_left > 0; ++s) K.issue_load();   // fill the pipeline
  }
#pragma unroll
  for (int j = 0; j < KC; ++j) K.cv[j] = (q * KC + j) < P.W;
  const int c = cb * CB + 2 * p;
  const bool chan_ok = c < P.C;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    K.w9[i] = chan_ok ? f2(P.wv[(int64_t)c * 9 + i], P.wv[(int64_t)(c + 1) * 9 + i]) : f2(0.f, 0.f);
    K.dw[i] = f2(0.f, 0.f);
  }
  K.lm = (chan_ok && P.lam) ? f2(P.lam[c], P.lam[c + 1]) : f2(0.f, 0.f);
  // ring column pairs 0 and npairs-1 (image columns -2,-1 and KC*NQ, KC*NQ+1) are never written: zero them once
  if (q == 0) {
    for (int sl = 0; sl < 4; ++sl) {
      sts_v4(ring_s + sl * ring_slot_bytes + (uint32_t)p * 16, f2(0.f, 0.f), f2(0.f, 0.f));
      sts_v4(ring_s + sl * ring_slot_bytes + (uint32_t)(npairs - 1) * Cons::PS + (uint32_t)p * 16, f2(0.f, 0.f), f2(0.f, 0.f));
    }
  }
  __syncthreads();   // mbarrier init + ring halo visible to everyone

  const int64_t BC = (int64_t)P.B * P.C;
  int b_prev = 0;
  for (int it = 0; it < n_my; ++it) {
    const int b = P.B - 1 - m - it * P.cpc;
    K.b = b;
    K.q0 = K.q1 = K.q2 = K.q3 = K.ta = K.dyc = f2(0.f, 0.f);
    if (chan_ok) {
      const float* cp = P.bcoef + (int64_t)b * P.C + c;
      K.q0 = *reinterpret_cast<const float2*>(cp);
      K.q1 = *reinterpret_cast<const float2*>(cp + BC);
      K.q2 = *reinterpret_cast<const float2*>(cp + 2 * BC);
      K.q3 = *reinterpret_cast<const float2*>(cp + 3 * BC);
      K.ta = *reinterpret_cast<const float2*>(cp + 4 * BC);
      K.dyc = *reinterpret_cast<const float2*>(cp + 5 * BC);
    }
    K.image(b_prev);
    b_prev = b;
  }

  // last staged row; then the dWv partial of this CTA: reduce over the NQ column groups (scratch aliases the ring),
  // slot m of channel block cb
  K.template barrier_and_io<false>(b_prev);
  if (K.issuer) bulk_wait_all<0>();
  float2* red = reinterpret_cast<float2*>(ring);
#pragma unroll
  for (int i = 0; i < 9; ++i) red[((size_t)q * 9 + i) * NP + p] = K.dw[i];
  __syncthreads();
  for (int idx = ct; idx < 9 * NP; idx += blockDim.x) {
    const int tap = idx / NP, pp = idx - tap * NP;
    float2 s = f2(0.f, 0.f);
    for (int qq = 0; qq < P.NQ; ++qq) {
      const float2 v = red[((size_t)qq * 9 + tap) * NP + pp];
      s.x += v.x;
      s.y += v.y;
    }
    const int cc = cb * CB + 2 * pp;
    if (cc < P.C) {
      float* dst = P.wv_part + ((int64_t)m * P.C + cc) * 9 + tap;
      dst[0] = s.x;
      dst[9] = s.y;
    }
  }
}

}  // namespace mrla
