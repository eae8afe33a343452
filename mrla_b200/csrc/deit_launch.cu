// extern "C" entry points of the fused DeiT MRLA-light module (see include/mrla_b200.h, deit_fused.cuh).
#include "../../include/mrla_b200.h"
#include "deit_fused.cuh"
#include "light_v7_launch.cuh"

namespace mrla {

struct DeitPlan {
  int NQ, threads, nw;
  size_t smem_fwd, smem_bwd;
};

static bool deit_plan(const MrlaDeitArgs* a, DeitPlan* p) {
  if (a == nullptr) return false;
  if (a->B < 1 || a->C < 64 || a->C % 64 || a->C > 64 * kDeitMaxCJ) return false;
  if (a->S < 1 || a->S > 2 * kV7 || a->n != a->S * a->S + 1) return false;
  if (a->dim_perhead < 1 || a->C % a->dim_perhead || a->C / a->dim_perhead > 64) return false;
  if (a->k_size < 1 || a->k_size > 15 || a->k_size % 2 == 0) return false;
  if (a->dtype < MRLA_F32 || a->dtype > MRLA_F16) return false;
  const size_t es = a->dtype == MRLA_F32 ? 4 : 2;
  p->NQ = (a->S + kV7 - 1) / kV7;
  if (p->NQ * a->C / 2 > 384) return false;
  p->threads = (p->NQ * a->C / 2 <= 192) ? 192 : 384;   // the image march uses the first NQ*C/2 threads
  p->nw = p->threads / 32;
  const size_t tile = (size_t)a->n * a->C * es;
  p->smem_fwd = tile + ((size_t)2 * a->C + 64 + (size_t)p->nw * a->C) * sizeof(float) + (size_t)a->n * sizeof(float2);
  size_t red = (size_t)p->NQ * 10 * a->C;
  if ((size_t)p->nw * 5 * a->C > red) red = (size_t)p->nw * 5 * a->C;
  p->smem_bwd = tile + ((size_t)6 * a->C + 128 + red) * sizeof(float);
  return p->smem_fwd <= 227 * 1024 && p->smem_bwd <= 227 * 1024;
}

static void deit_fill(DeitParams* P, const MrlaDeitArgs& a, const DeitPlan& p) {
  P->B = a.B; P->n = a.n; P->C = a.C; P->S = a.S; P->d = a.dim_perhead; P->k = a.k_size; P->NQ = p.NQ; P->eps = a.eps;
  P->x = a.x; P->o = a.o; P->out = a.out;
  P->gx = a.normx_w; P->bx = a.normx_b; P->go = a.normo_w; P->bo = a.normo_b;
  P->wq = a.wq; P->wk = a.wk; P->wv = a.wv; P->lam = a.lam;
  P->stats_x = a.stats_x; P->stats_o = a.stats_o; P->gate = a.gate;
  P->dout = a.dout; P->dx = a.dx; P->dox = a.dox;
  P->part = a.scratch; P->PW = 14 * a.C + 2 * a.k_size;
}

template <typename T>
static int deit_forward_t(const MrlaDeitArgs& a, const DeitPlan& p, cudaStream_t st) {
  DeitParams P;
  deit_fill(&P, a, p);
  cudaError_t e;
  if (p.threads == 192) {
    auto k = k_deit_light_fwd<T, 192>;
    e = ensure_smem_once(k, p.smem_fwd);
    if (e != cudaSuccess) return (int)e;
    k<<<a.B, p.threads, p.smem_fwd, st>>>(P);
  } else {
    auto k = k_deit_light_fwd<T, 384>;
    e = ensure_smem_once(k, p.smem_fwd);
    if (e != cudaSuccess) return (int)e;
    k<<<a.B, p.threads, p.smem_fwd, st>>>(P);
  }
  MRLA_V7_CHECK();
  return MRLA_OK;
}

template <typename T>
static int deit_backward_t(const MrlaDeitArgs& a, const DeitPlan& p, cudaStream_t st) {
  DeitParams P;
  deit_fill(&P, a, p);
  cudaError_t e;
  if (p.threads == 192) {
    auto k = k_deit_light_bwd<T, 192>;
    e = ensure_smem_once(k, p.smem_bwd);
    if (e != cudaSuccess) return (int)e;
    k<<<a.B, p.threads, p.smem_bwd, st>>>(P);
  } else {
    auto k = k_deit_light_bwd<T, 384>;
    e = ensure_smem_once(k, p.smem_bwd);
    if (e != cudaSuccess) return (int)e;
    k<<<a.B, p.threads, p.smem_bwd, st>>>(P);
  }
  MRLA_V7_CHECK();
  k_deit_reduce<<<(P.PW + 255) / 256, 256, 0, st>>>(a.scratch, a.B, P.PW, a.dparams);
  MRLA_V7_CHECK();
  return MRLA_OK;
}

static int ln_grid(const MrlaLnArgs* a) {
  const int64_t blocks = ((int64_t)a->B * a->n + 7) / 8;
  return (int)(blocks < 592 ? blocks : 592);
}
static bool ln_ok(const MrlaLnArgs* a) {
  return a != nullptr && a->B >= 1 && a->n >= 1 && a->C >= 2 && a->C % 2 == 0 && a->C <= 768 && a->dtype >= MRLA_F32 &&
         a->dtype <= MRLA_F16;
}
static void ln_fill(LnParams* P, const MrlaLnArgs& a) {
  P->B = a.B; P->n = a.n; P->C = a.C; P->eps = a.eps;
  P->x = a.x; P->xn = a.xn; P->cls_out = a.cls_out; P->bs_cls = a.bs_cls;
  P->gamma = a.gamma; P->beta = a.beta; P->stats = a.stats;
  P->g_cls = a.g_cls; P->bs_gcls = a.bs_gcls; P->g_img = a.g_img; P->bs_gimg = a.bs_gimg;
  P->dx = a.dx; P->part = a.scratch;
}
// partial [nparts, 2, C] -> out [2, C]
static __global__ void k_ln_reduce(const float* __restrict__ part, int nparts, int n2c, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2c) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(int64_t)p * n2c + i];
  out[i] = s;
}

}  // namespace mrla

using namespace mrla;

extern "C" {

size_t mrla_sizeof_ln_args(void) { return sizeof(MrlaLnArgs); }

size_t mrla_layernorm_scratch_bytes(const MrlaLnArgs* a) {
  if (!ln_ok(a)) return 0;
  return (size_t)ln_grid(a) * 2 * a->C * sizeof(float);
}

int mrla_layernorm_forward(const MrlaLnArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_layernorm_forward");
  g_launch_count = 0;
  if (a == nullptr) return MRLA_ERR_NULL;
  if (!ln_ok(a)) return MRLA_ERR_UNSUPPORTED;
  if (!a->x || !a->xn || !a->gamma || !a->beta || !a->stats) return MRLA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LnParams P;
  ln_fill(&P, *a);
  const int grid = ln_grid(a);
  if (a->dtype == MRLA_F32) k_ln_tokens_fwd<float><<<grid, 256, 0, st>>>(P);
  else if (a->dtype == MRLA_BF16) k_ln_tokens_fwd<__nv_bfloat16><<<grid, 256, 0, st>>>(P);
  else k_ln_tokens_fwd<__half><<<grid, 256, 0, st>>>(P);
  MRLA_V7_CHECK();
  return MRLA_OK;
}

int mrla_layernorm_backward(const MrlaLnArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_layernorm_backward");
  g_launch_count = 0;
  if (a == nullptr) return MRLA_ERR_NULL;
  if (!ln_ok(a)) return MRLA_ERR_UNSUPPORTED;
  if (!a->x || !a->gamma || !a->stats || !a->g_cls || (a->n > 1 && !a->g_img) || !a->dx || !a->dparams || !a->scratch)
    return MRLA_ERR_NULL;
  if (a->scratch_bytes < mrla_layernorm_scratch_bytes(a)) return MRLA_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LnParams P;
  ln_fill(&P, *a);
  const int grid = ln_grid(a);
  const size_t sm = (size_t)2 * a->C * sizeof(float);
  if (a->dtype == MRLA_F32) k_ln_tokens_bwd<float><<<grid, 256, sm, st>>>(P);
  else if (a->dtype == MRLA_BF16) k_ln_tokens_bwd<__nv_bfloat16><<<grid, 256, sm, st>>>(P);
  else k_ln_tokens_bwd<__half><<<grid, 256, sm, st>>>(P);
  MRLA_V7_CHECK();
  k_ln_reduce<<<(2 * a->C + 255) / 256, 256, 0, st>>>(a->scratch, grid, 2 * a->C, a->dparams);
  MRLA_V7_CHECK();
  return MRLA_OK;
}

size_t mrla_sizeof_deit_args(void) { return sizeof(MrlaDeitArgs); }

int mrla_deit_light_supported(const MrlaDeitArgs* a) {
  DeitPlan p;
  return deit_plan(a, &p) ? 1 : 0;
}

size_t mrla_deit_light_scratch_bytes(const MrlaDeitArgs* a) {
  DeitPlan p;
  if (!deit_plan(a, &p)) return 0;
  return (size_t)a->B * (14 * (size_t)a->C + 2 * a->k_size) * sizeof(float);
}

int mrla_deit_light_forward(const MrlaDeitArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_deit_light_forward");
  g_launch_count = 0;
  DeitPlan p;
  if (a == nullptr) return MRLA_ERR_NULL;
  if (!deit_plan(a, &p)) return MRLA_ERR_UNSUPPORTED;
  if (!a->x || !a->o || !a->out || !a->normx_w || !a->normx_b || !a->normo_w || !a->normo_b || !a->wq || !a->wk || !a->wv ||
      !a->lam || !a->stats_x || !a->stats_o || !a->gate)
    return MRLA_ERR_NULL;
  if (((uintptr_t)a->x % 4) || ((uintptr_t)a->o % 4) || ((uintptr_t)a->out % 4)) return MRLA_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return deit_forward_t<float>(*a, p, st);
    case MRLA_BF16: return deit_forward_t<__nv_bfloat16>(*a, p, st);
    default: return deit_forward_t<__half>(*a, p, st);
  }
}

int mrla_deit_light_backward(const MrlaDeitArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_deit_light_backward");
  g_launch_count = 0;
  DeitPlan p;
  if (a == nullptr) return MRLA_ERR_NULL;
  if (!deit_plan(a, &p)) return MRLA_ERR_UNSUPPORTED;
  if (!a->x || !a->o || !a->dout || !a->dx || !a->dox || !a->dparams || !a->scratch || !a->stats_x || !a->stats_o || !a->gate ||
      !a->normx_w || !a->normx_b || !a->normo_w || !a->normo_b || !a->wq || !a->wk || !a->wv || !a->lam)
    return MRLA_ERR_NULL;
  if (a->scratch_bytes < mrla_deit_light_scratch_bytes(a)) return MRLA_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return deit_backward_t<float>(*a, p, st);
    case MRLA_BF16: return deit_backward_t<__nv_bfloat16>(*a, p, st);
    default: return deit_backward_t<__half>(*a, p, st);
  }
}

}  // extern "C"
