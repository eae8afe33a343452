// extern "C" entry points of the channels-last BatchNorm(+ReLU) producer op (see include/mrla_b200.h).
#include "../../include/mrla_b200.h"
#include "bn_kernels.cuh"
#include "light_launch.cuh"

namespace mrla {

static int bn_plan(const MrlaBnArgs* a, BnShape* s) {
  if (a == nullptr) return MRLA_ERR_NULL;
  if (a->M < 1 || a->C < 8 || a->C > 2048) return MRLA_ERR_SHAPE;
  if (a->C % 8) return MRLA_ERR_ALIGN;
  if (a->dtype < MRLA_F32 || a->dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  s->M = a->M; s->C = a->C;
  s->CL = a->C / 8;
  s->RL = 256 / s->CL;
  if (s->RL < 1) return MRLA_ERR_SHAPE;
  const int64_t rows_per_pass = s->RL;
  int64_t grid = (a->M + rows_per_pass * 8 - 1) / (rows_per_pass * 8);   // >= 8 rows per thread
  if (grid > kNumSMs * 3) grid = kNumSMs * 3;
  if (grid < 1) grid = 1;
  s->nparts = (int)grid;
  return MRLA_OK;
}

template <typename T>
static int bn_forward_t(const MrlaBnArgs& a, const BnShape& s, cudaStream_t st) {
  const T* x = static_cast<const T*>(a.x);
  T* y = static_cast<T*>(a.y);
  float* pivot = a.scratch + (size_t)s.nparts * 2 * a.C + (size_t)3 * a.C;   // [C] shift of the statistics sums
  if (a.training) {
    k_bn_stats<T><<<s.nparts, 256, 256 * 2 * kSV * sizeof(float), st>>>(x, a.scratch, pivot, s);
    MRLA_CHECK_LAUNCH();
  }
  k_bn_finalize<<<(a.C + 31) / 32, 1024, 0, st>>>(a.scratch, pivot, s.nparts, a.C, (double)a.M, a.gamma, a.beta, a.running_mean,
                                                 a.running_var, a.stats, a.coef, a.eps, a.momentum, a.training,
                                                 a.update_running);
  MRLA_CHECK_LAUNCH();
  if (a.stats_only) return MRLA_OK;   // the consumer applies a_c x + b_c itself (MrlaLightArgs.z_coef)
  if (a.relu) k_bn_apply<T, true><<<s.nparts, 256, 0, st>>>(x, y, a.coef, s);
  else k_bn_apply<T, false><<<s.nparts, 256, 0, st>>>(x, y, a.coef, s);
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

template <typename T>
static int bn_backward_t(const MrlaBnArgs& a, const BnShape& s, cudaStream_t st) {
  const T* x = static_cast<const T*>(a.x);
  const T* dy = static_cast<const T*>(a.dy);
  T* dx = static_cast<T*>(a.dx);
  float* bcoef = a.scratch + (size_t)s.nparts * 2 * a.C;
  const size_t sm = 256 * 2 * kSV * sizeof(float);
  if (a.sums == nullptr) {
    if (a.relu) k_bn_bwd_reduce<T, true><<<s.nparts, 256, sm, st>>>(dy, x, a.coef, a.stats, a.scratch, s);
    else k_bn_bwd_reduce<T, false><<<s.nparts, 256, sm, st>>>(dy, x, a.coef, a.stats, a.scratch, s);
    MRLA_CHECK_LAUNCH();
  }
  // a.sums: the producer of dy (sweep B of the MRLA tail, MrlaLightArgs.dz_sums) already reduced sum dy, sum dy*x
  k_bn_bwd_finalize<<<(a.C + 31) / 32, 1024, 0, st>>>(a.sums ? a.sums : a.scratch, a.sums ? 1 : s.nparts, a.C, (double)a.M,
                                                     a.gamma, a.stats, bcoef, a.dgamma, a.dbeta, a.training, a.sums ? 0 : 1);
  MRLA_CHECK_LAUNCH();
  if (a.relu) k_bn_bwd_apply<T, true><<<s.nparts, 256, 0, st>>>(dy, x, dx, a.coef, bcoef, s);
  else k_bn_bwd_apply<T, false><<<s.nparts, 256, 0, st>>>(dy, x, dx, a.coef, bcoef, s);
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}
}  // namespace mrla

using namespace mrla;

extern "C" {

size_t mrla_sizeof_bn_args(void) { return sizeof(MrlaBnArgs); }

size_t mrla_bn_scratch_bytes(const MrlaBnArgs* a) {
  BnShape s;
  if (bn_plan(a, &s)) return 0;
  return ((size_t)s.nparts * 2 * a->C + (size_t)4 * a->C) * sizeof(float);   // partials | bcoef [3,C] | pivot [C]
}

int mrla_bn_forward(const MrlaBnArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_bn_forward");
  g_launch_count = 0;
  BnShape s;
  int rc = bn_plan(a, &s);
  if (rc) return rc;
  if (!a->x || (!a->y && !a->stats_only) || !a->stats || !a->coef || !a->scratch) return MRLA_ERR_NULL;
  if (a->stats_only && a->relu) return MRLA_ERR_UNSUPPORTED;
  if (!a->training && (!a->running_mean || !a->running_var)) return MRLA_ERR_NULL;
  if (((uintptr_t)a->x % 16) || (a->y && ((uintptr_t)a->y % 16))) return MRLA_ERR_ALIGN;
  if (a->scratch_bytes < mrla_bn_scratch_bytes(a)) return MRLA_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return bn_forward_t<float>(*a, s, st);
    case MRLA_BF16: return bn_forward_t<__nv_bfloat16>(*a, s, st);
    default: return bn_forward_t<__half>(*a, s, st);
  }
}

int mrla_bn_backward(const MrlaBnArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_bn_backward");
  g_launch_count = 0;
  BnShape s;
  int rc = bn_plan(a, &s);
  if (rc) return rc;
  if (!a->x || !a->dy || !a->dx || !a->stats || !a->coef || !a->scratch) return MRLA_ERR_NULL;
  if (a->sums && a->relu) return MRLA_ERR_UNSUPPORTED;
  if (((uintptr_t)a->x % 16) || ((uintptr_t)a->dy % 16) || ((uintptr_t)a->dx % 16)) return MRLA_ERR_ALIGN;
  if (a->scratch_bytes < mrla_bn_scratch_bytes(a)) return MRLA_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return bn_backward_t<float>(*a, s, st);
    case MRLA_BF16: return bn_backward_t<__nv_bfloat16>(*a, s, st);
    default: return bn_backward_t<__half>(*a, s, st);
  }
}

}  // extern "C"
