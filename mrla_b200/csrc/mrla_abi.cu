// extern "C" entry points of libmrla_b200.so (see include/mrla_b200.h).
#include "base_launch.cuh"
#include "layout_kernels.cuh"

namespace mrla {
thread_local int g_launch_count = 0;
thread_local const char* g_err_site = "";
thread_local char g_err_detail[192] = "";
static thread_local char g_err_msg[512];
extern template int light_forward_t<float>(const MrlaLightArgs&, cudaStream_t);
extern template int light_backward_t<float>(const MrlaLightArgs&, cudaStream_t);
extern template int light_forward_t<__nv_bfloat16>(const MrlaLightArgs&, cudaStream_t);
extern template int light_backward_t<__nv_bfloat16>(const MrlaLightArgs&, cudaStream_t);
extern template int light_forward_t<__half>(const MrlaLightArgs&, cudaStream_t);
extern template int light_backward_t<__half>(const MrlaLightArgs&, cudaStream_t);

extern template int base_forward_t<float>(const MrlaBaseArgs&, cudaStream_t);
extern template int base_backward_t<float>(const MrlaBaseArgs&, cudaStream_t);
extern template int base_forward_t<__nv_bfloat16>(const MrlaBaseArgs&, cudaStream_t);
extern template int base_backward_t<__nv_bfloat16>(const MrlaBaseArgs&, cudaStream_t);
extern template int base_forward_t<__half>(const MrlaBaseArgs&, cudaStream_t);
extern template int base_backward_t<__half>(const MrlaBaseArgs&, cudaStream_t);

static size_t esize(int dtype) { return dtype == MRLA_F32 ? 4 : 2; }

static int check_common(const MrlaLightArgs* a, bool bwd) {
  if (a == nullptr) return MRLA_ERR_NULL;
  if (a->B < 1 || a->C < 1 || a->H < 1 || a->W < 1) return MRLA_ERR_SHAPE;
  if (a->dim_perhead < 1 || a->C % a->dim_perhead) return MRLA_ERR_SHAPE;
  if (a->k_size < 1 || a->k_size > 15 || a->k_size % 2 == 0) return MRLA_ERR_SHAPE;
  if (a->dtype < MRLA_F32 || a->dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  if (a->layout != MRLA_NCHW && a->layout != MRLA_NHWC) return MRLA_ERR_UNSUPPORTED;
  if (a->act != MRLA_ACT_NONE && a->act != MRLA_ACT_GELU) return MRLA_ERR_UNSUPPORTED;
  if (a->bn_mode < MRLA_BN_NONE || a->bn_mode > MRLA_BN_EVAL) return MRLA_ERR_UNSUPPORTED;
  if ((!a->x && !a->x_virtual) || !a->wq || !a->wk || !a->wv || !a->mom || !a->gate || !a->mean || !a->rstd) return MRLA_ERR_NULL;
  if (a->x_virtual && (!a->z || !a->z_coef || !a->o)) return MRLA_ERR_NULL;
  if (a->o && !a->lam) return MRLA_ERR_NULL;
  if (a->bn_mode != MRLA_BN_NONE && (!a->gamma || (!bwd && !a->beta))) return MRLA_ERR_NULL;
  if (!bwd && a->bn_mode == MRLA_BN_EVAL && (!a->running_mean || !a->running_var)) return MRLA_ERR_NULL;
  if (a->layout == MRLA_NHWC) {
    // 4-channel vectors: base pointers and batch strides must keep them aligned
    const size_t al = 4 * esize(a->dtype);
    if (a->C % 4) return MRLA_ERR_ALIGN;
    const void* ptrs[] = {a->x, a->o, a->y, a->dy, a->dx, a->dout};
    for (const void* p : ptrs)
      if (p && ((uintptr_t)p % al)) return MRLA_ERR_ALIGN;
    const int64_t bss[] = {a->bs_x, a->bs_o, a->bs_y, a->bs_dy, a->bs_dx, a->bs_do};
    for (int64_t s : bss)
      if (s % 4) return MRLA_ERR_ALIGN;
  }
  return MRLA_OK;
}
static int check_base(const MrlaBaseArgs* a, bool bwd) {
  if (a == nullptr) return MRLA_ERR_NULL;
  if (a->B < 1 || a->C < 1 || a->H < 1 || a->W < 1) return MRLA_ERR_SHAPE;
  if (a->dim_perhead < 1 || a->C % a->dim_perhead) return MRLA_ERR_SHAPE;
  if (a->k_size < 1 || a->k_size > 15 || a->k_size % 2 == 0) return MRLA_ERR_SHAPE;
  if (a->t < 1 || a->t > a->t_cap) return MRLA_ERR_SHAPE;
  if (a->dtype < MRLA_F32 || a->dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  if (a->layout != MRLA_NCHW && a->layout != MRLA_NHWC) return MRLA_ERR_UNSUPPORTED;
  if (a->bn_mode < MRLA_BN_NONE || a->bn_mode > MRLA_BN_EVAL) return MRLA_ERR_UNSUPPORTED;
  if (!a->x || !a->v || !a->s || !a->kcache || !a->wq || !a->wk || !a->wv || !a->sx || !a->q || !a->p || !a->chan)
    return MRLA_ERR_NULL;
  if (a->bn_mode != MRLA_BN_NONE && (!a->gamma || (!bwd && !a->beta))) return MRLA_ERR_NULL;
  if (!bwd && a->bn_mode == MRLA_BN_EVAL && (!a->running_mean || !a->running_var)) return MRLA_ERR_NULL;
  if (a->layout == MRLA_NHWC) {
    const size_t al = 4 * esize(a->dtype);
    if (a->C % 4) return MRLA_ERR_ALIGN;
    const void* ptrs[] = {a->x, a->v, a->s, a->y, a->dy, a->dx, a->dv};
    for (const void* p : ptrs)
      if (p && ((uintptr_t)p % al)) return MRLA_ERR_ALIGN;
    const int64_t bss[] = {a->bs_x, a->bs_y, a->bs_s, a->bs_dy, a->bs_dx, a->bs_v, a->ts_v, a->bs_dv, a->ts_dv};
    for (int64_t s : bss)
      if (s % 4) return MRLA_ERR_ALIGN;
  }
  return MRLA_OK;
}
}  // namespace mrla

using namespace mrla;

extern "C" {

int mrla_abi_version(void) { return MRLA_ABI_VERSION; }

const char* mrla_build_info(void) {
#define MRLA_STR2(x) #x
#define MRLA_STR(x) MRLA_STR2(x)
  return "mrla_b200 sm_100a nvcc " MRLA_STR(__CUDACC_VER_MAJOR__) "." MRLA_STR(__CUDACC_VER_MINOR__) " built " __DATE__ " " __TIME__;
}

int mrla_last_launch_count(void) { return g_launch_count; }
const char* mrla_last_error_site(void) {
  snprintf(g_err_msg, sizeof(g_err_msg), "%s%s%s", g_err_site, g_err_detail[0] ? " " : "", g_err_detail);
  return g_err_msg;
}

size_t mrla_sizeof_light_args(void) { return sizeof(MrlaLightArgs); }

size_t mrla_light_bwd_scratch_bytes(const MrlaLightArgs* a) {
  if (a == nullptr) return 0;
  return light_bwd_scratch_floats(*a) * sizeof(float);
}

int mrla_light_bwd_fuses_relu(const MrlaLightArgs* a) {
  if (a == nullptr) return 0;
  return light_bwd_can_fuse_relu(*a) ? 1 : 0;
}

int mrla_light_fwd_folds_bn(const MrlaLightArgs* a) {
  if (a == nullptr) return 0;
  return light_fwd_can_fold_bn(*a) ? 1 : 0;
}

int mrla_light_virtual_x(const MrlaLightArgs* a) {
  if (a == nullptr) return 0;
  return v7_virtual_x_ok(*a) ? 1 : 0;
}

int mrla_light_v7_plan(const MrlaLightArgs* a, int kind, int xf, int64_t out[12]) {
  if (a == nullptr || out == nullptr || kind < 0 || kind > 3) return 0;
  V7Plan p{};
  const bool ok = v7_plan(*a, kind, xf != 0, &p);
  const int64_t v[12] = {ok ? 1 : 0, p.CB, p.NQ, p.NT, p.U, p.TPU, p.S, p.cpc, p.grid, p.threads, p.ctas, (int64_t)p.smem};
  for (int i = 0; i < 12; ++i) out[i] = ok ? v[i] : (i == 0 ? 0 : -1);
  return ok ? 1 : 0;
}

int mrla_light_forward(const MrlaLightArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_light_forward");
  g_launch_count = 0;
  g_err_site = "";
  g_err_detail[0] = 0;
  int rc = check_common(a, false);
  if (rc) return rc;
  if (!a->y || !a->coef) return MRLA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return light_forward_t<float>(*a, st);
    case MRLA_BF16: return light_forward_t<__nv_bfloat16>(*a, st);
    default: return light_forward_t<__half>(*a, st);
  }
}

int mrla_light_backward(const MrlaLightArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_light_backward");
  g_launch_count = 0;
  g_err_site = "";
  g_err_detail[0] = 0;
  int rc = check_common(a, true);
  if (rc) return rc;
  if (!a->dy || !a->dx || !a->gmom || !a->bcoef || !a->scratch) return MRLA_ERR_NULL;
  if (a->o && !a->dout) return MRLA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return light_backward_t<float>(*a, st);
    case MRLA_BF16: return light_backward_t<__nv_bfloat16>(*a, st);
    default: return light_backward_t<__half>(*a, st);
  }
}

int mrla_nchw_to_nhwc(const void* src, void* dst, int B, int C, int HW, int dtype, int64_t bs_src, int64_t bs_dst,
                      void* stream) {
  NvtxRange nvtx_("mrla_nchw_to_nhwc");
  g_launch_count = 0;
  if (!src || !dst) return MRLA_ERR_NULL;
  if (B < 1 || C < 1 || HW < 1 || B > 65535) return MRLA_ERR_SHAPE;
  if (dtype < MRLA_F32 || dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  if (C % 2 || ((uintptr_t)dst % (2 * esize(dtype))) || bs_dst % 2) return MRLA_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((HW + 63) / 64, (C + 63) / 64, B);
  if (dtype == MRLA_F32)
    k_nchw_to_nhwc<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), static_cast<float*>(dst), C, HW, bs_src, bs_dst);
  else
    k_nchw_to_nhwc<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), C, HW, bs_src, bs_dst);
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

int mrla_add_relu(const void* z, const void* idt, void* x, int64_t n, int dtype, void* stream) {
  NvtxRange nvtx_("mrla_add_relu");
  g_launch_count = 0;
  if (!z || !idt || !x) return MRLA_ERR_NULL;
  if (n < 1) return MRLA_ERR_SHAPE;
  if (dtype < MRLA_F32 || dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  const int64_t v = 16 / (int64_t)esize(dtype);
  if (n % v || ((uintptr_t)z % 16) || ((uintptr_t)idt % 16) || ((uintptr_t)x % 16)) return MRLA_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t nv = n / v;
  int64_t blocks = (nv + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == MRLA_F32)
    k_add_relu<float><<<(int)blocks, 256, 0, st>>>(static_cast<const float*>(z), static_cast<const float*>(idt), static_cast<float*>(x), nv);
  else if (dtype == MRLA_BF16)
    k_add_relu<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(idt), static_cast<__nv_bfloat16*>(x), nv);
  else
    k_add_relu<__half><<<(int)blocks, 256, 0, st>>>(static_cast<const __half*>(z), static_cast<const __half*>(idt), static_cast<__half*>(x), nv);
  MRLA_CHECK_LAUNCH();
  return MRLA_OK;
}

size_t mrla_sizeof_base_args(void) { return sizeof(MrlaBaseArgs); }

size_t mrla_base_bwd_scratch_bytes(const MrlaBaseArgs* a) {
  if (a == nullptr) return 0;
  return base_bwd_scratch_floats(*a) * sizeof(float);
}

int mrla_base_forward(const MrlaBaseArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_base_forward");
  g_launch_count = 0;
  int rc = check_base(a, false);
  if (rc) return rc;
  if (!a->y || !a->smom) return MRLA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return base_forward_t<float>(*a, st);
    case MRLA_BF16: return base_forward_t<__nv_bfloat16>(*a, st);
    default: return base_forward_t<__half>(*a, st);
  }
}

int mrla_base_backward(const MrlaBaseArgs* a, void* stream) {
  NvtxRange nvtx_("mrla_base_backward");
  g_launch_count = 0;
  int rc = check_base(a, true);
  if (rc) return rc;
  if (!a->dy || !a->dx || !a->dv || !a->dkcache || !a->gmom || !a->dpm || !a->dyc || !a->scratch) return MRLA_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case MRLA_F32: return base_backward_t<float>(*a, st);
    case MRLA_BF16: return base_backward_t<__nv_bfloat16>(*a, st);
    default: return base_backward_t<__half>(*a, st);
  }
}

}  // extern "C"
