/*
 * mrla_b200.h — C ABI of libmrla_b200.so (hand-written sm_100a kernels for the MRLA block tail).
 *
 * The reference (joyfang1106/MRLA) is pure Python/PyTorch and has no FFI of its own; the
 * "interface each entry point replaces" is therefore the sequence of ATen calls issued by the
 * reference modules.  Paths below are relative to /root/reference.
 *
 *   mrla_light_forward / mrla_light_backward
 *       replace  resnet/models/modules/mrla_light_module.py:52-74   (mrla_light_layer.forward)
 *                resnet/models/resnet_mrla_light.py:40-43            (mrla_module.forward, lambda recurrence)
 *                resnet/models/resnet_mrla_light.py:116              (bn_mrla + drop_path + residual)
 *                resnet/models/utils/drop.py:17-23                   (DropPath, consumed as a [B] scale)
 *                mmdetection/mmdet/models/backbones/resnet_mrlal.py:116 (eval-BN variant, bn_mode=2)
 *                deit/deit_mrla_light.py:157-180,204-206             (token layout, GELU on V, act=1, bn_mode=0)
 *       and the autograd graph PyTorch builds for them.
 *
 *   mrla_base_forward / mrla_base_backward
 *       replace  resnet/models/modules/mrla_base_module.py:54-89    (mrla_base_layer.forward)
 *                resnet/models/resnet_mrla_base.py:124-127           (bn_mrla + relu + drop_path + residual)
 *                deit/deit_mrla_base.py:166-201,224-243              (token layout, bn_mode=0)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates,
 *     frees or retains memory, never synchronises the host and only enqueues work on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph capturable and re-entrant.
 *   - return value: 0 = success; negative = argument / shape / alignment / unsupported-config
 *     error detected before any launch (see MRLA_ERR_*); positive = cudaError_t from a launch.
 *   - activations x,o,y,dy,dx,dout use `dtype` (fp32 / bf16 / fp16) in `layout`; all parameters,
 *     statistics, [B,C] side tensors and gradients of parameters are fp32.
 *   - NCHW: element (b,c,h,w) lives at  b*bs + (c*H + h)*W + w ;
 *     NHWC: element (b,c,h,w) lives at  b*bs + (h*W + w)*C + c   (channels_last / token-major).
 *     `bs_*` are per-tensor batch strides in elements (lets DeiT pass the [B,197,C] token
 *     buffer offset by one token, and MRLA-base pass slots of a [B,T,C,H,W] cache).
 */
#ifndef MRLA_B200_H_
#define MRLA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRLA_ABI_VERSION 4

enum { MRLA_F32 = 0, MRLA_BF16 = 1, MRLA_F16 = 2 };
enum { MRLA_NCHW = 0, MRLA_NHWC = 1 };
enum { MRLA_ACT_NONE = 0, MRLA_ACT_GELU = 1 };
enum { MRLA_BN_NONE = 0, MRLA_BN_TRAIN = 1, MRLA_BN_EVAL = 2 };

enum {
  MRLA_OK = 0,
  MRLA_ERR_NULL = -1,        /* a required pointer is NULL                          */
  MRLA_ERR_SHAPE = -2,       /* B,C,H,W,d,k out of the supported range             */
  MRLA_ERR_ALIGN = -3,       /* pointer / stride not aligned for the vector width  */
  MRLA_ERR_UNSUPPORTED = -4, /* dtype / layout / flag combination not implemented  */
  MRLA_ERR_WORKSPACE = -5    /* scratch buffer too small                           */
};

/* One MRLA-light block tail (forward and backward share the struct; unused fields may be NULL). */
typedef struct MrlaLightArgs {
  /* ---- problem ---- */
  int32_t B, C, H, W;
  int32_t dim_perhead;  /* d; heads g = C/d (mrla_light_module.py:32-38)                       */
  int32_t k_size;       /* ECA kernel size (mrla_light_module.py:40-43), odd                   */
  int32_t dtype;        /* MRLA_F32 / MRLA_BF16 / MRLA_F16                                       */
  int32_t layout;       /* MRLA_NCHW / MRLA_NHWC                                                */
  int32_t act;          /* MRLA_ACT_GELU: V = gelu(dwconv(x)) (deit_mrla_light.py:166-167)      */
  int32_t bn_mode;      /* MRLA_BN_*                                                           */
  int32_t residual;     /* 1: y = x + branch (resnet_mrla_light.py:116); 0: y = branch          */
  int32_t update_running; /* 1: BN train mode also updates running_mean/var in place            */
  int32_t fuse_relu_bwd;  /* backward only: x was produced as relu(z + o) in front of the tail
                             (resnet_mrla_light.py:113-114).  Then `dx` receives dz = dx_total*[x>0] and
                             `dout` receives the TOTAL identity gradient lam*dS + dz, which replaces the
                             reference's threshold_backward and gradient-accumulation passes.          */
  int32_t x_virtual;      /* 1: x = relu(z_coef.a*z + z_coef.b + o) is NEVER materialised (round 2): forward and backward
                             re-form it from the raw conv3 output `z`, `z_coef` and `o` inside every sweep; `x` may be NULL.
                             Needs z, z_coef, bn_mode TRAIN and (backward) fuse_relu_bwd = 1; only where
                             mrla_light_virtual_x() says so.  Backward then also fills `dz_sums`.                        */
  float eps, momentum;  /* BatchNorm2d eps / momentum                                          */
  int64_t bs_x, bs_o, bs_y, bs_dy, bs_dx, bs_do; /* batch strides (elements)                    */
  /* ---- forward tensors ---- */
  const void* x;            /* xt  [B,C,H,W]                                                    */
  const void* o;            /* ot_1 [B,C,H,W] or NULL (layer only: no lambda term)              */
  void* y;                  /* out [B,C,H,W]                                                    */
  const float* wq;          /* [k]                                                              */
  const float* wk;          /* [k]                                                              */
  const float* wv;          /* [C,3,3]                                                          */
  const float* lam;         /* [C] or NULL                                                      */
  const float* gamma;       /* [C] BN weight (NULL when bn_mode == NONE)                        */
  const float* beta;        /* [C] BN bias                                                      */
  float* running_mean;      /* [C] (read in eval, updated in train when update_running)         */
  float* running_var;       /* [C]                                                              */
  const float* drop_scale;  /* [B] DropPath scale m_b (0 or 1/keep) or NULL                     */
  /* ---- saved for backward (written by forward, read by backward), fp32 ---- */
  float* mom;   /* [6,B,C]: sum_hw of x, V, V^2, V*o, o, o^2                                     */
  float* gate;  /* [B,C/d] sigmoid gate                                                         */
  float* mean;  /* [C] BN batch (or running) mean actually used                                 */
  float* rstd;  /* [C] 1/sqrt(var+eps)                                                          */
  float* coef;  /* [3,B,C] forward per-(b,c) coefficients (scratch, not needed by backward)     */
  /* ---- backward tensors ---- */
  const void* dy;  /* [B,C,H,W] */
  void* dx;        /* [B,C,H,W] */
  void* dout;      /* [B,C,H,W] grad of o, or NULL */
  float* dwq;      /* [k]  (overwritten) */
  float* dwk;      /* [k]    */
  float* dwv;      /* [C,9]  */
  float* dlam;     /* [C] or NULL */
  float* dgamma;   /* [C] or NULL */
  float* dbeta;    /* [C] or NULL */
  /* ---- backward scratch, fp32 ---- */
  float* gmom;     /* [3,B,C] */
  float* bcoef;    /* [7,B,C] */
  float* scratch;  /* partial reductions; size from mrla_light_bwd_scratch_bytes() */
  size_t scratch_bytes;
  /* ---- optional producer fold (forward only) ---- */
  const void* z;   /* if non-NULL: pre-activation; forward first forms x = relu(z + o) (resnet_mrla_light.py:113-114)
                      and WRITES it to the buffer `x` points to (which the caller keeps for backward)            */
  int64_t bs_z;    /* batch stride of z (elements)                                                               */
  const float* z_coef; /* optional [2,C] (a_c, b_c): z is the RAW conv3 output and the bottleneck's bn3 apply
                      (resnet_mrla_light.py:101-102) is folded in front of the add: z' = round_dtype(a_c*z + b_c),
                      x = relu(z' + o).  Only where mrla_light_fwd_folds_bn() says so; NULL otherwise   */
  float* dz_sums;  /* backward with x_virtual: [2,C] sum_{b,h,w} dz and sum dz*z — the two reductions of bn3's backward
                      (resnet_mrla_light.py:101-102), accumulated in the sweep-B epilogue so that mrla_bn_backward
                      (MrlaBnArgs.sums) needs no reduction pass over dz and z.  NULL: not produced.                  */
} MrlaLightArgs;

int mrla_abi_version(void);

/* last CUDA error string / library build info (static storage). */
const char* mrla_build_info(void);

/* sizeof(MrlaLightArgs) as compiled — lets a foreign-language binding verify its struct mirror. */
size_t mrla_sizeof_light_args(void);

/* bytes of `scratch` mrla_light_backward needs for this problem. */
size_t mrla_light_bwd_scratch_bytes(const MrlaLightArgs* a);

/* 1 if mrla_light_backward honours `fuse_relu_bwd` for these arguments (layout / shape / alignment), else 0
 * (the caller then applies the ReLU mask and the identity-gradient sum itself). */
int mrla_light_bwd_fuses_relu(const MrlaLightArgs* a);

/* 1 if mrla_light_forward folds a per-channel affine on z (`z_coef`) for these arguments (TMA sweep-1 path:
 * NHWC, train-mode BN, o present, no GELU, aligned pointers), else 0 (the caller then applies bn3 itself). */
int mrla_light_fwd_folds_bn(const MrlaLightArgs* a);

/* 1 if forward AND backward of these arguments (z, z_coef, o, shapes, alignment as they will be passed) run the sweeps
 * that re-form x on the fly (`x_virtual`), else 0 (the caller then lets forward materialise x as before). */
int mrla_light_virtual_x(const MrlaLightArgs* a);

/* Host-only description of the launch plan the v7 sweeps would use for these arguments (no CUDA call; for tests and
 * diagnostics).  kind: 0 sweep 1, 1 sweep 2, 2 sweep A, 3 sweep B; xf: x re-formed from z (x_virtual).  Fills
 * out[12] = { eligible, CB, NQ, NT, U, TPU, stages, cpc, grid, threads, ctas_per_sm, smem_bytes } and returns `eligible`. */
int mrla_light_v7_plan(const MrlaLightArgs* a, int kind, int xf, int64_t out[12]);

/* y = residual*x + m_b*( BN( gate(x)*act(dwconv3x3(x)) + lambda*o ) ), plus saved statistics. */
int mrla_light_forward(const MrlaLightArgs* a, void* stream);

/* dx, dout and all parameter gradients of the same expression. */
int mrla_light_backward(const MrlaLightArgs* a, void* stream);


/* One MRLA-base block tail at cache depth t (this block is the t-th of its stage, 1-based).
 * The stage-scoped caches are caller-owned and written in place (no concatenation):
 *   v / dv     : slot j (0-based) of sample b starts at  base + j*ts + b*bs  (elements, activation dtype)
 *   kcache / dkcache : fp32 [B, t_cap, C]
 * forward writes slot t-1 of v and row t-1 of kcache; backward adds p_j*dS into dv slots 0..t-1 and
 * dlogit*q into dkcache rows 0..t-1 (accumulate=0: overwrite, used by the last block of the stage, whose
 * backward runs first), then consumes dv slot t-1 / dkcache row t-1 as the total gradient of v_t / k_t. */
typedef struct MrlaBaseArgs {
  int32_t B, C, H, W;
  int32_t dim_perhead, k_size, dtype, layout;
  int32_t t, t_cap;
  int32_t bn_mode;        /* MRLA_BN_* */
  int32_t relu;           /* 1: ReLU after BN (resnet_mrla_base.py:126); 0: none (base22 / DeiT) */
  int32_t residual, update_running;
  int32_t accumulate;     /* backward only */
  float eps, momentum;
  int64_t bs_x, bs_y, bs_s, bs_dy, bs_dx;
  int64_t bs_v, ts_v, bs_dv, ts_dv;
  const void* x;          /* [B,C,H,W] */
  void* v;                /* V cache base */
  void* s;                /* [B,C,H,W] attention output S (saved for backward) */
  void* y;                /* [B,C,H,W] */
  float* kcache;          /* [B,t_cap,C] */
  const float* wq;        /* [k] */
  const float* wk;        /* [k] */
  const float* wv;        /* [C,9] */
  const float* gamma;     /* [C] or NULL */
  const float* beta;      /* [C] or NULL */
  float* running_mean;
  float* running_var;
  const float* drop_scale;/* [B] or NULL */
  float* sx;              /* [B,C]   sum_hw x            (saved) */
  float* q;               /* [B,C]   query               (saved) */
  float* p;               /* [B,C/d,t] softmax weights   (saved) */
  float* smom;            /* [2,B,C] sum S, sum S^2      (scratch) */
  float* chan;            /* [4,C]   cA, cD, mean, rstd  (saved) */
  const void* dy;         /* [B,C,H,W] */
  void* dx;               /* [B,C,H,W] */
  void* dv;               /* dV cache base */
  float* dkcache;         /* [B,t_cap,C] */
  float* dwq;
  float* dwk;
  float* dwv;
  float* dgamma;
  float* dbeta;
  float* gmom;            /* [2,B,C] scratch */
  float* dpm;             /* [t,B,C] scratch */
  float* dyc;             /* [B,C]   scratch */
  float* scratch;
  size_t scratch_bytes;
} MrlaBaseArgs;

size_t mrla_sizeof_base_args(void);
size_t mrla_base_bwd_scratch_bytes(const MrlaBaseArgs* a);
/* y = residual*x + m_b * act( BN( sum_j softmax_j(q.K_j/sqrt(d)) V_j ) ), caches updated in place. */
int mrla_base_forward(const MrlaBaseArgs* a, void* stream);
int mrla_base_backward(const MrlaBaseArgs* a, void* stream);

/* Repack a dense NCHW activation [B,C,HW] (batch stride bs_src elements) into NHWC [B,HW,C] (batch stride bs_dst).
 * Used to promote NCHW callers onto the TMA (channels_last) kernels; replaces at::contiguous(channels_last). */
int mrla_nchw_to_nhwc(const void* src, void* dst, int B, int C, int HW, int dtype, int64_t bs_src, int64_t bs_dst,
                      void* stream);

/* x = relu(z + idt) over n contiguous elements (16-byte aligned, n multiple of the 16-byte vector width). */
int mrla_add_relu(const void* z, const void* idt, void* x, int64_t n, int dtype, void* stream);

/* Channels-last BatchNorm2d (+ optional ReLU): the bottleneck's bn1/bn2/bn3 and the stem BN on either side of the MRLA
 * tail.  Replaces at::batch_norm (+ at::relu) and their backward for NHWC activations
 * viewed as x[M = B*H*W, C]; PyTorch runs bf16 channels_last BN on its native (non-cuDNN) kernels. */
typedef struct MrlaBnArgs {
  int64_t M;               /* rows = B*H*W                                                        */
  int32_t C;               /* channels, multiple of 8, <= 2048                                    */
  int32_t dtype;           /* MRLA_F32 / MRLA_BF16 / MRLA_F16                                     */
  int32_t relu;            /* 1: y = relu(bn(x))                                                  */
  int32_t training;        /* 1: batch statistics; 0: running statistics                          */
  int32_t update_running;  /* 1: update running_mean / running_var in place (training only)       */
  int32_t stats_only;      /* forward: compute stats / coef / running statistics but do not write y (the consumer
                              applies a_c*x + b_c itself: MrlaLightArgs.z_coef)                                   */
  float eps, momentum;
  const void* x;           /* [M,C]                                                               */
  void* y;                 /* [M,C] (forward)                                                     */
  const float* gamma;      /* [C] or NULL                                                         */
  const float* beta;       /* [C] or NULL                                                         */
  float* running_mean;     /* [C]                                                                 */
  float* running_var;      /* [C]                                                                 */
  float* stats;            /* [2,C] mean, rstd actually used   (saved for backward)               */
  float* coef;             /* [2,C] a = gamma*rstd, b = beta - a*mean (saved for backward)        */
  const void* dy;          /* [M,C] (backward)                                                    */
  void* dx;                /* [M,C]                                                               */
  float* dgamma;           /* [C] or NULL                                                         */
  float* dbeta;            /* [C] or NULL                                                         */
  float* scratch;          /* mrla_bn_scratch_bytes()                                             */
  size_t scratch_bytes;
  const float* sums;       /* backward, optional: [2,C] precomputed sum dy, sum dy*x (MrlaLightArgs.dz_sums): the
                              reduction pass over dy and x is skipped (relu must be 0)                            */
} MrlaBnArgs;

size_t mrla_sizeof_bn_args(void);
size_t mrla_bn_scratch_bytes(const MrlaBnArgs* a);
int mrla_bn_forward(const MrlaBnArgs* a, void* stream);
int mrla_bn_backward(const MrlaBnArgs* a, void* stream);

/* 3x3 / stride 2 / pad 1 max pooling of a dense NHWC activation [B,H,W,C] -> [B,OH,OW,C], OH = (H-1)/2 + 1 (the
 * ResNet stem's nn.MaxPool2d: resnet/models/resnet_mrla_light.py `self.maxpool`; whole-step path of bench.py only,
 * SURVEY.md section 8f rank 3).  idx [B,OH,OW,C] bytes keeps the winning tap kh*3+kw; backward gathers from it.
 * Replaces at::max_pool2d_with_indices / max_pool2d_with_indices_backward (same tie and NaN rules).  C % 8 == 0;
 * pointers aligned to 8 elements. */
int mrla_maxpool3x3s2_forward(const void* x, void* y, unsigned char* idx, int B, int C, int H, int W, int dtype,
                              void* stream);
int mrla_maxpool3x3s2_backward(const void* dy, const unsigned char* idx, void* dx, int B, int C, int H, int W, int dtype,
                               void* stream);


/* Fused DeiT MRLA-light module (token layout), one kernel per direction, one CTA per sample (round 2):
 *   xn = LayerNorm_x(x), on = LayerNorm_o(o)      deit/deit_mrla_light.py:195-196
 *   out[:,0] = xn[:,0] ; out[:,1+t] = gate(img) * GELU(dwconv3x3(img))[t] + lambda * on[:,1+t], img = xn[:,1:] as [B,C,S,S]
 *                                                deit/deit_mrla_light.py:157-180,199-207
 * replacing both nn.LayerNorm calls, the cls split / torch.cat and the module's ATen sequence.  x, o, out, dout, dx, dox
 * are dense [B, n = S*S+1, C] in `dtype`; C % 64 == 0, C*n*sizeof(dtype) must fit one CTA's shared memory
 * (mrla_deit_light_supported()).  Parameters and gradients fp32. */
typedef struct MrlaDeitArgs {
  int32_t B, n, C, S;
  int32_t dim_perhead, k_size, dtype, reserved0;
  float eps;               /* LayerNorm eps (1e-6 in the reference) */
  float reserved1;
  const void* x;           /* [B,n,C] */
  const void* o;           /* [B,n,C] */
  void* out;               /* [B,n,C] */
  const float* normx_w; const float* normx_b; const float* normo_w; const float* normo_b;   /* [C] */
  const float* wq; const float* wk;   /* [k] */
  const float* wv;         /* [C,9] */
  const float* lam;        /* [C] */
  float* stats_x;          /* [B,n,2] mean, rstd of LN_x   (saved by forward) */
  float* stats_o;          /* [B,n,2] */
  float* gate;             /* [B,C/d] */
  const void* dout;        /* [B,n,C] (backward) */
  void* dx;                /* [B,n,C] */
  void* dox;               /* [B,n,C] */
  float* dparams;          /* [14*C + 2*k]: dWv[C,9] | dlam | dnormx_w | dnormx_b | dnormo_w | dnormo_b [C each] | dwq[k] | dwk[k] */
  float* scratch;          /* mrla_deit_light_scratch_bytes() : per-sample partials */
  size_t scratch_bytes;
} MrlaDeitArgs;

size_t mrla_sizeof_deit_args(void);
int mrla_deit_light_supported(const MrlaDeitArgs* a);          /* 1 if the fused kernels take this problem */
size_t mrla_deit_light_scratch_bytes(const MrlaDeitArgs* a);
int mrla_deit_light_forward(const MrlaDeitArgs* a, void* stream);
int mrla_deit_light_backward(const MrlaDeitArgs* a, void* stream);

/* Token LayerNorm over the last dimension of a dense [B, n, C] tensor (one warp per token), used by the DeiT MRLA-base module
 * (deit/deit_mrla_base.py:224-229, 242): forward also copies the normalised cls row (t = 0) to `cls_out[b*bs_cls + c]` (the
 * module output's row 0), backward takes d(xn) of token 0 from `g_cls` and of tokens t >= 1 from `g_img[b*bs_gimg + (t-1)*C + c]`
 * (the tail's image gradient), so neither torch.cat nor a gradient scatter is needed.  Replaces at::layer_norm (+ backward).
 * C even, C <= 768; dparams [2,C] = d(gamma), d(beta); scratch from mrla_layernorm_scratch_bytes(). */
typedef struct MrlaLnArgs {
  int32_t B, n, C, dtype;
  float eps, reserved0;
  const void* x; void* xn; void* cls_out; int64_t bs_cls;
  const float* gamma; const float* beta;
  float* stats;            /* [B,n,2] mean, rstd (saved) */
  const void* g_cls; int64_t bs_gcls; const void* g_img; int64_t bs_gimg;
  void* dx;
  float* dparams;          /* [2,C] */
  float* scratch; size_t scratch_bytes;
} MrlaLnArgs;
size_t mrla_sizeof_ln_args(void);
size_t mrla_layernorm_scratch_bytes(const MrlaLnArgs* a);
int mrla_layernorm_forward(const MrlaLnArgs* a, void* stream);
int mrla_layernorm_backward(const MrlaLnArgs* a, void* stream);

/* Number of kernel launches the last forward / backward call on this thread enqueued
 * (bench.py reports it as gpu_launches). */
int mrla_last_launch_count(void);

/* "file:line" inside the library where this thread's last MRLA_ERR_UNSUPPORTED was raised ("" if none): tells a
 * tensor-map refusal from a planner refusal.  Diagnostic only. */
const char* mrla_last_error_site(void);

#ifdef __cplusplus
}
#endif
#endif /* MRLA_B200_H_ */
