"""Autograd ops over the C ABI (explicit forward / backward, no autograd tracing of the math).

`light_tail` is the fused MRLA-light block tail

    y = residual*x + m_b * BN( gate(x) * act(dwconv3x3(x)) + lambda * o )

that replaces, in one op, the ATen sequence issued by
  resnet/models/modules/mrla_light_module.py:52-74   (mrla_light_layer.forward)
  resnet/models/resnet_mrla_light.py:40-43,116        (lambda recurrence, bn_mrla, drop_path, residual)
  deit/deit_mrla_light.py:157-180,204-206             (token layout, GELU on V, no BN)
(paths relative to /root/reference).  Every variant (layer only, module, eval-mode BN for the
mmdet backbone, DeiT) is the same op with different flags.
"""
from __future__ import annotations

import ctypes
import functools
from typing import NamedTuple, Optional

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


class LightCfg(NamedTuple):
    dim_perhead: int
    k_size: int
    act: int = _lib.ACT_NONE
    bn_mode: int = _lib.BN_NONE
    residual: bool = False
    update_running: bool = True
    eps: float = 1e-5
    momentum: float = 0.1
    fuse_add_relu: bool = False   # first tensor is z: x = relu(z + o) is formed inside the op (bottleneck :113-114)


# --------------------------------------------------------------------------------- layout helpers
def _layout_of(t: torch.Tensor):
    """Classify a logical [B,C,H,W] tensor: (layout, batch_stride) or None if it needs a copy."""
    B, C, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    dense_nchw = (sw == 1 or W == 1) and (sh == W or H == 1) and (sc == H * W or C == 1)
    if dense_nchw and (sb >= C * H * W or B == 1):
        return _lib.NCHW, (sb if B > 1 else C * H * W)
    dense_nhwc = (sc == 1 or C == 1) and (sw == C or W == 1) and (sh == W * C or H == 1)
    if dense_nhwc and (sb >= C * H * W or B == 1):
        return _lib.NHWC, (sb if B > 1 else C * H * W)
    return None


def _canon(t: torch.Tensor, want_layout: Optional[int] = None):
    """Return (tensor, layout, batch_stride) with the tensor in a layout the kernels read directly."""
    lay = _layout_of(t)
    if lay is not None and (want_layout is None or lay[0] == want_layout):
        return t, lay[0], lay[1]
    if want_layout == _lib.NHWC or (want_layout is None and t.is_contiguous(memory_format=torch.channels_last)):
        t = t.contiguous(memory_format=torch.channels_last)
        # channels_last .contiguous() may keep ambiguous strides for size-1 dims; rebuild explicitly
        B, C, H, W = t.shape
        if _layout_of(t) is None or _layout_of(t)[0] != _lib.NHWC:
            t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        return t, _lib.NHWC, C * H * W
    t = t.contiguous()
    B, C, H, W = t.shape
    return t, _lib.NCHW, C * H * W


def _empty_like_layout(ref: torch.Tensor, layout: int) -> torch.Tensor:
    B, C, H, W = ref.shape
    if layout == _lib.NHWC:
        return torch.empty((B, H, W, C), dtype=ref.dtype, device=ref.device).permute(0, 3, 1, 2)
    return torch.empty((B, C, H, W), dtype=ref.dtype, device=ref.device)


def _f32(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if p is None:
        return None
    p = p.detach().reshape(-1)
    if p.dtype != torch.float32:
        p = p.float()
    return p.contiguous()


def _like_param(g: torch.Tensor, meta) -> torch.Tensor:
    """Shape / dtype / stride a flat fp32 gradient like its parameter.  `model.to(channels_last)` gives the
    depthwise weight [C,1,3,3] strides (9,1,3,1): the memory order is that of the contiguous tensor (the
    permuted dim has size 1), so only the stride metadata is adjusted — DDP's bucket views then match."""
    shape, dtype, stride = meta
    g = g.reshape(shape).to(dtype)
    if g.stride() != stride and all(sz == 1 or a == b for sz, a, b in zip(shape, g.stride(), stride)):
        g = g.as_strided(shape, stride)
    return g


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev=None) -> int:
    """Raw handle of the current stream of `dev` (default: the current device; every op runs under `_on(dev)`)."""
    return torch.cuda.current_stream(dev).cuda_stream


def _on(t: torch.Tensor):
    """Device guard: kernels, TMA descriptors and scratch allocations must live on the tensor's device, not on whatever
    torch.cuda.current_device() happens to be (model.to('cuda:1') with current device 0)."""
    return torch.cuda.device(t.device)


def _guard(fn):
    """Run an autograd.Function forward / backward on the device (and its current stream) of the first CUDA tensor
    argument — everything inside (`_stream()`, scratch `torch.empty(device=...)`, tensor-map encoding, launches) then
    agrees with where the data lives."""
    @functools.wraps(fn)
    def wrapper(ctx, *args, **kw):
        t = next((a for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(ctx, *args, **kw)
        with torch.cuda.device(t.device):
            return fn(ctx, *args, **kw)
    return wrapper


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"mrla_b200: `{name}` must be a CUDA tensor — the MRLA kernels are sm_100a only "
                           "and there is no CPU fallback")
    if t.dtype not in _DTYPES:
        raise RuntimeError(f"mrla_b200: unsupported activation dtype {t.dtype} (fp32 / bf16 / fp16)")


# --------------------------------------------------------------------------------- profiling hook
class _Prof:
    """Optional CUDA-event timing of every C-ABI call (bench.py's live roofline)."""
    enabled = False
    records = []  # (tag, shape-key, start_event, end_event)

    @classmethod
    def begin(cls):
        if not cls.enabled:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    @classmethod
    def end(cls, tag, key, ev0):
        if ev0 is None:
            return
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        cls.records.append((tag, key, ev0, ev1))


launch_counter = {"fwd": 0, "bwd": 0}

# The TMA-pipelined kernels are NHWC.  A dense NCHW activation that is large enough to matter is therefore
# promoted to channels_last on first contact (one library copy per tensor); the result comes back channels_last,
# so everything downstream (cuDNN convs, the next tails) stays in the fast layout and pays nothing.  Small or
# TMA-ineligible NCHW inputs, and PROMOTE_NCHW = False, use the generic NCHW kernels directly.
PROMOTE_NCHW = True
PROMOTE_MIN_ELEMS = 1 << 20
# bn3-folded tails never materialise x where the library can re-form it in every sweep (tests switch this off to compare
# against the materialising kernels)
ALLOW_VIRTUAL_X = True


def _to_nhwc(t: torch.Tensor) -> torch.Tensor:
    """Dense NCHW -> channels_last copy through the library's tiled-transpose kernel."""
    B, C, H, W = t.shape
    out = torch.empty((B, H, W, C), dtype=t.dtype, device=t.device).permute(0, 3, 1, 2)
    L = _lib.lib()
    with _on(t):
        _lib.check(L.mrla_nchw_to_nhwc(t.data_ptr(), out.data_ptr(), B, C, H * W, _DTYPES[t.dtype], t.stride(0),
                                       C * H * W, _stream()), "mrla_nchw_to_nhwc")
    return out


class _ToNHWC(torch.autograd.Function):
    """Differentiable wrapper: the gradient simply flows back in whatever layout it arrives."""

    @staticmethod
    @_guard
    def forward(ctx, t):
        return _to_nhwc(t)

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, g):
        return g


def _want_nhwc(x: torch.Tensor) -> bool:
    B, C, H, W = x.shape
    # W > 56 runs the column-tiled v7 sweeps, which exist for 16-bit storage only
    wmax = 512 if x.dtype in (torch.bfloat16, torch.float16) else 56
    return PROMOTE_NCHW and x.numel() >= PROMOTE_MIN_ELEMS and C % 8 == 0 and W <= wmax and H * W > 1


def promote_images(x: torch.Tensor) -> torch.Tensor:
    """The image batch at the network input, made channels_last (one copy of a 3-channel tensor) when it arrives as a
    dense NCHW CUDA tensor: the reference's train.py never asks for channels_last (resnet/train.py:387), and with this the
    stem conv already produces NHWC, so every BatchNorm / pooling / tail kernel of the model runs its fast layout
    without touching train.py.  Values are unchanged; only strides differ."""
    if (PROMOTE_NCHW and x.is_cuda and x.dim() == 4 and x.shape[1] > 1
            and not x.is_contiguous(memory_format=torch.channels_last)):
        return x.contiguous(memory_format=torch.channels_last)
    return x


# --------------------------------------------------------------------------------- the fused op
class _LightTail(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x, o, wq, wk, wv, lam, gamma, beta, running_mean, running_var, drop_scale, cfg: LightCfg, out,
                z_coef=None, z_coef_fn=None):
        # z_coef / z_coef_fn are only used by _Bn3LightTail (which calls this method with its own context object):
        # z_coef_fn(folds) is called once the library has said whether it folds the bn3 affine on z, and returns the
        # [2,C] coefficients (folds) or None after applying bn3 itself into `x`'s storage
        _require_cuda(x, "x")
        L = _lib.lib()
        x_c, layout, bs_x = _canon(x)
        has_o = o is not None
        promote = layout == _lib.NCHW and has_o and out is None and _want_nhwc(x_c)
        if promote:
            x_c, layout, bs_x = _to_nhwc(x_c), _lib.NHWC, x_c[0].numel()
        B, C, H, W = x_c.shape
        if has_o:
            _require_cuda(o, "o")
            if o.shape != x.shape or o.dtype != x.dtype:
                raise RuntimeError("mrla_b200: o_prev must match x in shape and dtype")
            lay_o = _layout_of(o)
            if promote and lay_o is not None and lay_o[0] == _lib.NCHW and o.stride(1) == H * W:
                o_c, bs_o = _to_nhwc(o), C * H * W
            else:
                o_c, _, bs_o = _canon(o, layout)
        else:
            o_c, bs_o = None, 0
        ev = _Prof.begin()
        z_c = None
        if cfg.fuse_add_relu:
            if not has_o:
                raise RuntimeError("mrla_b200: fuse_add_relu needs o_prev (x = relu(z + o_prev))")
            n = C * H * W
            vec = 16 // x_c.element_size()
            if (bs_x == n and bs_o == n and (B * n) % vec == 0 and x_c.data_ptr() % 16 == 0
                    and o_c.data_ptr() % 16 == 0):
                # the library forms x = relu(z + o) itself: either on the fly in every sweep (x_virtual, decided below)
                # or inside sweep 1, which then writes it into a buffer allocated below
                z_c, x_c = x_c, None
            else:
                x_c, bs_x = _canon(torch.relu(x_c + o_c), layout)[0], n
        like = x_c if x_c is not None else z_c
        if out is not None:
            lay = _layout_of(out)
            if lay is None or lay[0] != layout or out.dtype != x.dtype or out.shape != x.shape:
                raise RuntimeError("mrla_b200: `out` buffer must have the layout / dtype / shape of x")
            y, bs_y = out, lay[1]
        else:
            y = _empty_like_layout(like, layout)
            bs_y = C * H * W
        dev = x.device
        g = C // cfg.dim_perhead
        f32 = dict(dtype=torch.float32, device=dev)
        mom = torch.empty((6, B, C), **f32)
        gate = torch.empty((B, g), **f32)
        stats = torch.empty((2, C), **f32)
        coef = torch.empty((3, B, C), **f32)
        wq32, wk32, wv32, lam32 = _f32(wq), _f32(wk), _f32(wv), _f32(lam)
        ga32, be32 = _f32(gamma), _f32(beta)
        ds32 = _f32(drop_scale)
        rm = running_mean if (running_mean is None or running_mean.dtype == torch.float32) else running_mean.float()
        rv = running_var if (running_var is None or running_var.dtype == torch.float32) else running_var.float()

        a = _lib.MrlaLightArgs()
        a.B, a.C, a.H, a.W = B, C, H, W
        a.dim_perhead, a.k_size = cfg.dim_perhead, cfg.k_size
        a.dtype, a.layout, a.act, a.bn_mode = _DTYPES[x.dtype], layout, cfg.act, cfg.bn_mode
        a.residual, a.update_running = int(cfg.residual), int(cfg.update_running and rm is not None)
        a.eps, a.momentum = cfg.eps, cfg.momentum
        a.bs_x, a.bs_o, a.bs_y = bs_x, bs_o, bs_y
        a.o, a.y = _ptr(o_c), _ptr(y)
        a.wq, a.wk, a.wv, a.lam = _ptr(wq32), _ptr(wk32), _ptr(wv32), _ptr(lam32)
        a.gamma, a.beta, a.running_mean, a.running_var = _ptr(ga32), _ptr(be32), _ptr(rm), _ptr(rv)
        a.drop_scale = _ptr(ds32)
        a.mom, a.gate, a.mean, a.rstd, a.coef = _ptr(mom), _ptr(gate), _ptr(stats[0]), _ptr(stats[1]), _ptr(coef)
        if z_c is not None:
            a.z, a.bs_z = _ptr(z_c), C * H * W
        # x is never materialised where the library re-forms it in every sweep (round 2)
        virtual = (z_c is not None and (z_coef is not None or z_coef_fn is not None) and ALLOW_VIRTUAL_X
                   and bool(L.mrla_light_virtual_x(ctypes.byref(a))))
        if z_c is not None and not virtual:
            x_c = _empty_like_layout(z_c, layout)
        a.x = _ptr(x_c)
        if z_coef_fn is not None:
            folds = virtual or (z_c is not None and bool(L.mrla_light_fwd_folds_bn(ctypes.byref(a))))
            z_coef = z_coef_fn(folds, z_c if z_c is not None else x_c)
            ctx.bn3_folded = z_coef is not None
        if z_coef is not None:
            a.z_coef = _ptr(z_coef)
        a.x_virtual = int(virtual)
        _lib.check(L.mrla_light_forward(ctypes.byref(a), _stream()), "mrla_light_forward")
        _Prof.end("light_fwd", (B, C, H, W, x.dtype, layout, bool(cfg.fuse_add_relu)), ev)
        launch_counter["fwd"] += L.mrla_last_launch_count()
        if rm is not None and rm is not running_mean and cfg.bn_mode == _lib.BN_TRAIN and cfg.update_running:
            running_mean.copy_(rm)
            running_var.copy_(rv)

        ctx.cfg, ctx.layout, ctx.has_o = cfg, layout, has_o
        ctx.bs = (bs_x, bs_o)
        ctx.param_meta = [(p.shape, p.dtype, p.stride()) if p is not None else None
                          for p in (wq, wk, wv, lam, gamma, beta)]
        ctx.virtual = virtual
        # virtual: the raw conv3 output (alive anyway as bn3's saved input) stands in for x, plus the [2,C] coefficients
        ctx.save_for_backward(z_c if virtual else x_c, o_c, wq32, wk32, wv32, lam32, ga32, ds32, mom, gate, stats,
                              z_coef if virtual else None)
        if out is not None:
            ctx.mark_dirty(out)
        return y

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dy):
        L = _lib.lib()
        x_c, o_c, wq32, wk32, wv32, lam32, ga32, ds32, mom, gate, stats, z_coef = ctx.saved_tensors
        cfg, layout = ctx.cfg, ctx.layout
        virtual = ctx.virtual
        B, C, H, W = x_c.shape
        if dy.dtype != x_c.dtype:
            dy = dy.to(x_c.dtype)
        lay_dy = _layout_of(dy)
        if layout == _lib.NHWC and lay_dy is not None and lay_dy[0] == _lib.NCHW and C % 2 == 0 and H * W > 1:
            dy_c, bs_dy = _to_nhwc(dy), C * H * W
        else:
            dy_c, _, bs_dy = _canon(dy, layout)
        dx = _empty_like_layout(x_c, layout)
        dout = _empty_like_layout(x_c, layout) if ctx.has_o else None
        dev = x_c.device
        f32 = dict(dtype=torch.float32, device=dev)
        k = cfg.k_size
        dwqk = torch.empty((2, k), **f32)
        dwv = torch.empty((C, 9), **f32)
        dch = torch.empty((3, C), **f32)  # dlam, dgamma, dbeta
        gmom = torch.empty((3, B, C), **f32)
        bcoef = torch.empty((7, B, C), **f32)

        a = _lib.MrlaLightArgs()
        a.B, a.C, a.H, a.W = B, C, H, W
        a.dim_perhead, a.k_size = cfg.dim_perhead, cfg.k_size
        a.dtype, a.layout, a.act, a.bn_mode = _DTYPES[x_c.dtype], layout, cfg.act, cfg.bn_mode
        a.residual, a.update_running = int(cfg.residual), 0
        a.eps, a.momentum = cfg.eps, cfg.momentum
        a.bs_x, a.bs_o = ctx.bs
        a.bs_dy, a.bs_dx, a.bs_do = bs_dy, C * H * W, C * H * W
        a.o = _ptr(o_c)
        dz_sums = None
        if virtual:
            dz_sums = torch.empty((2, C), **f32)
            a.z, a.bs_z, a.z_coef, a.x_virtual, a.dz_sums = _ptr(x_c), C * H * W, _ptr(z_coef), 1, _ptr(dz_sums)
        else:
            a.x = _ptr(x_c)
        ctx.dz_sums = dz_sums
        a.wq, a.wk, a.wv, a.lam, a.gamma = _ptr(wq32), _ptr(wk32), _ptr(wv32), _ptr(lam32), _ptr(ga32)
        a.drop_scale = _ptr(ds32)
        a.mom, a.gate, a.mean, a.rstd = _ptr(mom), _ptr(gate), _ptr(stats[0]), _ptr(stats[1])
        a.dy, a.dx, a.dout = _ptr(dy_c), _ptr(dx), _ptr(dout)
        a.dwq, a.dwk, a.dwv = _ptr(dwqk[0]), _ptr(dwqk[1]), _ptr(dwv)
        a.dlam = _ptr(dch[0]) if ctx.has_o else None
        has_bn = cfg.bn_mode != _lib.BN_NONE
        a.dgamma = _ptr(dch[1]) if has_bn else None
        a.dbeta = _ptr(dch[2]) if has_bn else None
        a.gmom, a.bcoef = _ptr(gmom), _ptr(bcoef)
        fused_epilogue = virtual or bool(cfg.fuse_add_relu and L.mrla_light_bwd_fuses_relu(ctypes.byref(a)))
        a.fuse_relu_bwd = int(fused_epilogue)
        nbytes = L.mrla_light_bwd_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        ev = _Prof.begin()
        _lib.check(L.mrla_light_backward(ctypes.byref(a), _stream()), "mrla_light_backward")
        _Prof.end("light_bwd", (B, C, H, W, x_c.dtype, layout, bool(cfg.fuse_add_relu)), ev)
        launch_counter["bwd"] += L.mrla_last_launch_count()
        if cfg.fuse_add_relu and not fused_epilogue:
            # shapes the ring kernel does not cover: ReLU mask + identity-gradient sum with library ops
            dx = dx * (x_c > 0).to(dx.dtype)
            dout = dout + dx

        def back(i, g):
            meta = ctx.param_meta[i]
            if meta is None or g is None:
                return None
            return _like_param(g, meta)

        grads = (dx, dout,
                 back(0, dwqk[0]), back(1, dwqk[1]), back(2, dwv),
                 back(3, dch[0]) if ctx.has_o else None,
                 back(4, dch[1]) if has_bn else None, back(5, dch[2]) if has_bn else None,
                 None, None, None, None, None, None, None)
        return grads


class _Ctx:
    """Stand-in for an autograd context so that one Function can run another's forward / backward inside its own."""

    def __init__(self, saved=()):
        self.saved_tensors = tuple(saved)

    def save_for_backward(self, *tensors):
        self.saved_tensors = tuple(tensors)

    def mark_dirty(self, *tensors):
        pass


class _Bn3LightTail(torch.autograd.Function):
    """bn3 + residual add + ReLU + MRLA-light tail of one bottleneck as ONE autograd node:
    `c3` is the raw conv3 output.  Forward: BatchNorm statistics of c3 (mrla_bn_forward, stats_only) -> sweep 1 applies
    a_c*c3 + b_c, adds the identity, ReLUs, stores x and takes the moments (MODE 6) -> mid -> sweep 2.  Backward: the
    tail's backward leaves dz (= gradient of the bn3 output) and the total identity gradient, then mrla_bn_backward turns
    dz into d(c3), d(gamma3), d(beta3).  Replaces resnet_mrla_light.py:101-102 + :113-116.  Only for tails the library
    folds (bn3_tail_eligible); every other case keeps bn3 as its own op (ops.bn_act) in front of light_tail."""

    @staticmethod
    @_guard
    def forward(ctx, c3, o, w3, b3, rm3, rv3, bn3_args, wq, wk, wv, lam, gamma, beta, running_mean, running_var,
                drop_scale, cfg: LightCfg):
        L = _lib.lib()
        training3, update3, momentum3, eps3 = bn3_args
        B, C, H, W = c3.shape
        f32 = dict(dtype=torch.float32, device=c3.device)
        stats3 = torch.empty((2, C), **f32)
        coef3 = torch.empty((2, C), **f32)
        w3_32, b3_32 = _f32(w3), _f32(b3)
        keep = {}

        def bn3(folds, src):
            # src: the dense NHWC buffer holding c3; statistics only — sweep 1 applies the affine
            if not folds:
                raise RuntimeError("mrla_b200: bn3 fold requested for a tail the library does not fold "
                                   "(callers check bn3_tail_eligible first)")
            a = _lib.MrlaBnArgs()
            a.M, a.C, a.dtype = B * H * W, C, _DTYPES[src.dtype]
            a.relu, a.training, a.update_running = 0, int(training3), int(update3)
            a.stats_only = int(folds)
            a.eps, a.momentum = eps3, momentum3
            a.x, a.gamma, a.beta = _ptr(src), _ptr(w3_32), _ptr(b3_32)
            a.running_mean, a.running_var = _ptr(rm3), _ptr(rv3)
            a.stats, a.coef = _ptr(stats3), _ptr(coef3)
            nbytes = L.mrla_bn_scratch_bytes(ctypes.byref(a))
            scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
            a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
            _lib.check(L.mrla_bn_forward(ctypes.byref(a), _stream()), "mrla_bn_forward")
            launch_counter["fwd"] += L.mrla_last_launch_count()
            keep["src"] = src
            return coef3

        inner = _Ctx()
        y = _LightTail.forward(inner, c3, o, wq, wk, wv, lam, gamma, beta, running_mean, running_var, drop_scale, cfg,
                               None, None, bn3)
        ctx.inner = {k: v for k, v in inner.__dict__.items() if k != "saved_tensors"}
        ctx.n_inner = len(inner.saved_tensors)
        ctx.bn3_meta = (bool(training3), eps3, momentum3)
        ctx.bn3_param_meta = [(p.shape, p.dtype, p.stride()) if p is not None else None for p in (w3, b3)]
        ctx.save_for_backward(*inner.saved_tensors, keep["src"], w3_32, stats3, coef3)
        return y

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dy):
        L = _lib.lib()
        saved = ctx.saved_tensors
        inner = _Ctx(saved[:ctx.n_inner])
        inner.__dict__.update(ctx.inner)
        g = _LightTail.backward(inner, dy)
        dz, d_id = g[0], g[1]
        c3, w3_32, stats3, coef3 = saved[ctx.n_inner:]
        training3, eps3, momentum3 = ctx.bn3_meta
        B, C, H, W = c3.shape
        f32 = dict(dtype=torch.float32, device=c3.device)
        dc3 = _empty_like_layout(c3, _lib.NHWC)
        dgb = torch.empty((2, C), **f32)
        a = _lib.MrlaBnArgs()
        a.M, a.C, a.dtype = B * H * W, C, _DTYPES[c3.dtype]
        a.relu, a.training = 0, int(training3)
        a.eps, a.momentum = eps3, momentum3
        a.x, a.gamma, a.stats, a.coef = _ptr(c3), _ptr(w3_32), _ptr(stats3), _ptr(coef3)
        a.dy, a.dx, a.dgamma, a.dbeta = _ptr(dz), _ptr(dc3), _ptr(dgb[0]), _ptr(dgb[1])
        a.sums = _ptr(getattr(inner, "dz_sums", None))   # sweep B already reduced sum dz, sum dz*c3 (x_virtual path)
        nbytes = L.mrla_bn_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        _lib.check(L.mrla_bn_backward(ctypes.byref(a), _stream()), "mrla_bn_backward")
        launch_counter["bwd"] += L.mrla_last_launch_count()
        dw3 = _like_param(dgb[0], ctx.bn3_param_meta[0]) if ctx.bn3_param_meta[0] is not None else None
        db3 = _like_param(dgb[1], ctx.bn3_param_meta[1]) if ctx.bn3_param_meta[1] is not None else None
        return (dc3, d_id, dw3, db3, None, None, None) + tuple(g[2:8]) + (None, None, None, None)


def bn3_tail_eligible(c3: torch.Tensor, o: torch.Tensor, bn3: "torch.nn.BatchNorm2d", cfg: LightCfg) -> bool:
    """Can bn3 be folded into the tail op for these tensors?  (Asks the library's own planner through
    mrla_light_fwd_folds_bn with the layout / strides / alignment the op would use.)"""
    if not (cfg.fuse_add_relu and cfg.bn_mode == _lib.BN_TRAIN and cfg.act == _lib.ACT_NONE):
        return False
    if not is_plain_batchnorm(bn3):
        return False
    if not (bn3.training or bn3.running_mean is None) or bn3.weight is None or bn3.bias is None:
        return False
    if bn3.running_mean is not None and bn3.running_mean.dtype != torch.float32:
        return False
    if not bn_act_eligible(c3) or o is None or o.shape != c3.shape or o.dtype != c3.dtype or not o.is_cuda:
        return False
    lay_o = _layout_of(o)
    B, C, H, W = c3.shape
    n = C * H * W
    if lay_o is None or lay_o[0] != _lib.NHWC or lay_o[1] != n or o.data_ptr() % 16:
        return False
    a = _lib.MrlaLightArgs()
    a.B, a.C, a.H, a.W = B, C, H, W
    a.dim_perhead, a.k_size = cfg.dim_perhead, cfg.k_size
    a.dtype, a.layout, a.act, a.bn_mode = _DTYPES[c3.dtype], _lib.NHWC, cfg.act, cfg.bn_mode
    a.bs_x = a.bs_o = a.bs_y = a.bs_z = n
    a.x, a.o, a.z = _ptr(c3), _ptr(o), _ptr(c3)   # x / y buffers are fresh allocations with the same alignment
    L = _lib.lib()
    # the round-1 materialising kernels (W <= 56), or the v7 sweeps that never materialise x (column-tiled, W <= 512)
    return bool(L.mrla_light_fwd_folds_bn(ctypes.byref(a))) or (ALLOW_VIRTUAL_X and bool(L.mrla_light_virtual_x(ctypes.byref(a))))


def bn3_light_tail(c3: torch.Tensor, o: torch.Tensor, bn3: "torch.nn.BatchNorm2d", wq, wk, wv, lam, gamma, beta,
                   running_mean, running_var, drop_scale, *, cfg: LightCfg) -> torch.Tensor:
    """`tail(bn3(c3), o)` with bn3's apply folded into sweep 1; callers check bn3_tail_eligible() first."""
    update = bn3.training and bn3.track_running_stats and bn3.running_mean is not None
    if update:
        bn3.num_batches_tracked += 1
        momentum = effective_momentum(bn3)
    else:
        momentum = 0.0
    return _Bn3LightTail.apply(c3, o, bn3.weight, bn3.bias, bn3.running_mean, bn3.running_var,
                               (True, update, momentum, bn3.eps), wq, wk, wv, lam, gamma, beta, running_mean, running_var,
                               drop_scale, cfg)


def light_tail(x: torch.Tensor, o: Optional[torch.Tensor], wq, wk, wv, lam=None, gamma=None, beta=None,
               running_mean=None, running_var=None, drop_scale=None, *, cfg: LightCfg, out=None,
               z_coef: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused MRLA-light tail (see module docstring).  `x`/`o` are logical [B,C,H,W] tensors, either
    NCHW-contiguous or channels-last (any batch stride); the result has the layout of `x`.

    `z_coef` ([2,C] fp32, with cfg.fuse_add_relu): constant per-channel affine applied to `x` (= z) in front of the add,
    as bn3_light_tail does with the bn3 coefficients; no gradient flows to it (kernel-level benchmarks use it to time
    the sweep-1 variant the model runs)."""
    return _LightTail.apply(x, o, wq, wk, wv, lam, gamma, beta, running_mean, running_var, drop_scale, cfg, out,
                            z_coef, None)


# ===================================================================================== DeiT fused module
class _DeitLightModule(torch.autograd.Function):
    """deit/deit_mrla_light.py:194-209 (`mrlal_module.forward`) as ONE kernel per direction: both LayerNorms, cls
    pass-through, GAP + ECA gate, depthwise conv + GELU, lambda recurrence (csrc/deit_fused.cuh)."""

    @staticmethod
    def forward(ctx, x, o, nx_w, nx_b, no_w, no_b, wq, wk, wv, lam, dim_perhead, k_size, eps):
        L = _lib.lib()
        B, n, C = x.shape
        S = int(round((n - 1) ** 0.5))
        out = torch.empty_like(x)
        f32 = dict(dtype=torch.float32, device=x.device)
        stats = torch.empty((2, B, n, 2), **f32)
        gate = torch.empty((B, C // dim_perhead), **f32)
        P = [_f32(t) for t in (nx_w, nx_b, no_w, no_b, wq, wk, wv, lam)]
        a = _lib.MrlaDeitArgs()
        a.B, a.n, a.C, a.S = B, n, C, S
        a.dim_perhead, a.k_size, a.dtype, a.eps = dim_perhead, k_size, _DTYPES[x.dtype], eps
        a.x, a.o, a.out = _ptr(x), _ptr(o), _ptr(out)
        (a.normx_w, a.normx_b, a.normo_w, a.normo_b, a.wq, a.wk, a.wv, a.lam) = [_ptr(t) for t in P]
        a.stats_x, a.stats_o, a.gate = _ptr(stats[0]), _ptr(stats[1]), _ptr(gate)
        _lib.check(L.mrla_deit_light_forward(ctypes.byref(a), _stream()), "mrla_deit_light_forward")
        launch_counter["fwd"] += L.mrla_last_launch_count()
        ctx.meta = (dim_perhead, k_size, eps)
        ctx.param_meta = [(p.shape, p.dtype, p.stride()) for p in (nx_w, nx_b, no_w, no_b, wq, wk, wv, lam)]
        ctx.save_for_backward(x, o, stats, gate, *P)
        return out

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dout):
        L = _lib.lib()
        x, o, stats, gate, *P = ctx.saved_tensors
        dim_perhead, k_size, eps = ctx.meta
        B, n, C = x.shape
        dout = dout.contiguous()
        if dout.dtype != x.dtype:
            dout = dout.to(x.dtype)
        dx, do = torch.empty_like(x), torch.empty_like(o)
        f32 = dict(dtype=torch.float32, device=x.device)
        dpar = torch.empty(14 * C + 2 * k_size, **f32)
        a = _lib.MrlaDeitArgs()
        a.B, a.n, a.C, a.S = B, n, C, int(round((n - 1) ** 0.5))
        a.dim_perhead, a.k_size, a.dtype, a.eps = dim_perhead, k_size, _DTYPES[x.dtype], eps
        a.x, a.o = _ptr(x), _ptr(o)
        (a.normx_w, a.normx_b, a.normo_w, a.normo_b, a.wq, a.wk, a.wv, a.lam) = [_ptr(t) for t in P]
        a.stats_x, a.stats_o, a.gate = _ptr(stats[0]), _ptr(stats[1]), _ptr(gate)
        a.dout, a.dx, a.dox, a.dparams = _ptr(dout), _ptr(dx), _ptr(do), _ptr(dpar)
        nbytes = L.mrla_deit_light_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        _lib.check(L.mrla_deit_light_backward(ctypes.byref(a), _stream()), "mrla_deit_light_backward")
        launch_counter["bwd"] += L.mrla_last_launch_count()
        k = k_size
        seg = [dpar[9 * C + 1 * C:9 * C + 2 * C], dpar[9 * C + 2 * C:9 * C + 3 * C], dpar[9 * C + 3 * C:9 * C + 4 * C],
               dpar[9 * C + 4 * C:9 * C + 5 * C], dpar[14 * C:14 * C + k], dpar[14 * C + k:14 * C + 2 * k], dpar[:9 * C],
               dpar[9 * C:10 * C]]   # order of the forward arguments: nx_w nx_b no_w no_b wq wk wv lam
        grads = [_like_param(g_, m_) for g_, m_ in zip(seg, ctx.param_meta)]
        return (dx, do, *grads, None, None, None)


def deit_light_supported(x: torch.Tensor, dim_perhead: int, k_size: int) -> bool:
    """Does the fused DeiT kernel take this token tensor ([B, S*S+1, C] dense, C % 64 == 0, one sample fits a CTA)?"""
    if not (x.is_cuda and x.dim() == 3 and x.is_contiguous() and x.dtype in _DTYPES):
        return False
    B, n, C = x.shape
    S = int(round((n - 1) ** 0.5))
    if S * S + 1 != n:
        return False
    a = _lib.MrlaDeitArgs()
    a.B, a.n, a.C, a.S = B, n, C, S
    a.dim_perhead, a.k_size, a.dtype = dim_perhead, k_size, _DTYPES[x.dtype]
    return bool(_lib.lib().mrla_deit_light_supported(ctypes.byref(a)))


def deit_light_module(x, o, nx_w, nx_b, no_w, no_b, wq, wk, wv, lam, *, dim_perhead: int, k_size: int, eps: float):
    return _DeitLightModule.apply(x, o, nx_w, nx_b, no_w, no_b, wq, wk, wv, lam, dim_perhead, k_size, eps)


# ===================================================================================== MRLA-base
class BaseCfg(NamedTuple):
    dim_perhead: int
    k_size: int
    bn_mode: int = _lib.BN_NONE
    relu: bool = False
    residual: bool = False
    update_running: bool = True
    eps: float = 1e-5
    momentum: float = 0.1


class StageCache:
    """Stage-scoped K/V cache of MRLA-base, written IN PLACE.

    The reference grows K [B,t,C] and V [B,t,C,H,W] with `torch.cat` in every block
    (resnet/models/modules/mrla_base_module.py:65-70), re-copying the whole cache each time (O(T^2) traffic per
    stage) and making autograd slice it back apart.  Here one buffer per stage holds the slots
    (`v[j]` is a dense [B,C,H,W] activation in the layout of x, `k` is fp32 [B,T,C]); gradients w.r.t. the cached
    slots are accumulated in place in `dv` / `dk` by the later blocks of the stage (their backward runs first
    because the blocks are chained through the residual stream) and consumed by the block that produced the slot.
    """

    def __init__(self, like: torch.Tensor, layout: int, cap: int):
        B, C, H, W = like.shape
        self.shape, self.layout, self.dtype, self.device = (B, C, H, W), layout, like.dtype, like.device
        self.cap = max(int(cap), 1)
        self.t = 0
        self.v = self._alloc_v(self.cap)
        self.k = torch.empty((B, self.cap, C), dtype=torch.float32, device=like.device)
        self.dv = None
        self.dk = None
        self.bwd_started = False
        self.n_ext = 0  # slots that came from foreign prev_K / prev_V tensors

    def _alloc_v(self, cap):
        B, C, H, W = self.shape
        shape = (cap, B, H, W, C) if self.layout == _lib.NHWC else (cap, B, C, H, W)
        return torch.empty(shape, dtype=self.dtype, device=self.device)

    def reserve(self, t: int):
        if t <= self.cap:
            return
        cap = max(t, 2 * self.cap)
        v = self._alloc_v(cap)
        v[: self.cap].copy_(self.v)
        k = torch.empty((self.shape[0], cap, self.shape[1]), dtype=torch.float32, device=self.device)
        k[:, : self.cap].copy_(self.k)
        self.v, self.k, self.cap = v, k, cap

    def ensure_grads(self):
        if self.dv is None:
            self.dv = torch.empty_like(self.v)
            self.dk = torch.empty_like(self.k)

    # reference-shaped views ------------------------------------------------------------------
    def K_view(self, t):
        kv = self.k[:, :t]
        return kv if self.dtype == torch.float32 else kv.to(self.dtype)

    def V_view(self, t):
        v = self.v[:t]
        if self.layout == _lib.NHWC:
            v = v.permute(0, 1, 4, 2, 3)  # [t,B,C,H,W]
        return v.transpose(0, 1)         # [B,t,C,H,W]


class _BaseTail(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x, wq, wk, wv, gamma, beta, ext_k, ext_v, token, running_mean, running_var, drop_scale, cache,
                t, cfg: BaseCfg, out):
        _require_cuda(x, "x")
        L = _lib.lib()
        x_c, layout, bs_x = _canon(x, cache.layout)
        B, C, H, W = x_c.shape
        if out is not None:
            lay = _layout_of(out)
            if lay is None or lay[0] != layout or out.dtype != x.dtype or out.shape != x.shape:
                raise RuntimeError("mrla_b200: `out` buffer must have the layout / dtype / shape of x")
            y, bs_y = out, lay[1]
        else:
            y, bs_y = _empty_like_layout(x_c, layout), C * H * W
        s = _empty_like_layout(x_c, layout)
        g = C // cfg.dim_perhead
        f32 = dict(dtype=torch.float32, device=x.device)
        sxq = torch.empty((2, B, C), **f32)
        p = torch.empty((B, g, t), **f32)
        smom = torch.empty((2, B, C), **f32)
        chan = torch.empty((4, C), **f32)
        wq32, wk32, wv32, ga32, be32, ds32 = _f32(wq), _f32(wk), _f32(wv), _f32(gamma), _f32(beta), _f32(drop_scale)
        rm = running_mean if (running_mean is None or running_mean.dtype == torch.float32) else running_mean.float()
        rv = running_var if (running_var is None or running_var.dtype == torch.float32) else running_var.float()
        N = C * H * W
        a = _lib.MrlaBaseArgs()
        a.B, a.C, a.H, a.W = B, C, H, W
        a.dim_perhead, a.k_size, a.dtype, a.layout = cfg.dim_perhead, cfg.k_size, _DTYPES[x.dtype], layout
        a.t, a.t_cap, a.bn_mode, a.relu = t, cache.cap, cfg.bn_mode, int(cfg.relu)
        a.residual, a.update_running = int(cfg.residual), int(cfg.update_running and rm is not None)
        a.eps, a.momentum = cfg.eps, cfg.momentum
        a.bs_x, a.bs_y, a.bs_s = bs_x, bs_y, N
        a.bs_v, a.ts_v = N, B * N
        a.x, a.v, a.s, a.y, a.kcache = _ptr(x_c), _ptr(cache.v), _ptr(s), _ptr(y), _ptr(cache.k)
        a.wq, a.wk, a.wv, a.gamma, a.beta = _ptr(wq32), _ptr(wk32), _ptr(wv32), _ptr(ga32), _ptr(be32)
        a.running_mean, a.running_var, a.drop_scale = _ptr(rm), _ptr(rv), _ptr(ds32)
        a.sx, a.q, a.p, a.smom, a.chan = _ptr(sxq[0]), _ptr(sxq[1]), _ptr(p), _ptr(smom), _ptr(chan)
        ev = _Prof.begin()
        _lib.check(L.mrla_base_forward(ctypes.byref(a), _stream()), "mrla_base_forward")
        _Prof.end("base_fwd", (B, C, H, W, x.dtype, layout), ev)
        launch_counter["fwd"] += L.mrla_last_launch_count()
        if rm is not None and rm is not running_mean and cfg.bn_mode == _lib.BN_TRAIN and cfg.update_running:
            running_mean.copy_(rm)
            running_var.copy_(rv)
        ctx.cfg, ctx.layout, ctx.cache, ctx.t, ctx.bs_x = cfg, layout, cache, t, bs_x
        ctx.has_ext = ext_k is not None
        ctx.param_meta = [(q_.shape, q_.dtype, q_.stride()) if q_ is not None else None
                          for q_ in (wq, wk, wv, gamma, beta)]
        ctx.ext_meta = (ext_k.dtype, ext_v.dtype) if ctx.has_ext else None
        ctx.save_for_backward(x_c, s, wq32, wk32, wv32, ga32, ds32, sxq, p, chan)
        ctx.has_token = token is not None
        if out is not None:
            ctx.mark_dirty(out)
        # `token` chains the blocks of a stage in the autograd graph (block t consumes the token block t-1
        # emitted), so the backward of block t is guaranteed to run before that of block t-1 even when the
        # activations themselves are not chained — the in-place dV / dK accumulation relies on that order.
        return y, torch.zeros((), dtype=torch.float32, device=x.device)

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dy, _dtoken):
        L = _lib.lib()
        x_c, s, wq32, wk32, wv32, ga32, ds32, sxq, p, chan = ctx.saved_tensors
        cfg, layout, cache, t = ctx.cfg, ctx.layout, ctx.cache, ctx.t
        B, C, H, W = x_c.shape
        N = C * H * W
        dy_c, _, bs_dy = _canon(dy, layout)
        if dy_c.dtype != x_c.dtype:
            dy_c = dy_c.to(x_c.dtype)
        cache.ensure_grads()
        dx = _empty_like_layout(x_c, layout)
        f32 = dict(dtype=torch.float32, device=x_c.device)
        k = cfg.k_size
        dwqk = torch.empty((2, k), **f32)
        dwv = torch.empty((C, 9), **f32)
        dch = torch.empty((2, C), **f32)
        gmom = torch.empty((2, B, C), **f32)
        dpm = torch.empty((t, B, C), **f32)
        dyc = torch.empty((B, C), **f32)
        a = _lib.MrlaBaseArgs()
        a.B, a.C, a.H, a.W = B, C, H, W
        a.dim_perhead, a.k_size, a.dtype, a.layout = cfg.dim_perhead, cfg.k_size, _DTYPES[x_c.dtype], layout
        a.t, a.t_cap, a.bn_mode, a.relu = t, cache.cap, cfg.bn_mode, int(cfg.relu)
        a.residual, a.update_running, a.accumulate = int(cfg.residual), 0, int(cache.bwd_started)
        a.eps, a.momentum = cfg.eps, cfg.momentum
        a.bs_x, a.bs_s, a.bs_dy, a.bs_dx = ctx.bs_x, N, bs_dy, N
        a.bs_v, a.ts_v, a.bs_dv, a.ts_dv = N, B * N, N, B * N
        a.x, a.v, a.s, a.kcache = _ptr(x_c), _ptr(cache.v), _ptr(s), _ptr(cache.k)
        a.wq, a.wk, a.wv, a.gamma, a.drop_scale = _ptr(wq32), _ptr(wk32), _ptr(wv32), _ptr(ga32), _ptr(ds32)
        a.sx, a.q, a.p, a.chan = _ptr(sxq[0]), _ptr(sxq[1]), _ptr(p), _ptr(chan)
        a.dy, a.dx, a.dv, a.dkcache = _ptr(dy_c), _ptr(dx), _ptr(cache.dv), _ptr(cache.dk)
        a.dwq, a.dwk, a.dwv = _ptr(dwqk[0]), _ptr(dwqk[1]), _ptr(dwv)
        has_bn = cfg.bn_mode != _lib.BN_NONE
        a.dgamma = _ptr(dch[0]) if has_bn else None
        a.dbeta = _ptr(dch[1]) if has_bn else None
        a.gmom, a.dpm, a.dyc = _ptr(gmom), _ptr(dpm), _ptr(dyc)
        nbytes = L.mrla_base_bwd_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        ev = _Prof.begin()
        _lib.check(L.mrla_base_backward(ctypes.byref(a), _stream()), "mrla_base_backward")
        _Prof.end("base_bwd", (B, C, H, W, x_c.dtype, layout), ev)
        launch_counter["bwd"] += L.mrla_last_launch_count()
        # the blocks of a stage run their backward last-to-first; once the block that opened the cache (t = n_ext + 1) is
        # through, a later backward over the same graph (retain_graph=True) starts from overwritten, not stale, dV / dK
        cache.bwd_started = t > cache.n_ext + 1
        def back(i, g_):
            meta = ctx.param_meta[i]
            return None if meta is None or g_ is None else _like_param(g_, meta)

        dk_ext = dv_ext = None
        if ctx.has_ext:
            n = cache.n_ext
            dk_ext = cache.dk[:, :n].to(ctx.ext_meta[0])
            dvx = cache.dv[:n]
            if layout == _lib.NHWC:
                dvx = dvx.permute(0, 1, 4, 2, 3)
            dv_ext = dvx.transpose(0, 1).to(ctx.ext_meta[1])
        dtok = torch.zeros((), dtype=torch.float32, device=x_c.device) if ctx.has_token else None
        return (dx, back(0, dwqk[0]), back(1, dwqk[1]), back(2, dwv), back(3, dch[0]) if has_bn else None,
                back(4, dch[1]) if has_bn else None, dk_ext, dv_ext, dtok, None, None, None, None, None, None, None)


def base_tail(x, prev_k, prev_v, wq, wk, wv, gamma=None, beta=None, running_mean=None, running_var=None,
              drop_scale=None, *, init_cell: bool, cfg: BaseCfg, cap_hint: int = 8, out=None):
    """Fused MRLA-base tail.  Returns (y, K, V) with K [B,t,C] / V [B,t,C,H,W] views of the in-place stage cache
    (reference signature: mrla_base_module.py:54-89 `forward(x, prev_K, prev_V) -> (out, K, V)`).

    K and V are handles for the NEXT block of the stage: their gradients flow through the stage cache (dV / dK are
    accumulated in place by the later blocks and consumed by the block that produced the slot), not through an autograd
    edge of these views.  Feeding K / V to anything other than the next `base_tail` of the same stage (an auxiliary loss
    on the keys, say) gets no gradient — the reference's `torch.cat` results would; no MRLA model does that."""
    _require_cuda(x, "x")
    x_c, layout, _ = _canon(x)
    if layout == _lib.NCHW and out is None and _want_nhwc(x_c):
        prev_cache = None if (init_cell or prev_k is None) else getattr(prev_k, "_mrla_cache", None)
        if prev_cache is None or prev_cache.layout == _lib.NHWC:
            x_c, layout = _ToNHWC.apply(x_c), _lib.NHWC   # promote onto the NHWC fast path (see PROMOTE_NCHW)
    ext_k = ext_v = None
    if init_cell or prev_k is None:
        cache = StageCache(x_c, layout, cap_hint)
    else:
        cache = getattr(prev_k, "_mrla_cache", None)
        if cache is None or cache.shape != tuple(x_c.shape) or cache.dtype != x_c.dtype or cache.layout != layout:
            # foreign tensors (not produced by this op): import them into a fresh cache; their gradients are
            # returned through the autograd edge of this call
            n = prev_k.shape[1]
            cache = StageCache(x_c, layout, max(cap_hint, n + 1))
            cache.k[:, :n].copy_(prev_k.detach().float())
            pv = prev_v.detach().transpose(0, 1)  # [n,B,C,H,W]
            if layout == _lib.NHWC:
                pv = pv.permute(0, 1, 3, 4, 2)
            cache.v[:n].copy_(pv)
            cache.t = cache.n_ext = n
            ext_k, ext_v = prev_k, prev_v
    t = cache.t + 1
    cache.reserve(t)
    prev_token = getattr(prev_k, "_mrla_token", None) if (prev_k is not None and ext_k is None) else None
    y, token = _BaseTail.apply(x_c, wq, wk, wv, gamma, beta, ext_k, ext_v, prev_token, running_mean, running_var,
                               drop_scale, cache, t, cfg, out)
    cache.t = t
    K, V = cache.K_view(t), cache.V_view(t)
    # The cache and the ordering token ride on the returned K tensor (K -> cache, K -> token -> autograd node ->
    # cache).  The cache itself holds no reference back, so nothing forms a cycle and the multi-GB slot buffers are
    # released as soon as the graph and the K/V views die.
    K._mrla_cache = cache
    K._mrla_token = token if token.requires_grad else None
    return y, K, V


# ===================================================================================== DeiT MRLA-base module
def _tok_img(tok: torch.Tensor) -> torch.Tensor:
    """[B, S*S, C] token slice -> logical [B, C, S, S] view with channels-last strides (no copy)."""
    b, m, c = tok.shape
    s = int(round(m ** 0.5))
    return tok.as_strided((b, c, s, s), (tok.stride(0), 1, s * tok.stride(1), tok.stride(1)), tok.storage_offset())


class _DeitBaseModule(torch.autograd.Function):
    """deit/deit_mrla_base.py:224-243 (`mrlab_module.forward`) as ONE autograd node without library calls: token LayerNorm
    kernel (writes xn and the cls row of the output) -> MRLA-base tail kernels on the token image, writing straight into
    the output's image rows -> (backward) tail backward, then LayerNorm backward fed from the cls row of d_out and the
    tail's image gradient.  No at::layer_norm, no permute / reshape / cat copies, no gradient scatter for the views."""

    @staticmethod
    @_guard
    def forward(ctx, xt, nx_w, nx_b, wq, wk, wv, token, cache, t, cfg, eps):
        L = _lib.lib()
        B, n, C = xt.shape
        xn, out = torch.empty_like(xt), torch.empty_like(xt)
        stats = torch.empty((B, n, 2), dtype=torch.float32, device=xt.device)
        g32, b32 = _f32(nx_w), _f32(nx_b)
        a = _lib.MrlaLnArgs()
        a.B, a.n, a.C, a.dtype, a.eps = B, n, C, _DTYPES[xt.dtype], eps
        a.x, a.xn, a.cls_out, a.bs_cls = _ptr(xt), _ptr(xn), _ptr(out), n * C
        a.gamma, a.beta, a.stats = _ptr(g32), _ptr(b32), _ptr(stats)
        _lib.check(L.mrla_layernorm_forward(ctypes.byref(a), _stream()), "mrla_layernorm_forward")
        launch_counter["fwd"] += L.mrla_last_launch_count()
        inner = _Ctx()
        _, tok = _BaseTail.forward(inner, _tok_img(xn[:, 1:]), wq, wk, wv, None, None, None, None, token, None, None, None,
                                   cache, t, cfg, _tok_img(out[:, 1:]))
        ctx.inner = {k: v for k, v in inner.__dict__.items() if k != "saved_tensors"}
        ctx.n_inner = len(inner.saved_tensors)
        ctx.eps = eps
        ctx.ln_meta = [(p.shape, p.dtype, p.stride()) for p in (nx_w, nx_b)]
        ctx.save_for_backward(*inner.saved_tensors, xt, g32, stats)
        return out, tok

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dout, _dtoken):
        L = _lib.lib()
        saved = ctx.saved_tensors
        inner = _Ctx(saved[:ctx.n_inner])
        inner.__dict__.update(ctx.inner)
        xt, g32, stats = saved[ctx.n_inner:]
        B, n, C = xt.shape
        dout = dout.contiguous()
        if dout.dtype != xt.dtype:
            dout = dout.to(xt.dtype)
        g = _BaseTail.backward(inner, _tok_img(dout[:, 1:]), None)
        dx_img = g[0]                                   # logical [B,C,S,S], memory [B,S,S,C] dense
        dxt = torch.empty_like(xt)
        f32 = dict(dtype=torch.float32, device=xt.device)
        dgb = torch.empty((2, C), **f32)
        a = _lib.MrlaLnArgs()
        a.B, a.n, a.C, a.dtype, a.eps = B, n, C, _DTYPES[xt.dtype], ctx.eps
        a.x, a.gamma, a.stats = _ptr(xt), _ptr(g32), _ptr(stats)
        a.g_cls, a.bs_gcls, a.g_img, a.bs_gimg = _ptr(dout), n * C, _ptr(dx_img), (n - 1) * C
        a.dx, a.dparams = _ptr(dxt), _ptr(dgb)
        nbytes = L.mrla_layernorm_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        _lib.check(L.mrla_layernorm_backward(ctypes.byref(a), _stream()), "mrla_layernorm_backward")
        launch_counter["bwd"] += L.mrla_last_launch_count()
        # _BaseTail.backward: (dx, dwq, dwk, dwv, dgamma, dbeta, dk_ext, dv_ext, dtok, ...)
        return (dxt, _like_param(dgb[0], ctx.ln_meta[0]), _like_param(dgb[1], ctx.ln_meta[1]), g[1], g[2], g[3], g[8],
                None, None, None, None)


def deit_base_module(xt, prev_k, prev_v, nx_w, nx_b, wq, wk, wv, *, init_cell: bool, cfg: "BaseCfg", eps: float,
                     cap_hint: int = 8):
    """Fused `mrlab_module.forward` (see _DeitBaseModule) -> (out [B,n,C], K, V), or None if this call needs the generic
    path (foreign K/V tensors, odd shapes)."""
    if not (xt.is_cuda and xt.dim() == 3 and xt.is_contiguous() and xt.dtype in _DTYPES):
        return None
    B, n, C = xt.shape
    S = int(round((n - 1) ** 0.5))
    if S * S + 1 != n or S < 1 or C % 4 or C > 768:
        return None
    if init_cell or prev_k is None:
        cache = StageCache(_MetaLike((B, C, S, S), xt.dtype, xt.device), _lib.NHWC, cap_hint)
        prev_token = None
    else:
        cache = getattr(prev_k, "_mrla_cache", None)
        if cache is None or cache.shape != (B, C, S, S) or cache.dtype != xt.dtype or cache.layout != _lib.NHWC:
            return None
        prev_token = getattr(prev_k, "_mrla_token", None)
    t = cache.t + 1
    cache.reserve(t)
    out, token = _DeitBaseModule.apply(xt, nx_w, nx_b, wq, wk, wv, prev_token, cache, t, cfg, eps)
    cache.t = t
    K, V = cache.K_view(t), cache.V_view(t)
    K._mrla_cache = cache
    K._mrla_token = token if token.requires_grad else None
    return out, K, V


class _MetaLike:
    """shape / dtype / device of the token image for StageCache (no allocation)."""

    def __init__(self, shape, dtype, device):
        self.shape, self.dtype, self.device = shape, dtype, device


# ===================================================================================== BatchNorm (+ReLU) producer
def is_plain_batchnorm(bn) -> bool:
    """The fused paths implement nn.BatchNorm2d and nothing else.  `norm_layer=nn.SyncBatchNorm` /
    `convert_sync_batchnorm` (statistics all-reduced across ranks), GroupNorm, FrozenBatchNorm or a user subclass with
    its own forward must keep their own semantics: callers route them through the module itself."""
    return type(bn) is torch.nn.BatchNorm2d


def effective_momentum(bn, pending: int = 0) -> float:
    """nn.BatchNorm2d momentum for this step (momentum=None: cumulative moving average 1/num_batches_tracked, counting
    this step: `pending` = 1 if the caller increments the counter only after the op).  The counter lives on the device, so momentum=None costs a host sync per call and
    cannot be captured in a CUDA graph."""
    if bn.momentum is not None:
        return bn.momentum
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("mrla_b200: BatchNorm2d(momentum=None) reads num_batches_tracked on the host every step and "
                           "cannot be captured in a CUDA graph; give the module a numeric momentum")
    return 1.0 / float(int(bn.num_batches_tracked) + pending)


def bn_act_eligible(x: torch.Tensor) -> bool:
    """Channels-last dense [B,C,H,W] (or [M,C]) CUDA activation the NHWC BatchNorm kernels can take."""
    if not x.is_cuda or x.dtype not in _DTYPES:
        return False
    if x.dim() == 2:
        return x.is_contiguous() and x.shape[1] % 8 == 0 and 8 <= x.shape[1] <= 2048 and x.data_ptr() % 16 == 0
    if x.dim() != 4:
        return False
    B, C, H, W = x.shape
    lay = _layout_of(x)
    return (lay is not None and lay[0] == _lib.NHWC and lay[1] == C * H * W and C % 8 == 0 and 8 <= C <= 2048
            and H * W > 1 and x.data_ptr() % 16 == 0)


class _BnAct(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x, weight, bias, running_mean, running_var, training, update_running, momentum, eps, relu):
        L = _lib.lib()
        if x.dim() == 4:
            B, C, H, W = x.shape
            M = B * H * W
            y = _empty_like_layout(x, _lib.NHWC)
        else:
            M, C = x.shape
            y = torch.empty_like(x)
        f32 = dict(dtype=torch.float32, device=x.device)
        stats = torch.empty((2, C), **f32)
        coef = torch.empty((2, C), **f32)
        w32, b32 = _f32(weight), _f32(bias)
        a = _lib.MrlaBnArgs()
        a.M, a.C, a.dtype = M, C, _DTYPES[x.dtype]
        a.relu, a.training, a.update_running = int(relu), int(training), int(update_running)
        a.eps, a.momentum = eps, momentum
        a.x, a.y, a.gamma, a.beta = _ptr(x), _ptr(y), _ptr(w32), _ptr(b32)
        a.running_mean, a.running_var = _ptr(running_mean), _ptr(running_var)
        a.stats, a.coef = _ptr(stats), _ptr(coef)
        nbytes = L.mrla_bn_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        _lib.check(L.mrla_bn_forward(ctypes.byref(a), _stream()), "mrla_bn_forward")
        launch_counter["fwd"] += L.mrla_last_launch_count()
        ctx.meta = (M, C, bool(relu), bool(training), eps, momentum)
        ctx.param_meta = [(p.shape, p.dtype, p.stride()) if p is not None else None for p in (weight, bias)]
        ctx.save_for_backward(x, w32, stats, coef)
        return y

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dy):
        L = _lib.lib()
        x, w32, stats, coef = ctx.saved_tensors
        M, C, relu, training, eps, momentum = ctx.meta
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        if x.dim() == 4:
            lay = _layout_of(dy)
            if lay is None or lay[0] != _lib.NHWC or lay[1] != x[0].numel():
                dy = dy.contiguous(memory_format=torch.channels_last)
                if _layout_of(dy) is None or _layout_of(dy)[0] != _lib.NHWC:
                    dy = dy.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            dx = _empty_like_layout(x, _lib.NHWC)
        else:
            dy = dy.contiguous()
            dx = torch.empty_like(x)
        f32 = dict(dtype=torch.float32, device=x.device)
        dgb = torch.empty((2, C), **f32)
        a = _lib.MrlaBnArgs()
        a.M, a.C, a.dtype = M, C, _DTYPES[x.dtype]
        a.relu, a.training = int(relu), int(training)
        a.eps, a.momentum = eps, momentum
        a.x, a.gamma, a.stats, a.coef = _ptr(x), _ptr(w32), _ptr(stats), _ptr(coef)
        a.dy, a.dx, a.dgamma, a.dbeta = _ptr(dy), _ptr(dx), _ptr(dgb[0]), _ptr(dgb[1])
        nbytes = L.mrla_bn_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        _lib.check(L.mrla_bn_backward(ctypes.byref(a), _stream()), "mrla_bn_backward")
        launch_counter["bwd"] += L.mrla_last_launch_count()
        dw = _like_param(dgb[0], ctx.param_meta[0]) if ctx.param_meta[0] is not None else None
        db = _like_param(dgb[1], ctx.param_meta[1]) if ctx.param_meta[1] is not None else None
        return dx, dw, db, None, None, None, None, None, None, None


def bn_act(x, bn: "torch.nn.BatchNorm2d", relu: bool = False) -> torch.Tensor:
    """BatchNorm2d (+ReLU) of the bottleneck on the fused NHWC kernels, with nn.BatchNorm2d semantics (train / eval,
    running statistics, momentum=None cumulative average, affine or not).  Activations that are not channels-last
    (or have C % 8 != 0) go through the library's own batch_norm (+relu) — they are not on the MRLA path."""
    if not is_plain_batchnorm(bn):
        y = bn(x)   # SyncBatchNorm / GroupNorm / frozen variants: the module's own semantics
        return torch.relu(y) if relu else y
    use_batch = bn.training or bn.running_mean is None
    if not bn_act_eligible(x) or (not use_batch and bn.running_mean is None):
        y = bn(x)
        return torch.relu(y) if relu else y
    update = bn.training and bn.track_running_stats and bn.running_mean is not None
    if update:
        bn.num_batches_tracked += 1
        momentum = effective_momentum(bn)
    else:
        momentum = 0.0
    rm, rv = bn.running_mean, bn.running_var
    if rm is not None and rm.dtype != torch.float32:
        y = bn(x)   # exotic buffer dtype: leave to the library
        return torch.relu(y) if relu else y
    return _BnAct.apply(x, bn.weight, bn.bias, rm, rv, use_batch, update, momentum, bn.eps, relu)


# ===================================================================================== stem max pooling (3x3, s2, p1)
def _max_pool_eligible(x: torch.Tensor) -> bool:
    if not x.is_cuda or x.dtype not in _DTYPES or x.dim() != 4:
        return False
    B, C, H, W = x.shape
    lay = _layout_of(x)
    align = 32 if x.dtype == torch.float32 else 16
    return (lay is not None and lay[0] == _lib.NHWC and lay[1] == C * H * W and C % 8 == 0 and H > 1 and W > 1
            and x.data_ptr() % align == 0)


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x):
        L = _lib.lib()
        B, C, H, W = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, OH, OW, C), dtype=x.dtype, device=x.device).permute(0, 3, 1, 2)
        idx = torch.empty((B, OH, OW, C), dtype=torch.uint8, device=x.device)
        _lib.check(L.mrla_maxpool3x3s2_forward(_ptr(x), _ptr(y), _ptr(idx), B, C, H, W, _DTYPES[x.dtype], _stream()),
                   "mrla_maxpool3x3s2_forward")
        launch_counter["fwd"] += L.mrla_last_launch_count()
        ctx.shape = (B, C, H, W)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    @once_differentiable
    @_guard
    def backward(ctx, dy):
        L = _lib.lib()
        (idx,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        lay = _layout_of(dy)
        if lay is None or lay[0] != _lib.NHWC or lay[1] != dy[0].numel():
            dy = dy.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        dx = torch.empty((B, H, W, C), dtype=dy.dtype, device=dy.device).permute(0, 3, 1, 2)
        _lib.check(L.mrla_maxpool3x3s2_backward(_ptr(dy), _ptr(idx), _ptr(dx), B, C, H, W, _DTYPES[dy.dtype], _stream()),
                   "mrla_maxpool3x3s2_backward")
        launch_counter["bwd"] += L.mrla_last_launch_count()
        return dx


def max_pool(x: torch.Tensor, pool: "torch.nn.MaxPool2d") -> torch.Tensor:
    """The stem's nn.MaxPool2d(3, 2, 1) on the NHWC kernels (byte tap index, gather backward); any other pooling
    configuration or activation layout goes through the module itself — it is backbone, not the MRLA path."""
    def _pair(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    if (_pair(pool.kernel_size) == (3, 3) and _pair(pool.stride) == (2, 2) and _pair(pool.padding) == (1, 1)
            and _pair(pool.dilation) == (1, 1) and not pool.ceil_mode and not pool.return_indices
            and _max_pool_eligible(x)):
        return _MaxPool3x3s2.apply(x)
    return pool(x)
