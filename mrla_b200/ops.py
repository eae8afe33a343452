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
from typing import NamedTuple, Optional

import torch

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


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


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


# --------------------------------------------------------------------------------- the fused op
class _LightTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, o, wq, wk, wv, lam, gamma, beta, running_mean, running_var, drop_scale, cfg: LightCfg, out):
        _require_cuda(x, "x")
        L = _lib.lib()
        x_c, layout, bs_x = _canon(x)
        B, C, H, W = x_c.shape
        has_o = o is not None
        if has_o:
            _require_cuda(o, "o")
            if o.shape != x.shape or o.dtype != x.dtype:
                raise RuntimeError("mrla_b200: o_prev must match x in shape and dtype")
            o_c, _, bs_o = _canon(o, layout)
        else:
            o_c, bs_o = None, 0
        if out is not None:
            lay = _layout_of(out)
            if lay is None or lay[0] != layout or out.dtype != x.dtype or out.shape != x.shape:
                raise RuntimeError("mrla_b200: `out` buffer must have the layout / dtype / shape of x")
            y, bs_y = out, lay[1]
        else:
            y = _empty_like_layout(x_c, layout)
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
        a.x, a.o, a.y = _ptr(x_c), _ptr(o_c), _ptr(y)
        a.wq, a.wk, a.wv, a.lam = _ptr(wq32), _ptr(wk32), _ptr(wv32), _ptr(lam32)
        a.gamma, a.beta, a.running_mean, a.running_var = _ptr(ga32), _ptr(be32), _ptr(rm), _ptr(rv)
        a.drop_scale = _ptr(ds32)
        a.mom, a.gate, a.mean, a.rstd, a.coef = _ptr(mom), _ptr(gate), _ptr(stats[0]), _ptr(stats[1]), _ptr(coef)
        ev = _Prof.begin()
        _lib.check(L.mrla_light_forward(ctypes.byref(a), _stream()), "mrla_light_forward")
        _Prof.end("light_fwd", (B, C, H, W, x.dtype, layout), ev)
        launch_counter["fwd"] += L.mrla_last_launch_count()
        if rm is not None and rm is not running_mean and cfg.bn_mode == _lib.BN_TRAIN and cfg.update_running:
            running_mean.copy_(rm)
            running_var.copy_(rv)

        ctx.cfg, ctx.layout, ctx.has_o = cfg, layout, has_o
        ctx.bs = (bs_x, bs_o)
        ctx.param_meta = [(p.shape, p.dtype) if p is not None else None for p in (wq, wk, wv, lam, gamma, beta)]
        ctx.save_for_backward(x_c, o_c, wq32, wk32, wv32, lam32, ga32, ds32, mom, gate, stats)
        if out is not None:
            ctx.mark_dirty(out)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x_c, o_c, wq32, wk32, wv32, lam32, ga32, ds32, mom, gate, stats = ctx.saved_tensors
        cfg, layout = ctx.cfg, ctx.layout
        B, C, H, W = x_c.shape
        dy_c, _, bs_dy = _canon(dy, layout)
        if dy_c.dtype != x_c.dtype:
            dy_c = dy_c.to(x_c.dtype)
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
        a.x, a.o = _ptr(x_c), _ptr(o_c)
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
        nbytes = L.mrla_light_bwd_scratch_bytes(ctypes.byref(a))
        scratch = torch.empty((max(nbytes, 4) + 3) // 4, **f32)
        a.scratch, a.scratch_bytes = _ptr(scratch), scratch.numel() * 4
        ev = _Prof.begin()
        _lib.check(L.mrla_light_backward(ctypes.byref(a), _stream()), "mrla_light_backward")
        _Prof.end("light_bwd", (B, C, H, W, x_c.dtype, layout), ev)
        launch_counter["bwd"] += L.mrla_last_launch_count()

        def back(i, g):
            meta = ctx.param_meta[i]
            if meta is None or g is None:
                return None
            return g.reshape(meta[0]).to(meta[1])

        grads = (dx, dout,
                 back(0, dwqk[0]), back(1, dwqk[1]), back(2, dwv),
                 back(3, dch[0]) if ctx.has_o else None,
                 back(4, dch[1]) if has_bn else None, back(5, dch[2]) if has_bn else None,
                 None, None, None, None, None)
        return grads


def light_tail(x: torch.Tensor, o: Optional[torch.Tensor], wq, wk, wv, lam=None, gamma=None, beta=None,
               running_mean=None, running_var=None, drop_scale=None, *, cfg: LightCfg, out=None) -> torch.Tensor:
    """Fused MRLA-light tail (see module docstring).  `x`/`o` are logical [B,C,H,W] tensors, either
    NCHW-contiguous or channels-last (any batch stride); the result has the layout of `x`."""
    return _LightTail.apply(x, o, wq, wk, wv, lam, gamma, beta, running_mean, running_var, drop_scale, cfg, out)
