"""Host-side planner of the v7 sweeps (mrla_b200/csrc/light_v7_launch.cuh::v7_plan) through the C ABI's host-only query
mrla_light_v7_plan — no CUDA call is made, so this runs without a GPU.  Covers: the BASELINE stage shapes, column tiles for
W > 56 (mmdet feature maps, mmdetection/mmdet/models/backbones/resnet_mrlal.py:116), one-unit-per-tile splitting for small
batches, shared-memory / thread budgets, and the shapes the planner must refuse."""
import ctypes

import pytest

from mrla_b200 import _lib

KEYS = ("ok", "CB", "NQ", "NT", "U", "TPU", "S", "cpc", "grid", "threads", "ctas", "smem")
SMS = 148


def plan(B, C, H, W, kind, xf, dtype=None, layout=None, act=0):
    L = _lib.lib()
    a = _lib.MrlaLightArgs()
    a.B, a.C, a.H, a.W = B, C, H, W
    a.dim_perhead, a.k_size = 32, 3
    a.dtype = _lib.BF16 if dtype is None else dtype
    a.layout = _lib.NHWC if layout is None else layout
    a.act, a.bn_mode = act, _lib.BN_TRAIN
    n = C * H * W
    a.bs_x = a.bs_o = a.bs_y = a.bs_z = n
    a.x = a.o = a.z = 0x7F0000000000          # never dereferenced: the query is host-only
    out = (ctypes.c_int64 * 12)()
    L.mrla_light_v7_plan.restype = ctypes.c_int
    L.mrla_light_v7_plan.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    ok = L.mrla_light_v7_plan(ctypes.byref(a), kind, int(xf), out)
    d = dict(zip(KEYS, list(out)))
    assert bool(d["ok"]) == bool(ok)
    return d


@pytest.mark.parametrize("C,HW,nq,threads_b", [(256, 56, 8, 256), (512, 28, 4, 256), (1024, 14, 2, 256), (2048, 7, 1, 128)])
@pytest.mark.parametrize("xf", [False, True])
def test_baseline_stage_shapes(C, HW, nq, threads_b, xf):
    """BASELINE configs[1] shapes, B = 256 bf16: one tile, one unit per image, the grid covers the SMs, sweep B has no
    producer warp, sweeps 1 / 2 / A have one."""
    for kind in range(4):
        d = plan(256, C, HW, HW, kind, xf)
        assert d["ok"] == 1, (kind, d)
        assert d["NQ"] == nq and d["NT"] == 1 and d["TPU"] == 1 and d["U"] == 256
        assert d["NQ"] * d["CB"] // 2 <= 256                       # 32 channel pairs per warp, <= 8 consumer warps
        assert d["threads"] == (threads_b if kind == 3 else threads_b + 32)
        assert d["grid"] == (C // d["CB"]) * d["cpc"] and SMS // 2 < d["grid"] <= SMS * d["ctas"]
        assert d["smem"] * d["ctas"] <= 227 * 1024
        assert d["S"] >= (5 if kind == 3 else 4)


@pytest.mark.parametrize("B,C,H,W", [(64, 256, 112, 112), (40, 256, 12, 70), (160, 64, 6, 113)])
def test_wide_maps_are_walked_in_column_tiles(B, C, H, W):
    """W > 56 with enough images to fill the SMs: tiles of 8 x 7 columns, one unit per image (deterministic moments)."""
    for kind in range(4):
        d = plan(B, C, H, W, kind, True)
        assert d["ok"] == 1
        assert d["NQ"] == 8 and d["NT"] == -(-W // 56) and d["TPU"] == d["NT"] and d["U"] == B


@pytest.mark.parametrize("B,C,H,W", [(2, 256, 200, 304), (2, 512, 100, 152), (2, 1024, 50, 76), (1, 256, 200, 304), (2, 64, 20, 100)])
def test_small_batches_of_wide_maps_split_into_tile_units(B, C, H, W):
    """Detection batches (2 images per GPU): the work unit is one column tile, tiles narrow to 4 x 7 columns while the
    units do not cover the SMs, channel blocks stay at 64, and all four sweeps agree on the geometry."""
    geo = set()
    for kind in range(4):
        for xf in (False, True):
            d = plan(B, C, H, W, kind, xf)
            assert d["ok"] == 1
            assert d["TPU"] == 1 and d["U"] == B * d["NT"] and d["CB"] == 64
            assert d["NT"] == -(-W // (7 * d["NQ"]))
            assert d["cpc"] <= d["U"]
            geo.add((d["NQ"], d["NT"], d["U"]))
    assert len(geo) == 1
    nq = next(iter(geo))[0]
    assert nq == (4 if B * (-(-W // 56)) * (C // 64) < SMS else 8)


def test_stage4_of_a_detection_batch_needs_no_tiles():
    d = plan(2, 2048, 25, 38, 3, True)
    assert d["ok"] == 1 and d["NT"] == 1 and d["U"] == 2


@pytest.mark.parametrize("kw", [dict(C=96), dict(H=2), dict(W=513), dict(layout="nchw"), dict(act=1)])
def test_planner_refuses_what_the_kernels_do_not_cover(kw):
    """C % 64 != 0 (EfficientNet widths), H < 3, W > 512, NCHW, the GELU variant: the generic kernels serve these."""
    args = dict(B=8, C=256, H=14, W=14)
    layout = None
    act = 0
    for k, v in kw.items():
        if k == "layout":
            layout = _lib.NCHW
        elif k == "act":
            act = v
        else:
            args[k] = v
    for kind in range(4):
        d = plan(args["B"], args["C"], args["H"], args["W"], kind, False, layout=layout, act=act)
        assert d["ok"] == 0 and d["CB"] == -1


def test_fp32_virtual_x_at_eight_column_groups_does_not_fit():
    """fp32 stages of 8 column groups with the o-halo of the x re-formation exceed shared memory in sweep B: the folded fp32
    op keeps the round-1 kernels at W = 56 (tests/test_v7_gpu.py runs them), while bf16 / fp16 fit."""
    assert plan(32, 256, 56, 56, 3, True, dtype=_lib.F32)["ok"] == 0
    assert plan(32, 256, 56, 56, 3, True, dtype=_lib.BF16)["ok"] == 1
    assert plan(32, 256, 56, 56, 3, True, dtype=_lib.F16)["ok"] == 1


def test_multiply_high_reciprocal_is_exact_in_the_planned_range():
    """light_v7.cuh::v7_fdiv computes a / d as umulhi(a, ceil(2^32 / d)); v7_plan refuses shapes whose pipeline-row index
    could leave the range where that is exact (rows * d < 2^32).  Check the arithmetic the kernels rely on, including the
    largest admitted operands."""
    import random
    rnd = random.Random(7)

    def fdiv(a, d):
        m = ((1 << 32) + d - 1) // d
        assert m < (1 << 32)
        return (a * m) >> 32

    for _ in range(20000):
        d = rnd.randint(2, 1 << rnd.randint(2, 20))
        amax = ((1 << 32) - 1) // d
        for a in (0, 1, d - 1, d, d + 1, amax // 2, amax - 1, amax, rnd.randint(0, amax)):
            if 0 <= a <= amax:
                assert fdiv(a, d) == a // d, (a, d)
    # the bound is tight-ish: far outside it the shortcut does go wrong, which is why the planner checks it
    assert any(fdiv(a, 3) != a // 3 for a in range((1 << 32) - 64, 1 << 32))
