#!/usr/bin/env python
"""bench.py — resnet50_mrlal training throughput on N B200s (BASELINE.json configs[1]) + MRLA-tail roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm (CUDA kernels via the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]       # reference arm: the reference's CPU path
    python bench.py --config mrlab|deit_tail|effnet_tail            # the secondary BASELINE configs[2..4] (own JSON line)
    python bench.py --compile-baseline                              # adds torch.compile to the same-GPU eager baseline

One "step" = forward + loss + backward + SGD update of `resnet50_mrlal` on one synthetic batch
(256 x 3 x 224 x 224 per GPU, bf16 autocast, channels_last, random-init weights, drop_path 0.2 as
resnet/train.py:67 of the reference).  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

METRIC = "resnet50_mrlal_train_images_per_sec"

# Libraries (NCCL's version banner, cuDNN warnings) write to the C-level stdout; the contract is ONE JSON line there.
# fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
STAGE_BLOCKS = {(256, 56): 3, (512, 28): 4, (1024, 14): 6, (2048, 7): 3}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            # "under load" = samples with power in the upper half of the observed range
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            load = [s for s, p_ in zip(sm, pw) if p_ >= thr] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       power_w_max=max(pw), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(steps: int, warmup: int, batch: int = 0):
    """The reference's CPU path (oracle port of resnet50_mrlal; /root/reference is absent on the GPU box):
    fwd+bwd of BASELINE.json configs[0] (batch 32 x 3 x 224 x 224 fp32) on all host cores."""
    from oracle.resnet_oracle import resnet50_mrlal_oracle
    batch = batch or int(os.environ.get("MRLA_BENCH_CPU_BATCH", "32"))  # env override: CPU test-suite only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = resnet50_mrlal_oracle(drop_path=0.0).train()
    x = torch.randn(batch, 3, 224, 224)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        y = model(x)
        y.sum().backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    med = statistics.median(times)
    cpu_name = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_name = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return dict(img_per_s=batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                threads=torch.get_num_threads(), cpu=cpu_name, batch=batch, img_per_s_median=batch / med,
                ms_median=1e3 * med, iters=len(times))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_reference_run(steps, warmup)
    sample = (f"oracle port of reference resnet50_mrlal (resnet/models/resnet_mrla_light.py), fwd+bwd, batch {r['batch']}"
              f" x3x224x224 fp32 (BASELINE configs[0]), {steps} timed + {warmup} warm-up iterations, {r['cpu']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["img_per_s"], 3), "unit": "img/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(r["ms_per_step"], 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "resnet50_mrlal train step (fwd+bwd), 224x224 synthetic, CPU sample batch 32 fp32",
                   "model": "resnet50_mrlal", "device": "host CPU"},
        "cpu_baseline": {"value": round(r["img_per_s"], 3), "unit": "img/s", "cores": r["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(r["img_per_s"], 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ product arm
def measure_tail_group(B, C, HW, dev, iters=10, bn3=True):
    """Device time (ms) of the folded MRLA-light tail op, forward and backward kernel groups, at one stage shape."""
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    bf = torch.bfloat16
    k = 7 if C == 2048 else 5
    mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
    z, idt, dy = mk(), mk(), mk()
    z.requires_grad_(); idt.requires_grad_()
    P = [torch.randn(k, device=dev).requires_grad_(), torch.randn(k, device=dev).requires_grad_(),
         (torch.randn(C, 1, 3, 3, device=dev) * 0.05).requires_grad_(), torch.randn(C, 1, 1, device=dev).requires_grad_(),
         torch.ones(C, device=dev).requires_grad_(), torch.zeros(C, device=dev).requires_grad_()]
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    cfg = LightCfg(dim_perhead=32, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
    # the model hands the op the raw conv3 output plus the bn3 coefficients (sweep-1 MODE 6): time that variant
    zc = torch.stack([torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1]).contiguous() if bn3 else None

    def fwd():
        return light_tail(z, idt, *P, rm, rv, None, cfg=cfg, z_coef=zc)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fwd().backward(dy)
    torch.cuda.current_stream().wait_stream(side)
    for t in [z, idt] + P:
        t.grad = None
    g_f, g_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_f):
        y = fwd()
    with torch.cuda.graph(g_b, pool=g_f.pool()):
        y.backward(dy)
    out = []
    for g in (g_f, g_b):
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / iters)
    del g_f, g_b, y
    return out[0], out[1]


def gpu_eager_run(dev, B, steps, warmup, drop_path, compiled=False):
    """SURVEY.md §2.2's real bar: eager PyTorch running the reference graph on the SAME B200 — the oracle port of
    resnet50_mrlal (the reference's own ~14 ATen calls per tail; /root/reference does not travel to the GPU box), bf16
    autocast, channels_last, SGD, same batch.  `compiled`: the same model through torch.compile (inductor)."""
    from oracle.resnet_oracle import resnet50_mrlal_oracle
    torch.manual_seed(0)
    model = resnet50_mrlal_oracle(drop_path=drop_path).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    crit = nn.CrossEntropyLoss().to(dev)
    net = torch.compile(model) if compiled else model
    img = torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    lbl = torch.randint(0, 1000, (B,), device=dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(img)
        loss = crit(out.float(), lbl)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, net, opt
    torch.cuda.empty_cache()
    return dict(img_per_s=B / (ms / 1e3), ms_per_step=ms)


def traffic_record():
    """Measured DRAM bytes of the stage-1 tail op (ncu --set full, tools/profile_v7.sh) with the commit it was taken at."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tpath))
    except Exception:
        return {}


def run_product_arm(args):
    from mrla_b200 import ops
    from mrla_b200.train import GraphedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    if args.config == "mrlab":
        from mrla_b200.resnet_mrla_base import resnet50_mrlab as factory
        metric, model_name = "resnet50_mrlab_train_images_per_sec", "resnet50_mrlab"
    else:
        from mrla_b200.resnet_mrla_light import resnet50_mrlal as factory
        metric, model_name = METRIC, "resnet50_mrlal"
    model = factory(drop_path=args.drop_path).to(dev).train()
    if not args.nchw:   # --nchw: the stock reference call path (train.py never asks for channels_last)
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)  # reference train.py:199
    net = model
    if world > 1 and args.no_graph:   # eager path: the reference's DDP wrap; the graph path all-reduces a flat buffer
        net = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True)
    crit = nn.CrossEntropyLoss().to(dev)
    gen = torch.Generator(device="cpu").manual_seed(1 + rank)
    # synthetic ImageNet-shaped batch.  Host side (e2e): decoded uint8 HWC images in pinned memory, normalised to
    # bf16 channels_last ON THE DEVICE after the copy (timm fast_collate / PrefetchLoader convention; the reference's
    # torchvision path ships fp32 CHW = 4x the PCIe bytes).  `value` uses the normalised batch resident in HBM.
    host_img = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(2)]
    host_lbl = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(2)]
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).mul_(255).view(1, 3, 1, 1)
    inv_std = (1.0 / (torch.tensor([0.229, 0.224, 0.225], device=dev) * 255)).view(1, 3, 1, 1)

    def normalise(u8_nhwc):   # [B,224,224,3] uint8 on device -> [B,3,224,224] bf16, channels_last strides (no transpose)
        return u8_nhwc.permute(0, 3, 1, 2).float().sub_(mean).mul_(inv_std).to(torch.bfloat16)

    if args.nchw:
        normalise_cl = normalise
        normalise = lambda u8: normalise_cl(u8).contiguous()   # dense NCHW, as torchvision's ToTensor pipeline delivers
    dev_img = normalise(host_img[0].to(dev))
    assert args.nchw or dev_img.is_contiguous(memory_format=torch.channels_last)
    dev_lbl = host_lbl[0].to(dev)

    def step(img, lbl):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(img)
        loss = crit(out.float(), lbl)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = args.warmup, args.steps
    for _ in range(W):
        step(dev_img, dev_lbl)
    barrier()

    # ---- the product's graphed step (mrla_b200.train.GraphedStep): graph A = forward + loss + backward into ONE flat
    # fp32 gradient buffer with the bucketed NCCL all-reduce (AVG) overlapped with backward INSIDE the graph,
    # graph B = SGD update.  Replaces the reference's eager loop + DDP wrap (train.py:172-174, 387-409).
    gstep = None
    launches_per_step = None
    if not args.no_graph:
        l0 = ops.launch_counter["fwd"] + ops.launch_counter["bwd"]
        gstep = GraphedStep(model, opt, crit, dev_img, dev_lbl, warmup=3, overlap=args.overlap_comm,
                            capture_collective=args.capture_comm)
        # launches of one captured step = (warm-up + capture) launches / their count
        launches_per_step = (ops.launch_counter["fwd"] + ops.launch_counter["bwd"] - l0) // 4
        dev_img, dev_lbl = gstep.inputs, gstep.targets
    barrier()

    def run_step():
        if gstep is not None:
            return gstep()
        return step(dev_img, dev_lbl)

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ops.launch_counter["fwd"] + ops.launch_counter["bwd"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` then captures exactly the timed steps (all threads)
    e0.record()
    for _ in range(K):
        run_step()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = (launches_per_step * K) if gstep is not None else (ops.launch_counter["fwd"] + ops.launch_counter["bwd"] - l0)
    clocks = sampler.stop() if sampler else None

    # ---- device time of the gradient exchange alone (the buckets of the flat buffer, back to back on one stream) ----
    collective_ms = None
    if world > 1 and gstep is not None:
        for _ in range(2):
            for s_, e_, _ in gstep.buckets:
                dist.all_reduce(gstep.flat[s_:e_], op=dist.ReduceOp.AVG)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            for s_, e_, _ in gstep.buckets:
                dist.all_reduce(gstep.flat[s_:e_], op=dist.ReduceOp.AVG)
        c1.record()
        torch.cuda.synchronize()
        collective_ms = max_over_ranks(c0.elapsed_time(c1) / 5)

    # ---- timed region 2 (e2e): pinned host batch -> H2D every step, loss read back every step ----
    copy_stream = torch.cuda.Stream()
    slots = [None, None]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            img = host_img[i % 2].to(dev, non_blocking=True)
            lbl = host_lbl[i % 2].to(dev, non_blocking=True)
            img = normalise(img)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        slots[i % 2] = (img, lbl, ev)

    host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_loop(n):
        prefetch(0)
        last = 0.0
        pending = None  # (pinned host scalar, event): the loss of step i is read on the host while step i+1 runs
        for i in range(n):
            img, lbl, ev = slots[i % 2]
            torch.cuda.current_stream().wait_event(ev)
            img.record_stream(torch.cuda.current_stream())
            if gstep is not None:
                gstep.load(img, lbl)      # the graph reads its static input buffers
            if i + 1 < n:
                prefetch(i + 1)
            loss = run_step() if gstep is not None else step(img, lbl)
            host_loss[i % 2].copy_(loss.detach(), non_blocking=True)   # D2H read of this step's result
            done = torch.cuda.Event()
            done.record()
            if pending is not None:
                pending[1].synchronize()
                last = float(pending[0])
            pending = (host_loss[i % 2], done)
        pending[1].synchronize()
        return float(pending[0])

    e2e_loop(3)   # untimed: the copy stream's allocator pool, the pinned-copy path and the D2H slot are warm afterwards
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    last_loss = e2e_loop(K)
    s1.record()
    barrier()
    ms_e2e = max_over_ranks(s0.elapsed_time(s1))
    h2d = host_img[0].numel() * host_img[0].element_size() + host_lbl[0].numel() * 8

    # ---- roofline region: the MRLA tail kernel group of every ResNet-50 stage shape, exactly as the model calls it
    # (same public op, same tensors sizes / dtype / layout), captured in a CUDA graph and replayed back to back
    # between two CUDA events on the launching stream: pure device time of the kernel group, independent of how
    # fast the host can enqueue (operands are 3 x 0.4 GB at stage 1, far beyond L2, so every replay is cold).
    tail_times = {}
    if rank == 0 and args.config == "mrlal":
        for (C, HW), nblk in STAGE_BLOCKS.items():
            tail_times[(C, HW)] = measure_tail_group(B, C, HW, dev)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = peaks()
    roof = None
    if tail_times:
        per_shape, tot_bytes, tot_ms = {}, 0.0, 0.0
        for (C, HW), (f_ms, b_ms) in tail_times.items():
            # 8*N*sizeof — fwd R c3,id W y ; bwd R dy,c3,id W d(c3-side),d(id).  x = relu(bn3(c3)+id) is
            # never materialised (round 2), so the op's algorithmic bytes ARE the survey's 8 N; round 1 stored x (9 N).
            nbytes = 8.0 * B * C * HW * HW * 2
            per_shape[(C, HW)] = dict(fwd_ms=f_ms, bwd_ms=b_ms, bytes=nbytes, gbs=nbytes / (f_ms + b_ms) / 1e6)
            tot_bytes += STAGE_BLOCKS[(C, HW)] * nbytes
            tot_ms += STAGE_BLOCKS[(C, HW)] * (f_ms + b_ms)
        d = per_shape[(256, 56)]
        tr = traffic_record()
        roof = {"bound": "hbm", "achieved": round(d["gbs"], 1), "peak": peak, "unit": "GB/s",
                "frac": round(d["gbs"] / peak, 4), "traffic": tr.get("stage1_tail_fwd_bwd_dram_bytes"),
                "traffic_source": {k_: tr.get(k_) for k_ in ("measured_at_commit", "file", "how") if k_ in tr},
                "frac_on_round1_9N_basis": round(d["gbs"] * 9.0 / 8.0 / peak, 4),
                "frac_of_nominal_8TBps": round(d["gbs"] / 8000.0, 4),
                "kernel": "MRLA-light block tail fwd+bwd kernel group with the bottleneck's bn3 affine and residual "
                          "add+ReLU folded in and x never materialised, stage-1 shape (B,256,56,56) bf16 NHWC (sweep 1 = "
                          "re-form x + moments, cluster mid kernel, sweep 2; sweep A, cluster mid + gate kernels, sweep B "
                          "incl. bn3's backward sums, finish); algorithmic bytes 8*N*2 per block (fwd R c3,id W y; bwd R "
                          "dy,c3,id W d_c3side,d_id) = SURVEY.md 8(d)",
                "peak_kind": peak_kind,
                "launch_ms": {"fwd": round(d["fwd_ms"], 4), "bwd": round(d["bwd_ms"], 4)},
                "how": "the op's forward / backward kernel groups captured in CUDA graphs and replayed 10x between CUDA "
                       "events on the launching stream, same shapes/dtype/layout as the model's calls, inputs randn (the "
                       "post-ReLU sparsity of the real identity does not change the bytes moved); stage-1 operands "
                       "(3 x 0.41 GB) exceed L2, stages 3-4 are partially L2-resident as they are inside the model",
                "all_16_tails": {"ms_per_step": round(tot_ms, 3), "alg_GB_per_step": round(tot_bytes / 1e9, 3),
                                 "achieved": round(tot_bytes / tot_ms / 1e6, 1) if tot_ms else None,
                                 "frac": round(tot_bytes / tot_ms / 1e6 / peak, 4) if tot_ms else None},
                "per_stage": {f"{c}x{h}x{h}": {"fwd_ms": round(v["fwd_ms"], 4), "bwd_ms": round(v["bwd_ms"], 4),
                                               "GBps": round(v["gbs"], 1), "frac": round(v["gbs"] / peak, 4)}
                              for (c, h), v in per_shape.items()}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.config == "mrlal":
        r = cpu_reference_run(steps=3, warmup=1)
        cpu = {"value": round(r["img_per_s_median"], 3), "unit": "img/s", "cores": r["threads"], "kind": "port",
               "sample": f"oracle port of reference resnet50_mrlal fwd+bwd, batch {r['batch']} x3x224x224 fp32 "
                         f"(BASELINE configs[0]), median of {r['iters']} timed iterations after 1 warm-up, "
                         f"{r['ms_median']:.0f} ms/iter, {r['cpu']}"}
    eager = None
    if world == 1 and not args.no_eager_baseline and args.config == "mrlal":
        # the GPU memory of the product model is still held: free what the graphs do not need first
        torch.cuda.empty_cache()
        try:
            r = gpu_eager_run(dev, B, steps=5, warmup=3, drop_path=args.drop_path)
            eager = {"value": round(r["img_per_s"], 1), "unit": "img/s", "ms_per_step": round(r["ms_per_step"], 2),
                     "what": "oracle port of the reference resnet50_mrlal (its own ATen call sequence) on the same B200: eager "
                             "PyTorch, bf16 autocast, channels_last, SGD, same per-GPU batch; 5 timed steps after 3 warm-up"}
            if args.compile_baseline:
                rc = gpu_eager_run(dev, B, steps=5, warmup=3, drop_path=args.drop_path, compiled=True)
                eager["torch_compile"] = {"value": round(rc["img_per_s"], 1), "ms_per_step": round(rc["ms_per_step"], 2)}
            else:   # inductor needs minutes to compile the training graph: measured on its own, quoted with its source
                try:
                    rec = json.load(open(os.path.join(ROOT, "profiles", "r02_compile_baseline.json")))
                    eager["torch_compile"] = {"value": rec["value"], "ms_per_step": rec["ms_per_step"],
                                              "measured": "separately (tools/compile_baseline.py, not in this run)",
                                              "source": "profiles/r02_compile_baseline.json"}
                except Exception:
                    pass
        except Exception as exc:   # out of memory next to the captured graphs, inductor missing a toolchain, ...
            eager = {"unavailable": repr(exc)[:200]}
    line = {
        "metric": metric, "value": round(world * B * K / (ms / 1e3), 2), "unit": "img/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{model_name} training step (fwd+loss+bwd+SGD), BASELINE configs[{2 if args.config == 'mrlab' else 1}]",
                   "model": model_name, "per_gpu_batch": B, "global_batch": B * world, "image": "3x224x224",
                   "precision": "bf16 autocast, fp32 master weights", "memory_format": "nchw model + images (promoted by the product model)" if args.nchw else "channels_last",
                   "drop_path": args.drop_path,
                   "host_batch": "e2e: pinned uint8 HWC images + int64 labels copied H2D every step, normalised to bf16 "
                                 "channels_last on the device",
                   "parallelism": f"dp{world}" + ((" (flat-gradient NCCL all-reduce " +
                                                   ("bucketed and overlapped with backward inside the step graph)" if args.overlap_comm
                                                    else ("captured in the step graph)" if args.capture_comm
                                                          else "between the fwd+bwd graph and the SGD graph)"))
                                                   if gstep is not None else " (DDP/NCCL)") if world > 1 else ""),
                   "l2": "inputs exceed L2 (per-step activations >> 126 MB); no explicit flush",
                   "launch": "mrla_b200.train.GraphedStep: CUDA graphs (fwd+bwd+all-reduce | SGD)" if gstep is not None
                             else "eager launches"},
        "clocks": clocks,
        "e2e": {"value": round(world * B * K / (ms_e2e / 1e3), 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 3), "last_loss": round(last_loss, 4)},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
    }
    if collective_ms is not None:
        line["collective_ms"] = round(collective_ms, 4)
        line["collective"] = {"buckets": len(gstep.buckets), "bytes": int(gstep.flat.numel() * 4),
                              "what": "device time of the gradient all-reduce (AVG, fp32 flat buffer) run alone; in the step it "
                                      + ("is launched per bucket from backward hooks on a side stream (--overlap-comm)"
                                         if args.overlap_comm else
                                         ("is captured at the end of graph A (--capture-comm)" if args.capture_comm
                                          else "runs between graph A (fwd+bwd) and graph B (SGD)"))}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ secondary configs
def _time_fn(fn, iters=20, warm=5, leaves=()):
    """Device time (ms) of `fn` (one forward + backward): captured in a CUDA graph after warm-up on a side stream and replayed
    back to back between CUDA events — these modules are a few tens of microseconds of GPU work, so an eager loop would time
    the Python / launch path instead.  Returns (graph_ms, eager_loop_ms)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / iters
    for t in leaves:
        t.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    del g
    return ms, eager_ms


def run_tail_config(args):
    """--config deit_tail / effnet_tail: BASELINE configs[4] / [3] at the module level (SURVEY.md 8d metric 3): device
    time of the MRLA tail fwd+bwd at the model's shapes, algorithmic GB/s on 8*N*sizeof, next to eager PyTorch running the
    oracle restatement (the reference's ATen sequence) on the same GPU."""
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    from oracle import mrla_oracle as O
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak, peak_kind = peaks()
    bf = torch.bfloat16
    rows = {}
    tot_bytes = tot_ms = tot_eager = 0.0
    if args.config == "deit_tail":
        from mrla_b200.deit_mrla_light import mrlal_module
        B = args.batch
        mod = mrlal_module(192, 16).to(dev).to(bf)
        x = torch.randn(B, 197, 192, device=dev, dtype=bf).requires_grad_()
        o = torch.randn(B, 197, 192, device=dev, dtype=bf).requires_grad_()
        dy = torch.randn(B, 197, 192, device=dev, dtype=bf)
        P = dict(mod.named_parameters())

        def mine():
            y = x + mod(x, o)
            y.backward(dy)

        def eager():
            y = O.deit_light_block_tail(x, o, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                                        12, P["normx.weight"], P["normx.bias"], P["normo.weight"], P["normo.bias"])
            y.backward(dy)

        leaves = [x, o] + list(P.values())
        (t_m, t_m_loop), (t_e, t_e_loop) = _time_fn(mine, leaves=leaves), _time_fn(eager, leaves=leaves)
        nbytes = 8.0 * B * 197 * 192 * 2
        rows["256x197x192"] = dict(ms=round(t_m, 4), eager_ms=round(t_e, 4), GBps=round(nbytes / t_m / 1e6, 1),
                                   python_loop_ms=round(t_m_loop, 4), eager_python_loop_ms=round(t_e_loop, 4))
        tot_bytes, tot_ms, tot_eager = 12 * nbytes, 12 * t_m, 12 * t_e
        metric, unit, value = "deit_mrlal_tiny_mrla_tails_ms_per_step", "ms", 12 * t_m
        workload = "deit_mrlal_tiny_patch16_224: the 12 mrlal_module calls of one training step (fwd+bwd), B=%d, 197x192 bf16" % B
    else:
        B = args.batch if args.batch != 256 else 384
        shapes = [(16, 112, 1), (24, 56, 2), (40, 28, 2), (80, 14, 3), (112, 14, 3), (192, 7, 4), (320, 7, 1)]
        for C, HW, nblk in shapes:
            k = O.eca_kernel_size(C)
            mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
            x, o_, dy = torch.relu(mk()).requires_grad_(), mk().requires_grad_(), mk()
            Pm = [torch.randn(k, device=dev).requires_grad_(), torch.randn(k, device=dev).requires_grad_(),
                  (torch.randn(C, 1, 3, 3, device=dev) * 0.3).requires_grad_(), torch.randn(C, 1, 1, device=dev).requires_grad_(),
                  torch.ones(C, device=dev).requires_grad_(), torch.zeros(C, device=dev).requires_grad_()]
            rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
            cfg = LightCfg(dim_perhead=8, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True)
            Pe = [p.detach().to(bf).requires_grad_() for p in Pm]

            def mine():
                light_tail(x, o_, *Pm, rm, rv, None, cfg=cfg).backward(dy)

            def eager():
                y, _, _ = O.light_tail(x, o_, Pe[0], Pe[1], Pe[2], Pe[3], C // 8, Pe[4], Pe[5], rm.to(bf), rv.to(bf))
                y.backward(dy)

            leaves = [x, o_] + Pm + Pe
            (t_m, _), (t_e, _) = _time_fn(mine, 10, 3, leaves), _time_fn(eager, 5, 3, leaves)
            nbytes = 8.0 * B * C * HW * HW * 2
            rows[f"{C}x{HW}x{HW}"] = dict(ms=round(t_m, 4), eager_ms=round(t_e, 4), GBps=round(nbytes / t_m / 1e6, 1), blocks=nblk)
            tot_bytes += nblk * nbytes
            tot_ms += nblk * t_m
            tot_eager += nblk * t_e
        metric, unit, value = "efficientnet_b0_mrlal_tails_ms_per_step", "ms", tot_ms
        workload = "EfficientNet-B0 block-output shapes @224 (SURVEY 8a A9), MRLA-light tail fwd+bwd per block, B=%d bf16 NHWC, d=8" % B
    emit({"metric": metric, "value": round(value, 4), "unit": unit, "n_gpus": 1, "steps": 20, "warmup": 5,
          "ms_per_step": round(value, 4), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
          "data": "synthetic", "config": {"workload": workload, "l2": "tensors of the small shapes are L2-resident: latency-bound, "
                                          "the HBM figure is for reference (SURVEY 8d)"},
          "roofline": {"bound": "hbm", "achieved": round(tot_bytes / tot_ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                       "frac": round(tot_bytes / tot_ms / 1e6 / peak, 4), "traffic": None, "peak_kind": peak_kind,
                       "per_shape": rows},
          "gpu_eager_baseline": {"value": round(tot_eager, 4), "unit": "ms", "speedup": round(tot_eager / tot_ms, 2),
                                 "what": "oracle restatement (the reference's ATen sequence) on the same GPU, same dtype, also "
                                         "replayed from a CUDA graph (device time, no launch overhead)"},
          "timing": "fwd+bwd of each module captured in a CUDA graph, replayed 10-20x between CUDA events (device time)",
          "gpu_launches": None})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--drop-path", type=float, default=0.2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--config", default="mrlal", choices=["mrlal", "mrlab", "deit_tail", "effnet_tail"],
                    help="mrlal = BASELINE configs[1] (default); mrlab = configs[2]; deit_tail / effnet_tail = configs[4] / [3] "
                         "at the module level")
    ap.add_argument("--overlap-comm", action="store_true",
                    help="N > 1: bucketed gradient all-reduce overlapped with backward instead of one all-reduce after it")
    ap.add_argument("--capture-comm", action="store_true",
                    help="N > 1: capture the flat-gradient all-reduce inside graph A instead of issuing it between the graphs")
    ap.add_argument("--nchw", action="store_true",
                    help="leave model and images in the stock NCHW layout of the reference's train.py (the product model "
                         "promotes the image batch to channels_last itself)")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the same-GPU eager PyTorch baseline")
    ap.add_argument("--compile-baseline", action="store_true", help="also time the eager baseline under torch.compile")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != args.gpus and world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun when invoked directly with --gpus N
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd, stdout=_REAL_STDOUT))
        if args.config in ("deit_tail", "effnet_tail"):
            run_tail_config(args)
        else:
            run_product_arm(args)


if __name__ == "__main__":
    main()
