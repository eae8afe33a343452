"""World-size-2 `gloo` tests (CPU) for the host-side logic of the multi-GPU path:
the path shards by batch with NO collective inside the MRLA op (plain per-rank BatchNorm statistics, as the
reference's DDP does — resnet/train.py:172-174); the only exchange is the gradient all-reduce.  The CUDA kernels
cannot run here, so the per-rank arithmetic is the oracle's; what is exercised is the sharding / reduction
contract bench.py relies on and its rank-0-only reporting."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical weights on every rank (DDP would broadcast rank 0's)
    from oracle.resnet_oracle import ResNetMrlalOracle
    model = ResNetMrlalOracle([1, 1, 1, 1], num_classes=10).train()
    for m in model.modules():
        if hasattr(m, "bn3"):
            torch.nn.init.normal_(m.bn3.weight, 1.0, 0.1)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    g = torch.Generator().manual_seed(100 + rank)  # each rank gets its own shard of the global batch
    x = torch.randn(2, 3, 64, 64, generator=g)
    y = ddp(x)
    y.square().mean().backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    # reference semantics: per-rank BN statistics -> running stats differ across ranks, gradients are averaged
    rm = model.layer1[0].bn_mrla.running_mean.clone()
    gathered = [torch.zeros_like(rm) for _ in range(world)]
    dist.all_gather(gathered, rm)
    lam_grad = grads["layer1.0.mrla.lambda_t"]
    glist = [torch.zeros_like(lam_grad) for _ in range(world)]
    dist.all_gather(glist, lam_grad)
    # local (un-reduced) gradient of this rank's shard, recomputed without DDP
    model.zero_grad()
    model(x).square().mean().backward()
    local = model.layer1[0].mrla.lambda_t.grad.clone()
    llist = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(llist, local)
    if rank == 0:
        q.put(dict(same_across_ranks=bool(torch.allclose(glist[0], glist[1])),
                   is_mean_of_locals=bool(torch.allclose(glist[0], (llist[0] + llist[1]) / world, rtol=1e-4, atol=1e-7)),
                   bn_stats_per_rank=bool(not torch.allclose(gathered[0], gathered[1]))))
    dist.destroy_process_group()


def test_batch_sharding_contract_gloo_ws2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == dict(same_across_ranks=True, is_mean_of_locals=True, bn_stats_per_rank=True)


def test_reference_arm_prints_on_rank0_only(monkeypatch):
    """`bench.py --impl reference` under a 2-rank launch: rank 0 alone runs and prints ONE JSON line, rank 1 exits 0."""
    env = dict(os.environ, MRLA_BENCH_CPU_BATCH="2", OMP_NUM_THREADS="4")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29655")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "1"], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("{")])
    assert len(outs[0]) == 1 and outs[1] == []
    line = json.loads(outs[0][0])
    assert line["impl"] == "reference" and line["metric"] == "resnet50_mrlal_train_images_per_sec"
    assert line["unit"] == "img/s" and line["value"] > 0 and line["n_gpus"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_product_arm_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
