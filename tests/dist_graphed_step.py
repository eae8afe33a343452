"""Two-rank NCCL check of mrla_b200.train.GraphedStep (run under torchrun; see tests/test_train_gpu.py).

 * the in-graph bucketed all-reduce leaves every rank with the AVERAGE of the per-rank gradients (checked against
   gradients computed without any exchange and averaged with an explicit all_gather);
 * three graph-replayed steps == three eager steps of the same sequence; ranks end with identical weights."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def main():
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # PyTorch's note for NCCL inside CUDA graphs
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    print(f"[rank {rank}] start", flush=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
    from mrla_b200.train import GraphedStep
    torch.backends.cudnn.benchmark = False
    torch.manual_seed(0)
    base = ResNet_mrlal(MRLA_Bottleneck, [1, 1, 1, 1], num_classes=10).to(dev).to(memory_format=torch.channels_last).train()
    for n, p in base.named_parameters():
        if n.endswith("bn3.weight"):
            torch.nn.init.normal_(p, 1.0, 0.2)
    crit = torch.nn.CrossEntropyLoss()
    g = torch.Generator(device="cpu").manual_seed(10 + rank)   # different data per rank
    batches = [(torch.randn(8, 3, 64, 64, generator=g).to(dev).contiguous(memory_format=torch.channels_last),
                torch.randint(0, 10, (8,), generator=g).to(dev)) for _ in range(3)]

    # (1) averaged gradients: no-exchange gradients of this rank, averaged by hand
    ref = copy.deepcopy(base)
    out = ref(batches[0][0])
    crit(out.float(), batches[0][1]).backward()
    mine = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    gathered = [torch.empty_like(mine) for _ in range(2)]
    log(rank, "reference gradients done")
    dist.all_gather(gathered, mine)
    log(rank, "all_gather done")
    want = (gathered[0] + gathered[1]) / 2
    m1 = copy.deepcopy(base)
    opt1 = torch.optim.SGD(m1.parameters(), lr=0.0)   # lr 0: the warm-up / capture executions leave the weights alone
    overlap = os.environ.get("MRLA_TEST_OVERLAP", "0") == "1"
    st1 = GraphedStep(m1, opt1, crit, batches[0][0], batches[0][1], autocast_dtype=None, bucket_mb=1.0, warmup=2, overlap=overlap)
    log(rank, "GraphedStep captured")
    assert len(st1.buckets) >= (3 if overlap else 1), len(st1.buckets)
    st1(batches[0][0], batches[0][1])
    torch.cuda.synchronize()
    got = torch.cat([p.grad.reshape(-1) for p in m1.parameters()])
    # running statistics moved during warm-up but train-mode BN uses batch statistics, so the gradients are comparable
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-5, err
    st1.close()
    print(f"[rank {rank}] averaged-gradient check ok ({err:.2e})", flush=True)

    # (2) graph replay == eager, identical weights on both ranks
    losses, states = [], []
    for capture in (True, False):
        m = copy.deepcopy(base)
        opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
        st = GraphedStep(m, opt, crit, batches[0][0], batches[0][1], autocast_dtype=None, bucket_mb=1.0, capture=capture, warmup=2, overlap=overlap)
        m.load_state_dict(base.state_dict())
        for s_ in opt.state.values():
            for v in s_.values():
                if torch.is_tensor(v):
                    v.zero_()
        log(rank, f"capture={capture} constructed")
        losses.append([float(st(x, y).detach()) for x, y in batches])
        log(rank, f"capture={capture} 3 steps done")
        states.append({k: v.clone() for k, v in m.state_dict().items()})
        st.close()
    for a, b in zip(*losses):
        assert abs(a - b) < 1e-5 * max(1.0, abs(b)), losses
    for k in states[0]:
        if states[0][k].dtype.is_floating_point:
            assert (states[0][k] - states[1][k]).abs().max().item() <= 1e-5 * max(1.0, states[1][k].abs().max().item()), k
    flat = torch.cat([v.reshape(-1).float() for k, v in states[0].items() if "running" not in k and "num_batches" not in k])
    other = [torch.empty_like(flat) for _ in range(2)]
    dist.all_gather(other, flat)
    assert torch.equal(other[0], other[1]), "ranks diverged"
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("DIST_GRAPHED_STEP_OK", flush=True)
    # captured graphs still reference the NCCL communicator; tearing the process group down under them blocks in this
    # PyTorch / NCCL combination, so the checker leaves without the (irrelevant) orderly shutdown
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
