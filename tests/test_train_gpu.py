"""GPU tests of mrla_b200.train.GraphedStep — the product's replacement for the hot loop of the reference trainer
(resnet/train.py:387-409) and its DDP wrap (train.py:172-174): graph-replayed steps must reproduce eager steps, and with
two ranks (NCCL) the bucketed in-graph all-reduce must leave every rank with the average of the per-rank gradients."""
import copy
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_model(dev, drop_path=0.0):
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
    torch.manual_seed(0)
    m = ResNet_mrlal(MRLA_Bottleneck, [1, 1, 1, 1], num_classes=10, drop_path=drop_path).to(dev)
    for n, p in m.named_parameters():
        if n.endswith("bn3.weight"):
            torch.nn.init.normal_(p, 1.0, 0.2)
    return m.to(memory_format=torch.channels_last).train()


@pytest.mark.parametrize("autocast", [None, torch.bfloat16])
def test_graphed_step_matches_eager(autocast, cuda_device):
    """3 graph-replayed steps == 3 eager steps (same sequence executed without capture): loss of every step and all
    weights / BatchNorm buffers afterwards."""
    from mrla_b200.train import GraphedStep
    dev = cuda_device
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    ma = _small_model(dev)
    mb = copy.deepcopy(ma)
    crit = torch.nn.CrossEntropyLoss()
    g = torch.Generator(device="cpu").manual_seed(3)
    batches = [(torch.randn(8, 3, 64, 64, generator=g).to(dev).contiguous(memory_format=torch.channels_last),
                torch.randint(0, 10, (8,), generator=g).to(dev)) for _ in range(3)]
    losses = []
    for model, capture in ((ma, True), (mb, False)):
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        step = GraphedStep(model, opt, crit, batches[0][0], batches[0][1], autocast_dtype=autocast, capture=capture, warmup=2)
        # the warm-up / capture executions above already updated the weights of the captured variant: reset both
        model.load_state_dict(_small_model(dev).state_dict())
        for st in opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()
        ls = []
        for x, y in batches:
            ls.append(float(step(x, y).detach()))
        losses.append(ls)
        step.close()
    tol = 1e-5 if autocast is None else 2e-3
    for a, b in zip(*losses):
        assert abs(a - b) <= tol * max(1.0, abs(b)), losses
    sa, sb = ma.state_dict(), mb.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point:
            d = (sa[k] - sb[k]).abs().max().item()
            assert d <= (2e-4 if autocast is None else 5e-3) * max(1.0, sb[k].abs().max().item()), k   # cuDNN picks algorithms per call site; 3 SGD steps amplify the last-bit differences
        else:
            assert torch.equal(sa[k], sb[k]), k


@pytest.mark.parametrize("overlap", ["0", "1"])
def test_graphed_step_two_ranks_nccl(overlap):
    """torchrun --nproc-per-node 2 tests/dist_graphed_step.py (needs two GPUs): the captured flat all-reduce (default) and
    the bucketed all-reduce overlapped with backward."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if os.environ.get("MRLA_RUN_DIST_TESTS", "0") != "1":
        pytest.skip("set MRLA_RUN_DIST_TESTS=1 (spawns torchrun with 2 ranks; run it through `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29533 + int(overlap)), os.path.join(ROOT, "tests", "dist_graphed_step.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, MRLA_TEST_OVERLAP=overlap))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_GRAPHED_STEP_OK" in r.stdout
