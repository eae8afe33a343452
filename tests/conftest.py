import gzip
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(kind=None):
    names = sorted(f[:-6] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt.gz"))
    if kind is not None:
        names = [n for n in names if n.startswith(kind)]
    return names


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".pt.gz"), "rb") as f:
        return torch.load(f, map_location="cpu", weights_only=False)


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the parity measure used throughout)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def rel_err_l2(a, b):
    """||a-b||_2 / ||b||_2 — used for bf16 paths gated by a ReLU, where a sign flip of a near-zero
    pre-activation (pure rounding) moves single elements by O(1) and makes the max-norm meaningless."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
