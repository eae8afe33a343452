#!/usr/bin/env python
"""Device time of the folded MRLA-light tail op (forward / backward kernel groups, CUDA-graph replays between CUDA
events) at the four ResNet-50 stage shapes — the measurement bench.py's `roofline` object uses, without the model."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
peak, _ = bench.peaks()
res = {}
# "bn3_folded": z = raw conv3 output + bn3 coefficients (sweep-1 MODE 6, what resnet50_mrlal runs);
# "bn3_separate": z = bn3 output (sweep-1 MODE 5)
for name, bn3 in (("bn3_folded", True), ("bn3_separate", False)):
    out = {}
    tot_b = tot_ms = 0.0
    for (C, HW), n in bench.STAGE_BLOCKS.items():
        f, b = bench.measure_tail_group(B, C, HW, dev, iters=20, bn3=bn3)
        nbytes = 9.0 * B * C * HW * HW * 2
        out[f"{C}x{HW}"] = dict(fwd_ms=round(f, 4), bwd_ms=round(b, 4), GBps=round(nbytes / (f + b) / 1e6, 1))
        tot_b += n * nbytes
        tot_ms += n * (f + b)
    out["all_16"] = dict(ms=round(tot_ms, 3), GBps=round(tot_b / tot_ms / 1e6, 1), frac=round(tot_b / tot_ms / 1e6 / peak, 4))
    res[name] = out
os.write(bench._REAL_STDOUT, (json.dumps(res) + "\n").encode())
