#!/usr/bin/env python
"""The stronger same-GPU baseline of SURVEY.md §2.2: the oracle port of the reference resnet50_mrlal under torch.compile
(inductor), bf16 autocast, channels_last, SGD, batch 256 — bench.py's `--compile-baseline` leg on its own (inductor needs
minutes to compile the training graph, so it is not part of the default bench run)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
t0 = time.time()
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
out = {"what": "oracle port of the reference resnet50_mrlal under torch.compile (inductor default mode), bf16 autocast, "
               "channels_last, SGD, B=256, 5 timed steps after 3 warm-up (compile happens in warm-up)"}
try:
    r = bench.gpu_eager_run(dev, 256, steps=5, warmup=3, drop_path=0.2, compiled=True)
    out.update(value=round(r["img_per_s"], 1), unit="img/s", ms_per_step=round(r["ms_per_step"], 2))
except Exception as e:   # noqa: BLE001
    out["error"] = f"{type(e).__name__}: {str(e)[:400]}"
out["wall_s"] = round(time.time() - t0, 1)
bench.emit(out) if hasattr(bench, 'emit') else print(json.dumps(out))   # bench.py points fd 1 at stderr on import
