#!/usr/bin/env python
"""torch.profiler breakdown of one resnet50_mrlal training step (kernel time by name)."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200.resnet_mrla_light import resnet50_mrlal
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
model = resnet50_mrlal(drop_path=0.2).to(dev).to(memory_format=torch.channels_last).train()
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
x = torch.randn(256, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (256,), device=dev)
crit = nn.CrossEntropyLoss()
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x)
    loss = crit(out.float(), y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted([(e.device_time_total / 3e3, e.count // 3, e.key) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"], reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device kernel time per step: {tot:.2f} ms")
for t, n, k in rows[:40]:
    print(f"{t:8.3f} ms  x{n:4d}  {k[:110]}")
