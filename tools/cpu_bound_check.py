import os, sys, time, torch, torch.nn as nn
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
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time per step {1e3*(t1-t0)/10:.2f} ms ; wall per step {1e3*(t2-t0)/10:.2f} ms")
