import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200 import _lib
from mrla_b200.ops import BaseCfg, base_tail
dev = torch.device("cuda:0")
B, C, HW, d, T, k = 256, 256, 56, 16, 3, 5
dt = torch.bfloat16
mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
xs = [torch.relu(mk()).requires_grad_() for _ in range(T)]
dys = [mk() for _ in range(T)]
Ps = [dict(wq=torch.randn(k, device=dev), wk=torch.randn(k, device=dev), wv=torch.randn(C, 1, 3, 3, device=dev) * 0.3,
           gamma=torch.ones(C, device=dev), beta=torch.zeros(C, device=dev)) for _ in range(T)]
for P in Ps:
    for v in P.values(): v.requires_grad_()
cfg = BaseCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, relu=True, residual=True)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    kk = vv = None; ys = []
    for t in range(T):
        P = Ps[t]
        y, kk, vv = base_tail(xs[t], kk, vv, P["wq"], P["wk"], P["wv"], P["gamma"], P["beta"], torch.zeros(C, device=dev), torch.ones(C, device=dev), None, init_cell=(t == 0), cfg=cfg, cap_hint=T)
        ys.append(y)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    torch.autograd.backward(ys, dys)
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"it{it}: fwd cpu {1e3*(t1-t0):.2f} ms, fwd total {1e3*(t2-t0):.2f}; bwd cpu {1e3*(t3-t2):.2f}, bwd total {1e3*(t4-t2):.2f}; mem {torch.cuda.memory_allocated()/1e9:.1f} GB reserved {torch.cuda.memory_reserved()/1e9:.1f}")
    for x in xs: x.grad = None
