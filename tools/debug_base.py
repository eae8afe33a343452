import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200 import _lib
from mrla_b200.ops import BaseCfg, base_tail
from oracle import mrla_oracle as O
dev = torch.device("cuda:0")
def run(dt, B, C, HW, d, T, k, layout="nchw", chain=False):
    torch.manual_seed(11)
    mk = lambda: torch.randn(B, C, HW, HW, device=dev)
    xs = [torch.relu(mk()).to(dt).requires_grad_() for _ in range(T)]
    dys = [mk().to(dt) for _ in range(T)]
    Ps = [dict(wq=torch.randn(k, device=dev) * 0.5, wk=torch.randn(k, device=dev) * 0.5,
               wv=torch.randn(C, 1, 3, 3, device=dev) * 0.47, gamma=1 + 0.3 * torch.randn(C, device=dev),
               beta=0.2 * torch.randn(C, device=dev)) for _ in range(T)]
    for P in Ps:
        for v_ in P.values(): v_.requires_grad_()
    kk = vv = None; ys = []
    for t in range(T):
        P = Ps[t]
        cfg = BaseCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, relu=True, residual=True)
        y, kk, vv = base_tail(xs[t], kk, vv, P["wq"], P["wk"], P["wv"], P["gamma"], P["beta"],
                              torch.zeros(C, device=dev), torch.ones(C, device=dev), None, init_cell=(t == 0), cfg=cfg)
        ys.append(y)
    torch.autograd.backward(ys, dys)
    xd = [x.detach().double().requires_grad_() for x in xs]
    Pd = [{n: v_.detach().double().requires_grad_() for n, v_ in P.items()} for P in Ps]
    kr = vr = None; yr = []
    for t in range(T):
        P = Pd[t]
        y, kr, vr, _, _ = O.base_tail(xd[t], kr, vr, P["wq"], P["wk"], P["wv"], C // d, t == 0, P["gamma"], P["beta"],
                                      torch.zeros(C, dtype=torch.float64, device=dev), torch.ones(C, dtype=torch.float64, device=dev))
        yr.append(y)
    torch.autograd.backward(yr, [g_.double() for g_ in dys])
    for t in range(T):
        e = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
        print(dt, f"t={t} y {e(ys[t], yr[t]):.2e} dx {e(xs[t].grad, xd[t].grad):.2e}", " ".join(f"{n} {e(Ps[t][n].grad, Pd[t][n].grad):.2e}" for n in Ps[t]))
run(torch.float32, 8, 1024, 14, 16, 6, 5)
run(torch.bfloat16, 8, 1024, 14, 16, 6, 5)
run(torch.bfloat16, 64, 1024, 14, 16, 6, 5)
