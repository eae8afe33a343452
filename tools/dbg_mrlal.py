"""Whole-model gradient agreement of the small resnet_mrlal used by tests/test_model_gpu.py over seeds and layouts:
product (NCHW weights / channels_last) against the oracle run in NCHW and in channels_last, and the oracle against itself
across the two layouts (the noise floor of a chaotic fp32 network: cuDNN picks different algorithms per layout)."""
import sys, torch
import os; R = os.environ.get('MRLA_ROOT', '/root/repo'); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
from oracle.resnet_oracle import ResNetMrlalOracle
from conftest import rel_err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')


def grads(model, x):
    for p in model.parameters():
        p.grad = None
    y = model(x)
    y.square().sum().backward()
    return y.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}


def cmp(tag, a, b):
    ya, ga = a
    yb, gb = b
    scale = max(g.abs().max().item() for g in gb.values())
    mx = sorted((((ga[n] - gb[n]).abs().max() / gb[n].abs().max().clamp_min(1e-3 * scale)).item(), n) for n in gb)[::-1]
    l2 = sorted((((ga[n] - gb[n]).norm() / gb[n].norm().clamp_min(1e-3 * scale)).item(), n) for n in gb)[::-1]
    tot = (sum((ga[n] - gb[n]).square().sum() for n in gb).sqrt() / sum(gb[n].square().sum() for n in gb).sqrt()).item()
    print(f'  {tag}: y {rel_err(ya, yb):.2e}  max-norm worst {mx[0][0]:.2e} {mx[0][1]}  L2 worst {l2[0][0]:.2e} {l2[0][1]} / {l2[1][0]:.2e} {l2[1][1]}  global L2 {tot:.2e}')


for seed in (0, 1, 2, 3):
    torch.manual_seed(seed)
    prod = ResNet_mrlal(MRLA_Bottleneck, [2, 1, 1, 1], num_classes=10).to(dev).train()
    for n, p in prod.named_parameters():
        if n.endswith("bn3.weight"):
            torch.nn.init.normal_(p, 1.0, 0.2)
    orc = ResNetMrlalOracle([2, 1, 1, 1], num_classes=10).to(dev).train()
    orc.load_state_dict(prod.state_dict(), strict=True)
    x = torch.randn(4, 3, 96, 96, device=dev)
    xcl = x.contiguous(memory_format=torch.channels_last)
    sd = {k: v.clone() for k, v in prod.state_dict().items()}
    print('seed', seed)
    o_nchw = grads(orc, x); orc.load_state_dict(sd)
    o_cl = grads(orc, xcl); orc.load_state_dict(sd)
    p_nchw = grads(prod, x); prod.load_state_dict(sd)
    prod_cl = prod.to(memory_format=torch.channels_last)
    p_cl = grads(prod_cl, xcl)
    cmp('oracle NCHW vs oracle CL   ', o_nchw, o_cl)
    cmp('product NCHW-w vs oracle NCHW', p_nchw, o_nchw)
    cmp('product NCHW-w vs oracle CL  ', p_nchw, o_cl)
    cmp('product CL vs oracle CL      ', p_cl, o_cl)
    cmp('product NCHW-w vs product CL ', p_nchw, p_cl)
