import sys, torch
import os; R = os.environ.get('MRLA_ROOT', '/root/repo'); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
from mrla_b200.resnet_mrla_base import MRLA_Bottleneck, ResNet_mrlab
from oracle.resnet_oracle import ResNetMrlabOracle
from conftest import rel_err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
def run(seed):
    torch.manual_seed(seed)
    prod = ResNet_mrlab(MRLA_Bottleneck, [2, 2, 1, 1], num_classes=10).to(dev).train()
    for n, p in prod.named_parameters():
        if n.endswith("bn3.weight"):
            torch.nn.init.normal_(p, 1.0, 0.2)
    orc = ResNetMrlabOracle([2, 2, 1, 1], num_classes=10).to(dev).train()
    orc.load_state_dict(prod.state_dict(), strict=True)
    x = torch.randn(4, 3, 96, 96, device=dev).contiguous(memory_format=torch.channels_last)
    prod = prod.to(memory_format=torch.channels_last)
    yp, yo = prod(x), orc(x)
    print('seed', seed, 'y err', rel_err(yp, yo))
    yp.square().sum().backward(); yo.square().sum().backward()
    go = dict(orc.named_parameters())
    scale = max(p.grad.abs().max().item() for p in go.values())
    errs = []
    for n, p in prod.named_parameters():
        err = (p.grad - go[n].grad).abs().max().item() / max(go[n].grad.abs().max().item(), 1e-3 * scale)
        errs.append((err, n))
    errs.sort(reverse=True)
    print(errs[:3])
    l2 = sorted(((( p.grad - go[n].grad).norm() / go[n].grad.norm().clamp_min(1e-3 * scale)).item(), n) for n, p in prod.named_parameters())[::-1]
    print('L2', l2[:3])
for s in (1, 2, 3):
    run(s)
