import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import network_blocks
torch.manual_seed(0)
for shape, offset in (((64, 64, 32, 64), 0.0), ((64, 64, 32, 64), 3.0), ((1, 8, 32, 32, 64), 0.5), ((1, 8, 32, 32, 64), 5.0)):
    C = shape[1]
    x0 = torch.randn(shape, device='cuda') * 1.7 + offset
    g0 = torch.rand(C, device='cuda') + 0.5
    b0 = torch.randn(C, device='cuda')
    probe = torch.randn(shape, device='cuda')
    def run(kind):
        dt = torch.float64 if kind == 'f64' else torch.float32
        x, g, b = (t.detach().to(dt).requires_grad_(True) for t in (x0, g0, b0))
        if kind == 'kernel':
            y = network_blocks._LeakyInstanceNorm.apply(x, g, b, 1e-5, 0.1)
        else:
            y = F.instance_norm(F.leaky_relu(x, 0.1), weight=g, bias=b, eps=1e-5)
        (y * probe.to(dt)).sum().backward()
        return [t.double() for t in (y.detach(), x.grad, g.grad, b.grad)]
    ref, ker, aten = run('f64'), run('kernel'), run('aten')
    rel = lambda a, r: float((a - r).norm() / r.norm())
    print(shape, 'offset', offset, 'kernel:', ' '.join('%.2e' % rel(a, r) for a, r in zip(ker, ref)), '| ATen:', ' '.join('%.2e' % rel(a, r) for a, r in zip(aten, ref)))
