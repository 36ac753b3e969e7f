import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork, loss as pds_loss, matching
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
left = torch.rand(1, 3, 128, 256, device='cuda') * 255
right = torch.rand(1, 3, 128, 256, device='cuda') * 255
gt = torch.rand(1, 128, 256, device='cuda') * 120
torch.manual_seed(0)
net = PdsNetwork.default(127).cuda().train()
def run(kernels):
    matching.USE_TRAINING_KERNELS = kernels
    net.zero_grad()
    hook = {}
    h = net._matching.register_forward_hook(lambda m, i, o: hook.setdefault('sig', o.detach().clone()))
    cost = net(left, right)
    h.remove()
    v = pds_loss.SubpixelCrossEntropy()(cost, gt)
    v.backward()
    return hook['sig'], cost.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
def cmp(a, b, tag):
    top = max(float(g.abs().max()) for g in b[2].values())
    worst, wk = 0, None
    for k in a[2]:
        scale = max(float(b[2][k].abs().max()), 1e-3 * top)
        rel = float((a[2][k] - b[2][k]).abs().max()) / scale
        if rel > worst: worst, wk = rel, k
    print(tag, 'sig diff %.3e (scale %.2f)' % (float((a[0]-b[0]).abs().max()), float(b[0].abs().max())),
          'cost diff %.3e (scale %.2f)' % (float((a[1]-b[1]).abs().max()), float(b[1].abs().max())), 'worst grad rel %.3e' % worst, wk)
print('no grad:', [k for k, p in net.named_parameters() if p.grad is None][:5])
l1, l2, k1, k2 = run(False), run(False), run(True), run(True)
cmp(l1, l2, 'loop vs loop    ')
cmp(k1, k2, 'kernel vs kernel')
cmp(k1, l1, 'kernel vs loop  ')
rows = sorted(((float((k1[2][k] - l1[2][k]).abs().max()), float(l1[2][k].abs().max()), float(l1[2][k].norm()), float((k1[2][k] - l1[2][k]).norm()), k) for k in l1[2]), key=lambda r: -r[0] / max(r[1], 1e-12))
for r in rows[:8]: print('  abs diff %.3e own max %.3e own norm %.3e diff norm %.3e %s' % r)
sys.exit(0)
torch.use_deterministic_algorithms(True, warn_only=True)
torch.backends.cudnn.benchmark = False
d1, d2, e1 = run(False), run(False), run(True)
cmp(d1, d2, 'deterministic loop vs loop  ')
cmp(e1, d1, 'deterministic kernel vs loop')
