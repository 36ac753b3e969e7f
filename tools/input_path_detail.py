import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork, _capi
torch.manual_seed(0)
net = PdsNetwork.default(191, precision='fp16x2').cuda().eval()
for kind in ('f32', 'u8nhwc', 'u8nchw'):
    if kind == 'f32':
        l, r = torch.rand(1, 3, 540, 960).cuda() * 255, torch.rand(1, 3, 540, 960).cuda() * 255
    elif kind == 'u8nhwc':
        l, r = (torch.rand(1, 540, 960, 3) * 255).to(torch.uint8).cuda(), (torch.rand(1, 540, 960, 3) * 255).to(torch.uint8).cuda()
    else:
        l, r = (torch.rand(1, 3, 540, 960) * 255).to(torch.uint8).cuda(), (torch.rand(1, 3, 540, 960) * 255).to(torch.uint8).cuda()
    with torch.no_grad():
        for _ in range(3): net(l, r)
        torch.cuda.synchronize()
        _capi.profiler_reset(); _capi.profiler_enable(True)
        for _ in range(5): net(l, r)
        torch.cuda.synchronize(); _capi.profiler_enable(False)
    rep = _capi.profiler_report()
    print(kind, {k: round(v[1] / v[0] * 1e3, 1) for k, v in rep.items() if k.startswith('image')}, 'total %.3f' % (sum(v[1] for v in rep.values()) / 5))
