"""Per-layer kernel times of one network forward (PDS_B200_PROFILE_DETAIL=1 names every conv_tcg
launch by its layer geometry).   PDS_B200_PROFILE_DETAIL=1 python tools/bench_detail.py [workload] [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork, _capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'C2'
precision = sys.argv[2] if len(sys.argv) > 2 else 'fp16x2'
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255), 'C4': (375, 1242, 191)}[wl]
torch.manual_seed(0)
net = PdsNetwork.default(md, precision=precision).cuda().eval()
left, right = torch.rand(1, 3, H, W).cuda() * 255, torch.rand(1, 3, H, W).cuda() * 255
reps = 5
with torch.no_grad():
    for _ in range(3):
        net(left, right)
    torch.cuda.synchronize()
    _capi.profiler_reset(); _capi.profiler_enable(True)
    for _ in range(reps):
        net(left, right)
    torch.cuda.synchronize()
    _capi.profiler_enable(False)
rep = _capi.profiler_report()
tot = 0.0
for name, (n, ms, fl, by) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    tot += ms / reps
    print('%-58s n=%4.1f %8.1f us/launch %8.3f ms/step %8.1f TF/s %7.0f GB/s' % (
        name, n / reps, ms / n * 1e3, ms / reps, fl / ms / 1e9 if ms else 0, by / ms / 1e6 if ms else 0))
print('total %.3f ms/step' % tot)
