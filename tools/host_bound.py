"""Is the step host-bound?  Compares the time the Python loop takes to ENQUEUE n forwards with the
time the GPU takes to finish them.   python tools/host_bound.py [workload] [precision] [batch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'C2'
precision = sys.argv[2] if len(sys.argv) > 2 else 'fp16x2'
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255), 'C4': (375, 1242, 191)}[wl]
torch.manual_seed(0)
net = PdsNetwork.default(md, precision=precision).cuda().eval()
left, right = torch.rand(B, 3, H, W).cuda() * 255, torch.rand(B, 3, H, W).cuda() * 255
n = 30
with torch.no_grad():
    for _ in range(5):
        net(left, right)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        net(left, right)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f'{wl} {precision} batch {B}: enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, complete {1e3 * (t2 - t0) / n:.3f} ms/step')
