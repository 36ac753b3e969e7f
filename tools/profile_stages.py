"""Runs each stage of the hot path a few times at a BASELINE workload size, for
use under ncu (launch lists / --set full captures) and for quick stage timings.

    python tools/profile_stages.py [--workload C2] [--precision fp32] [--reps 3]
                                   [--stages estimator,concat,matching,regularization,network]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork, matching  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C2')
ap.add_argument('--precision', default='fp32')
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--batch', type=int, default=1)
ap.add_argument('--stages', default='estimator,concat,matching,regularization,network')
ap.add_argument('--nvtx', action='store_true',
                help='wrap ONE extra call of the last stage in the NVTX range "capture" '
                     '(ncu --nvtx --nvtx-include "capture/")')
args = ap.parse_args()
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255),
            'C4': (375, 1242, 191)}[args.workload]
Hp, Wp = H + (-H) % 64, W + (-W) % 64
Hq, Wq, Dq, Dc = Hp // 4, Wp // 4, (md + 1) // 4, (md + 1) // 2
B = args.batch
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
net = PdsNetwork.default(md, precision=args.precision).cuda().eval()
dev = 'cuda'


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    print(f'{name:16s} {ev[0].elapsed_time(ev[1]) / args.reps:9.3f} ms', flush=True)


with torch.no_grad():
    stages = args.stages.split(',')
    if 'estimator' in stages:
        cost = torch.randn(B, Dc, Hp, Wp, device=dev)
        timed('estimator', lambda: net._estimator(cost, crop_top=Hp - H, crop_left=Wp - W))
        del cost
    ld, rd = torch.randn(B, 64, Hq, Wq, device=dev), torch.randn(B, 64, Hq, Wq, device=dev)
    if 'concat' in stages:
        m = matching.Matching(Dq - 1, lambda x: x[:, :8])
        timed('concat(generic)', lambda: m(ld, rd))
    if 'matching' in stages:
        timed('matching', lambda: net._matching(ld, rd))
    if 'regularization' in stages:
        sig, sc = torch.randn(B, 8, Dq, Hq, Wq, device=dev), torch.randn(B, 8, Hq, Wq, device=dev)
        timed('regularization', lambda: net._regularization(sig, sc))
        del sig, sc
    if 'network' in stages:
        left, right = torch.rand(B, 3, H, W, device=dev) * 255, torch.rand(B, 3, H, W, device=dev) * 255
        timed('embedding', lambda: net._embed(net._size_adapter.pad(left), net._size_adapter.pad(right)))
        timed('network', lambda: net(left, right))
        if args.nvtx:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push('capture')
            net(left, right)
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
