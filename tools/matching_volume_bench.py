"""Micro-benchmark of the generic-operation matching path (a1): pds_matching_concat (shift and
concatenate, matching.py:50-62) and pds_matching_stack (th.stack on the disparity axis,
matching.py:63) at a BASELINE workload, in achieved HBM GB/s against MEASURED_PEAKS.json.  Buffers
rotate so that every launch misses the 126 MB L2.

    python tools/matching_volume_bench.py [--workload C2] [--json profiles/r02_matching_volume_C2.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C2')
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--json', default=None)
args = ap.parse_args()
Hq, Wq, Dq = {'C2': (144, 240, 48), 'C3': (144, 240, 64), 'C4': (96, 320, 48), 'C1': (16, 32, 16)}[args.workload]
B, C, F = 1, 64, 8
dev = torch.device('cuda', 0)
lib = _capi.lib()
peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))
st = _capi.stream_ptr(dev)
rows = {}
for dtype, code, size in ((torch.float32, _capi.PDS_F32, 4), (torch.bfloat16, _capi.PDS_BF16, 2)):
    left = [torch.randn(B, C, Hq, Wq, device=dev).to(dtype) for _ in range(2)]
    right = [torch.randn(B, C, Hq, Wq, device=dev).to(dtype) for _ in range(2)]
    volumes = [torch.empty(B, Dq, 2 * C, Hq, Wq, device=dev, dtype=dtype) for _ in range(2)]   # 849 MB each (fp32, C2)

    def concat(i):
        _capi.check(lib.pds_matching_concat(_capi.ptr(left[i % 2]), _capi.ptr(right[i % 2]),
                                            _capi.ptr(volumes[i % 2]), B, C, Hq, Wq, Dq, code, st))
    sig = [torch.randn(B * Dq, F, Hq, Wq, device=dev).to(dtype) for _ in range(4)]
    stacked = [torch.empty(B, F, Dq, Hq, Wq, device=dev, dtype=dtype) for _ in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def stack(i):
        _capi.check(lib.pds_matching_stack(_capi.ptr(sig[i % 4]), _capi.ptr(stacked[i % 4]), B, F, Dq, Hq, Wq,
                                           code, st))
    for name, fn, nbytes, flush_l2 in (
            ('matching_concat', concat, (2 * B * C * Hq * Wq + B * Dq * 2 * C * Hq * Wq) * size, False),
            ('matching_stack', stack, 2 * B * Dq * F * Hq * Wq * size, True)):
        for i in range(3):
            fn(i)
        times = []
        for i in range(args.reps):
            if flush_l2:
                flush.zero_()                       # 53 MB tensors would otherwise sit in L2
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i); b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        times.sort()
        ms = times[len(times) // 2]
        gbs = nbytes / ms / 1e6
        rows[f'{name}_{str(dtype).split(".")[-1]}'] = {
            'median_us': ms * 1e3, 'algorithmic_bytes': nbytes, 'achieved_gbs': gbs,
            'peak_gbs_measured': peaks['hbm_gbs'], 'frac': gbs / peaks['hbm_gbs']}
        print(name, dtype, f'{ms * 1e3:.1f} us  {gbs:.0f} GB/s  {gbs / peaks["hbm_gbs"]:.3f} of measured peak', flush=True)
    del left, right, volumes, sig, stacked, flush
    torch.cuda.empty_cache()
if args.json:
    with open(args.json, 'w') as fh:
        json.dump({'workload': args.workload, 'shape': {'B': B, 'C': C, 'Hq': Hq, 'Wq': Wq, 'Dq': Dq, 'F': F},
                   'rows': rows}, fh, indent=1)
