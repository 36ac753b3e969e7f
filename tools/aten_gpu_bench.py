"""Stock PyTorch on the same B200: the reference forward (oracle/torch_port.py = the reference's
operator sequence on ATen / cuDNN, cudnn.benchmark=True as trainer.py:32-34 sets it) timed with the
reference's own protocol (one forward bracketed by cuda.synchronize(), trainer.py:141-148), TF32
off and TF32 on, next to this package's PdsNetwork.forward under the same protocol.  SURVEY.md 8(d)
calls this "the real bar".  A tool run, not the reference arm of bench.py.

    python tools/aten_gpu_bench.py [--workload C2] [--reps 10] [--json profiles/r02_aten_gpu_C2.json]
"""
import argparse
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, torch_port  # noqa: E402
from practicaldeepstereo_nips2018_b200 import PdsNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C2')
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--json', default=None)
args = ap.parse_args()
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255), 'C4': (375, 1242, 191)}[args.workload]
dev = torch.device('cuda', 0)
torch.backends.cudnn.benchmark = True
params = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_params(synth.network_specs(), 61).items()}
g = torch.Generator().manual_seed(5)
left = (torch.rand(1, 3, H, W, generator=g) * 255).to(dev)
right = (torch.rand(1, 3, H, W, generator=g) * 255).to(dev)


def protocol(fn, reps):
    for _ in range(3):
        fn()
    lat = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return {'latency_ms_median': statistics.median(lat), 'latency_ms_min': min(lat),
            'back_to_back_ms': start.elapsed_time(stop) / reps,
            'pairs_per_s_sync_protocol': 1e3 / statistics.median(lat)}


rows = {}
with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        name = 'aten_cudnn_tf32' if tf32 else 'aten_cudnn_fp32'
        rows[name] = protocol(lambda: torch_port.network_forward(left, right, params, md), args.reps)
        print(name, json.dumps(rows[name]), flush=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for precision in ('fp16x2', 'bf16'):
        net = PdsNetwork.default(md, precision=precision)
        net.load_state_dict({k: v.cpu() for k, v in params.items()})
        net = net.to(dev).eval()
        rows[f'pds_b200_{precision}'] = protocol(lambda: net(left, right), max(args.reps, 20))
        print(f'pds_b200_{precision}', json.dumps(rows[f'pds_b200_{precision}']), flush=True)
        del net
out = {'workload': args.workload, 'gpu': torch.cuda.get_device_name(0), 'torch': torch.__version__,
       'cudnn': torch.backends.cudnn.version(), 'reps': args.reps,
       'protocol': 'one forward bracketed by torch.cuda.synchronize() (trainer.py:141-148), 3 warm-ups, '
                   'median; back_to_back_ms = CUDA events around `reps` consecutive forwards',
       'rows': rows}
if args.json:
    with open(args.json, 'w') as fh:
        json.dump(out, fh, indent=1)
