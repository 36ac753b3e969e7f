"""Repeats the value / e2e loops of bench.py several times in one process (run-to-run variance).
    python tools/pipeline_variance.py [workload] [precision] [streams] [steps] [repeats]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork  # noqa: E402
from practicaldeepstereo_nips2018_b200.pipeline import HostPipeline  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'C2'
precision = sys.argv[2] if len(sys.argv) > 2 else 'fp16x2'
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
repeats = int(sys.argv[5]) if len(sys.argv) > 5 else 6
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255), 'C4': (375, 1242, 191)}[wl]
torch.manual_seed(0)
net = PdsNetwork.default(md, precision=precision).cuda().eval()
dev_pairs = [(torch.rand(1, 3, H, W).cuda() * 255, torch.rand(1, 3, H, W).cuda() * 255) for _ in range(4)]
host_pairs = [((torch.rand(1, 3, H, W) * 255).pin_memory(), (torch.rand(1, 3, H, W) * 255).pin_memory()) for _ in range(4)]
pipe = HostPipeline(net, streams=streams)
d2h = [torch.empty((1, H, W)).pin_memory() for _ in range(2 * streams)]
pipe.run([dev_pairs[i % 4] for i in range(6)], download=False)
pipe.run([host_pairs[i % 4] for i in range(6)], out=d2h)
torch.cuda.synchronize()
for rep in range(repeats):
    res = []
    for name, pairs, kw in (('value', dev_pairs, dict(download=False)), ('e2e', host_pairs, dict(out=d2h))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = pipe.run((pairs[i % 4] for i in range(steps)), **kw)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res.append(f'{name} {steps / (t2 - t0):7.1f} pairs/s (enqueue {1e3 * (t1 - t0) / steps:.2f} ms/step)')
        del outs
    print(f'rep {rep}: ' + ' | '.join(res), flush=True)
print('max memory allocated %.1f GB, reserved %.1f GB' % (torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
