"""Runs one layer case of tests/test_gpu_tcg.py (for compute-sanitizer / timing on the GPU box).
    python tools/debug_tcg.py <case index> [S fp16]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import test_gpu_tcg as T  # noqa: E402

if sys.argv[1] == 'custom':      # custom kind nd cin cout n Z Y X   (PDS_B200_TCG_TRACE=1 prints cycle stamps)
    kind, nd, cin, cout, n = (int(a) for a in sys.argv[2:7])
    spatial = tuple(int(a) for a in sys.argv[7:7 + 3])[3 - nd:]
    S, fp16 = 2, 1
else:
    kind, nd, cin, cout, n, spatial = T.CASES[int(sys.argv[1])]
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    fp16 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
k = {0: 3, 1: 3, 2: 4, 3: 5, 4: 4}[kind]
g = torch.Generator(device='cpu').manual_seed(1)
x = torch.randn((n, cin) + spatial, generator=g).cuda()
wshape = ((cin, cout) if kind in (2, 4) else (cout, cin)) + (k,) * nd
w = (torch.randn(wshape, generator=g) / np.sqrt(cin * k ** nd)).cuda()
b = torch.randn(cout, generator=g).cuda()
t0 = time.time()
try:
    out, stats = T.run_layer(kind, nd, x, w, b, S=S, fp16=fp16)
    torch.cuda.synchronize()
    ref = T.aten(kind, nd, x.double(), w.double(), b.double())
    print('case', sys.argv[1], 'ok in %.2fs' % (time.time() - t0), 'max err',
          float((out - ref).abs().max()), 'scale', float(ref.abs().max()))
except Exception as e:
    print('case', sys.argv[1], 'FAILED after %.2fs:' % (time.time() - t0), str(e)[:300])
