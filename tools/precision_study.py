"""Error of every arithmetic option against an fp64 run of the same network.

Runs PdsNetwork.forward stage by stage on the GPU at a BASELINE workload and
prints, for each precision of the convolution stacks (fp16x2 -- the default --,
bf16x3, fp32 CUDA cores, bf16x2, bf16, fp16) and for plain ATen fp32 (the
reference's own arithmetic, TF32 off), the max-abs / mean-abs error of the
matching signatures and of the cost volume against the fp64 restatement
(oracle/torch_port.py in double), the arg-max flip fraction, and the max-abs error
of the final disparity on margin-safe pixels (margin > 4 x the row's cost error) and
on every pixel whose arg-max agrees.  The ATen-fp32 row is the noise floor any "fp32" claim
has to be measured against.

    python tools/precision_study.py [--workload C2] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port  # noqa: E402
from practicaldeepstereo_nips2018_b200 import PdsNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C2')
ap.add_argument('--json', default=None)
ap.add_argument('--precisions', default='fp16x2,bf16x3,fp32,bf16x2,bf16,fp16')
args = ap.parse_args()
H, W, md = {'C1': (64, 128, 63), 'C2': (540, 960, 191), 'C3': (540, 960, 255),
            'C4': (375, 1242, 191), 'S': (256, 512, 127)}[args.workload]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'

torch.manual_seed(0)
net = PdsNetwork.default(md, precision='fp32').to(dev).eval()
g = torch.Generator().manual_seed(7)
left = torch.rand(1, 3, H, W, generator=g) * 255
right = torch.rand(1, 3, H, W, generator=g) * 255
right[..., :-11] = 0.8 * left[..., 11:] + 0.2 * right[..., :-11]
left, right = left.to(dev), right.to(dev)
p32 = {k: v.detach() for k, v in net.state_dict().items()}
p64 = {k: v.double() for k, v in p32.items()}

with torch.no_grad():
    ref = torch_port.network_stages(left.double(), right.double(), p64, md)
    rows = {}

    top2 = ref['cost'].topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    ph, pw = -H % 64, -W % 64

    def report(name, sig, cost):
        se = (sig.double() - ref['signatures']).abs()
        ce = (cost.double() - ref['cost']).abs()
        idx = cost.argmax(dim=1)
        flipped = idx != ref['argmax']
        flips = flipped.double().mean().item()
        safe = margin > 4 * ce.max()
        safe_flips = (flipped & safe).double().sum().item()
        # final disparity (SubpixelMap on this row's own cost volume, torch port) against the fp64
        # run's, on the un-padded image: max-abs on margin-safe pixels and wherever the arg-max agrees
        disp = torch_port.subpixel_map(cost.float())[0][..., ph:, pw:].double()
        de = (disp - ref['disparity']).abs()
        safe_c, agree_c = safe[..., ph:, pw:], ~flipped[..., ph:, pw:]
        d_safe = de[safe_c].max().item() if safe_c.any() else 0.0
        d_agree = de[agree_c].max().item()
        rows[name] = {'sig_max': se.max().item(), 'sig_mean': se.mean().item(),
                      'cost_max': ce.max().item(), 'cost_mean': ce.mean().item(),
                      'argmax_flip_frac': flips, 'safe_frac': safe.double().mean().item(),
                      'flips_on_safe_pixels': safe_flips,
                      'disparity_max_safe': d_safe, 'disparity_max_agree': d_agree,
                      'disparity_mean': de.mean().item()}
        print(f'{name:10s} signatures max {se.max().item():.3e} mean {se.mean().item():.3e} | '
              f'cost max {ce.max().item():.3e} mean {ce.mean().item():.3e} | argmax flips '
              f'{flips:.3e} (safe pixels {safe.double().mean().item():.4f}, flips there {int(safe_flips)}) | '
              f'disparity max-abs: safe pixels {d_safe:.3e}, arg-max-agreeing pixels {d_agree:.3e}, '
              f'mean over all {de.mean().item():.3e}', flush=True)

    print(f'workload {args.workload}: {W}x{H} md={md}; reference = fp64 ATen; signature scale '
          f'{ref["signatures"].abs().max().item():.2f}, cost scale {ref["cost"].abs().max().item():.2f}')
    st = torch_port.network_stages(left, right, p32, md)
    report('aten-fp32', st['signatures'], st['cost'])
    del st
    for prec in args.precisions.split(','):
        n = PdsNetwork.default(md, precision=prec).to(dev).eval()
        n.load_state_dict(net.state_dict())
        lp, rp = n._size_adapter.pad(left), n._size_adapter.pad(right)
        ld, rd, sc = n._embed(lp, rp)
        sig = n._matching(ld, rd)
        cost = n._regularization(sig, sc)
        report(prec, sig, cost)
        # the matching stage alone, on the fp64 run's own descriptors
        sig2 = n._matching(ref['left_descriptor'].float(), ref['right_descriptor'].float())
        e = (sig2.double() - ref['signatures']).abs()
        rows[prec]['sig_only_max'] = e.max().item()
        rows[prec]['sig_only_mean'] = e.mean().item()
        print(f'{"":10s} matching alone on identical descriptors: max {e.max().item():.3e} '
              f'mean {e.mean().item():.3e}', flush=True)
        del n, sig, cost, sig2
if args.json:
    with open(args.json, 'w') as fh:
        json.dump({'workload': args.workload, 'rows': rows}, fh, indent=1)
