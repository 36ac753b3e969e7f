"""f4: one training step of PdsNetwork (forward, SubpixelCrossEntropy, backward, RMSprop step --
train_on_flyingthings3d.py:56-70, trainer.py:207-227) at 960x540, md = 255, batch 1, on one B200:

  reference composition  -- per-disparity Python loop over MatchingOperation (matching.py:53-63) and
                            the loss as tensor expressions (loss.py:30-78): what the reference runs;
  training kernels       -- Matching as volume kernel -> ONE batched operation call -> stack kernel
                            (adjoint kernels in the backward) + fused SubpixelCrossEntropy kernels.

Convolutions / InstanceNorm are ATen (cuDNN, TF32 off) under autograd in both.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port  # noqa: E402
from practicaldeepstereo_nips2018_b200 import PdsNetwork, loss as pds_loss, network_blocks  # noqa: E402


def step_ms(kernels, H, W, md, steps):
    network_blocks.USE_TRAINING_KERNELS = kernels
    torch.manual_seed(0)
    net = PdsNetwork.default(md).cuda().train()
    opt = torch.optim.RMSprop(net.parameters(), lr=1e-2)
    fused = pds_loss.SubpixelCrossEntropy()
    left = torch.rand(1, 3, H, W, device='cuda') * 255
    right = torch.rand(1, 3, H, W, device='cuda') * 255
    gt = torch.rand(1, H, W, device='cuda') * (md - 1)

    def one():
        opt.zero_grad()
        cost = net(left, right)
        value = fused(cost, gt) if kernels else torch_port.subpixel_cross_entropy(cost, gt, None, 1.0, 2)
        value.backward()
        opt.step()
        return value

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    times = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        value = one()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    times.sort()
    out = {'ms_per_step': times[len(times) // 2], 'min_ms': times[0], 'loss': float(value.detach()),
           'peak_memory_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
    del net, opt
    torch.cuda.empty_cache()
    return out


def main():
    torch.backends.cudnn.benchmark = True            # as the reference trainer does (trainer.py:32-34)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    H, W, md = int(os.environ.get('H', 540)), int(os.environ.get('W', 960)), int(os.environ.get('MD', 255))
    steps = int(os.environ.get('STEPS', 5))
    ref = step_ms(False, H, W, md, steps)
    ker = step_ms(True, H, W, md, steps)
    print(json.dumps({'workload': f'training step {W}x{H} md={md} batch 1 (fp32 ATen convolutions, TF32 off)',
                      'steps': steps, 'reference_composition': ref, 'training_kernels': ker,
                      'speedup': ref['ms_per_step'] / ker['ms_per_step']}))


if __name__ == '__main__':
    main()
