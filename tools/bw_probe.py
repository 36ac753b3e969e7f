"""Write-only / read-only / copy bandwidth of this GPU with plain ATen kernels (context for the
store-stream kernels' rooflines).   python tools/bw_probe.py"""
import torch

n = 1 << 28                      # 1 GiB of fp32
a = torch.empty(n, device='cuda')
b = torch.empty(n, device='cuda')


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t = timed(lambda: a.fill_(1.0)); print('write-only  %.0f GB/s' % (4 * n / t / 1e9))
t = timed(lambda: torch.cuda.memset if False else a.zero_()); print('memset      %.0f GB/s' % (4 * n / t / 1e9))
t = timed(lambda: a.sum()); print('read-only   %.0f GB/s' % (4 * n / t / 1e9))
t = timed(lambda: b.copy_(a)); print('copy        %.0f GB/s (read + write)' % (8 * n / t / 1e9))
t = timed(lambda: torch.add(a, 1.0, out=b)); print('add         %.0f GB/s (read + write)' % (8 * n / t / 1e9))
