import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from practicaldeepstereo_nips2018_b200 import PdsNetwork
from practicaldeepstereo_nips2018_b200.pipeline import HostPipeline
torch.manual_seed(0)
net = PdsNetwork.default(191).cuda().eval()
pairs = [(torch.rand(1, 3, 540, 960).cuda() * 255, torch.rand(1, 3, 540, 960).cuda() * 255) for _ in range(4)]
def timed(name, fn, n=60):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); fn(n); b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f'{name:40s} {a.elapsed_time(b) / n:7.3f} ms/pair (device)  host enqueue {1e3 * (t1 - t0) / n:7.3f} ms/pair', flush=True)
with torch.no_grad():
    for i in range(5):
        net(*pairs[i % 4])
    def loop(n):
        for i in range(n):
            net(*pairs[i % 4])
    timed('plain loop, default stream', loop)
    timed('plain loop, default stream (again)', loop)
    outs = []
    def loop_keep(n):
        for i in range(n):
            outs.append(net(*pairs[i % 4]))
    timed('plain loop keeping outputs', loop_keep)
    s = torch.cuda.Stream()
    def loop_side(n):
        with torch.cuda.stream(s):
            for i in range(n):
                net(*pairs[i % 4])
        torch.cuda.current_stream().wait_stream(s)
    timed('plain loop, side stream', loop_side)
    timed('plain loop, side stream (again)', loop_side)
    p1 = HostPipeline(net, streams=1)
    timed('HostPipeline streams=1', lambda n: p1.run((pairs[i % 4] for i in range(n)), download=False))
    p4 = HostPipeline(net, streams=4)
    timed('HostPipeline streams=4', lambda n: p4.run((pairs[i % 4] for i in range(n)), download=False))
    timed('HostPipeline streams=4 (again)', lambda n: p4.run((pairs[i % 4] for i in range(n)), download=False))
    timed('HostPipeline streams=1 (again)', lambda n: p1.run((pairs[i % 4] for i in range(n)), download=False))
    timed('plain loop, default stream (last)', loop)
