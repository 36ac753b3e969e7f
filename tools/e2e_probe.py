"""Which direction of the host traffic costs the e2e leg at N ranks?  Same protocol as bench.py
(HostPipeline, 4 streams, CUDA graphs, barrier + max over ranks), four legs x 3 repeats:
value (device inputs, device outputs), upload only, download only, e2e (both).
torchrun --nproc-per-node N tools/e2e_probe.py"""
import os
import sys
import json

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    rank, world, local = bench.env_int('RANK', 0), bench.env_int('WORLD_SIZE', 1), bench.env_int('LOCAL_RANK', 0)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from practicaldeepstereo_nips2018_b200 import PdsNetwork
    from practicaldeepstereo_nips2018_b200.pipeline import HostPipeline
    H, W, md, _ = bench.WORKLOADS['C2']
    torch.manual_seed(0)
    net = PdsNetwork.default(md).to(dev).eval()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    steps = int(os.environ.get('STEPS', 30))
    pairs = bench.synthetic_pairs(4, 1, H, W, dev, seed=1000 + rank)
    host = bench.synthetic_pairs(4, 1, H, W, dev, seed=2000 + rank, pinned=True, images='u8')
    d2h = [torch.empty((1, H, W), dtype=torch.float32).pin_memory() for _ in range(8)]
    rows = {}
    with torch.no_grad():
        for i in range(3):
            net(*pairs[i % 4])
        pipe = HostPipeline(net, dev, streams=4, graphs=True)
        for rep in range(3):
            for name, items, kw in (('value', pairs, dict(download=False)), ('upload', host, dict(download=False)),
                                    ('download', pairs, dict(out=d2h)), ('e2e', host, dict(out=d2h))):
                ms = bench.time_pipeline(pipe, items, steps, barrier, max_over_ranks, **kw)
                rows.setdefault(name, []).append(round(steps * world / (ms / 1e3), 1))
                rows.setdefault(name + '_host_ms', []).append(round(bench.time_pipeline.host_ms, 1))
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'steps': steps, **rows}))


if __name__ == '__main__':
    main()
