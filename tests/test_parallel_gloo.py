"""N > 1 host logic on CPU: world_size-2 gloo processes -- weight broadcast,
batch sharding, result gather (replicas only, no hot-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from practicaldeepstereo_nips2018_b200 import PdsNetwork, parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(100 + rank)               # different init on every rank
    net = PdsNetwork.default(63)
    before = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    nbytes = parallel.broadcast_parameters(net, src=0)
    after = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    changed = not torch.equal(before, after)
    b, e = parallel.shard_range(7, rank, world)
    local = torch.arange(b, e, dtype=torch.float32).view(-1, 1, 1).expand(-1, 2, 3).contiguous()
    full = parallel.gather_results(local)
    ok_gather = True
    if rank == 0:
        ok_gather = torch.equal(full[:, 0, 0], torch.arange(7, dtype=torch.float32))
    ret[rank] = (nbytes, same, changed, (b, e), ok_gather)
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == ret[1][0] == 2217717 * 4
    assert ret[0][1] and ret[1][1]              # identical parameters everywhere
    assert not ret[0][2] and ret[1][2]          # rank 0 kept its weights, rank 1 received them
    assert ret[0][3] == (0, 4) and ret[1][3] == (4, 7)
    assert ret[0][4]


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
