"""Multi-GPU plumbing: replicas only.

Stereo pairs are independent (InstanceNorm has no cross-sample statistics), so
the path shards by batch with NO collective on the hot path.
The only communication is one broadcast of the parameter blob (2 217 717 fp32 =
8.9 MB) from rank 0 at start-up -- NCCL over NVLink 5 / NVSwitch on the GPU
box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [begin, end) slice of `total` items owned by `rank`; ranks with
    index < total % world get one extra item."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_parameters(module, src=0, group=None):
    """ONE collective: every parameter and buffer of `module` flattened into a
    single blob, broadcast from `src`, scattered back.  Returns the blob size in
    bytes."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return 0
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    offset = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[offset:offset + n].view_as(t).to(t.dtype))
        offset += n
    return flat.numel() * 4


def gather_results(local, group=None):
    """Optional final gather of per-rank disparity maps on rank 0 (outside any
    timed region).  `local` is [b_local, H, W]; returns the concatenation on rank
    0 and None elsewhere."""
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.size(0)], dtype=torch.int64, device=local.device),
                    group=group)
    out = []
    for r in range(world):
        buf = local if r == dist.get_rank(group) else \
            local.new_empty((int(sizes[r].item()),) + tuple(local.shape[1:]))
        dist.broadcast(buf, src=r, group=group)
        out.append(buf)
    return torch.cat(out) if dist.get_rank(group) == 0 else None
