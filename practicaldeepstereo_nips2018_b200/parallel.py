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
    tensors = list(module.parameters()) + list(module.buffers())
    if not tensors:
        return 0
    with torch.no_grad():
        flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
        dist.broadcast(flat, src=src, group=group)
        offset = 0
        for t in tensors:
            n = t.numel()
            # copy_ on the parameter itself (not .data): bumps the version counter the kernel
            # handles key their packed weights on
            t.copy_(flat[offset:offset + n].view_as(t).to(t.dtype))
            offset += n
    for m in module.modules():                     # belt and braces: drop packed weights explicitly
        handle = m.__dict__.get('_kernel')
        if handle is not None:
            handle.invalidate()
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


def _parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (sysfs cpulist format)."""
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(pci_bus_id, sysfs='/sys'):
    """CPUs of the NUMA node a GPU hangs off ('0000:1b:00.0' -> cpu list), or None when the
    platform does not say (single-node hosts report -1)."""
    import os
    try:
        with open(os.path.join(sysfs, 'bus/pci/devices', pci_bus_id.lower(), 'numa_node')) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, 'devices/system/node', f'node{node}', 'cpulist')) as fh:
            return _parse_cpulist(fh.read()) or None
    except (OSError, ValueError):
        return None


def bind_to_gpu_numa_node(device_index):
    """Pins the calling process (and the pinned host buffers it allocates afterwards: first
    touch) to the CPUs next to `device_index`'s PCIe root, so that each replica's H2D / D2H
    copies do not cross the socket interconnect.  Returns the CPU list, or None if unknown."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = device_index
        if visible:
            entry = visible.split(',')[device_index].strip()
            if entry.isdigit():
                index = int(entry)
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        if len(bus.split(':')[0]) == 8:          # NVML pads the PCI domain to 8 hex digits
            bus = bus[4:]
        cpus = gpu_numa_cpus(bus)
        if cpus and hasattr(os, 'sched_setaffinity'):
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                return allowed
    except Exception:
        return None
    return None
