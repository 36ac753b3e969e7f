"""Host-buffer inference pipeline: the call a serving loop makes.

The reference moves every example to the GPU and runs the network back to back
(trainer.py:241-244: ``_move_tensors_to_cuda`` then ``network(left, right)``), so the
upload of a pair is serialised with its forward.  ``HostPipeline`` keeps the same
per-pair work -- upload of both images from pinned host memory, ``PdsNetwork.forward``,
download of the disparity map -- but issues the upload of pair i+1 on a copy stream while
pair i computes; results land in pinned host buffers.
"""
import torch

from . import _capi, matching

# kernels launched through graph replays (pds_launch_count only sees launches issued by the host)
_replayed_launches = [0]


def replayed_launches():
    return _replayed_launches[0]


class GraphedNetwork(object):
    """``PdsNetwork.forward`` (eval, no-grad) captured as ONE CUDA graph per input signature and
    replayed: the ~75 kernel launches of a forward become a single graph launch, so the gaps
    between the latency-bound kernels (the deep hourglass levels run a few dozen CTAs for ~10 us
    each) no longer depend on the host.  Inputs are copied into the graph's static buffers; the
    returned disparity is the graph's static output, VALID UNTIL THE NEXT CALL on this object
    (``clone()`` it to keep it).  Scratch memory and outputs live in the graph's private pool.

    The reference runs one forward per ``network(left, right)`` call (pds_trainer.py:35-38); this
    wrapper is called the same way."""

    def __init__(self, network, warmup=2):
        self._network, self._warmup, self._graphs = network, warmup, {}

    def _signature(self, left, right):
        return (tuple(left.shape), left.dtype, tuple(right.shape), right.dtype, str(left.device),
                getattr(self._network, '_maximum_disparity', None))

    def _capture(self, left, right):
        device = left.device
        static_left, static_right = torch.empty_like(left), torch.empty_like(right)
        static_left.copy_(left)
        static_right.copy_(right)
        tag = ('graph', id(self), len(self._graphs))
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.no_grad(), torch.cuda.stream(side), matching.workspace_scope(tag):
            for _ in range(max(1, self._warmup)):      # packs weights, plans layers, sizes the scratch buffers
                self._network(static_left, static_right)
        torch.cuda.current_stream(device).wait_stream(side)
        matching.release_workspaces(self._network, tag)          # re-allocated inside the capture, in the graph's pool
        graph = torch.cuda.CUDAGraph()
        launches = _capi.launch_count()
        with torch.no_grad(), torch.cuda.graph(graph), matching.workspace_scope(tag):
            static_out = self._network(static_left, static_right)
        launches = _capi.launch_count() - launches       # kernels recorded into the graph
        # the scratch buffers allocated during the capture belong to THIS graph from now on (they would
        # otherwise stay referenced by the kernel handles for the life of the network)
        scratch = matching.release_workspaces(self._network, tag)
        entry = (graph, static_left, static_right, static_out, launches, scratch)
        self._graphs[self._signature(left, right)] = entry
        return entry

    def close(self):
        """Drops every captured graph with its static buffers and scratch memory."""
        self._graphs.clear()

    def __call__(self, left_image, right_image):
        if not (left_image.is_cuda and right_image.is_cuda):
            raise RuntimeError('GraphedNetwork needs CUDA tensors')
        entry = self._graphs.get(self._signature(left_image, right_image))
        if entry is None:
            entry = self._capture(left_image.contiguous(), right_image.contiguous())
        graph, static_left, static_right, static_out, launches, _ = entry
        static_left.copy_(left_image, non_blocking=True)
        static_right.copy_(right_image, non_blocking=True)
        graph.replay()
        _replayed_launches[0] += launches
        return static_out


class HostPipeline(object):
    def __init__(self, network, device=None, depth=2, streams=1, graphs=False):
        """depth: uploads issued ahead of the forward that consumes them.  streams > 1: pairs are
        dealt round-robin to that many compute streams, so that the latency-bound layers of one
        pair (the deep hourglass levels occupy a few dozen SMs) overlap the other pair's work;
        every stream has its own scratch memory (matching._KernelHandle.workspace).  graphs: every
        compute stream replays its own CUDA graph of the forward (GraphedNetwork)."""
        self._network = network
        self._graphed = [GraphedNetwork(network) for _ in range(max(1, streams))] if graphs else None
        self._device = device if device is not None else next(network.parameters()).device
        self._copy_stream = torch.cuda.Stream(self._device)
        self._streams = [torch.cuda.Stream(self._device) for _ in range(streams)] if streams > 1 else []
        self._depth = max(1, depth, streams)
        self._download = True
        # staging slots: uploads in flight + pairs being computed
        self._slots = [{'left': None, 'right': None, 'consumed': None}
                       for _ in range(self._depth + max(1, streams) + 1)]
        self._next_slot = 0

    def close(self):
        """Releases the captured graphs (if any) and the staging slots."""
        if self._graphed is not None:
            for g in self._graphed:
                g.close()
        for slot in self._slots:
            slot['left'] = slot['right'] = slot['consumed'] = None

    def _upload(self, pair):
        """H2D of one (left, right) pair on the copy stream into a preallocated staging slot
        (no allocation per pair: the caching allocator would otherwise fall back to cudaMalloc
        whenever a block shared between streams is not yet reusable); returns tensors, the event
        that marks the upload complete and the slot (device-resident pairs are passed through)."""
        if pair[0].is_cuda and pair[1].is_cuda:
            return pair[0], pair[1], None, None
        slot = self._slots[self._next_slot % len(self._slots)]
        self._next_slot += 1
        with torch.cuda.stream(self._copy_stream):
            if (slot['left'] is None or slot['left'].shape != pair[0].shape or slot['right'].shape != pair[1].shape
                    or slot['left'].dtype != pair[0].dtype or slot['right'].dtype != pair[1].dtype):
                # allocated ON the copy stream: a block the caching allocator hands out here was last
                # used on this stream, not by a forward that is still in flight on a compute stream
                slot['left'] = torch.empty(pair[0].shape, dtype=pair[0].dtype, device=self._device)
                slot['right'] = torch.empty(pair[1].shape, dtype=pair[1].dtype, device=self._device)
                slot['consumed'] = None
            if slot['consumed'] is not None:          # the forward that last read this slot is done
                self._copy_stream.wait_event(slot['consumed'])
            slot['left'].copy_(pair[0], non_blocking=True)
            slot['right'].copy_(pair[1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        return slot['left'], slot['right'], ready, slot

    def run(self, host_pairs, out=None, download=True):
        """host_pairs: iterable of (left, right) pinned CPU tensors [B, 3, H, W] float32 -- or uint8
        [B, 3, H, W] / [B, H, W, 3], a quarter of the upload (PdsNetwork.forward) -- (CUDA tensors
        are taken as they are).  out: optional list of pinned CPU tensors [B, H, W] (reused
        round-robin).  Returns the list of disparity tensors, one per pair -- on the host
        (valid after ``torch.cuda.current_stream().synchronize()``), or on the device when
        ``download=False``."""
        self._download = download
        caller = torch.cuda.current_stream(self._device)
        for s in self._streams:                    # work queued by the caller comes first
            s.wait_stream(caller)
        pairs = iter(host_pairs)
        inflight, results, k = [], [], 0
        with torch.no_grad():
            for pair in pairs:
                inflight.append(self._upload(pair))
                if len(inflight) < self._depth:
                    continue
                k = self._dispatch(inflight.pop(0), caller, out, results, k)
            while inflight:
                k = self._dispatch(inflight.pop(0), caller, out, results, k)
        for s in self._streams:                    # the caller's stream sees every result
            caller.wait_stream(s)
        return results

    def _dispatch(self, item, caller, out, results, k):
        if not self._streams:
            return self._step(item, caller, out, results, k, 0)
        lane = k % len(self._streams)
        stream = self._streams[lane]
        with torch.cuda.stream(stream):
            return self._step(item, stream, out, results, k, lane)

    def _step(self, item, compute, out, results, k, lane=0):
        left, right, ready, slot = item
        if ready is not None:
            compute.wait_event(ready)
        if self._graphed is not None:
            disparity = self._graphed[lane](left, right)
            if not self._download:
                disparity = disparity.clone()      # the graph's static output is rewritten by the next replay
        else:
            disparity = self._network(left, right)
        if slot is not None:
            left.record_stream(compute)            # staging memory is not recycled under the forward
            right.record_stream(compute)
            slot['consumed'] = torch.cuda.Event()
            slot['consumed'].record(compute)
        if not self._download:
            disparity.record_stream(torch.cuda.current_stream(self._device))
            results.append(disparity)
            return k + 1
        if out is not None:
            host = out[k % len(out)]
        else:
            host = torch.empty(disparity.shape, dtype=disparity.dtype).pin_memory()
        host.copy_(disparity, non_blocking=True)
        results.append(host)
        return k + 1
