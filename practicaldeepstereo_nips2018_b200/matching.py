"""Matching + MatchingOperation (reference matching.py:16-112) on sm_100a kernels.

Two kernel paths sit behind the reference API:

* ``Matching`` with an arbitrary ``operation`` callable: one kernel builds the
  disparity-stacked operation input ``(B*D, 2C, H, W)`` (csrc/matching_volume.cu),
  ``operation`` runs ONCE on it, one kernel re-stacks the result to
  ``(B, F, D, H, W)``.  Exact for any per-sample operation (the reference's own
  MatchingOperation uses InstanceNorm, whose statistics are per sample).
* ``Matching`` with a ``MatchingOperation``: the fused pipeline of
  csrc/matching_op.cu -- the concatenated volume is never materialised.

Gradient-enabled calls (training) are outside the inference hot path: the
convolution stacks run the plain ATen composition of the same modules under
autograd; ``Matching`` itself keeps its volume / stack kernels, with adjoint
kernels in the backward (``_ConcatVolume``, ``_StackDisparities``).
"""
import copy
import ctypes
import threading

import torch
from torch import nn

from . import _capi, network_blocks


def _needs_autograd(*tensors_and_modules):
    if not torch.is_grad_enabled():
        return False
    for obj in tensors_and_modules:
        if isinstance(obj, torch.Tensor):
            if obj.requires_grad:
                return True
        elif isinstance(obj, nn.Module):
            if any(p.requires_grad for p in obj.parameters()):
                return True
    return False


_WORKSPACE_SCOPE = threading.local()


class workspace_scope(object):
    """``with workspace_scope(tag):`` -- kernel scratch buffers requested inside are keyed by `tag`
    (not by the current stream); ``release_workspaces(module, tag)`` drops them."""

    def __init__(self, tag):
        self._tag, self._previous = tag, None

    def __enter__(self):
        self._previous = getattr(_WORKSPACE_SCOPE, 'tag', None)
        _WORKSPACE_SCOPE.tag = self._tag

    def __exit__(self, *exc):
        _WORKSPACE_SCOPE.tag = self._previous


def release_workspaces(module, tag=None):
    """Detaches the scratch buffers keyed by `tag` (all of them when `tag` is None) from the kernel
    handles of `module` and returns them: the caller decides how long they live (a captured CUDA
    graph keeps its own; dropping the list frees the memory)."""
    detached = []
    for m in module.modules():
        handle = m.__dict__.get('_kernel')
        if handle is not None and handle._workspace:
            for key in [k for k in handle._workspace if tag is None or k[1] == ('scope', tag)]:
                detached.append(handle._workspace.pop(key))
    return detached


class _KernelHandle(object):
    """Owns the C-ABI handles built from a module's parameters, one per device (DataParallel
    replicas share the module's ``__dict__`` and therefore this object).  A handle is rebuilt
    whenever a parameter is replaced or modified in place through autograd-visible operations
    (optimizer step, ``load_state_dict``, ``copy_`` under ``no_grad``: they bump ``_version``).

    Writes through ``param.data`` (``p.data.copy_()``, ``p.data.mul_()``) do NOT bump the version
    counter: call ``invalidate()`` (or ``PdsNetwork.invalidate_kernels()``) after such an update,
    as ``parallel.broadcast_parameters`` does."""

    def __init__(self, create, destroy):
        self._create, self._destroy = create, destroy
        self._handles = {}       # device string -> (handle, key)
        self._workspace = None
        self._lock = threading.Lock()

    def get(self, params, precision, device):
        dev = str(device)
        key = (precision,) + tuple((p.data_ptr(), p._version) for p in params)
        with self._lock:
            entry = self._handles.get(dev)
            if entry is not None and entry[1] == key:
                return entry[0]
            if entry is not None:
                self._destroy(entry[0])
                del self._handles[dev]
            handle = ctypes.c_void_p()
            with torch.cuda.device(device):
                self._create(handle, [p.detach().contiguous().float() for p in params],
                             precision, device)
            self._handles[dev] = (handle, key)
            return handle

    def workspace(self, nbytes, device):
        """Scratch buffer for the calling stream (one per CUDA stream, so that pipelines which
        keep several pairs in flight on different streams do not share scratch memory).  Under
        ``workspace_scope(tag)`` the buffer is keyed by the tag instead (a CUDA-graph capture owns
        its scratch memory: it lives in the graph's pool and its address is baked into the graph)."""
        tag = getattr(_WORKSPACE_SCOPE, 'tag', None)
        key = (str(device), ('scope', tag) if tag is not None else torch.cuda.current_stream(device).cuda_stream)
        if self._workspace is None:
            self._workspace = {}
        ws = self._workspace.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._workspace[key] = ws
        return ws

    def invalidate(self):
        """Forget every packed copy of the parameters: the next forward re-packs them."""
        self.release()

    def release(self):
        with self._lock:
            for handle, _ in self._handles.values():
                self._destroy(handle)
            self._handles = {}

    # copy.deepcopy(module) / torch.save(module): a fresh, empty handle bound to the copy
    # (ctypes pointers cannot be pickled and must not be shared between modules)
    def __deepcopy__(self, memo):
        return _KernelHandle(copy.deepcopy(self._create, memo), self._destroy)

    def __getstate__(self):
        return {'_create': self._create, '_destroy': self._destroy}

    def __setstate__(self, state):
        self.__init__(state['_create'], state['_destroy'])

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def check_fp16_weight_range(params, precision):
    """fp16-term precisions scale convolution weights by 2^8 before splitting them into half
    terms (include/pds_b200.h): |w| * 256 must stay below the largest finite half."""
    if precision not in ('fp16x2', 'fp16'):
        return
    for p in params:
        if p.dim() >= 4 and p.numel() and float(p.detach().abs().max()) * 256.0 >= 65504.0:
            raise ValueError(
                f'a convolution weight of magnitude {float(p.detach().abs().max()):.3g} does not fit the '
                f"'{precision}' operand format (|w| < 255); use precision='bf16x3'")


class MatchingOperation(nn.Module):
    """Per-disparity 2-D network applied to cat[left, shifted right]:
    conv3x3 2C->F, `number_of_residual_blocks` residual blocks, conv3x3 F->S."""

    def __init__(self, number_of_concatenated_descriptor_features=128, number_of_features=64,
                 number_of_compact_matching_signature_features=8,
                 number_of_residual_blocks=2, precision=None):
        super().__init__()
        precision = precision or _capi.DEFAULT_PRECISION
        if precision not in _capi.PRECISIONS:
            raise ValueError(f'precision should be one of {sorted(_capi.PRECISIONS)}')
        self._shape = (number_of_concatenated_descriptor_features, number_of_features,
                       number_of_compact_matching_signature_features, number_of_residual_blocks)
        self.precision = precision
        layers = [network_blocks.convolution_3x3(number_of_concatenated_descriptor_features,
                                                 number_of_features)]
        layers += [network_blocks.ResidualBlock(number_of_features)
                   for _ in range(number_of_residual_blocks)]
        layers += [network_blocks.convolution_3x3(
            number_of_features, number_of_compact_matching_signature_features)]
        self._matching_operation_modules = nn.ModuleList(layers)
        self.__dict__['_kernel'] = _KernelHandle(self._create_handle, self._destroy_handle)

    # -- C-ABI plumbing -------------------------------------------------------
    def _create_handle(self, handle, params, precision, device):
        cin, f, s, n_res = self._shape
        check_fp16_weight_range(params, precision)
        arr = _capi.pointer_array(params)
        _capi.check(_capi.lib().pds_matching_op_create(
            ctypes.byref(handle), arr, len(params), cin // 2, f, s, n_res,
            _capi.PRECISIONS[precision], _capi.stream_ptr(device)))
        torch.cuda.current_stream(device).synchronize()   # params may be temporaries

    @staticmethod
    def _destroy_handle(handle):
        _capi.lib().pds_matching_op_destroy(handle)

    def match_all_disparities(self, left, right, number_of_disparities):
        """left, right [B, C, H, W] -> signatures [B, S, D, H, W] for d = 0..D-1,
        i.e. th.stack([op(cat[left, shift_d(right)]) for d], dim=2)."""
        _capi.require_cuda(left, right)
        cin, _, s, _ = self._shape
        if left.shape != right.shape or left.dim() != 4 or 2 * left.size(1) != cin:
            raise ValueError(f'descriptors should be two [B, {cin // 2}, H, W] tensors')
        left = left.detach().contiguous().float()
        right = right.detach().contiguous().float()
        B, _, H, W = left.shape
        D = int(number_of_disparities)
        params = list(self.parameters())
        lib = _capi.lib()
        handle = self._kernel.get(params, self.precision, left.device)
        out = torch.empty((B, s, D, H, W), dtype=torch.float32, device=left.device)
        with torch.cuda.device(left.device):
            nbytes = lib.pds_matching_op_workspace_bytes(handle, B, H, W, D)
            ws = self._kernel.workspace(nbytes, left.device)
            _capi.check(lib.pds_matching_op_forward(
                handle, _capi.ptr(left), _capi.ptr(right), _capi.ptr(out), B, H, W, D,
                _capi.ptr(ws), ws.numel(), _capi.stream_ptr(left.device)))
        return out

    def forward(self, concatenated_descriptors):
        """[N, 2C, H, W] -> compact matching signature [N, S, H, W]."""
        if _needs_autograd(concatenated_descriptors, self):
            out = concatenated_descriptors            # training: ATen composition
            for module in self._matching_operation_modules:
                out = module(out)
            return out
        half = self._shape[0] // 2
        left, right = concatenated_descriptors[:, :half], concatenated_descriptors[:, half:]
        return self.match_all_disparities(left, right, 1)[:, :, 0]


class _ConcatVolume(torch.autograd.Function):
    """f4: the reference's per-disparity pad / slice / cat (matching.py:53-60) as ONE differentiable
    node.  forward: pds_matching_concat -> [B, D, 2C, H, W]; backward: pds_matching_concat_backward
    (sum over disparities of the left half, shifted sum of the right half)."""

    @staticmethod
    def forward(ctx, left, right, number_of_disparities):
        left, right = left.contiguous(), right.contiguous()
        B, C, H, W = left.shape
        D = int(number_of_disparities)
        ctx.dims = (B, C, H, W, D)
        volume = torch.empty((B, D, 2 * C, H, W), dtype=left.dtype, device=left.device)
        with torch.cuda.device(left.device):
            _capi.check(_capi.lib().pds_matching_concat(
                _capi.ptr(left), _capi.ptr(right), _capi.ptr(volume), B, C, H, W, D,
                _capi.dtype_code(left), _capi.stream_ptr(left.device)))
        return volume

    @staticmethod
    def backward(ctx, grad_volume):
        B, C, H, W, D = ctx.dims
        grad_volume = grad_volume.contiguous()
        grad_left = torch.empty((B, C, H, W), dtype=grad_volume.dtype, device=grad_volume.device)
        grad_right = torch.empty_like(grad_left)
        with torch.cuda.device(grad_volume.device):
            _capi.check(_capi.lib().pds_matching_concat_backward(
                _capi.ptr(grad_volume), _capi.ptr(grad_left), _capi.ptr(grad_right), B, C, H, W, D,
                _capi.dtype_code(grad_volume), _capi.stream_ptr(grad_volume.device)))
        return grad_left, grad_right, None


class _StackDisparities(torch.autograd.Function):
    """th.stack(signatures, dim=2) (matching.py:63) of a batched operation output:
    [B * D, F, H, W] -> [B, F, D, H, W]; the backward is the inverse permutation."""

    @staticmethod
    def forward(ctx, signatures, batch):
        signatures = signatures.contiguous()
        N, F, H, W = signatures.shape
        D = N // batch
        ctx.dims = (batch, F, D, H, W)
        out = torch.empty((batch, F, D, H, W), dtype=signatures.dtype, device=signatures.device)
        with torch.cuda.device(signatures.device):
            _capi.check(_capi.lib().pds_matching_stack(
                _capi.ptr(signatures), _capi.ptr(out), batch, F, D, H, W, _capi.dtype_code(signatures),
                _capi.stream_ptr(signatures.device)))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, F, D, H, W = ctx.dims
        grad_out = grad_out.contiguous()
        grad_in = torch.empty((B * D, F, H, W), dtype=grad_out.dtype, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _capi.check(_capi.lib().pds_matching_unstack(
                _capi.ptr(grad_out), _capi.ptr(grad_in), B, F, D, H, W, _capi.dtype_code(grad_out),
                _capi.stream_ptr(grad_out.device)))
        return grad_in, None


class Matching(nn.Module):
    def __init__(self, maximum_disparity, operation, batched_operation=True):
        """maximum_disparity: disparity range is [0, maximum_disparity];
        operation: module or function applied to the concatenated left / shifted
        right descriptors of every disparity.  ``batched_operation=False`` calls a
        generic operation once per disparity (as the reference does) instead of
        once on the disparity-stacked batch."""
        super().__init__()
        self._maximum_disparity = maximum_disparity
        self._operation = operation
        self._batched_operation = batched_operation

    def set_maximum_disparity(self, maximum_disparity):
        self._maximum_disparity = maximum_disparity

    def _autograd_forward(self, left, right):
        """Training (gradients required).  CUDA float32 / bfloat16 descriptors with a batched
        operation: volume kernel (+ its adjoint in the backward) -> ONE call of the operation on the
        disparity-stacked batch [B * D, 2C, H, W] (InstanceNorm statistics are per sample, so the
        stacked call equals the reference's D calls, matching.py:53-62) -> stack kernel.  The
        reference's loop costs D x ~20 operator launches and autograd nodes forward and again
        backward; this is ~20.  Otherwise (CPU, other dtypes, ``batched_operation=False``) the
        reference's per-disparity composition."""
        md = self._maximum_disparity
        if (self._batched_operation and network_blocks.USE_TRAINING_KERNELS and left.is_cuda and right.is_cuda
                and left.dtype == right.dtype and left.dtype in (torch.float32, torch.bfloat16)
                and left.shape == right.shape and left.dim() == 4):
            volume = _ConcatVolume.apply(left, right, md + 1)
            signatures = self._operation(volume.view(-1, *volume.shape[2:]))
            return _StackDisparities.apply(signatures, left.size(0))
        padded = nn.functional.pad(right, (md, 0, 0, 0))
        width = right.size(-1)
        out = [self._operation(torch.cat(
            [left, padded[..., md - d:md - d + width]], dim=1)) for d in range(md + 1)]
        return torch.stack(out, dim=2)

    def forward(self, left_embedding, right_embedding):
        """[B, C, H, W] x 2 -> matching signatures [B, F, maximum_disparity + 1, H, W]."""
        op = self._operation
        if _needs_autograd(left_embedding, right_embedding, op):
            return self._autograd_forward(left_embedding, right_embedding)
        D = self._maximum_disparity + 1
        if isinstance(op, MatchingOperation):
            return op.match_all_disparities(left_embedding, right_embedding, D)
        # generic operation: volume kernel -> operation -> stack kernel
        _capi.require_cuda(left_embedding, right_embedding)
        left = left_embedding.detach().contiguous()
        right = right_embedding.detach().contiguous()
        if left.shape != right.shape or left.dim() != 4:
            raise ValueError('descriptors should be two [B, C, H, W] tensors')
        B, C, H, W = left.shape
        lib, dt = _capi.lib(), _capi.dtype_code(left)
        volume = torch.empty((B, D, 2 * C, H, W), dtype=left.dtype, device=left.device)
        with torch.cuda.device(left.device):
            st = _capi.stream_ptr(left.device)
            _capi.check(lib.pds_matching_concat(_capi.ptr(left), _capi.ptr(right),
                                                _capi.ptr(volume), B, C, H, W, D, dt, st))
            if self._batched_operation:
                sig = op(volume.view(B * D, 2 * C, H, W))
            else:
                sig = torch.stack([op(volume[:, d]) for d in range(D)], dim=1)
                sig = sig.reshape(B * D, *sig.shape[2:])
            sig = sig.contiguous()
            F = sig.size(1)
            out = torch.empty((B, F, D, H, W), dtype=sig.dtype, device=sig.device)
            _capi.check(lib.pds_matching_stack(_capi.ptr(sig), _capi.ptr(out), B, F, D, H, W,
                                               _capi.dtype_code(sig), st))
        return out
