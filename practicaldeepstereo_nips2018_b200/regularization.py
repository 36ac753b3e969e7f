"""3-D hourglass regularizer (reference regularization.py:11-126) on sm_100a
kernels (csrc/regularization.cu).  Module structure and state_dict keys are the
reference's; gradient-enabled calls run the ATen composition (training is
outside the inference hot path)."""
import ctypes

import torch
from torch import nn

from . import _capi, network_blocks
from .matching import _KernelHandle, _needs_autograd, check_fp16_weight_range


def _f32(t):
    return t.detach().contiguous().float()


class _BlockWorkspace(object):
    """Scratch memory of a standalone ContractionBlock3d / ExpansionBlock3d call, kept per
    (device, stream) and grown on demand instead of being allocated on every call."""

    def __init__(self):
        self._buffers = {}

    def get(self, nbytes, device):
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        ws = self._buffers.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buffers[key] = ws
        return ws

    def __deepcopy__(self, memo):
        return _BlockWorkspace()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self._buffers = {}


class ContractionBlock3d(nn.Module):
    """stride-2 "downsampling" block followed by a "smoothing" block; returns both."""

    def __init__(self, number_of_features):
        super().__init__()
        n = number_of_features
        self._downsampling_2x = network_blocks.convolutional_block_3x3x3_stride_2(n, 2 * n)
        self._smoothing = network_blocks.convolutional_block_3x3x3(2 * n, 2 * n)
        self.__dict__['_scratch'] = _BlockWorkspace()

    def forward(self, block_input):
        if _needs_autograd(block_input, self):
            down = self._downsampling_2x(block_input)
            return down, self._smoothing(down)
        _capi.require_cuda(block_input)
        x = _f32(block_input)
        B, C, D, H, W = x.shape
        o = (B, 2 * C, (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1)
        down, smooth = x.new_empty(o), x.new_empty(o)
        params = [_f32(p) for p in self.parameters()]
        lib = _capi.lib()
        with torch.cuda.device(x.device):
            nbytes = lib.pds_contraction_block_workspace_bytes(B, C, D, H, W)
            ws = self._scratch.get(nbytes, x.device)
            _capi.check(lib.pds_contraction_block_forward(
                _capi.pointer_array(params), _capi.ptr(x), _capi.ptr(down), _capi.ptr(smooth),
                B, C, D, H, W, _capi.ptr(ws), ws.numel(), _capi.stream_ptr(x.device)))
        return down, smooth


class ExpansionBlock3d(nn.Module):
    """transposed-conv 2x "upsampling" block, + shortcut, "smoothing" block."""

    def __init__(self, number_of_features):
        super().__init__()
        n = number_of_features
        self._upsampling_2x = network_blocks.transposed_convolutional_block_4x4x4_stride_2(
            n, n // 2)
        self._smoothing = network_blocks.convolutional_block_3x3x3(n // 2, n // 2)
        self.__dict__['_scratch'] = _BlockWorkspace()

    def forward(self, block_input, shortcut_from_contraction):
        if _needs_autograd(block_input, shortcut_from_contraction, self):
            return self._smoothing(self._upsampling_2x(block_input) + shortcut_from_contraction)
        _capi.require_cuda(block_input, shortcut_from_contraction)
        x, skip = _f32(block_input), _f32(shortcut_from_contraction)
        B, C, D, H, W = x.shape
        out = x.new_empty((B, C // 2, 2 * D, 2 * H, 2 * W))
        if tuple(skip.shape) != tuple(out.shape):
            raise ValueError('shortcut_from_contraction should have the up-sampled shape')
        params = [_f32(p) for p in self.parameters()]
        lib = _capi.lib()
        with torch.cuda.device(x.device):
            nbytes = lib.pds_expansion_block_workspace_bytes(B, C, D, H, W)
            ws = self._scratch.get(nbytes, x.device)
            _capi.check(lib.pds_expansion_block_forward(
                _capi.pointer_array(params), _capi.ptr(x), _capi.ptr(skip), _capi.ptr(out),
                B, C, D, H, W, _capi.ptr(ws), ws.numel(), _capi.stream_ptr(x.device)))
        return out


class Regularization(nn.Module):
    """Hourglass over the (B, 8, D, H, W) signature volume: 16x contraction,
    then expansion to (B, 2D, 4H, 4W) -- matching cost for even disparities."""

    def __init__(self, number_of_features=8, precision=None):
        super().__init__()
        precision = precision or _capi.DEFAULT_PRECISION
        if precision not in _capi.PRECISIONS:
            raise ValueError(f'precision should be one of {sorted(_capi.PRECISIONS)}')
        n = number_of_features
        self._number_of_features = n
        self.precision = precision
        self._smoothing = network_blocks.convolutional_block_3x3x3(n, n)
        self._contraction_blocks = nn.ModuleList(
            [ContractionBlock3d(n * scale) for scale in (1, 2, 4, 8)])
        self._expansion_blocks = nn.ModuleList(
            [ExpansionBlock3d(n * scale) for scale in (16, 8, 4, 2)])
        self._upsample_to_halfsize = \
            network_blocks.transposed_convolutional_block_4x4x4_stride_2(n, n // 2)
        self._upsample_to_fullsize = \
            network_blocks.transposed_convolution_3x4x4_stride_122(n // 2, 1)
        self.__dict__['_kernel'] = _KernelHandle(self._create_handle, self._destroy_handle)

    def _create_handle(self, handle, params, precision, device):
        check_fp16_weight_range(params, precision)
        arr = _capi.pointer_array(params)
        _capi.check(_capi.lib().pds_regularization_create(
            ctypes.byref(handle), arr, len(params), self._number_of_features,
            _capi.PRECISIONS[precision], _capi.stream_ptr(device)))
        torch.cuda.current_stream(device).synchronize()

    @staticmethod
    def _destroy_handle(handle):
        _capi.lib().pds_regularization_destroy(handle)

    def _autograd_forward(self, matching_signatures, shortcut_from_left_image):
        skips = []
        shortcut = shortcut_from_left_image.unsqueeze(2)
        output = self._smoothing(matching_signatures)
        for block in self._contraction_blocks:
            skips.append(output)
            shortcut, output = block(shortcut + output)
        for block in self._expansion_blocks:
            output = block(output, skips.pop())
        return self._upsample_to_fullsize(self._upsample_to_halfsize(output)).squeeze(1)

    def forward_disparity(self, matching_signatures, shortcut_from_left_image,
                          half_support_window, disparity_step, crop_top=0, crop_left=0,
                          return_argmax=False):
        """Regularization.forward followed by SubpixelMap and SizeAdapter.unpad as ONE
        pipeline whose cost volume is never written (bit-identical to the separate calls):
        signatures [B, F, D, H, W], shortcut [B, F, H, W] -> disparity
        [B, 4H - crop_top, 4W - crop_left]."""
        _capi.require_cuda(matching_signatures, shortcut_from_left_image)
        sig, sc = _f32(matching_signatures), _f32(shortcut_from_left_image)
        B, F, D, H, W = sig.shape
        if F != self._number_of_features or tuple(sc.shape) != (B, F, H, W):
            raise ValueError('expected signatures [B, F, D, H, W] and shortcut [B, F, H, W]')
        if D % 16 or H % 16 or W % 16:
            raise ValueError('D, H and W of the signature volume should be multiples of 16')
        lib = _capi.lib()
        handle = self._kernel.get(list(self.parameters()), self.precision, sig.device)
        out = sig.new_empty((B, 4 * H - crop_top, 4 * W - crop_left))
        idx = torch.empty_like(out, dtype=torch.int64) if return_argmax else None
        with torch.cuda.device(sig.device):
            nbytes = lib.pds_regularization_workspace_bytes(handle, B, D, H, W)
            ws = self._kernel.workspace(nbytes, sig.device)
            _capi.check(lib.pds_regularization_forward_disparity(
                handle, _capi.ptr(sig), _capi.ptr(sc), _capi.ptr(out),
                _capi.ptr(idx) if idx is not None else None, B, D, H, W, half_support_window,
                disparity_step, crop_top, crop_left, _capi.ptr(ws), ws.numel(),
                _capi.stream_ptr(sig.device)))
        return (out, idx) if return_argmax else out

    def can_fuse_estimator(self, estimator_module):
        """True when forward_disparity serves this estimator (window radius <= 4)."""
        hsw = getattr(estimator_module, '_half_support_window', None)
        step = getattr(estimator_module, '_disparity_step', None)
        return (self._number_of_features == 8 and hsw is not None and step is not None
                and hsw // step <= 4)

    def forward(self, matching_signatures, shortcut_from_left_image):
        """signatures [B, F, D, H, W], shortcut [B, F, H, W] -> cost [B, 2D, 4H, 4W]."""
        if _needs_autograd(matching_signatures, shortcut_from_left_image, self):
            return self._autograd_forward(matching_signatures, shortcut_from_left_image)
        _capi.require_cuda(matching_signatures, shortcut_from_left_image)
        sig, sc = _f32(matching_signatures), _f32(shortcut_from_left_image)
        B, F, D, H, W = sig.shape
        if F != self._number_of_features or tuple(sc.shape) != (B, F, H, W):
            raise ValueError('expected signatures [B, F, D, H, W] and shortcut [B, F, H, W]')
        if D % 16 or H % 16 or W % 16:
            raise ValueError('D, H and W of the signature volume should be multiples of 16')
        lib = _capi.lib()
        handle = self._kernel.get(list(self.parameters()), self.precision, sig.device)
        cost = sig.new_empty((B, 2 * D, 4 * H, 4 * W))
        with torch.cuda.device(sig.device):
            nbytes = lib.pds_regularization_workspace_bytes(handle, B, D, H, W)
            ws = self._kernel.workspace(nbytes, sig.device)
            _capi.check(lib.pds_regularization_forward(
                handle, _capi.ptr(sig), _capi.ptr(sc), _capi.ptr(cost), B, D, H, W,
                _capi.ptr(ws), ws.numel(), _capi.stream_ptr(sig.device)))
        return cost
