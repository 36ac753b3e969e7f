"""Embedding tower (reference embedding.py:11-65) on the tcgen05 convolution
engine (csrc/embedding.cu).  Structure and state_dict keys are the reference's.

Eval / no-grad CUDA calls in a tensor-core precision run the kernels; the fp32
precision and gradient-enabled calls (training, outside the inference hot path)
run the plain ATen composition of the same modules."""
import ctypes
import warnings

import torch
from torch import nn

from . import _capi, network_blocks
from .matching import _KernelHandle, _needs_autograd, check_fp16_weight_range


class Embedding(nn.Module):
    _warned_aten = False

    def __init__(self, number_of_input_features=3, number_of_embedding_features=64,
                 number_of_shortcut_features=8, number_of_residual_blocks=2, precision=None):
        super().__init__()
        precision = precision or _capi.DEFAULT_PRECISION
        if precision not in _capi.PRECISIONS:
            raise ValueError(f'precision should be one of {sorted(_capi.PRECISIONS)}')
        f = number_of_embedding_features
        self._shape = (number_of_input_features, f, number_of_shortcut_features,
                       number_of_residual_blocks)
        self.precision = precision
        tower = [nn.InstanceNorm2d(number_of_input_features),
                 network_blocks.convolutional_block_5x5_stride_2(number_of_input_features, f),
                 network_blocks.convolutional_block_5x5_stride_2(f, f)]
        tower += [network_blocks.ResidualBlock(f) for _ in range(number_of_residual_blocks)]
        self._embedding_modules = nn.ModuleList(tower)
        self._shortcut = network_blocks.convolutional_block_3x3(f, number_of_shortcut_features)
        self.__dict__['_kernel'] = _KernelHandle(self._create_handle, self._destroy_handle)

    # -- C-ABI plumbing -------------------------------------------------------
    def _create_handle(self, handle, params, precision, device):
        cin, f, fs, n_res = self._shape
        check_fp16_weight_range(params, precision)
        _capi.check(_capi.lib().pds_embedding_create(
            ctypes.byref(handle), _capi.pointer_array(params), len(params), cin, f, fs, n_res,
            _capi.PRECISIONS[precision], _capi.stream_ptr(device)))
        torch.cuda.current_stream(device).synchronize()   # params may be temporaries

    @staticmethod
    def _destroy_handle(handle):
        _capi.lib().pds_embedding_destroy(handle)

    def uses_kernels(self, image):
        cin, f, fs, _ = self._shape
        return (self.precision != 'fp32' and image.is_cuda and not _needs_autograd(image, self)
                and cin <= 8 and f == 64 and fs in (4, 8, 12, 16)
                and image.dim() == 4 and image.size(2) % 4 == 0 and image.size(3) % 4 == 0)

    def embed(self, images, number_of_shortcuts):
        """images (N,3,H,W) -> descriptors (N,64,H/4,W/4) and the shortcut
        (number_of_shortcuts,8,H/4,W/4) of the first `number_of_shortcuts` images."""
        cin, f, fs, _ = self._shape
        images = images.detach().contiguous().float()
        n, c, H, W = images.shape
        if c != cin:
            raise ValueError(f'images should have {cin} channels')
        lib = _capi.lib()
        handle = self._kernel.get(list(self.parameters()), self.precision, images.device)
        descriptor = images.new_empty((n, f, H // 4, W // 4))
        shortcut = images.new_empty((number_of_shortcuts, fs, H // 4, W // 4))
        with torch.cuda.device(images.device):
            nbytes = lib.pds_embedding_workspace_bytes(handle, n, H, W)
            ws = self._kernel.workspace(nbytes, images.device)
            _capi.check(lib.pds_embedding_forward(
                handle, _capi.ptr(images), _capi.ptr(descriptor),
                _capi.ptr(shortcut) if number_of_shortcuts else None, n, number_of_shortcuts,
                H, W, _capi.ptr(ws), ws.numel(), _capi.stream_ptr(images.device)))
        return descriptor, shortcut

    # -- f3: un-padded float / uint8 images straight into the first operand planes ----------
    def image_geometry(self, image):
        """(layout, h, w) of a batch of images the fused input path accepts: float32
        (B,C,h,w), uint8 (B,C,h,w) or uint8 (B,h,w,C) (decoder order, dataset.py:67-72);
        None for anything else."""
        cin = self._shape[0]
        if not torch.is_tensor(image) or image.dim() != 4:
            return None
        if image.dtype == torch.float32 and image.size(1) == cin:
            return _capi.IMAGE_LAYOUTS['f32_nchw'], image.size(2), image.size(3)
        if image.dtype == torch.uint8 and image.size(1) == cin:
            return _capi.IMAGE_LAYOUTS['u8_nchw'], image.size(2), image.size(3)
        if image.dtype == torch.uint8 and image.size(3) == cin:
            return _capi.IMAGE_LAYOUTS['u8_nhwc'], image.size(1), image.size(2)
        return None

    def can_embed_images(self, left_images, right_images):
        cin, f, fs, _ = self._shape
        return (self.precision != 'fp32' and self.image_geometry(left_images) is not None
                and left_images.is_cuda and left_images.shape == right_images.shape
                and left_images.dtype == right_images.dtype
                and left_images.device == right_images.device
                and not _needs_autograd(left_images, right_images, self)
                and cin <= 8 and f == 64 and fs in (4, 8, 12, 16))

    def embed_images(self, left_images, right_images, pad_top, pad_left):
        """SizeAdapter.pad + Embedding of both batches in one pipeline
        (pds_embedding_forward_images): returns the left descriptors, the right
        descriptors and the shortcut of the left images, all at the PADDED extent / 4."""
        cin, f, fs, _ = self._shape
        layout, h, w = self.image_geometry(left_images)
        H, W = h + pad_top, w + pad_left
        if H % 4 or W % 4:
            raise ValueError('padded height and width should be multiples of 4')
        left_images, right_images = left_images.detach().contiguous(), right_images.detach().contiguous()
        batch = left_images.size(0)
        device = left_images.device
        lib = _capi.lib()
        handle = self._kernel.get(list(self.parameters()), self.precision, device)
        descriptor = torch.empty((2 * batch, f, H // 4, W // 4), dtype=torch.float32, device=device)
        shortcut = torch.empty((batch, fs, H // 4, W // 4), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            nbytes = lib.pds_embedding_workspace_bytes(handle, 2 * batch, H, W)
            ws = self._kernel.workspace(nbytes, device)
            _capi.check(lib.pds_embedding_forward_images(
                handle, _capi.ptr(left_images), batch, _capi.ptr(right_images), batch, layout, h, w,
                pad_top, pad_left, _capi.ptr(descriptor), _capi.ptr(shortcut), batch,
                _capi.ptr(ws), ws.numel(), _capi.stream_ptr(device)))
        return descriptor[:batch], descriptor[batch:], shortcut

    def forward(self, image, with_shortcut=True):
        """image (B,3,H,W) -> descriptor (B,64,H/4,W/4), shortcut (B,8,H/4,W/4).

        ``with_shortcut=False`` skips the shortcut block, whose result the
        reference computes for the right image and throws away (network.py:40)."""
        if self.uses_kernels(image):
            descriptor, shortcut = self.embed(image, image.size(0) if with_shortcut else 0)
            return descriptor, (shortcut if with_shortcut else None)
        if (self.precision != 'fp32' and image.is_cuda and not _needs_autograd(image, self)
                and not Embedding._warned_aten):
            # never reached from PdsNetwork (SizeAdapter pads to multiples of 64); a standalone
            # call on other extents keeps the reference's semantics on ATen operators -- loudly
            Embedding._warned_aten = True
            warnings.warn(
                f'Embedding: image extent {tuple(image.shape[2:])} is not a multiple of 4 (or the '
                'module is not the 3->64->8 tower): this call runs the ATen composition, not the '
                'sm_100a kernels', RuntimeWarning, stacklevel=2)
        descriptor = image
        for module in self._embedding_modules:
            descriptor = module(descriptor)
        if not with_shortcut:
            return descriptor, None
        return descriptor, self._shortcut(descriptor)
