"""PdsNetwork: drop-in for practical_deep_stereo.network.PdsNetwork
(reference network.py:14-65) wired to the B200 kernel modules."""
import os
import warnings

import torch
from torch import nn

from . import _capi, embedding, estimator, matching, regularization, size_adapter


# Regularization.forward_disparity: hourglass tail + estimator + crop as one pipeline whose cost
# volume is never written (bit-identical to the separate calls: per-segment SubpixelMap states +
# a merge kernel).  Built and tested, but carrying the estimator state through the transposed
# convolution costs more than the 424 MB round trip it saves at C2 (0.52 + 0.01 ms vs 0.28 +
# 0.11 ms): off unless PDS_B200_FUSE_TAIL=1.
FUSE_TAIL_AND_ESTIMATOR = os.environ.get('PDS_B200_FUSE_TAIL', '0') == '1'


class PdsNetwork(nn.Module):
    """Practical Deep Stereo network; same constructor (dependency injection of
    the five stages), methods and tensor shapes as the reference."""
    _warned_eval_grad = False

    def __init__(self, size_adapter_module, embedding_module, matching_module,
                 regularization_module, estimator_module):
        super().__init__()
        self._size_adapter = size_adapter_module
        self._embedding = embedding_module
        self._matching = matching_module
        self._regularization = regularization_module
        self._estimator = estimator_module

    def set_maximum_disparity(self, maximum_disparity):
        """Reconfigure the disparity range; (maximum_disparity + 1) % 64 == 0."""
        if (maximum_disparity + 1) % 64 != 0:
            raise ValueError(
                '"maximum_disparity" + 1 should be multiple of 64, e.g.,'
                '"maximum disparity" can be equal to 63, 191, 255, 319...')
        self._maximum_disparity = maximum_disparity
        # descriptors are 4x down-sampled -> matching range (md + 1) / 4 - 1
        self._matching.set_maximum_disparity((maximum_disparity + 1) // 4 - 1)

    def invalidate_kernels(self):
        """Drops every packed copy of the parameters held by the kernel handles.  Needed only
        after writes that bypass autograd's version counters (``param.data.copy_()`` and
        friends); optimizer steps, ``load_state_dict`` and ``copy_`` under ``no_grad`` are
        detected automatically."""
        for module in self.modules():
            handle = module.__dict__.get('_kernel')
            if handle is not None:
                handle.invalidate()

    def release_workspaces(self):
        """Frees the scratch buffers the kernel handles cache per CUDA stream (they are re-allocated
        on demand): useful between workloads of very different size."""
        del matching.release_workspaces(self)[:]

    def _check_inputs(self, left_image, right_image):
        """One clear error up front instead of a failure deep inside a stage: the inference
        kernels need CUDA tensors on one device (there is no CPU path); gradient-enabled calls
        run the ATen composition on any device."""
        for name, image in (('left_image', left_image), ('right_image', right_image)):
            if not torch.is_tensor(image) or image.dim() != 4:
                raise ValueError(f'"{name}" should be a [B, 3, H, W] tensor (or uint8 [B, H, W, 3])')
        if left_image.device != right_image.device:
            raise ValueError('"left_image" and "right_image" should be on the same device')
        needs_grad = matching._needs_autograd(left_image, right_image, self)
        if not needs_grad and not left_image.is_cuda:
            raise RuntimeError(
                'PdsNetwork.forward without gradients runs on the sm_100a kernels and needs CUDA '
                'tensors; there is no CPU path (enable gradients for the ATen composition)')
        if needs_grad and not self.training and not PdsNetwork._warned_eval_grad:
            PdsNetwork._warned_eval_grad = True
            warnings.warn(
                'PdsNetwork is in eval mode but gradients are enabled and parameters require grad: '
                'this forward runs the ATen composition (a Python loop over every disparity), not '
                'the sm_100a kernels; wrap inference in torch.no_grad()', RuntimeWarning, stacklevel=3)

    def _embed(self, left_image, right_image):
        """Embedding of both images.  With the kernel embedding the two images are
        samples of ONE batch and the shortcut block runs on the left image only (the
        reference computes the right image's and discards it, network.py:40).  A
        foreign / fp32 embedding module is called as the reference does, on ATen
        operators, with TF32 switched off while the convolution stacks are fp32-grade
        so that the pipeline is fp32 end to end."""
        emb = self._embedding
        if isinstance(emb, embedding.Embedding) and emb.uses_kernels(left_image) \
                and left_image.shape == right_image.shape:
            batch = left_image.size(0)
            descriptors, shortcut_from_left = emb.embed(torch.cat([left_image, right_image]), batch)
            return descriptors[:batch], descriptors[batch:], shortcut_from_left
        precision = getattr(getattr(self._matching, '_operation', None), 'precision', 'fp32')
        exact = left_image.is_cuda and precision in ('fp32', 'bf16x3', 'bf16x2', 'fp16x2')
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=not exact):
            left_descriptor, shortcut_from_left = emb(left_image)
            right_descriptor = emb(right_image)[0]
        return left_descriptor, right_descriptor, shortcut_from_left

    def pass_through_network(self, left_image, right_image):
        left_descriptor, right_descriptor, shortcut_from_left = self._embed(left_image,
                                                                            right_image)
        signatures = self._matching(left_descriptor, right_descriptor)
        return self._regularization(signatures, shortcut_from_left), shortcut_from_left

    def _embed_unpadded(self, left_image, right_image):
        """f3 -- the input path: un-padded float32 or uint8 images (planar, or interleaved
        as the decoder leaves them, dataset.py:67-72) go straight into the kernel that
        pads (size_adapter.py:29-43), normalises (embedding.py:32) and writes the first
        convolution's operands.  Returns None when that path does not apply."""
        emb, adapter = self._embedding, self._size_adapter
        if not (isinstance(emb, embedding.Embedding) and isinstance(adapter, size_adapter.SizeAdapter)
                and emb.can_embed_images(left_image, right_image)):
            return None
        _, height, width = emb.image_geometry(left_image)
        pad_top, pad_left = adapter.padding_for(height, width)
        adapter._pixels_pad_to_height, adapter._pixels_pad_to_width = pad_top, pad_left
        return emb.embed_images(left_image, right_image, pad_top, pad_left)

    @staticmethod
    def _as_float_planes(image):
        """uint8 images (planar or interleaved) as the float (B,3,H,W) tensor the reference's
        data loader produces (dataset.py:67-72)."""
        if torch.is_tensor(image) and image.dtype == torch.uint8:
            if image.dim() == 4 and image.size(1) != 3 and image.size(3) == 3:
                image = image.permute(0, 3, 1, 2)
            return image.float()
        return image

    def forward(self, left_image, right_image):
        """Sub-pixel disparity [B, H, W] in eval mode, matching cost
        [B, (md + 1) / 2, H, W] in training mode.  Besides the reference's float
        (B,3,H,W) images, uint8 images (B,3,H,W) or (B,H,W,3) are accepted."""
        reg, est = self._regularization, self._estimator
        self._check_inputs(left_image, right_image)
        embedded = self._embed_unpadded(left_image, right_image)
        if embedded is None:
            left_image, right_image = self._as_float_planes(left_image), self._as_float_planes(right_image)
            left, right = self._size_adapter.pad(left_image), self._size_adapter.pad(right_image)
            embedded = self._embed(left, right)
        left_descriptor, right_descriptor, shortcut = embedded
        signatures = self._matching(left_descriptor, right_descriptor)
        on_kernels = signatures.is_cuda and isinstance(est, estimator.SubpixelMap)
        if (FUSE_TAIL_AND_ESTIMATOR and not self.training and on_kernels
                and isinstance(reg, regularization.Regularization) and reg.can_fuse_estimator(est)
                and not matching._needs_autograd(signatures, shortcut, self)):
            # hourglass tail, estimator and SizeAdapter.unpad as one pipeline: the cost volume
            # (212 MB at 960x540) is never written (bit-identical to the separate calls below)
            return reg.forward_disparity(signatures, shortcut, est._half_support_window,
                                         est._disparity_step,
                                         crop_top=self._size_adapter._pixels_pad_to_height,
                                         crop_left=self._size_adapter._pixels_pad_to_width)
        cost = reg(signatures, shortcut)
        if self.training:
            return self._size_adapter.unpad(cost)
        if on_kernels:
            # SizeAdapter.unpad fused into the estimator's store
            return est(cost, crop_top=self._size_adapter._pixels_pad_to_height,
                       crop_left=self._size_adapter._pixels_pad_to_width)
        return self._size_adapter.unpad(est(cost))

    @staticmethod
    def default(maximum_disparity=255, precision=None):
        """Network with the default modules; `precision` selects the arithmetic of
        the convolution stacks (see include/pds_b200.h, enum pds_precision).  The default
        (``_capi.DEFAULT_PRECISION`` = 'fp16x2') is the fp32-grade split-operand tensor-core
        mode, so the reference's unchanged ``PdsNetwork.default().cuda()`` call
        (benchmark_on_flyingthings3d.py:56-59) gets the tcgen05 path; 'fp32' selects the
        CUDA-core FFMA kernels with the embedding on ATen."""
        precision = precision or _capi.DEFAULT_PRECISION
        network = PdsNetwork(
            size_adapter_module=size_adapter.SizeAdapter(),
            embedding_module=embedding.Embedding(precision=precision),
            matching_module=matching.Matching(
                operation=matching.MatchingOperation(precision=precision),
                maximum_disparity=0),
            regularization_module=regularization.Regularization(precision=precision),
            estimator_module=estimator.SubpixelMap())
        network.set_maximum_disparity(maximum_disparity)
        return network
