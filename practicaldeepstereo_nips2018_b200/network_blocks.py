"""Parameter containers for the PDS layers.

The B200 kernels read the parameters straight from these modules, so the only
hard requirement is that ``state_dict()`` keys, shapes and semantics equal the
reference's (network_blocks.py:9-144 there; key families in SURVEY.md A.3):
every block is Conv -> LeakyReLU(0.1) -> InstanceNorm(affine, eps 1e-5), stored
as an ``nn.Sequential`` whose index 0 is the convolution and index 2 the norm.
The modules stay callable for autograd / training, which is outside the
inference hot path: convolutions on ATen, LeakyReLU + InstanceNorm on fused
forward / backward kernels (``ConvBlock``).
"""
import os

import torch
from torch import nn

from . import _capi

LEAKY_SLOPE = 0.1

# f4: training-mode kernels (fused LeakyReLU + InstanceNorm forward / backward here, the volume /
# stack kernels and their adjoints in matching.py); 0: the plain ATen composition
USE_TRAINING_KERNELS = os.environ.get('PDS_B200_TRAIN_KERNELS', '1') == '1'


class _LeakyInstanceNorm(torch.autograd.Function):
    """InstanceNorm(affine)(LeakyReLU(x)) as one differentiable node on the kernels of
    csrc/instance_norm_train.cu (two flat HBM passes forward, two backward).  Saves x (the
    convolution output) and the per-(sample, channel) mean / rstd; the activation is recomputed."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope):
        x = x.contiguous()
        N, C = x.shape[:2]
        L = x.numel() // max(1, N * C)
        y = torch.empty_like(x)
        mean_rstd = torch.empty((N * C, 2), dtype=torch.float32, device=x.device)
        sums = torch.empty((N * C, 2), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.lib().pds_instance_norm_forward(
                _capi.ptr(x), _capi.ptr(gamma), _capi.ptr(beta), _capi.ptr(y), _capi.ptr(mean_rstd),
                _capi.ptr(sums), N, C, L, float(eps), float(slope), _capi.stream_ptr(x.device)))
        ctx.save_for_backward(x, gamma, mean_rstd)
        ctx.slope = float(slope)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, gamma, mean_rstd = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        N, C = x.shape[:2]
        L = x.numel() // max(1, N * C)
        grad_x = torch.empty_like(x)
        sums = torch.empty((N * C, 2), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(_capi.lib().pds_instance_norm_backward(
                _capi.ptr(x), _capi.ptr(grad_y), _capi.ptr(gamma), _capi.ptr(mean_rstd), _capi.ptr(grad_x),
                _capi.ptr(sums), N, C, L, ctx.slope, _capi.stream_ptr(x.device)))
        per_channel = sums.view(N, C, 2).sum(0)          # [C, 2]: sum dy, sum dy * zhat over samples
        return grad_x, per_channel[:, 1].float(), per_channel[:, 0].float(), None, None


class ConvBlock(nn.Sequential):
    """Conv -> LeakyReLU -> InstanceNorm(affine) with the reference's module indices (0, 1, 2 are
    part of the state_dict keys).  Gradient-enabled CUDA float32 calls run the convolution on ATen
    and LeakyReLU + InstanceNorm (forward and backward) on the fused kernels; everything else is
    the plain ``nn.Sequential`` composition."""

    def forward(self, block_input):
        norm = self[2]
        if (USE_TRAINING_KERNELS and torch.is_grad_enabled() and block_input.is_cuda
                and block_input.dtype == torch.float32 and norm.affine and not norm.track_running_stats
                and (block_input.requires_grad or self[0].weight.requires_grad or norm.weight.requires_grad)):
            out = self[0](block_input)
            if out.size(0) * out.size(1) <= 65535:
                return _LeakyInstanceNorm.apply(out, norm.weight, norm.bias, norm.eps, self[1].negative_slope)
            return norm(self[1](out))
        return super().forward(block_input)


def conv_block(dims, n_in, n_out, kernel_size, stride=1, transposed=False, padding=None):
    """[Conv | ConvTranspose]{dims}d -> LeakyReLU(0.1) -> InstanceNorm{dims}d(affine)."""
    if transposed:
        conv = {3: nn.ConvTranspose3d}[dims](n_in, n_out, kernel_size=kernel_size,
                                             stride=stride, padding=padding)
    else:
        conv = {2: nn.Conv2d, 3: nn.Conv3d}[dims](n_in, n_out, kernel_size=kernel_size,
                                                  stride=stride, padding=kernel_size // 2)
    norm = {2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[dims](n_out, affine=True)
    return ConvBlock(conv, nn.LeakyReLU(negative_slope=LEAKY_SLOPE, inplace=True), norm)


def convolution_3x3(n_in, n_out):
    return nn.Conv2d(n_in, n_out, kernel_size=3, padding=1)


def convolutional_block_3x3(n_in, n_out):
    return conv_block(2, n_in, n_out, 3)


def convolutional_block_5x5_stride_2(n_in, n_out):
    return conv_block(2, n_in, n_out, 5, stride=2)


def convolutional_block_3x3x3(n_in, n_out):
    return conv_block(3, n_in, n_out, 3)


def convolutional_block_3x3x3_stride_2(n_in, n_out):
    return conv_block(3, n_in, n_out, 3, stride=2)


def transposed_convolutional_block_4x4x4_stride_2(n_in, n_out):
    return conv_block(3, n_in, n_out, 4, stride=2, transposed=True, padding=1)


def transposed_convolution_3x4x4_stride_122(n_in, n_out):
    return nn.ConvTranspose3d(n_in, n_out, kernel_size=(3, 4, 4), stride=(1, 2, 2),
                              padding=(1, 1, 1))


class ResidualBlock(nn.Module):
    """x + block(block(x)); no activation after the sum (reference
    network_blocks.py:134-144).  The attribute name is part of the state_dict."""

    def __init__(self, number_of_features):
        super().__init__()
        self.convolutions = nn.Sequential(
            convolutional_block_3x3(number_of_features, number_of_features),
            convolutional_block_3x3(number_of_features, number_of_features))

    def forward(self, block_input):
        return self.convolutions(block_input) + block_input
