"""Parameter containers for the PDS layers.

The B200 kernels read the parameters straight from these modules, so the only
hard requirement is that ``state_dict()`` keys, shapes and semantics equal the
reference's (network_blocks.py:9-144 there; key families in SURVEY.md A.3):
every block is Conv -> LeakyReLU(0.1) -> InstanceNorm(affine, eps 1e-5), stored
as an ``nn.Sequential`` whose index 0 is the convolution and index 2 the norm.
The modules stay callable (plain ATen composition) for autograd / training,
which is outside the inference hot path.
"""
from torch import nn

LEAKY_SLOPE = 0.1


def conv_block(dims, n_in, n_out, kernel_size, stride=1, transposed=False, padding=None):
    """[Conv | ConvTranspose]{dims}d -> LeakyReLU(0.1) -> InstanceNorm{dims}d(affine)."""
    if transposed:
        conv = {3: nn.ConvTranspose3d}[dims](n_in, n_out, kernel_size=kernel_size,
                                             stride=stride, padding=padding)
    else:
        conv = {2: nn.Conv2d, 3: nn.Conv3d}[dims](n_in, n_out, kernel_size=kernel_size,
                                                  stride=stride, padding=kernel_size // 2)
    norm = {2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[dims](n_out, affine=True)
    return nn.Sequential(conv, nn.LeakyReLU(negative_slope=LEAKY_SLOPE, inplace=True), norm)


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
