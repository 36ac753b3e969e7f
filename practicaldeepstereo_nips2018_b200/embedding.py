"""Embedding tower (reference embedding.py:11-65).

Adjacent to the hot path (SURVEY.md 8 row a6, 4.6 % of the FLOPs, not named by
the north star): it is kept on stock ATen/cuDNN operators in this round and is
listed as "next" (f2) in DESIGN.md.  Structure and state_dict keys are the
reference's."""
from torch import nn

from . import network_blocks


class Embedding(nn.Module):
    def __init__(self, number_of_input_features=3, number_of_embedding_features=64,
                 number_of_shortcut_features=8, number_of_residual_blocks=2):
        super().__init__()
        f = number_of_embedding_features
        tower = [nn.InstanceNorm2d(number_of_input_features),
                 network_blocks.convolutional_block_5x5_stride_2(number_of_input_features, f),
                 network_blocks.convolutional_block_5x5_stride_2(f, f)]
        tower += [network_blocks.ResidualBlock(f) for _ in range(number_of_residual_blocks)]
        self._embedding_modules = nn.ModuleList(tower)
        self._shortcut = network_blocks.convolutional_block_3x3(f, number_of_shortcut_features)

    def forward(self, image, with_shortcut=True):
        """image (B,3,H,W) -> descriptor (B,64,H/4,W/4), shortcut (B,8,H/4,W/4).

        ``with_shortcut=False`` skips the shortcut block, whose result the
        reference computes for the right image and throws away (network.py:40)."""
        descriptor = image
        for module in self._embedding_modules:
            descriptor = module(descriptor)
        if not with_shortcut:
            return descriptor, None
        return descriptor, self._shortcut(descriptor)
