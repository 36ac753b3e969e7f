"""B200-native (sm_100a) stereo cost-volume pipeline behind
``PdsNetwork.forward`` of tlkvstepan/PracticalDeepStereo_NIPS2018.

Module names mirror the reference package ``practical_deep_stereo``:
``network.PdsNetwork``, ``matching.Matching`` / ``MatchingOperation``,
``regularization.Regularization``, ``estimator.SubpixelMap``,
``embedding.Embedding``, ``size_adapter.SizeAdapter``, ``errors``, ``loss.SubpixelCrossEntropy``.
"""
from . import (embedding, errors, estimator, loss, matching, network, network_blocks,  # noqa: F401
               regularization, size_adapter)
from .network import PdsNetwork  # noqa: F401

__version__ = '0.1.0'
