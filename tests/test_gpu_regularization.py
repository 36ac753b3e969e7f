"""a3 parity: the hourglass and its blocks vs golden vectors, the C oracle and
the torch port (B200 only)."""
import numpy as np
import pytest
import torch

from oracle import synth, torch_port
from practicaldeepstereo_nips2018_b200 import regularization
from gpu_util import cuda, load_module, max_abs, tdict

pytestmark = pytest.mark.gpu


def test_blocks_golden(golden):
    g = golden('regularization_blocks')
    x = synth.tensor((2, 6, 10, 14, 16), 42)
    skip = synth.tensor((2, 3, 20, 28, 32), 44)
    cb = load_module(regularization.ContractionBlock3d(6),
                     synth.make_params(synth.contraction_block_specs(6), 41))
    eb = load_module(regularization.ExpansionBlock3d(6),
                     synth.make_params(synth.expansion_block_specs(6), 43))
    with torch.no_grad():
        down, smooth = cb(cuda(x))
        out = eb(cuda(x), cuda(skip))
    assert down.shape == smooth.shape == (2, 12, 5, 7, 8)    # test_regularization.py:13-19
    assert out.shape == (2, 3, 20, 28, 32)                   # test_regularization.py:22-28
    assert max_abs(down, g['down']) <= 1e-4
    assert max_abs(smooth, g['smooth']) <= 1e-4
    assert max_abs(out, g['expansion']) <= 1e-4


def test_blocks_fast_channel_counts():
    """8 / 16-channel blocks take the vectorised kernels (the golden case uses 6)."""
    torch.backends.cudnn.allow_tf32 = False
    for c in (8, 16):
        pc = synth.make_params(synth.contraction_block_specs(c), 70 + c)
        pe = synth.make_params(synth.expansion_block_specs(c), 80 + c)
        x = cuda(synth.tensor((1, c, 8, 12, 20), 90 + c))
        skip = cuda(synth.tensor((1, c // 2, 16, 24, 40), 91 + c))
        cb = load_module(regularization.ContractionBlock3d(c), pc)
        eb = load_module(regularization.ExpansionBlock3d(c), pe)
        with torch.no_grad():
            down, smooth = cb(x)
            out = eb(x, skip)
            rd, rs = torch_port.contraction_block(x, tdict(pc), '')
            ro = torch_port.expansion_block(x, skip, tdict(pe), '')
        assert max_abs(down, rd) <= 1e-4 and max_abs(smooth, rs) <= 1e-4
        assert max_abs(out, ro) <= 1e-4


def test_hourglass_golden(golden):
    params = synth.make_params(synth.regularization_specs(), 45)
    reg = load_module(regularization.Regularization(precision='fp32'), params)
    sig, sc = synth.tensor((1, 8, 16, 16, 32), 46), synth.tensor((1, 8, 16, 32), 47)
    with torch.no_grad():
        out = reg(cuda(sig), cuda(sc))
    assert out.shape == (1, 32, 64, 128)
    # fp32 noise floor of this tiny-bottleneck config (1x1x2 voxels under InstanceNorm):
    # reference vs its own fp64 run = 1.8e-4, amplified up to 316x by the 2-voxel normalisation
    assert max_abs(out, golden('regularization')['out']) <= 3e-3


def test_hourglass_well_conditioned_vs_torch_port():
    """Bottleneck of 2x2x4 voxels: the hourglass (incl. the fused InstanceNorm + transposed
    convolution tail) against ATen fp32 on the same device at the fp32 noise floor."""
    torch.backends.cudnn.allow_tf32 = False
    params = synth.make_params(synth.regularization_specs(), 48)
    reg = load_module(regularization.Regularization(precision='fp32'), params)
    sig, sc = cuda(synth.tensor((2, 8, 32, 32, 64), 49)), cuda(synth.tensor((2, 8, 32, 64), 50))
    with torch.no_grad():
        out = reg(sig, sc)
        ref = torch_port.regularization(sig, sc, tdict(params))
    assert out.shape == ref.shape == (2, 64, 128, 256)
    assert max_abs(out, ref) <= 2e-4


def test_hourglass_output_size():
    # reference test/test_regularization.py:31-36
    torch.manual_seed(0)
    reg = regularization.Regularization(precision='fp32').cuda().eval()
    with torch.no_grad():
        cost = reg(torch.rand(2, 8, 32, 32, 32).cuda(), torch.rand(2, 8, 32, 32).cuda())
    assert cost.size() == (2, 64, 128, 128)


def test_hourglass_vs_torch_port():
    """Larger volume (bottleneck 2x3x4) with batch 2, against ATen on the same GPU."""
    torch.backends.cudnn.allow_tf32 = False
    params = synth.make_params(synth.regularization_specs(), 48)
    reg = load_module(regularization.Regularization(precision='fp32'), params)
    sig, sc = cuda(synth.tensor((2, 8, 32, 48, 64), 49)), cuda(synth.tensor((2, 8, 48, 64), 50))
    with torch.no_grad():
        out = reg(sig, sc)
        ref = torch_port.regularization(sig, sc, tdict(params))
    scale = float(ref.abs().max())
    assert max_abs(out, ref) <= 2e-5 * scale + 2e-4
    with torch.no_grad(), pytest.raises(ValueError):
        reg(sig[:, :, :24], sc)               # D not a multiple of 16
