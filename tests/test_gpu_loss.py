"""f4 (training half): the fused SubpixelCrossEntropy kernels (csrc/loss.cu) against the golden value
and gradients of the unmodified reference, the reference's own known-answer test and the oracle's
loop (oracle/torch_port.py) at a training-sized volume (B200 only)."""
import numpy as np
import pytest
import torch

from oracle import synth, torch_port
from practicaldeepstereo_nips2018_b200 import loss as pds_loss
from gpu_util import cuda, max_abs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('use_weights', [True, False])
def test_fused_loss_golden(golden, use_weights):
    g = golden('loss')
    tag = 'weighted' if use_weights else 'mean'
    sim = cuda(synth.tensor((2, 12, 9, 11), 71)).requires_grad_(True)      # 99 locations: scalar path
    gt = cuda(g['ground_truth'])
    w = cuda(g['weights'].copy()).requires_grad_(True) if use_weights else None
    value = pds_loss.SubpixelCrossEntropy(diversity=1.5, disparity_step=2)(sim, gt, w)
    value.backward()
    assert abs(value.item() - float(g[f'{tag}_loss'])) <= 2e-6 * abs(float(g[f'{tag}_loss']))
    assert max_abs(sim.grad, g[f'{tag}_grad_similarities']) <= 1e-7
    if use_weights:
        assert max_abs(w.grad, g['weighted_grad_weights']) <= 1e-6


def test_fused_loss_known_answer():
    # reference test/test_loss.py:13-38
    sim = torch.tensor([[0.1, 0.3, 0.2, 0.05], [0.2, 0.1, 0.4, 0.0], [0.2, 0.1, 0.4, 0.0]]).t().reshape(1, 4, 3, 1)
    sim = sim.contiguous().cuda().requires_grad_(True)
    gt = torch.tensor([1.3, float('inf'), 1.9]).view(1, 3, 1).cuda()
    w = torch.tensor([0.9, 0.0, 0.01]).view(1, 3, 1).cuda().requires_grad_(True)
    value = pds_loss.SubpixelCrossEntropy(diversity=2.0, disparity_step=1)(sim, gt, w)
    value.backward()
    expected = torch.tensor([[0.0262, -0.0567, -0.0219, 0.0524], [0.0, 0.0, 0.0, 0.0],
                             [0.0011, -0.0002, -0.0007, -0.0002]]).t().reshape(1, 4, 3, 1)
    assert abs(value.item() - 1.3654) <= 1e-3
    assert torch.allclose(sim.grad.cpu(), expected, atol=1e-3)


@pytest.mark.parametrize('use_weights', [True, False])
def test_fused_loss_vs_oracle_training_size(use_weights):
    """(2, 96, 120, 200): the shape class of a training crop at md=191; vectorised path, a strided
    (un-padded) view as PdsNetwork returns in training mode, a scaled upstream gradient."""
    torch.manual_seed(5)
    full = torch.randn(2, 96, 128, 256, device='cuda') * 3
    view = full[..., 8:, 56:].detach().requires_grad_(True)             # like SizeAdapter.unpad
    gt = torch.rand(2, 120, 200, device='cuda') * 190
    gt[0, :7] = float('inf')
    gt[1, 40:50, 100:] = float('inf')
    w = (torch.rand(2, 120, 200, device='cuda') + 0.05).requires_grad_(True) if use_weights else None
    value = pds_loss.SubpixelCrossEntropy(diversity=1.0, disparity_step=2)(view, gt, w)
    (3.0 * value).backward()
    ref_in = view.detach().double().requires_grad_(True)
    ref_w = w.detach().double().requires_grad_(True) if use_weights else None
    ref = torch_port.subpixel_cross_entropy(ref_in, gt.double(), ref_w, 1.0, 2)
    (3.0 * ref).backward()
    assert abs(value.item() - ref.item()) <= 2e-6 * abs(ref.item())
    scale = float(ref_in.grad.abs().max())
    assert max_abs(view.grad, ref_in.grad) <= 2e-5 * scale
    if use_weights:
        assert max_abs(w.grad, ref_w.grad) <= 2e-5 * float(ref_w.grad.abs().max())
