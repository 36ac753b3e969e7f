"""a4 parity: SubpixelMap kernel vs the oracle / golden vectors (B200 only).
Bar: arg-max indices bit-exact, disparity <= 1e-5 * scale on identical input."""
import numpy as np
import pytest
import torch

from oracle import oracle, synth, torch_port
from practicaldeepstereo_nips2018_b200.estimator import SubpixelMap
from gpu_util import cuda, max_abs

pytestmark = pytest.mark.gpu


def test_known_answer():
    # reference test/test_estimator.py:14-27
    sim = cuda(np.array([0.1, 0.4, 0.3, 0.2, 0.3], np.float32).reshape(1, 5, 1, 1))
    assert np.isclose(SubpixelMap(2, 1)(sim).item(), 1.52, atol=1e-4)
    assert np.isclose(SubpixelMap(2, 2)(sim).item(), 2.124, atol=1e-4)


@pytest.mark.parametrize('hsw,step', [(4, 2), (2, 1), (2, 2), (6, 3), (1, 1)])
def test_golden(golden, hsw, step):
    g = golden('estimator')
    for name, x in (('random', synth.tensor((2, 12, 9, 11), 21)),
                    ('adversarial', g['adversarial_in'])):
        d, idx = SubpixelMap(hsw, step)(cuda(x), return_argmax=True)
        assert np.array_equal(idx.cpu().numpy(), g[f'{name}_argmax'])
        assert max_abs(d, g[f'{name}_{hsw}_{step}']) <= 2e-6 * step * x.shape[1]


@pytest.mark.parametrize('shape', [(1, 32, 64, 128), (2, 96, 40, 60), (1, 7, 5, 13), (3, 1, 4, 4),
                                   (1, 2, 3, 2)])
@pytest.mark.parametrize('hsw,step', [(4, 2), (4, 1), (10, 2)])
def test_vs_oracle(shape, hsw, step):
    x = synth.tensor(shape, 7)
    d, idx = SubpixelMap(hsw, step)(cuda(x), return_argmax=True)
    rd, ridx = oracle.subpixel_map(x, hsw, step)
    assert np.array_equal(idx.cpu().numpy(), ridx)           # bit-exact indices
    assert max_abs(d, rd) <= 1e-5 * step * shape[1]


def test_ties_nan_and_edges():
    x = synth.tensor((1, 9, 6, 8), 3)
    x[0, :, 0, :] = 1.0                       # all equal -> index 0
    x[0, 2, 1, :] = x[0, 6, 1, :] = 50.0      # tie -> lowest index
    x[0, 0, 2, :] = 60.0                      # peak at index 0
    x[0, 8, 3, :] = 60.0                      # peak at index D-1
    x[0, 5, 4, 0] = np.nan
    x[0, 3, 4, 0] = np.nan                    # first NaN wins
    x[0, 4, 5, :] = np.inf
    d, idx = SubpixelMap(4, 2)(cuda(x), return_argmax=True)
    rd, ridx = oracle.subpixel_map(x, 4, 2)
    tidx = torch.max(torch.from_numpy(x), dim=1)[1].numpy()
    assert np.array_equal(ridx, tidx)
    assert np.array_equal(idx.cpu().numpy(), tidx)
    d = d.cpu().numpy()
    assert np.array_equal(np.isnan(d), np.isnan(rd))
    assert np.nanmax(np.abs(d - rd)) <= 1e-4
    assert idx[0, 0].eq(0).all() and idx[0, 1].eq(2).all() and idx[0, 4, 0] == 3


def test_fused_crop_and_unaligned():
    x = synth.tensor((2, 16, 20, 28), 9)
    full, fidx = SubpixelMap()(cuda(x), return_argmax=True)
    for top, left in ((0, 0), (3, 0), (0, 5), (4, 7), (19, 27)):
        d, idx = SubpixelMap()(cuda(x), crop_top=top, crop_left=left, return_argmax=True)
        assert torch.equal(d, full[..., top:, left:])
        assert torch.equal(idx, fidx[..., top:, left:])
    y = synth.tensor((1, 10, 6, 9), 10)       # W % 4 != 0 -> scalar path
    d, idx = SubpixelMap()(cuda(y), return_argmax=True)
    rd, ridx = oracle.subpixel_map(y)
    assert np.array_equal(idx.cpu().numpy(), ridx) and max_abs(d, rd) <= 1e-4
    z = cuda(synth.tensor((1, 10, 6, 12), 11))[:, :, :, ::1].transpose(2, 3)  # non-contiguous
    d = SubpixelMap()(z)
    rd, _ = oracle.subpixel_map(z.cpu().numpy())
    assert max_abs(d, rd) <= 1e-4
    empty = SubpixelMap()(torch.empty(0, 8, 4, 4, device='cuda'))
    assert empty.shape == (0, 4, 4)


def test_bf16_input_matches_torch_semantics():
    # reference semantics for bf16 cost volumes: float32 result, softmax in bf16
    x = torch.from_numpy(synth.tensor((1, 24, 16, 32), 13)).cuda().bfloat16()
    d, idx = SubpixelMap()(x, return_argmax=True)
    assert d.dtype == torch.float32
    tidx = torch.max(x, dim=1)[1]
    assert torch.equal(idx, tidx)
    rd, _ = oracle.subpixel_map(x.float().cpu().numpy())
    assert max_abs(d, rd) <= 0.25             # bf16 softmax rounding (2^-8 * 46)


def test_full_size_properties():
    """C2 shape (1, 96, 576, 960): compare with torch.max on the GPU and with the
    torch port; shift invariance: adding a constant per pixel changes nothing."""
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(1, 96, 576, 960, device='cuda', generator=g)
    d, idx = SubpixelMap()(x, crop_top=36, return_argmax=True)
    assert d.shape == (1, 540, 960)
    assert torch.equal(idx, torch.max(x, dim=1)[1][:, 36:])
    td, tidx = torch_port.subpixel_map(x)
    assert torch.equal(tidx[:, 36:], idx)
    assert max_abs(d, td[:, 36:]) <= 2e-4
    d2 = SubpixelMap()(x + 3.0, crop_top=36)
    assert max_abs(d, d2) <= 2e-3
    assert float(d.min()) >= 0.0 and float(d.max()) <= 190.0


def test_error_metrics_fused_kernel():
    """pds_disparity_errors (errors.py:9-74 in one pass) against the reference's known answers
    (test/test_errors.py) and against the tensor expressions on a full-size map with unknown
    (inf) ground truth, NaN estimates and ties at the threshold."""
    import math
    from practicaldeepstereo_nips2018_b200 import errors
    est = torch.tensor([[1.0, 2.0], [3.0, 4.0]]).cuda()
    gt = torch.tensor([[2.0, 2.0], [float('inf'), 1.0]]).cuda()
    pix, mean = errors.compute_absolute_error(est, gt)
    assert torch.equal(pix.cpu(), torch.tensor([[1.0, 0.0], [0.0, 3.0]])) and math.isclose(mean, 4.0 / 3.0, rel_tol=1e-6)
    pix, bad = errors.compute_n_pixels_error(est, gt, n=1.0)
    assert torch.equal(pix.cpu(), torch.tensor([[0.0, 0.0], [0.0, 1.0]])) and math.isclose(bad, 100.0 / 3.0, rel_tol=1e-6)
    nothing = torch.full((2, 2), float('inf')).cuda()
    assert errors.compute_absolute_error(est, nothing)[1] == 0.0
    assert errors.compute_n_pixels_error(est, nothing)[1] == 0.0
    assert math.isclose(errors.compute_absolute_error(est, gt, use_mean=False)[1], 1.0, rel_tol=1e-6)   # median: tensor path

    g = torch.Generator().manual_seed(7)
    est = (torch.rand(2, 540, 960, generator=g) * 190).cuda()
    gt = (est.cpu() + torch.randn(2, 540, 960, generator=g) * 2).cuda()
    gt[0, :50] = float('inf')
    gt[1, 100, 100] = float('-inf')
    gt[1, 7, 7] = est[1, 7, 7] + 3.0                       # exactly n: not an error (strict >)
    est[1, 9, 9] = float('nan')                            # NaN estimate: mean becomes NaN, gt() is false
    pa, ma, pb, bad = errors.compute_errors(est, gt, n=3.0)
    ra, rma = errors.compute_absolute_error(est.cpu(), gt.cpu())
    rb, rbad = errors.compute_n_pixels_error(est.cpu(), gt.cpu(), n=3.0)
    assert torch.equal(torch.nan_to_num(pa.cpu(), nan=-1.0), torch.nan_to_num(ra, nan=-1.0))
    assert torch.equal(pb.cpu(), rb)
    assert math.isnan(ma) and math.isnan(rma)
    assert math.isclose(bad, rbad, rel_tol=1e-5)
    est[1, 9, 9] = 0.0
    _, ma, _, _ = errors.compute_errors(est, gt, n=3.0)
    assert math.isclose(ma, errors.compute_absolute_error(est.cpu(), gt.cpu())[1], rel_tol=1e-5)
