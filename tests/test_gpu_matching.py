"""a1 / a2 parity: Matching (volume kernels) and MatchingOperation (fused conv
pipeline) vs golden vectors, the C oracle and the torch port (B200 only)."""
import numpy as np
import pytest
import torch

from oracle import oracle, synth, torch_port
from practicaldeepstereo_nips2018_b200 import matching
from gpu_util import cuda, load_module, max_abs, tdict

pytestmark = pytest.mark.gpu


def mockup_operation(x):
    return torch.max(x, dim=1, keepdim=True)[0]


def test_matching_known_answer():
    # reference test/test_matching.py:17-32, on the GPU kernels
    net = matching.Matching(maximum_disparity=2, operation=mockup_operation)
    left = torch.Tensor([0, 2, 1, 2]).view(1, 1, 1, 4).cuda()
    right = torch.Tensor([3, 4, 2, 4]).view(1, 1, 1, 4).cuda()
    expected = np.array([[3, 4, 2, 4], [0, 3, 4, 2], [0, 2, 3, 4]]).reshape(1, 1, 3, 1, 4)
    with torch.no_grad():
        assert np.array_equal(net(left, right).cpu().numpy(), expected)
        net.set_maximum_disparity(maximum_disparity=1)
        assert np.array_equal(net(left, right).cpu().numpy(), expected[:, :, :2])
        per_d = matching.Matching(2, mockup_operation, batched_operation=False)
        assert np.array_equal(per_d(left, right).cpu().numpy(), expected)


def test_matching_identity_golden(golden):
    g = golden('matching_known_answer')
    l, r = synth.tensor((2, 3, 4, 9), 11), synth.tensor((2, 3, 4, 9), 12)
    with torch.no_grad():
        out = matching.Matching(5, lambda x: x)(cuda(l), cuda(r))
    assert np.array_equal(out.cpu().numpy(), g['identity_md5'])   # bit-exact data movement


@pytest.mark.parametrize('shape,md', [((1, 64, 16, 32), 15), ((2, 5, 7, 13), 12), ((1, 8, 3, 40), 39),
                                      ((1, 2, 2, 3), 5)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_concat_vs_oracle(shape, md, dtype):
    l, r = synth.tensor(shape, 1), synth.tensor(shape, 2)
    lt, rt = cuda(l).to(dtype), cuda(r).to(dtype)
    with torch.no_grad():
        out = matching.Matching(md, lambda x: x)(lt, rt)
    ref = oracle.matching_concat(lt.float().cpu().numpy(), rt.float().cpu().numpy(), md)
    assert np.array_equal(out.float().cpu().numpy(), ref.transpose(0, 2, 1, 3, 4))


def test_concat_full_size_checksum():
    """C2 shape: 2 x (1, 64, 144, 240), 48 disparities -> 849 MB volume.  Checksum
    of checksums: sum over the volume == D*sum(left) + sum_d sum(right[..., :W-d])."""
    g = torch.Generator(device='cuda').manual_seed(1)
    left = torch.randn(1, 64, 144, 240, device='cuda', generator=g).double().float()
    right = torch.randn(1, 64, 144, 240, device='cuda', generator=g)
    with torch.no_grad():
        vol = matching.Matching(47, lambda x: x)(left, right)       # (1, 128, 48, 144, 240)
    assert vol.shape == (1, 128, 48, 144, 240)
    assert torch.equal(vol[:, :64, 17], left) and torch.equal(vol[:, :64, 47], left)
    for d in (0, 1, 13, 47):
        assert torch.equal(vol[:, 64:, d, :, d:], right[..., :240 - d])
        assert float(vol[:, 64:, d, :, :d].abs().sum()) == 0.0
    cs = right.double().cumsum(-1).sum(dim=(0, 1, 2))               # prefix sums over x
    expect = 48 * left.double().sum() + sum(cs[240 - d - 1] for d in range(48))
    assert abs(float(vol.double().sum() - expect)) <= 1e-6 * float(vol.double().abs().sum())


def test_matching_operation_golden(golden):
    params = synth.make_params(synth.matching_operation_specs(), 31)
    op = load_module(matching.MatchingOperation(precision='fp32'), params)
    x = synth.tensor((2, 128, 12, 14), 32)
    with torch.no_grad():
        out = op(cuda(x))
    assert out.shape == (2, 8, 12, 14)
    assert max_abs(out, golden('matching_operation')['out']) <= 1e-4
    assert max_abs(out, oracle.matching_operation(x, synth.flatten(params))) <= 1e-4


def test_matching_operation_output_size():
    # reference test/test_matching.py:35-40
    torch.manual_seed(0)
    op = matching.MatchingOperation(precision='fp32').cuda().eval()
    with torch.no_grad():
        assert op(torch.rand(2, 128, 25, 25).cuda()).size() == (2, 8, 25, 25)


def test_matching_golden(golden):
    params = synth.make_params(synth.matching_operation_specs(), 31)
    op = load_module(matching.MatchingOperation(precision='fp32'), params)
    l, r = synth.tensor((1, 64, 10, 24), 33), synth.tensor((1, 64, 10, 24), 34)
    with torch.no_grad():
        out = matching.Matching(7, op)(cuda(l), cuda(r))
    assert out.shape == (1, 8, 8, 10, 24)
    assert max_abs(out, golden('matching')['out']) <= 1e-4
    assert max_abs(out, oracle.matching(l, r, synth.flatten(params), 7)) <= 1e-4


@pytest.mark.parametrize('B,H,W,md', [(2, 16, 32, 15), (1, 21, 37, 9), (1, 48, 80, 23)])
def test_matching_vs_torch_port(B, H, W, md):
    """Fused pipeline vs the ATen restatement on the same device (TF32 disabled):
    ragged sizes, batch > 1, disparity larger than a tile."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = synth.make_params(synth.matching_operation_specs(), 35)
    op = load_module(matching.MatchingOperation(precision='fp32'), params)
    l, r = cuda(synth.tensor((B, 64, H, W), 36)), cuda(synth.tensor((B, 64, H, W), 37))
    p = tdict(params)
    with torch.no_grad():
        out = matching.Matching(md, op)(l, r)
        ref = torch_port.matching(l, r, lambda x: torch_port.matching_operation(x, p), md)
        generic = matching.Matching(md, lambda x: torch_port.matching_operation(x, p))(l, r)
    assert max_abs(out, ref) <= 2e-4
    assert max_abs(generic, ref) <= 2e-4      # volume kernel + batched generic operation


def test_parameter_update_rebuilds_kernel_weights():
    params = synth.make_params(synth.matching_operation_specs(), 31)
    op = load_module(matching.MatchingOperation(precision='fp32'), params)
    x = cuda(synth.tensor((1, 128, 8, 8), 5))
    with torch.no_grad():
        a = op(x)
        op._matching_operation_modules[3].bias.add_(1.0)
        b = op(x)
    assert max_abs(b - 1.0, a) <= 1e-5
