"""CPU-only tests: host-side logic of the drop-in modules, state_dict
compatibility with the reference, argument validation, and that the C-ABI
library loads and exports every symbol declared in include/pds_b200.h."""
import os
import re

import pytest
import torch

from oracle import synth
from practicaldeepstereo_nips2018_b200 import (PdsNetwork, _capi, embedding, estimator,
                                               matching, network, regularization,
                                               size_adapter)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'pds_b200.h')).read()
    declared = set(re.findall(r'\b(pds_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    lib = _capi.lib()                      # raises if a symbol is missing
    assert lib.pds_version() == 100
    assert lib.pds_status_string(1) == b'invalid argument'


def test_state_dict_matches_reference_layout():
    net = PdsNetwork.default(191)
    sd = net.state_dict()
    specs = synth.network_specs()
    assert list(sd.keys()) == [k for k, _ in specs]
    assert len(sd) == 122
    for k, shape in specs:
        assert tuple(sd[k].shape) == shape, k
    assert sum(v.numel() for v in sd.values()) == 2217717
    assert list(matching.MatchingOperation().state_dict().keys()) == \
        [k for k, _ in synth.matching_operation_specs()]
    assert list(regularization.Regularization().state_dict().keys()) == \
        [k for k, _ in synth.regularization_specs()]
    assert list(embedding.Embedding().state_dict().keys()) == \
        [k for k, _ in synth.embedding_specs()]


def test_set_maximum_disparity_validation():
    net = PdsNetwork.default(63)
    assert net._matching._maximum_disparity == 15
    net.set_maximum_disparity(255)
    assert net._matching._maximum_disparity == 63
    with pytest.raises(ValueError):
        net.set_maximum_disparity(100)     # reference network.py:28-31


def test_estimator_validation():
    for hsw, step in ((4, 0), (0, 1), (3, 2)):
        with pytest.raises(ValueError):    # reference estimator.py:34-41
            estimator.SubpixelMap(hsw, step)
    estimator.SubpixelMap(4, 2)
    est = estimator.SubpixelMap()
    assert not isinstance(est, torch.nn.Module)


def test_size_adapter_roundtrip():
    adapter = size_adapter.SizeAdapter()   # reference test_size_adapter.py
    x = torch.rand(1, 10, 63, 100)
    padded = adapter.pad(x)
    assert padded.size() == (1, 10, 64, 128)
    assert (padded[..., :1, :] == 0).all() and (padded[..., :28] == 0).all()
    assert (adapter.unpad(padded) == x).all()
    assert adapter.padding_for(540, 960) == (36, 0)
    assert adapter.padding_for(375, 1242) == (9, 38)


def test_inference_path_refuses_cpu_tensors():
    # no CPU fallback: the kernel path must fail loudly
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            estimator.SubpixelMap()(torch.rand(1, 8, 4, 4))
        with pytest.raises(RuntimeError):
            matching.Matching(2, lambda x: x)(torch.rand(1, 2, 3, 4), torch.rand(1, 2, 3, 4))
        with pytest.raises(RuntimeError):
            matching.MatchingOperation()(torch.rand(1, 128, 8, 8))
        with pytest.raises(RuntimeError):
            regularization.Regularization()(torch.rand(1, 8, 16, 16, 16), torch.rand(1, 8, 16, 16))


def test_training_mode_is_aten_composition():
    # gradient-enabled calls are outside the inference hot path: ATen ops, autograd works
    torch.manual_seed(0)
    op = matching.MatchingOperation()
    m = matching.Matching(maximum_disparity=3, operation=op)
    left = torch.rand(1, 64, 6, 8, requires_grad=True)
    out = m(left, torch.rand(1, 64, 6, 8))
    assert out.shape == (1, 8, 4, 6, 8)
    out.sum().backward()
    assert left.grad is not None and op._matching_operation_modules[0].weight.grad is not None


def test_training_regularization_shapes():
    torch.manual_seed(0)                   # reference test_regularization.py:31-36 (smaller)
    reg = regularization.Regularization()
    cost = reg(torch.rand(1, 8, 16, 16, 32), torch.rand(1, 8, 16, 32))
    assert cost.shape == (1, 32, 64, 128) and cost.requires_grad
    down, smooth = regularization.ContractionBlock3d(6)(torch.rand(2, 6, 10, 14, 16))
    assert down.shape == smooth.shape == (2, 12, 5, 7, 8)
    out = regularization.ExpansionBlock3d(6)(torch.rand(2, 6, 10, 14, 16), torch.rand(2, 3, 20, 28, 32))
    assert out.shape == (2, 3, 20, 28, 32)


def test_network_train_mode_matches_golden(golden):
    # train mode (cost volume output) runs the ATen composition on CPU: must equal the
    # reference bit-for-bit-ish since it is the same operator sequence
    g = golden('network_md63')
    params = synth.make_params(synth.network_specs(), 61)
    net = PdsNetwork.default(63)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net.train()
    li = synth.tensor((1, 3, 62, 100), 62, scale=255.0, uniform=True)
    ri = synth.tensor((1, 3, 62, 100), 63, scale=255.0, uniform=True)
    ri[..., :-6] = 0.8 * li[..., 6:] + 0.2 * ri[..., :-6]
    cost = net(torch.from_numpy(li), torch.from_numpy(ri))
    assert cost.shape == (1, 32, 62, 100)
    assert torch.allclose(cost.detach(), torch.from_numpy(g['cost_unpadded']), atol=1e-4)


def test_precision_does_not_change_the_state_dict():
    """Every precision builds the reference's 122-key state_dict (checkpoints load unchanged)."""
    ref = list(PdsNetwork.default(191).state_dict().keys())
    for precision in _capi.PRECISIONS:
        assert list(PdsNetwork.default(191, precision=precision).state_dict().keys()) == ref
    with pytest.raises(ValueError):
        embedding.Embedding(precision='int8')
    with pytest.raises(ValueError):
        regularization.Regularization(precision='int8')


def test_kernel_paths_refuse_cpu_tensors_and_keep_autograd():
    """No CPU fallback on the product path: eval / no-grad calls on CPU tensors raise; calls that
    need gradients run the ATen composition (training keeps working)."""
    torch.manual_seed(0)
    net = PdsNetwork.default(63, precision='fp16x2').eval()
    left, right = torch.rand(1, 3, 64, 128), torch.rand(1, 3, 64, 128)   # smallest legal size
    with torch.no_grad(), pytest.raises(RuntimeError):
        net(left, right)
    emb = embedding.Embedding(precision='fp16x2')
    assert not emb.uses_kernels(left)                       # CPU tensor: module composition
    with torch.no_grad():
        d, s = emb(left)
    assert d.shape == (1, 64, 16, 32) and s.shape == (1, 8, 16, 32)
    net.train()
    cost = net(left, right)                                 # autograd path
    assert cost.shape == (1, 32, 64, 128) and cost.requires_grad
    cost.mean().backward()
    assert all(p.grad is not None for p in net.parameters())


def test_host_pipeline_is_importable_without_a_gpu():
    from practicaldeepstereo_nips2018_b200 import pipeline
    assert callable(pipeline.HostPipeline)


def test_error_metrics_known_answers_on_cpu():
    """errors.py mirror: the reference's own known-answer cases (test/test_errors.py:13-66);
    CPU tensors take the reference's tensor expressions."""
    import math
    from practicaldeepstereo_nips2018_b200 import errors
    est = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    gt = torch.tensor([[2.0, 2.0], [float('inf'), 1.0]])
    pix, mean = errors.compute_absolute_error(est, gt, use_mean=True)
    assert torch.equal(pix, torch.tensor([[1.0, 0.0], [0.0, 3.0]])) and math.isclose(mean, 4.0 / 3.0, rel_tol=1e-6)
    pix, median = errors.compute_absolute_error(est, gt, use_mean=False)
    assert math.isclose(median, 1.0, rel_tol=1e-6)
    pix, bad = errors.compute_n_pixels_error(est, gt, n=1.0)
    assert torch.equal(pix, torch.tensor([[0.0, 0.0], [0.0, 1.0]])) and math.isclose(bad, 100.0 / 3.0, rel_tol=1e-6)
    nothing = torch.full((2, 2), float('inf'))
    assert errors.compute_absolute_error(est, nothing)[1] == 0.0
    assert errors.compute_n_pixels_error(est, nothing)[1] == 0.0
