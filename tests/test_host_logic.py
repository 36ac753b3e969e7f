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


def test_default_precision_is_the_tensor_core_mode():
    """The reference's unchanged `PdsNetwork.default().cuda()` (benchmark_on_flyingthings3d.py:
    56-59) must land on the tcgen05 path: every stage defaults to the fp32-grade split precision."""
    assert _capi.DEFAULT_PRECISION == os.environ.get('PDS_B200_PRECISION', 'fp16x2')
    net = PdsNetwork.default(191)
    assert net._embedding.precision == net._matching._operation.precision == \
        net._regularization.precision == _capi.DEFAULT_PRECISION
    assert matching.MatchingOperation().precision == _capi.DEFAULT_PRECISION
    assert PdsNetwork.default(63, precision='fp32')._regularization.precision == 'fp32'


def test_kernel_handles_survive_deepcopy_and_pickle():
    """the handle holds ctypes pointers -- copies get a fresh empty handle bound to the
    copied module, and invalidate() drops packed weights after `.data` writes."""
    import copy
    import io
    import ctypes
    net = PdsNetwork.default(63)
    op = net._matching._operation
    op._kernel._handles['cuda:9'] = (ctypes.c_void_p(0), ('fake',))      # as after a kernel forward
    destroyed = []
    op._kernel._destroy = destroyed.append
    clone = copy.deepcopy(net)
    cop = clone._matching._operation
    assert cop._kernel is not op._kernel and cop._kernel._handles == {}
    assert cop._kernel._create.__self__ is cop                          # re-bound to the copy
    buf = io.BytesIO()
    torch.save(op, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)
    assert loaded._kernel._handles == {} and loaded._kernel._create.__self__ is loaded
    net.invalidate_kernels()
    assert op._kernel._handles == {} and len(destroyed) == 1
    # standalone hourglass blocks keep their scratch buffer between calls (and copies do not share it)
    blk = regularization.ContractionBlock3d(4)
    assert copy.deepcopy(blk)._scratch is not blk._scratch


def test_handle_key_tracks_in_place_updates():
    """optimizer steps / load_state_dict / copy_ under no_grad bump `_version` (the handle's key);
    `.data` writes do not, which is why parallel.broadcast_parameters copies into the parameters."""
    p = torch.nn.Parameter(torch.zeros(3))
    v0 = p._version
    with torch.no_grad():
        p.copy_(torch.ones(3))
    assert p._version > v0
    v1 = p._version
    p.data.mul_(2.0)
    assert p._version == v1                # documented restriction -> invalidate_kernels()


def test_fp16_operand_range_check():
    w = [torch.zeros(8, 8, 3, 3), torch.zeros(8)]
    matching.check_fp16_weight_range(w, 'fp16x2')
    w[0][0, 0, 0, 0] = 300.0
    with pytest.raises(ValueError):
        matching.check_fp16_weight_range(w, 'fp16x2')
    matching.check_fp16_weight_range(w, 'bf16x3')      # bfloat16 terms have fp32's range


def test_forward_validates_inputs_up_front():
    net = PdsNetwork.default(63).eval()
    left = torch.rand(1, 3, 64, 128)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match='no CPU path'):
            net(left, left)
        with pytest.raises(ValueError):
            net(left[0], left[0])
    with pytest.warns(RuntimeWarning, match='eval mode but gradients are enabled'):
        network.PdsNetwork._warned_eval_grad = False
        net(left, left)                    # eval + grad: ATen composition, announced once


def test_estimator_gradient_path_known_answers():
    """reference test/test_estimator.py:14-27 on the gradient-enabled (tensor expression) path."""
    sim = torch.tensor([0.0, 1.0, 3.0, 2.0, 1.0, 0.5]).view(1, 6, 1, 1).repeat(1, 1, 2, 1)
    sim[0, :, 1, 0] = torch.tensor([5.0, 1.0, 0.0, 0.0, 0.0, 0.0])
    sim.requires_grad_(True)
    est = estimator.SubpixelMap(half_support_window=2, disparity_step=1)
    out, idx = est(sim, return_argmax=True)
    w = torch.softmax(torch.tensor([0.0, 1.0, 3.0, 2.0, 1.0]), 0)
    assert torch.allclose(out[0, 0, 0], (w * torch.arange(5.0)).sum(), atol=1e-6)
    w = torch.softmax(torch.tensor([5.0, 1.0, 0.0]), 0)        # taps -2, -1 fall outside: weight 0
    assert torch.allclose(out[0, 1, 0], (w * torch.arange(3.0)).sum(), atol=1e-6)
    assert idx.tolist() == [[[2], [0]]]
    out.sum().backward()
    assert sim.grad is not None


def test_numa_binding_helpers(tmp_path):
    from practicaldeepstereo_nips2018_b200 import parallel
    assert parallel._parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    dev = tmp_path / 'bus/pci/devices/0000:1b:00.0'
    dev.mkdir(parents=True)
    (dev / 'numa_node').write_text('1\n')
    node = tmp_path / 'devices/system/node/node1'
    node.mkdir(parents=True)
    (node / 'cpulist').write_text('16-31\n')
    assert parallel.gpu_numa_cpus('0000:1B:00.0', sysfs=str(tmp_path)) == list(range(16, 32))
    (dev / 'numa_node').write_text('-1\n')
    assert parallel.gpu_numa_cpus('0000:1b:00.0', sysfs=str(tmp_path)) is None
    assert parallel.gpu_numa_cpus('0000:ff:00.0', sysfs=str(tmp_path)) is None


def test_conv_block_is_a_plain_sequential_on_cpu():
    """network_blocks.ConvBlock keeps the reference's module indices and, on CPU tensors (or without
    gradients), runs the plain nn.Sequential composition: equal to Conv -> LeakyReLU -> InstanceNorm."""
    from practicaldeepstereo_nips2018_b200 import network_blocks
    torch.manual_seed(1)
    block = network_blocks.conv_block(3, 8, 8, 3)
    assert isinstance(block, torch.nn.Sequential)
    assert sorted(block.state_dict()) == ['0.bias', '0.weight', '2.bias', '2.weight']
    x = torch.randn(2, 8, 4, 6, 7, requires_grad=True)
    y = block(x)
    ref = torch.nn.functional.instance_norm(
        torch.nn.functional.leaky_relu(block[0](x), 0.1), weight=block[2].weight, bias=block[2].bias, eps=1e-5)
    assert torch.allclose(y, ref, atol=1e-6)
    y.sum().backward()
    assert x.grad is not None and block[2].weight.grad is not None


def test_training_entry_points_validate_arguments_without_a_gpu():
    """The f4 entry points reject bad arguments before any CUDA call (status 1 = invalid argument)."""
    lib = _capi.lib()
    assert lib.pds_matching_concat_backward(None, None, None, 1, 4, 3, 5, 2, 0, None) == 1
    assert lib.pds_matching_unstack(None, None, 1, 2, 3, 4, 5, 0, None) == 1
    assert lib.pds_instance_norm_forward(None, None, None, None, None, None, 1, 4, 10, 1e-5, 0.1, None) == 1
    assert lib.pds_instance_norm_backward(None, None, None, None, None, None, 1, 4, 10, 0.1, None) == 1
    assert b'null pointer' in lib.pds_last_error()
