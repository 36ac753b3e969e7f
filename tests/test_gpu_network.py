"""a5 parity: PdsNetwork.forward end to end vs the golden vectors of the
unmodified reference and vs the torch port (B200 only)."""
import numpy as np
import pytest
import torch

from oracle import oracle, synth, torch_port
from practicaldeepstereo_nips2018_b200 import PdsNetwork
from gpu_util import cuda, load_module, max_abs, tdict

pytestmark = pytest.mark.gpu


def _inputs():
    li = synth.tensor((1, 3, 62, 100), 62, scale=255.0, uniform=True)
    ri = synth.tensor((1, 3, 62, 100), 63, scale=255.0, uniform=True)
    ri[..., :-6] = 0.8 * li[..., 6:] + 0.2 * ri[..., :-6]
    return li, ri


def margin_aware(disp, idx, ref_disp, ref_cost, crop, cost_err):
    """Disparity parity on pixels whose reference top-1/top-2 margin exceeds the
    measured cost error; returns (flip fraction, max-abs there)."""
    top2 = np.sort(ref_cost, axis=1)[:, -2:]
    margin = (top2[:, 1] - top2[:, 0])[..., crop[0]:, crop[1]:]
    ref_idx = ref_cost.argmax(axis=1)[..., crop[0]:, crop[1]:]
    safe = margin > 4 * cost_err
    assert np.array_equal(idx[safe], ref_idx[safe])          # bit-exact where well-defined
    flips = float((idx != ref_idx).mean())
    return flips, float(np.abs(disp - ref_disp)[safe].max()), float(safe.mean())


def test_network_golden(golden):
    torch.backends.cudnn.allow_tf32 = False
    g = golden('network_md63')
    net = load_module(PdsNetwork.default(63, precision='fp32'), synth.make_params(synth.network_specs(), 61))
    li, ri = _inputs()
    with torch.no_grad():
        left, right = cuda(li), cuda(ri)
        disp = net(left, right)
        sig = net._matching(*[net._embedding(net._size_adapter.pad(t))[0] for t in (left, right)])
        cost = net.pass_through_network(net._size_adapter.pad(left), net._size_adapter.pad(right))[0]
        d2, idx = net._estimator(cost, crop_top=2, crop_left=28, return_argmax=True)
    assert disp.shape == (1, 62, 100) and torch.equal(disp, d2)
    assert max_abs(sig, g['signatures']) <= 2e-4
    cost_err = max_abs(cost, g['cost_padded'])
    # 62x100 / md=63 has a 1x1x2 hourglass bottleneck: InstanceNorm over TWO voxels amplifies
    # input noise by up to 1/sqrt(eps) = 316x, so the reference itself moves by 1.3e-3 between
    # fp32 and fp64 here (and the GPU embedding runs on cuDNN, the golden one on oneDNN).  This
    # fixture therefore pins the well-conditioned stages (signatures, estimator on an identical
    # volume, shapes); the END-TO-END bound of north_star (arg-max bit-exact, disparity 1e-3) is
    # asserted on the well-conditioned fixture in test_network_golden_well_conditioned.
    assert cost_err <= 1e-2
    # estimator on the identical volume is exact
    rd, ridx = oracle.subpixel_map(cost.cpu().numpy())
    assert np.array_equal(ridx[..., 2:, 28:], idx.cpu().numpy())
    assert max_abs(rd[..., 2:, 28:], disp) <= 1e-4


@pytest.mark.parametrize('precision', ['fp16x2', 'bf16x3', 'fp32'])
def test_network_golden_well_conditioned(golden, precision):
    """north_star's bound against the UNMODIFIED reference (tests/golden/network_md127.npz, 250x120,
    md=127, hourglass bottleneck 2x2x4): arg-max bit-exact and disparity <= 1e-3 on every pixel
    whose reference top-1/top-2 margin exceeds 4x the measured cost error; >= 99 % of the pixels
    are such pixels."""
    from test_oracle_golden import margin_checks, md127_inputs
    torch.backends.cudnn.allow_tf32 = False
    g = golden('network_md127')
    net = load_module(PdsNetwork.default(127, precision=precision), synth.make_params(synth.network_specs(), 61))
    li, ri = md127_inputs()
    with torch.no_grad():
        left, right = cuda(li), cuda(ri)
        disp = net(left, right)
        cost = net.pass_through_network(net._size_adapter.pad(left), net._size_adapter.pad(right))[0]
        d2, idx = net._estimator(cost, crop_top=8, crop_left=6, return_argmax=True)
    assert disp.shape == (1, 120, 250) and torch.equal(disp, d2)
    cost_err, safe, flips, err = margin_checks(g, cost.cpu().numpy(), disp.cpu().numpy(), idx.cpu().numpy(), 1e-3)
    print(f'network_md127 {precision}: cost max-abs {cost_err:.3e} safe {safe:.4f} flips {flips:.3e} '
          f'disparity max-abs on safe pixels {err:.3e}')
    assert safe >= 0.99 and flips <= 1e-3 and err <= 1e-3, (cost_err, safe, flips, err)


def test_network_shapes_and_modes():
    # reference test/test_network.py:11-27 with a legal size for torch >= 2
    torch.manual_seed(0)
    net = PdsNetwork.default(63).cuda()
    left, right = torch.rand(1, 3, 120, 200).cuda(), torch.rand(1, 3, 120, 200).cuda()
    net.train()
    assert net(left, right).size() == (1, 32, 120, 200)
    net.set_maximum_disparity(127)
    assert net(left, right).size() == (1, 64, 120, 200)
    net.eval()
    with torch.no_grad():
        assert net(left, right).size() == (1, 120, 200)
    with pytest.raises(ValueError):
        net.set_maximum_disparity(100)


@pytest.mark.parametrize('precision', [None, 'fp32'])
def test_network_vs_torch_port_kitti_like(precision):
    """C4-like aspect (ragged, needs both pads) at reduced size, batch 2; `None` = what the
    reference's unchanged PdsNetwork.default() call gets (the fp32-grade tensor-core mode)."""
    torch.backends.cudnn.allow_tf32 = False
    params = synth.make_params(synth.network_specs(), 71)
    net = load_module(PdsNetwork.default(63, precision=precision), params)
    from practicaldeepstereo_nips2018_b200 import _capi
    assert net._matching._operation.precision == (precision or _capi.DEFAULT_PRECISION)
    left = cuda(synth.tensor((2, 3, 100, 310), 72, scale=255.0, uniform=True))
    right = cuda(synth.tensor((2, 3, 100, 310), 73, scale=255.0, uniform=True))
    right[..., :-9] = 0.7 * left[..., 9:] + 0.3 * right[..., :-9]
    with torch.no_grad():
        disp = net(left, right)
        cost = net.pass_through_network(net._size_adapter.pad(left), net._size_adapter.pad(right))[0]
        _, idx = net._estimator(cost, crop_top=28, crop_left=10, return_argmax=True)
        st = torch_port.network_stages(left, right, tdict(params), 63)
    assert disp.shape == (2, 100, 310)
    cost_err = max_abs(cost, st['cost'])
    assert cost_err <= 1e-3
    flips, err, safe = margin_aware(disp.cpu().numpy(), idx.cpu().numpy(),
                                    st['disparity'].cpu().numpy(), st['cost'].cpu().numpy(),
                                    (28, 10), cost_err)
    print(f'kitti-like md63: cost max-abs {cost_err:.3e} safe {safe:.4f} flips {flips:.3e} disparity {err:.3e}')
    assert safe >= 0.99 and flips <= 1e-3 and err <= 1e-3, (cost_err, safe, flips, err)


def test_host_pipeline_matches_direct_calls():
    """pipeline.HostPipeline (uploads overlapped with the previous forward) returns exactly what
    PdsNetwork.forward returns for every pair, in order."""
    from practicaldeepstereo_nips2018_b200.pipeline import HostPipeline
    torch.manual_seed(3)
    net = PdsNetwork.default(63, precision='fp16x2').cuda().eval()
    pairs = [(torch.rand(1, 3, 64, 128).mul(255).pin_memory(), torch.rand(1, 3, 64, 128).mul(255).pin_memory())
             for _ in range(5)]
    with torch.no_grad():
        direct = [net(l.cuda(), r.cuda()).cpu() for l, r in pairs]
    out = HostPipeline(net).run(pairs)
    torch.cuda.synchronize()
    assert len(out) == len(pairs)
    for a, b in zip(out, direct):
        assert torch.equal(a, b)
    # reused output buffers (round-robin of 2): the last two results are still intact
    bufs = [torch.empty(1, 64, 128).pin_memory() for _ in range(2)]
    out2 = HostPipeline(net).run(pairs, out=bufs)
    torch.cuda.synchronize()
    assert torch.equal(out2[-1], direct[-1]) and torch.equal(out2[-2], direct[-2])
    # two compute streams (pairs dealt round-robin, one scratch buffer per stream): same results
    for _ in range(3):
        out3 = HostPipeline(net, streams=2).run(pairs)
        torch.cuda.synchronize()
        for a, b in zip(out3, direct):
            assert torch.equal(a, b)


def test_graphed_network_and_graph_pipeline_match_eager_calls():
    """pipeline.GraphedNetwork (one CUDA graph per input signature, replayed) and
    HostPipeline(graphs=True) return exactly what PdsNetwork.forward returns; a change of extent
    captures a second graph and the first one stays valid (the handles keep the plans and weight
    images of earlier extents)."""
    from practicaldeepstereo_nips2018_b200.pipeline import GraphedNetwork, HostPipeline
    torch.manual_seed(4)
    net = PdsNetwork.default(63).cuda().eval()
    small = [(torch.rand(1, 3, 64, 128).mul(255).cuda(), torch.rand(1, 3, 64, 128).mul(255).cuda()) for _ in range(3)]
    wide = [(torch.rand(2, 3, 100, 310).mul(255).cuda(), torch.rand(2, 3, 100, 310).mul(255).cuda()) for _ in range(2)]
    with torch.no_grad():
        eager_small = [net(l, r).clone() for l, r in small]
        eager_wide = [net(l, r).clone() for l, r in wide]
    g = GraphedNetwork(net)
    for (l, r), ref in zip(small, eager_small):
        assert torch.equal(g(l, r), ref)
    for (l, r), ref in zip(wide, eager_wide):          # second signature: second graph
        assert torch.equal(g(l, r), ref)
    for (l, r), ref in zip(small, eager_small):        # the first graph is still valid
        assert torch.equal(g(l, r), ref)
    with torch.no_grad():                              # eager calls still work next to the graphs
        assert torch.equal(net(*small[0]), eager_small[0])
    pinned = [(l.cpu().pin_memory(), r.cpu().pin_memory()) for l, r in small]
    for streams in (1, 2):
        out = HostPipeline(net, streams=streams, graphs=True).run(pinned + pinned)
        torch.cuda.synchronize()
        for a, ref in zip(out, eager_small + eager_small):
            assert torch.equal(a.cuda(), ref)
        dev = HostPipeline(net, streams=streams, graphs=True).run(small, download=False)
        torch.cuda.synchronize()
        for a, ref in zip(dev, eager_small):
            assert torch.equal(a, ref)


@pytest.mark.parametrize('name,H,W,md,precision', [
    ('C2', 540, 960, 191, 'fp16x2'),
    ('C4', 375, 1242, 191, 'fp16x2'),
    ('C3', 540, 960, 255, 'fp16x2'),
    ('C3-bf16', 540, 960, 255, 'bf16'),
])
def test_full_size_configs_vs_torch_port(name, H, W, md, precision):
    """BASELINE.json's full-size configurations against the ATen composition of the same
    operators on the same device (fp32, TF32 off): margin-aware parity for the fp32-grade
    precision, MAE / 3-pixel error (errors.py:9-74 semantics) for plain bf16 operands."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = synth.make_params(synth.network_specs(), 81)
    net = load_module(PdsNetwork.default(md, precision=precision), params)
    g = torch.Generator().manual_seed(82)
    left = (torch.rand(1, 3, H, W, generator=g) * 255).cuda()
    right = (torch.rand(1, 3, H, W, generator=g) * 255).cuda()
    right[..., :-17] = 0.8 * left[..., 17:] + 0.2 * right[..., :-17]
    ph, pw = -H % 64, -W % 64
    with torch.no_grad():
        disp = net(left, right)
        cost = net.pass_through_network(net._size_adapter.pad(left), net._size_adapter.pad(right))[0]
        _, idx = net._estimator(cost, crop_top=ph, crop_left=pw, return_argmax=True)
        st = torch_port.network_stages(left, right, tdict(params), md)
    assert disp.shape == (1, H, W) and cost.shape == (1, (md + 1) // 2, H + ph, W + pw)
    assert torch.isfinite(disp).all()
    ref = st['disparity']
    mae = float((disp - ref).abs().mean())
    bad3 = float(((disp - ref).abs() > 3).float().mean())
    if precision == 'bf16':
        # single-term operands: not an fp32-grade mode; judged on the reference's own metrics.  With
        # random weights the cost volume is nearly flat along the disparity axis, so every arg-max
        # flip moves the disparity by tens of pixels (the reference's own bf16 run
        # differs from its fp32 run by up to 55 px); measured here: MAE 3.8 px, 3PE 5.1 %
        assert mae < 6.0 and bad3 < 0.08, (mae, bad3)
        return
    cost_err = max_abs(cost, st['cost'])
    assert cost_err <= 1e-3, cost_err
    flips, err, safe = margin_aware(disp.cpu().numpy(), idx.cpu().numpy(), ref.cpu().numpy(),
                                    st['cost'].cpu().numpy(), (ph, pw), cost_err)
    print(f'{name} {precision}: cost max-abs {cost_err:.3e} safe {safe:.4f} flips {flips:.3e} '
          f'disparity max-abs on safe pixels {err:.3e} MAE {mae:.3e} 3PE {bad3:.3e}')
    # north_star: arg-max bit-exact (asserted inside margin_aware on the margin-safe pixels, >= 99 %
    # of the image) and disparity within 1e-3 there.  Flips elsewhere: at most the reference's OWN
    # self-flip rate between a batch-1 and a batch-2 run of the same pair, 1.2e-4 (SURVEY.md 0 / 8c;
    # ATen-fp32 against fp64 flips 1.3e-5, profiles/r02_precision_*.txt)
    assert safe >= 0.99 and flips <= 1.2e-4 and err <= 1e-3, (safe, flips, err)
    assert mae < 0.01 and bad3 < 5e-4, (mae, bad3)
