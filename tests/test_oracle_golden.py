"""Pins the CPU oracle (C restatement and torch port) against the reference:
its own known-answer tests and the golden vectors generated from the unmodified
reference modules by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle, synth, torch_port


def tdict(params):
    return {k: torch.from_numpy(v) for k, v in params.items()}


def close(a, b, atol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))
    assert err <= atol, f'max-abs {err:.3e} > {atol:.1e}'
    return err


# ----------------------------------------------------------------------------
# reference test/test_matching.py:17-32 (known answer, mock operation)
def test_matching_known_answer(golden):
    g = golden('matching_known_answer')
    left = np.array([0, 2, 1, 2], np.float32).reshape(1, 1, 1, 4)
    right = np.array([3, 4, 2, 4], np.float32).reshape(1, 1, 1, 4)
    expected2 = np.array([[3, 4, 2, 4], [0, 3, 4, 2], [0, 2, 3, 4]]).reshape(1, 1, 3, 1, 4)
    expected1 = expected2[:, :, :2]
    assert np.array_equal(g['md2'], expected2) and np.array_equal(g['md1'], expected1)
    for md, exp in ((2, expected2), (1, expected1)):
        vol = oracle.matching_concat(left, right, md)          # (B, D, 2C, H, W)
        out = vol.max(axis=2, keepdims=True).transpose(0, 2, 1, 3, 4)
        assert np.array_equal(out, exp)
        tp = torch_port.matching(torch.from_numpy(left), torch.from_numpy(right),
                                 lambda x: x.max(dim=1, keepdim=True)[0], md)
        assert np.array_equal(tp.numpy(), exp)


def test_matching_concat_identity(golden):
    g = golden('matching_known_answer')
    l, r = synth.tensor((2, 3, 4, 9), 11), synth.tensor((2, 3, 4, 9), 12)
    vol = oracle.matching_concat(l, r, 5)                      # (B, D, 2C, H, W)
    assert np.array_equal(vol.transpose(0, 2, 1, 3, 4), g['identity_md5'])


# reference test/test_estimator.py:14-27
def test_estimator_known_answer(golden):
    sim = np.array([0.1, 0.4, 0.3, 0.2, 0.3], np.float32).reshape(1, 5, 1, 1)
    d1, _ = oracle.subpixel_map(sim, 2, 1)
    d2, _ = oracle.subpixel_map(sim, 2, 2)
    assert np.isclose(d1.item(), 1.52, atol=1e-4)
    assert np.isclose(d2.item(), 2.124, atol=1e-4)
    ka = golden('estimator')['known_answer']
    assert np.allclose([d1.item(), d2.item()], ka, atol=1e-6)
    t1, _ = torch_port.subpixel_map(torch.from_numpy(sim), 2, 1)
    assert np.isclose(t1.item(), 1.52, atol=1e-4)


@pytest.mark.parametrize('hsw,step', [(4, 2), (2, 1), (2, 2), (6, 3), (1, 1)])
def test_estimator_golden(golden, hsw, step):
    g = golden('estimator')
    for name, x in (('random', synth.tensor((2, 12, 9, 11), 21)),
                    ('adversarial', g['adversarial_in'])):
        d, idx = oracle.subpixel_map(x, hsw, step)
        assert np.array_equal(idx, g[f'{name}_argmax'])       # bit-exact indices
        close(d, g[f'{name}_{hsw}_{step}'], 2e-6 * step * x.shape[1])
        td, tidx = torch_port.subpixel_map(torch.from_numpy(x), hsw, step)
        assert np.array_equal(tidx.numpy(), g[f'{name}_argmax'])
        close(td.numpy(), g[f'{name}_{hsw}_{step}'], 2e-6 * step * x.shape[1])


def test_estimator_nan_rule():
    # th.max: a NaN wins and the first NaN's index is returned
    x = synth.tensor((1, 6, 2, 3), 5)
    x[0, 4, 0, 0] = np.nan
    x[0, 2, 0, 0] = np.nan
    d, idx = oracle.subpixel_map(x)
    ref_idx = torch.max(torch.from_numpy(x), dim=1)[1].numpy()
    assert np.array_equal(idx, ref_idx) and idx[0, 0, 0] == 2
    assert np.isnan(d[0, 0, 0]) and not np.isnan(d[0, 1, 1])


def test_matching_operation_golden(golden):
    params = synth.make_params(synth.matching_operation_specs(), 31)
    x = synth.tensor((2, 128, 12, 14), 32)
    ref = golden('matching_operation')['out']
    close(oracle.matching_operation(x, synth.flatten(params)), ref, 1e-4)
    with torch.no_grad():
        close(torch_port.matching_operation(torch.from_numpy(x), tdict(params)).numpy(),
              ref, 1e-5)


def test_matching_golden(golden):
    params = synth.make_params(synth.matching_operation_specs(), 31)
    l, r = synth.tensor((1, 64, 10, 24), 33), synth.tensor((1, 64, 10, 24), 34)
    ref = golden('matching')['out']
    close(oracle.matching(l, r, synth.flatten(params), 7), ref, 1e-4)
    p = tdict(params)
    with torch.no_grad():
        out = torch_port.matching(torch.from_numpy(l), torch.from_numpy(r),
                                  lambda x: torch_port.matching_operation(x, p), 7)
    close(out.numpy(), ref, 1e-5)


def test_regularization_blocks_golden(golden):
    g = golden('regularization_blocks')
    pc = synth.make_params(synth.contraction_block_specs(6), 41)
    pe = synth.make_params(synth.expansion_block_specs(6), 43)
    x = synth.tensor((2, 6, 10, 14, 16), 42)
    skip = synth.tensor((2, 3, 20, 28, 32), 44)
    down, smooth = oracle.contraction_block(x, synth.flatten(pc))
    assert down.shape == (2, 12, 5, 7, 8)      # test/test_regularization.py:13-19
    close(down, g['down'], 1e-4)
    close(smooth, g['smooth'], 1e-4)
    out = oracle.expansion_block(x, skip, synth.flatten(pe))
    assert out.shape == (2, 3, 20, 28, 32)     # test/test_regularization.py:22-28
    close(out, g['expansion'], 1e-4)
    with torch.no_grad():
        td, ts = torch_port.contraction_block(torch.from_numpy(x), tdict(pc), '')
        close(td.numpy(), g['down'], 1e-5)
        close(ts.numpy(), g['smooth'], 1e-5)
        te = torch_port.expansion_block(torch.from_numpy(x), torch.from_numpy(skip),
                                        tdict(pe), '')
        close(te.numpy(), g['expansion'], 1e-5)


def test_regularization_golden(golden):
    params = synth.make_params(synth.regularization_specs(), 45)
    sig, sc = synth.tensor((1, 8, 16, 16, 32), 46), synth.tensor((1, 8, 16, 32), 47)
    ref = golden('regularization')['out']
    assert ref.shape == (1, 32, 64, 128)
    # fp32 noise floor here: reference-vs-fp64 is 1.8e-4 on outputs of scale 17.7
    close(oracle.regularization(sig, sc, synth.flatten(params)), ref, 1e-3)
    with torch.no_grad():
        out = torch_port.regularization(torch.from_numpy(sig), torch.from_numpy(sc),
                                        tdict(params))
    close(out.numpy(), ref, 1e-5)


def test_embedding_golden(golden):
    g = golden('embedding')
    params = synth.make_params(synth.embedding_specs(), 51)
    img = synth.tensor((1, 3, 64, 128), 52, scale=255.0, uniform=True)
    d, s = oracle.embedding(img, synth.flatten(params))
    close(d, g['descriptor'], 2e-4)
    close(s, g['shortcut'], 2e-4)
    with torch.no_grad():
        td, ts = torch_port.embedding(torch.from_numpy(img), tdict(params))
    close(td.numpy(), g['descriptor'], 1e-5)
    close(ts.numpy(), g['shortcut'], 1e-5)


def _network_inputs():
    li = synth.tensor((1, 3, 62, 100), 62, scale=255.0, uniform=True)
    ri = synth.tensor((1, 3, 62, 100), 63, scale=255.0, uniform=True)
    ri[..., :-6] = 0.8 * li[..., 6:] + 0.2 * ri[..., :-6]
    return li, ri


def test_network_golden(golden):
    g = golden('network_md63')
    params = synth.make_params(synth.network_specs(), 61)
    li, ri = _network_inputs()
    with torch.no_grad():
        st = torch_port.network_stages(torch.from_numpy(li), torch.from_numpy(ri),
                                       tdict(params), 63)
    close(st['signatures'].numpy(), g['signatures'], 1e-5)
    close(st['cost'].numpy(), g['cost_padded'], 1e-5)
    close(st['disparity'].numpy(), g['disparity'], 1e-3)
    assert np.array_equal(g['cost_padded'][..., 2:, 28:], g['cost_unpadded'])

    pe = synth.flatten({k: v for k, v in params.items() if k.startswith('_embedding.')})
    pm = synth.flatten({k: v for k, v in params.items() if k.startswith('_matching.')})
    pr = synth.flatten({k: v for k, v in params.items() if k.startswith('_regularization.')})
    disp, cost = oracle.network_forward(li, ri, pe, pm, pr, 63, return_cost=True)
    # fp32 noise floor of this config (1x1x2 bottleneck InstanceNorm): the reference
    # itself is 1.3e-3 away from its own fp64 run on a cost volume of scale 19
    err = close(cost, g['cost_padded'], 3e-3)
    # disparity: exact where the oracle's arg-max agrees (margin-aware)
    _, idx_ref = oracle.subpixel_map(g['cost_padded'])
    _, idx = oracle.subpixel_map(cost)
    agree = (idx == idx_ref)[..., 2:, 28:]
    assert agree.mean() > 0.995
    assert np.max(np.abs(disp - g['disparity'])[agree]) < 1e-2 + 100 * err
    with pytest.raises(ValueError):
        oracle.network_forward(li, ri, pe, pm, pr, 100)


def md127_inputs():
    li = synth.tensor((1, 3, 120, 250), 64, scale=255.0, uniform=True)
    ri = synth.tensor((1, 3, 120, 250), 65, scale=255.0, uniform=True)
    ri[..., :-9] = 0.8 * li[..., 9:] + 0.2 * ri[..., :-9]
    return li, ri


def margin_checks(g, cost_padded, disparity, argmax, tol_cost, crop=(8, 6)):
    """Parity against the well-conditioned fixture of the unmodified reference (network_md127):
    cost volume on the stored 4x sub-sampled grid, arg-max BIT-EXACT and disparity within 1e-3 on
    every pixel whose reference top-1/top-2 margin exceeds 4x the measured cost error.  Returns
    (cost error, safe fraction, flip fraction, disparity max-abs on safe pixels)."""
    cost_err = float(np.max(np.abs(cost_padded[..., ::4, ::4].astype(np.float64) - g['cost_padded_sub4'])))
    assert cost_err <= tol_cost, cost_err
    safe = g['margin'] > 4 * cost_err
    assert np.array_equal(argmax[safe], g['argmax'][safe].astype(argmax.dtype))
    flips = float((argmax != g['argmax']).mean())
    err = float(np.abs(disparity - g['disparity'])[safe].max())
    return cost_err, float(safe.mean()), flips, err


def test_network_golden_well_conditioned(golden):
    """250x120, md=127 (hourglass bottleneck 2x2x4): both oracles against the unmodified reference at
    north_star's own bound -- arg-max bit-exact and disparity <= 1e-3 on margin-safe pixels."""
    g = golden('network_md127')
    params = synth.make_params(synth.network_specs(), 61)
    li, ri = md127_inputs()
    with torch.no_grad():
        st = torch_port.network_stages(torch.from_numpy(li), torch.from_numpy(ri), tdict(params), 127)
    cost_err, safe, flips, err = margin_checks(
        g, st['cost'].numpy(), st['disparity'].numpy(), st['argmax'][..., 8:, 6:].numpy(), 1e-5)
    assert safe > 0.999 and flips == 0.0 and err <= 1e-4, (cost_err, safe, flips, err)

    sub = lambda p: synth.flatten({k: v for k, v in params.items() if k.startswith(p)})
    disp, cost = oracle.network_forward(li, ri, sub('_embedding.'), sub('_matching.'),
                                        sub('_regularization.'), 127, return_cost=True)
    _, idx = oracle.subpixel_map(cost)
    cost_err, safe, flips, err = margin_checks(g, cost, disp, idx[..., 8:, 6:], 2e-4)
    assert safe > 0.99 and flips < 1e-3 and err <= 1e-3, (cost_err, safe, flips, err)


def _loss_case(g, use_weights, module):
    sim = torch.from_numpy(synth.tensor((2, 12, 9, 11), 71)).requires_grad_(True)
    gt = torch.from_numpy(g['ground_truth'])
    w = torch.from_numpy(g['weights'].copy()).requires_grad_(True) if use_weights else None
    value = module(sim, gt, w)
    value.backward()
    return value, sim.grad, (w.grad if use_weights else None)


@pytest.mark.parametrize('use_weights', [True, False])
def test_subpixel_cross_entropy_golden(golden, use_weights):
    """loss.SubpixelCrossEntropy (loss.py:16-78): the oracle's loop and the package's CPU tensor
    expression against value and gradients of the unmodified reference; plus the reference's own
    known-answer test (test/test_loss.py:13-38)."""
    from practicaldeepstereo_nips2018_b200 import loss as pds_loss
    g = golden('loss')
    tag = 'weighted' if use_weights else 'mean'
    for module in (lambda s, t, w: torch_port.subpixel_cross_entropy(s, t, w, 1.5, 2),
                   pds_loss.SubpixelCrossEntropy(diversity=1.5, disparity_step=2)):
        value, grad, grad_w = _loss_case(g, use_weights, module)
        assert abs(value.item() - float(g[f'{tag}_loss'])) <= 1e-6
        close(grad.numpy(), g[f'{tag}_grad_similarities'], 1e-7)
        if use_weights:
            close(grad_w.numpy(), g['weighted_grad_weights'], 1e-7)
    sim = torch.tensor([[0.1, 0.3, 0.2, 0.05], [0.2, 0.1, 0.4, 0.0], [0.2, 0.1, 0.4, 0.0]]).t().reshape(1, 4, 3, 1)
    sim = sim.clone().requires_grad_(True)
    gt = torch.tensor([1.3, float('inf'), 1.9]).view(1, 3, 1)
    w = torch.tensor([0.9, 0.0, 0.01]).view(1, 3, 1)
    value = pds_loss.SubpixelCrossEntropy(diversity=2.0, disparity_step=1)(sim, gt, w)
    value.backward()
    expected = torch.tensor([[0.0262, -0.0567, -0.0219, 0.0524], [0.0, 0.0, 0.0, 0.0],
                             [0.0011, -0.0002, -0.0007, -0.0002]]).t().reshape(1, 4, 3, 1)
    assert abs(value.item() - 1.3654) <= 1e-3 and torch.allclose(sim.grad, expected, atol=1e-3)


# ----------------------------------------------------------------------------
# f4: gradients of Matching + MatchingOperation in training mode (reference matching.py:34-63 under
# autograd) -- the torch port and the package's CPU composition against the unmodified reference
def _matching_grad_case():
    params = synth.make_params(synth.matching_operation_specs(), 81)
    left = torch.from_numpy(synth.tensor((2, 64, 6, 13), 82)).requires_grad_(True)
    right = torch.from_numpy(synth.tensor((2, 64, 6, 13), 83)).requires_grad_(True)
    probe = torch.from_numpy(synth.tensor((2, 8, 5, 6, 13), 84))
    return params, left, right, probe


def test_torch_port_matching_gradients(golden):
    g = golden('matching_grad')
    params, left, right, probe = _matching_grad_case()
    p = {k: v.clone().requires_grad_(True) for k, v in tdict(params).items()}
    sig = torch_port.matching(left, right, lambda x: torch_port.matching_operation(x, p, 2), 4)
    (sig * probe).sum().backward()
    close(sig.detach().numpy(), g['signatures'], 2e-5)
    scale = float(np.abs(g['grad_left']).max())
    close(left.grad.numpy(), g['grad_left'], 2e-5 * scale)
    close(right.grad.numpy(), g['grad_right'], 2e-5 * scale)
    for k, v in p.items():
        ref = g['grad_param_' + k.replace('.', '__')]
        close(v.grad.numpy(), ref, 1e-4 * max(1.0, float(np.abs(ref).max())))


def test_package_matching_gradients_on_cpu(golden):
    """CPU tensors with gradients: the package runs the reference's per-disparity composition."""
    from practicaldeepstereo_nips2018_b200 import matching
    g = golden('matching_grad')
    params, left, right, probe = _matching_grad_case()
    op = matching.MatchingOperation(precision='fp32')
    op.load_state_dict(tdict(params))
    op.train()
    sig = matching.Matching(4, op)(left, right)
    (sig * probe).sum().backward()
    close(sig.detach().numpy(), g['signatures'], 2e-5)
    scale = float(np.abs(g['grad_left']).max())
    close(left.grad.numpy(), g['grad_left'], 2e-5 * scale)
    close(right.grad.numpy(), g['grad_right'], 2e-5 * scale)
    for k, v in op.named_parameters():
        ref = g['grad_param_' + k.replace('.', '__')]
        close(v.grad.numpy(), ref, 1e-4 * max(1.0, float(np.abs(ref).max())))


def test_oracle_adjoints_against_autograd():
    """numpy restatements of the f4 adjoints (oracle.py) against float64 autograd through the torch
    port's own forward operators: the volume adjoint for more disparities than columns, and
    LeakyReLU + InstanceNorm in 2-D and 3-D."""
    import torch.nn.functional as F
    rng = np.random.RandomState(7)
    for B, C, H, W, D in ((2, 3, 4, 9, 6), (1, 2, 3, 5, 9)):
        left = torch.from_numpy(rng.randn(B, C, H, W)).requires_grad_(True)
        right = torch.from_numpy(rng.randn(B, C, H, W)).requires_grad_(True)
        probe = rng.randn(B, 2 * C, D, H, W)
        out = torch_port.matching(left, right, lambda x: x, D - 1)            # (B, 2C, D, H, W)
        (out * torch.from_numpy(probe)).sum().backward()
        gl, gr = oracle.matching_concat_backward(probe.transpose(0, 2, 1, 3, 4))
        close(gl, left.grad.numpy().astype(np.float32), 1e-5)
        close(gr, right.grad.numpy().astype(np.float32), 1e-5)
    for shape in ((2, 4, 5, 7), (1, 3, 4, 5, 6)):
        x = torch.from_numpy(rng.randn(*shape) * 2 + 0.3).requires_grad_(True)
        gamma = torch.from_numpy(rng.rand(shape[1]) + 0.5).requires_grad_(True)
        beta = torch.from_numpy(rng.randn(shape[1])).requires_grad_(True)
        probe = rng.randn(*shape)
        y = F.instance_norm(F.leaky_relu(x, 0.1), weight=gamma, bias=beta, eps=1e-5)
        (y * torch.from_numpy(probe)).sum().backward()
        y_o, _ = oracle.leaky_instance_norm(x.detach().numpy(), gamma.detach().numpy(), beta.detach().numpy())
        dx, dg, db = oracle.leaky_instance_norm_backward(x.detach().numpy(), probe, gamma.detach().numpy())
        close(y_o, y.detach().numpy(), 1e-9)
        close(dx, x.grad.numpy(), 1e-9)
        close(dg, gamma.grad.numpy(), 1e-9)
        close(db, beta.grad.numpy(), 1e-9)
