"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Runs only in the build container (needs /root/reference); the fixtures it
writes are committed, so nothing at test / bench time reads the reference.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Inputs and weights come from oracle/synth.py (frozen numpy RandomState streams)
so the fixtures store outputs only (plus tiny inputs where convenient).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

from practical_deep_stereo import (embedding, estimator, loss, matching, network,  # noqa: E402
                                   regularization)
from oracle import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def load(module, specs, seed):
    params = synth.make_params(specs, seed)
    sd = module.state_dict()
    assert list(sd.keys()) == list(params.keys()), 'state_dict order mismatch'
    for k, v in sd.items():
        assert tuple(v.shape) == params[k].shape, (k, v.shape, params[k].shape)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    return module.eval()


def t(a):
    return torch.from_numpy(a)


ONLY = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None


def save(name, **arrays):
    if ONLY and name != ONLY:
        return
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


# --- SubpixelCrossEntropy (loss.py:16-78): value and gradients on a random volume with unknown
# locations, with and without weights (round 2; needs autograd, hence outside no_grad) -------------
if not ONLY or ONLY == 'loss':
    sim0 = synth.tensor((2, 12, 9, 11), 71)
    gt0 = np.abs(synth.tensor((2, 9, 11), 72)) * 9.0
    gt0[0, 2, 3:7] = np.inf
    gt0[1, 5, :] = np.inf
    w0 = np.abs(synth.tensor((2, 9, 11), 73)) + 0.1
    arrays = dict(ground_truth=gt0, weights=w0)
    for tag, use_w in (('weighted', True), ('mean', False)):
        sim = t(sim0.copy()).requires_grad_(True)
        w = t(w0.copy()).requires_grad_(True) if use_w else None
        value = loss.SubpixelCrossEntropy(diversity=1.5, disparity_step=2)(sim, t(gt0), w)
        value.backward()
        arrays[f'{tag}_loss'] = np.array(value.item(), np.float64)
        arrays[f'{tag}_grad_similarities'] = sim.grad.numpy()
        if use_w:
            arrays['weighted_grad_weights'] = w.grad.numpy()
    save('loss', **arrays)

# --- Matching + MatchingOperation in TRAINING mode (matching.py:34-63 under autograd, f4): the
# gradients the reference's per-disparity loop accumulates, for a fixed linear functional of the
# signatures (round 2) ---------------------------------------------------------------------------
if not ONLY or ONLY == 'matching_grad':
    opg = load(matching.MatchingOperation(), synth.matching_operation_specs(), 81).train()
    lg = t(synth.tensor((2, 64, 6, 13), 82)).requires_grad_(True)
    rg = t(synth.tensor((2, 64, 6, 13), 83)).requires_grad_(True)
    mg = matching.Matching(maximum_disparity=4, operation=opg)
    sig = mg(lg, rg)
    probe = t(synth.tensor(tuple(sig.shape), 84))
    (sig * probe).sum().backward()
    named = dict(opg.named_parameters())
    save('matching_grad', signatures=sig.detach().numpy(), grad_left=lg.grad.numpy(), grad_right=rg.grad.numpy(),
         **{'grad_param_' + k.replace('.', '__'): v.grad.numpy() for k, v in named.items()})

with torch.no_grad():
    # --- Matching with a mock operation (test/test_matching.py:13-32) ---------
    def mock(x):
        return torch.max(x, dim=1, keepdim=True)[0]
    m = matching.Matching(maximum_disparity=2, operation=mock)
    left = torch.Tensor([0, 2, 1, 2]).view(1, 1, 1, 4)
    right = torch.Tensor([3, 4, 2, 4]).view(1, 1, 1, 4)
    out2 = m(left, right).numpy()
    m.set_maximum_disparity(1)
    out1 = m(left, right).numpy()
    # generic shift/concat/stack on a random tensor with an identity operation
    l2, r2 = synth.tensor((2, 3, 4, 9), 11), synth.tensor((2, 3, 4, 9), 12)
    m5 = matching.Matching(maximum_disparity=5, operation=lambda x: x)
    ident = m5(t(l2), t(r2)).numpy()
    save('matching_known_answer', md2=out2, md1=out1, identity_md5=ident)

    # --- SubpixelMap (test/test_estimator.py:14-27 + random / adversarial) ----
    sim = torch.Tensor([0.1, 0.4, 0.3, 0.2, 0.3]).view(1, 5, 1, 1)
    ka = [estimator.SubpixelMap(2, 1)(sim).item(), estimator.SubpixelMap(2, 2)(sim).item()]
    rnd = synth.tensor((2, 12, 9, 11), 21)
    adv = synth.tensor((1, 8, 4, 16), 22)
    adv[0, :, 0, :] = 0.5                      # all equal -> index 0
    adv[0, 0, 1, :] = 9.0                      # peak at low edge
    adv[0, 7, 2, :] = 9.0                      # peak at high edge
    adv[0, 3, 3, :8] = 7.0
    adv[0, 5, 3, :8] = 7.0                     # exact tie -> lowest index
    adv[0, 1, 3, 8:] = 6.0
    adv[0, 2, 3, 8:] = 6.0
    arrays = dict(known_answer=np.array(ka, np.float32), adversarial_in=adv)
    for name, x in (('random', rnd), ('adversarial', adv)):
        for hsw, step in ((4, 2), (2, 1), (2, 2), (6, 3), (1, 1)):
            arrays[f'{name}_{hsw}_{step}'] = estimator.SubpixelMap(hsw, step)(t(x)).numpy()
        arrays[f'{name}_argmax'] = torch.max(t(x), dim=1)[1].numpy()
    save('estimator', **arrays)

    # --- MatchingOperation / Matching (matching.py) -----------------------------
    op = load(matching.MatchingOperation(), synth.matching_operation_specs(), 31)
    x = synth.tensor((2, 128, 12, 14), 32)
    save('matching_operation', out=op(t(x)).numpy())
    l, r = synth.tensor((1, 64, 10, 24), 33), synth.tensor((1, 64, 10, 24), 34)
    mm = matching.Matching(maximum_disparity=7, operation=op)
    save('matching', out=mm(t(l), t(r)).numpy())

    # --- Regularization blocks and hourglass (regularization.py) ---------------
    cb = load(regularization.ContractionBlock3d(6), synth.contraction_block_specs(6), 41)
    xin = synth.tensor((2, 6, 10, 14, 16), 42)
    down, smooth = cb(t(xin))
    eb = load(regularization.ExpansionBlock3d(6), synth.expansion_block_specs(6), 43)
    skip = synth.tensor((2, 3, 20, 28, 32), 44)
    save('regularization_blocks', down=down.numpy(), smooth=smooth.numpy(),
         expansion=eb(t(xin), t(skip)).numpy())
    reg = load(regularization.Regularization(), synth.regularization_specs(), 45)
    sig, sc = synth.tensor((1, 8, 16, 16, 32), 46), synth.tensor((1, 8, 16, 32), 47)
    save('regularization', out=reg(t(sig), t(sc)).numpy())

    # --- Embedding (embedding.py) ------------------------------------------------
    emb = load(embedding.Embedding(), synth.embedding_specs(), 51)
    img = synth.tensor((1, 3, 64, 128), 52, scale=255.0, uniform=True)
    d, s = emb(t(img))
    save('embedding', descriptor=d.numpy(), shortcut=s.numpy())

    # --- PdsNetwork.forward, eval (network.py:45-52), C1-like config -------------
    net = load(network.PdsNetwork.default(63), synth.network_specs(), 61)
    li = synth.tensor((1, 3, 62, 100), 62, scale=255.0, uniform=True)
    ri = synth.tensor((1, 3, 62, 100), 63, scale=255.0, uniform=True)
    # a more stereo-like pair: right = left shifted by 6 px + noise
    ri[..., :-6] = 0.8 * li[..., 6:] + 0.2 * ri[..., :-6]
    disp = net(t(li), t(ri)).numpy()
    net.train()
    cost = net(t(li), t(ri)).numpy()      # train mode returns the (un-padded) cost
    net.eval()
    padded = net.pass_through_network(net._size_adapter.pad(t(li)),
                                      net._size_adapter.pad(t(ri)))[0]
    ld, ls = net._embedding(net._size_adapter.pad(t(li)))
    rd = net._embedding(net._size_adapter.pad(t(ri)))[0]
    sigs = net._matching(ld, rd)
    save('network_md63', disparity=disp, cost_unpadded=cost, cost_padded=padded.numpy(),
         signatures=sigs.numpy(), left_descriptor=ld.numpy(), shortcut=ls.numpy(),
         right_descriptor=rd.numpy())
    # --- PdsNetwork.forward at a WELL-CONDITIONED small size (round 2): 250x120, md=127 ->
    # padded 256x128, hourglass bottleneck 2x2x4 voxels (the md=63 fixture above has a 1x1x2
    # bottleneck whose InstanceNorm amplifies rounding noise 300x).  Stored: the disparity, the
    # arg-max, the top-1/top-2 margin of every pixel (all on the un-padded image) and the padded
    # cost volume sub-sampled 4x in y and x.
    if not ONLY or ONLY == 'network_md127':
        net.set_maximum_disparity(127)
        li = synth.tensor((1, 3, 120, 250), 64, scale=255.0, uniform=True)
        ri = synth.tensor((1, 3, 120, 250), 65, scale=255.0, uniform=True)
        ri[..., :-9] = 0.8 * li[..., 9:] + 0.2 * ri[..., :-9]
        disp = net(t(li), t(ri)).numpy()
        padded = net.pass_through_network(net._size_adapter.pad(t(li)),
                                          net._size_adapter.pad(t(ri)))[0]
        top2 = torch.topk(padded, 2, dim=1)[0]
        save('network_md127', disparity=disp,
             argmax=torch.max(padded, dim=1)[1][..., 8:, 6:].numpy().astype(np.int16),
             margin=(top2[:, 0] - top2[:, 1])[..., 8:, 6:].numpy(),
             cost_padded_sub4=padded[..., ::4, ::4].numpy())
        net.set_maximum_disparity(63)
    try:
        net.set_maximum_disparity(100)
        raised = False
    except ValueError:
        raised = True
    assert raised
print('done')
