"""a6 parity: the embedding tower on the tcgen05 engine (csrc/embedding.cu) against the golden
vectors of the unmodified reference and the torch port in fp64 (B200 only), and the whole
network in the fp32-grade tensor-core precision."""
import numpy as np
import pytest
import torch

from oracle import synth, torch_port
from practicaldeepstereo_nips2018_b200 import PdsNetwork, embedding
from gpu_util import cuda, load_module, max_abs, tdict
from test_gpu_network import margin_aware

pytestmark = pytest.mark.gpu


def _f64(params):
    return {k: v.double() for k, v in tdict(params).items()}


@pytest.mark.parametrize('precision,tol', [('fp16x2', 2e-4), ('bf16x3', 2e-4)])
def test_embedding_golden(golden, precision, tol):
    g = golden('embedding')
    params = synth.make_params(synth.embedding_specs(), 51)
    emb = load_module(embedding.Embedding(precision=precision), params)
    img = cuda(synth.tensor((1, 3, 64, 128), 52, scale=255.0, uniform=True))
    with torch.no_grad():
        assert emb.uses_kernels(img)
        d, s = emb(img)
    assert d.shape == (1, 64, 16, 32) and s.shape == (1, 8, 16, 32)
    assert max_abs(d, g['descriptor']) <= tol * float(np.abs(g['descriptor']).max())
    assert max_abs(s, g['shortcut']) <= tol * float(np.abs(g['shortcut']).max())


def test_embedding_vs_torch_port_ragged_batch():
    """Quarter-resolution extent 25 x 33 (partial tiles everywhere), batch 3, shortcut for 2."""
    params = synth.make_params(synth.embedding_specs(), 53)
    emb = load_module(embedding.Embedding(precision='fp16x2'), params)
    img = cuda(synth.tensor((3, 3, 100, 132), 54, scale=255.0, uniform=True))
    with torch.no_grad():
        d, s = emb.embed(img, 2)
        rd, rs = torch_port.embedding(img.double(), _f64(params))
        d_all, s_all = emb(img)                      # module call: shortcut for every sample
        d_no, s_no = emb(img, with_shortcut=False)
    assert d.shape == (3, 64, 25, 33) and s.shape == (2, 8, 25, 33) and s_no is None
    assert max_abs(d, rd) <= 2e-4 * float(rd.abs().max())
    assert max_abs(s, rs[:2]) <= 2e-4 * float(rs.abs().max())
    assert torch.equal(d_all, d) and torch.equal(d_no, d) and torch.equal(s_all[:2], s)
    assert max_abs(s_all, rs) <= 2e-4 * float(rs.abs().max())
    # extents that are not multiples of 4 (never produced by PdsNetwork: SizeAdapter pads to
    # multiples of 64): the ATen composition, announced with a RuntimeWarning -- not silently
    embedding.Embedding._warned_aten = False
    with torch.no_grad(), pytest.warns(RuntimeWarning, match='not a multiple of 4'):
        d5, _ = emb(img[..., :98, :130])
    with torch.no_grad():
        r5, _ = torch_port.embedding(img[..., :98, :130], tdict(params))
    assert not emb.uses_kernels(img[..., :98, :130]) and max_abs(d5, r5) <= 1e-3


def test_network_fp16x2_vs_torch_port():
    """PdsNetwork.forward with every stage on the tensor-core kernels (the bench precision):
    margin-aware parity against the fp32 torch port on the same device."""
    torch.backends.cudnn.allow_tf32 = False
    params = synth.make_params(synth.network_specs(), 71)
    net = load_module(PdsNetwork.default(63, precision='fp16x2'), params)
    left = cuda(synth.tensor((2, 3, 100, 310), 72, scale=255.0, uniform=True))
    right = cuda(synth.tensor((2, 3, 100, 310), 73, scale=255.0, uniform=True))
    right[..., :-9] = 0.7 * left[..., 9:] + 0.3 * right[..., :-9]
    with torch.no_grad():
        disp = net(left, right)
        pl, pr = net._size_adapter.pad(left), net._size_adapter.pad(right)
        ld, rd, sc = net._embed(pl, pr)
        cost = net.pass_through_network(pl, pr)[0]
        d2, idx = net._estimator(cost, crop_top=28, crop_left=10, return_argmax=True)
        st = torch_port.network_stages(left.double(), right.double(), _f64(params), 63)
    assert disp.shape == (2, 100, 310)
    assert torch.equal(disp, d2)     # forward() fuses hourglass tail + estimator + crop: bit-identical
    assert max_abs(ld, st['left_descriptor']) <= 2e-4 * float(st['left_descriptor'].abs().max())
    assert max_abs(rd, st['right_descriptor']) <= 2e-4 * float(st['right_descriptor'].abs().max())
    assert max_abs(sc, st['shortcut']) <= 2e-4 * float(st['shortcut'].abs().max())
    cost_err = max_abs(cost, st['cost'])
    assert cost_err <= 1e-3
    flips, err, safe = margin_aware(disp.cpu().numpy(), idx.cpu().numpy(),
                                    st['disparity'].float().cpu().numpy(),
                                    st['cost'].float().cpu().numpy(), (28, 10), cost_err)
    print(f'fp16x2 vs fp64 port: cost max-abs {cost_err:.3e} safe {safe:.4f} flips {flips:.3e} disparity {err:.3e}')
    assert safe >= 0.99 and flips <= 1e-3 and err <= 1e-3, (cost_err, safe, flips, err)


def test_input_path_pad_and_uint8_are_bit_identical():
    """f3: pds_embedding_forward_images (un-padded images, SizeAdapter.pad + the first
    InstanceNorm2d fused, left / right as two pointers, float32 or uint8 planar or uint8
    interleaved as cv2 decodes them, dataset.py:67-72) == pad + cat + embed, bit for bit."""
    params = synth.make_params(synth.embedding_specs(), 55)
    emb = load_module(embedding.Embedding(precision='fp16x2'), params)
    g = torch.Generator().manual_seed(56)
    left8 = torch.randint(0, 256, (2, 3, 62, 100), dtype=torch.uint8, generator=g).cuda()
    right8 = torch.randint(0, 256, (2, 3, 62, 100), dtype=torch.uint8, generator=g).cuda()
    pad_top, pad_left = 2, 28                                    # SizeAdapter: 62 x 100 -> 64 x 128
    with torch.no_grad():
        padded = torch.nn.functional.pad(torch.cat([left8, right8]).float(), (pad_left, 0, pad_top, 0))
        d_ref, s_ref = emb.embed(padded, 2)
        for l, r in ((left8.float(), right8.float()), (left8, right8),
                     (left8.permute(0, 2, 3, 1).contiguous(), right8.permute(0, 2, 3, 1).contiguous())):
            assert emb.can_embed_images(l, r)
            dl, dr, s = emb.embed_images(l, r, pad_top, pad_left)
            assert dl.shape == (2, 64, 16, 32) and s.shape == (2, 8, 16, 32)
            assert torch.equal(torch.cat([dl, dr]), d_ref) and torch.equal(s, s_ref)
        # float images whose plane size is not a multiple of 4 (scalar statistics path)
        lo, ro = left8[..., :61, :99].float().contiguous(), right8[..., :61, :99].float().contiguous()
        padded = torch.nn.functional.pad(torch.cat([lo, ro]), (29, 0, 3, 0))
        d_ref, s_ref = emb.embed(padded, 2)
        dl, dr, s = emb.embed_images(lo, ro, 3, 29)
        assert torch.equal(torch.cat([dl, dr]), d_ref) and torch.equal(s, s_ref)
    with pytest.raises(ValueError):
        emb.embed_images(left8, right8, 1, 0)                    # padded extent not a multiple of 4
    assert not emb.can_embed_images(left8, right8[:1])
    assert not emb.can_embed_images(left8.double(), right8.double())


def test_network_accepts_uint8_images():
    """PdsNetwork.forward on uint8 images (planar and interleaved) == on their float copies."""
    params = synth.make_params(synth.network_specs(), 57)
    net = load_module(PdsNetwork.default(63, precision='fp16x2'), params)
    g = torch.Generator().manual_seed(58)
    left8 = torch.randint(0, 256, (1, 62, 100, 3), dtype=torch.uint8, generator=g).cuda()
    right8 = torch.randint(0, 256, (1, 62, 100, 3), dtype=torch.uint8, generator=g).cuda()
    left, right = left8.permute(0, 3, 1, 2).float().contiguous(), right8.permute(0, 3, 1, 2).float().contiguous()
    with torch.no_grad():
        ref = net(left, right)
        cost = net.pass_through_network(net._size_adapter.pad(left), net._size_adapter.pad(right))[0]
        ref2 = net._estimator(cost, crop_top=2, crop_left=28)    # the padded-float path of earlier rounds
        assert ref.shape == (1, 62, 100) and torch.equal(ref, ref2)
        assert torch.equal(net(left8, right8), ref)
        assert torch.equal(net(left8.permute(0, 3, 1, 2).contiguous(), right8.permute(0, 3, 1, 2).contiguous()), ref)
    fp32 = load_module(PdsNetwork.default(63, precision='fp32'), params)   # no fused input path: converted
    with torch.no_grad():
        assert torch.equal(fp32(left8, right8), fp32(left, right))
