"""a2 parity of the tcgen05 path: MatchingOperation over all disparities on the
tensor cores (split 16-bit operands) vs the ATen restatement (B200 only).

Tolerances (signature scale ~10): the fp32-grade modes (fp16x2: 22 significand
bits, bf16x3: 24) must stay within 2e-4 max-abs -- the bound the fp32 CUDA-core
path is held to in test_gpu_matching.py; bf16x2 (16 bits) 1e-3; single-term
modes are only checked for gross errors."""
import pytest
import torch

from oracle import synth, torch_port
from practicaldeepstereo_nips2018_b200 import matching
from gpu_util import cuda, load_module, max_abs, tdict

pytestmark = pytest.mark.gpu

TOL = {'fp16x2': 2e-4, 'bf16x3': 2e-4, 'bf16x2': 1e-3, 'bf16': 0.5, 'fp16': 0.1}


def _reference(l, r, params, md, n_res=2):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    p = tdict(params)
    return torch_port.matching(l, r, lambda x: torch_port.matching_operation(x, p, n_res), md)


@pytest.mark.parametrize('precision', sorted(TOL))
@pytest.mark.parametrize('B,H,W,md', [(1, 16, 16, 0), (2, 21, 37, 9), (1, 48, 80, 23)])
def test_tc_matching_vs_torch_port(precision, B, H, W, md):
    params = synth.make_params(synth.matching_operation_specs(), 35)
    op = load_module(matching.MatchingOperation(precision=precision), params)
    l, r = cuda(synth.tensor((B, 64, H, W), 36)), cuda(synth.tensor((B, 64, H, W), 37))
    with torch.no_grad():
        out = matching.Matching(md, op)(l, r)
        ref = _reference(l, r, params, md)
    assert out.shape == ref.shape == (B, 8, md + 1, H, W)
    assert max_abs(out, ref) <= TOL[precision]


@pytest.mark.parametrize('precision', ['fp16x2', 'bf16x3'])
def test_tc_disparity_beyond_width_and_groups(precision, monkeypatch):
    """More disparities than columns (fully shifted-out right image) and a slice
    count that is not a multiple of the L2 group size."""
    monkeypatch.setenv('PDS_B200_MATCH_GROUP', '3')
    params = synth.make_params(synth.matching_operation_specs(), 38)
    op = load_module(matching.MatchingOperation(precision=precision), params)
    l, r = cuda(synth.tensor((2, 64, 18, 11), 39)), cuda(synth.tensor((2, 64, 18, 11), 40))
    with torch.no_grad():
        out = matching.Matching(12, op)(l, r)            # 26 slices, groups of 3
        ref = _reference(l, r, params, 12)
    assert max_abs(out, ref) <= TOL[precision]


def test_tc_full_width_slice():
    """One C2-sized disparity group (144 x 240, 3 disparities): every tile column,
    the weight-resident kernel and the InstanceNorm sums at full size."""
    params = synth.make_params(synth.matching_operation_specs(), 41)
    op = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    l, r = cuda(synth.tensor((1, 64, 144, 240), 42)), cuda(synth.tensor((1, 64, 144, 240), 43))
    with torch.no_grad():
        out = matching.Matching(2, op)(l, r)
        ref = _reference(l, r, params, 2)
    assert max_abs(out, ref) <= 2e-4


@pytest.mark.parametrize('W', [600, 1100])
def test_tc_wide_images_use_fewer_rows_per_cta(W):
    """Quarter-resolution widths beyond what four staged rows of the per-sample terms fit in
    shared memory (csrc/matching_factor.cu picks 2 rows, then 1): few rows, 3 disparities,
    ragged height (rows per CTA do not divide it)."""
    params = synth.make_params(synth.matching_operation_specs(), 44)
    op = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    l, r = cuda(synth.tensor((1, 64, 7, W), 45)), cuda(synth.tensor((1, 64, 7, W), 46))
    with torch.no_grad():
        out = matching.Matching(2, op)(l, r)
        ref = _reference(l, r, params, 2)
    assert max_abs(out, ref) <= 2e-4


@pytest.mark.parametrize('B,H,W,md', [(1, 48, 80, 23), (2, 21, 37, 9), (1, 144, 240, 7), (1, 18, 11, 15)])
def test_dynamic_and_fused_kernels_match_the_round1_pipeline(B, H, W, md, monkeypatch):
    """Round 2 kernels of the 64->64 layers against the round-1 pipeline (statically scheduled
    convolution, one-pass tc_compose_second, separate normalisation passes):
      default  -- tc_compose_second as sums + normalised planes (no fp32 round trip); last convolution
                  (64 -> 8) with the taps on the M axis of the MMA tile (csrc/conv_last.cu: the nine taps
                  are summed in fp32 in the epilogue, so this one differs from round 1 by rounding);
      dynamic  -- PDS_B200_DYNAMIC_CONV=1: dynamically scheduled convolution (tiles claimed in slice
                  order, eight epilogue warps, one private copy of the InstanceNorm sums per epilogue
                  warp), separate normalisation passes;
      fused    -- PDS_B200_FUSE_NORM=1: the InstanceNorm (+ residual) passes run INSIDE the dynamic
                  convolution launches (trailing normalisation warps behind per-slice counters).
    The last two are built and parity-green but not faster than the static kernel + separate passes
    at 960x540 D=192, hence opt-in.
    Same arithmetic everywhere; only the order of the double-precision sum atomics differs."""
    params = synth.make_params(synth.matching_operation_specs(), 47)
    l, r = cuda(synth.tensor((B, 64, H, W), 48)), cuda(synth.tensor((B, 64, H, W), 49))
    default = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    with torch.no_grad():
        out = matching.Matching(md, default)(l, r)
        again = matching.Matching(md, default)(l, r)
        ref = _reference(l, r, params, md)
    monkeypatch.setenv('PDS_B200_DYNAMIC_CONV', '1')     # switches are read at handle creation
    dynamic = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    with torch.no_grad():
        dyn = matching.Matching(md, dynamic)(l, r)
    monkeypatch.setenv('PDS_B200_FUSE_NORM', '1')
    fused = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    with torch.no_grad():
        fus = matching.Matching(md, fused)(l, r)
    monkeypatch.setenv('PDS_B200_FUSE_NORM', '0')
    monkeypatch.setenv('PDS_B200_DYNAMIC_CONV', '0')
    monkeypatch.setenv('PDS_B200_COMPOSE_TWO_PASS', '0')
    monkeypatch.setenv('PDS_B200_LAST_TAPS', '0')
    round1 = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    with torch.no_grad():
        sep = matching.Matching(md, round1)(l, r)
    scale = float(ref.abs().max())
    print(f'default vs round-1: max-abs {max_abs(out, sep):.3e} (run-to-run {max_abs(out, again):.3e}), dynamic vs '
          f'round-1 {max_abs(dyn, sep):.3e}, fused vs round-1 {max_abs(fus, sep):.3e}; vs ATen: default '
          f'{max_abs(out, ref):.3e}, fused {max_abs(fus, ref):.3e}, round-1 {max_abs(sep, ref):.3e}, scale {scale:.2f}')
    assert max_abs(out, ref) <= 2e-4 and max_abs(fus, ref) <= 2e-4 and max_abs(dyn, ref) <= 2e-4
    assert max_abs(out, sep) <= 1e-5 * scale and max_abs(out, again) <= 2e-6 * scale
    assert max_abs(fus, out) <= 2e-6 * scale and max_abs(dyn, out) <= 2e-6 * scale


@pytest.mark.parametrize('fuse', ['0', '1'])
def test_dynamic_kernels_on_concurrent_streams(fuse, monkeypatch):
    """Dynamically scheduled (and, opt-in, fused) launches on several streams at once (HostPipeline's
    serving pattern): tiles are CLAIMED, never assigned to a CTA that may not be resident, and a
    normalisation warp only waits for tiles, so partially resident grids cannot wait on each other.  Eight forwards on four streams of a
    96-slice problem; every result must equal the single-stream one."""
    params = synth.make_params(synth.matching_operation_specs(), 50)
    monkeypatch.setenv('PDS_B200_DYNAMIC_CONV', '1')
    monkeypatch.setenv('PDS_B200_FUSE_NORM', fuse)
    op = load_module(matching.MatchingOperation(precision='fp16x2'), params)
    pairs = [(cuda(synth.tensor((2, 64, 64, 96), 51 + i)), cuda(synth.tensor((2, 64, 64, 96), 61 + i))) for i in range(4)]
    m = matching.Matching(47, op)
    with torch.no_grad():
        expected = [m(l, r) for l, r in pairs]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(4)]
        results = []
        for k in range(8):
            with torch.cuda.stream(streams[k % 4]):
                results.append(m(*pairs[k % 4]))
        torch.cuda.synchronize()
    for k, out in enumerate(results):
        assert max_abs(out, expected[k % 4]) <= 1e-5
