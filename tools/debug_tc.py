"""Step-by-step numerical check of the tcgen05 matching path against ATen."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, torch_port
from practicaldeepstereo_nips2018_b200 import matching

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'


def run(tag, n_res, B, H, W, D, precision, seed=0, special=None):
    specs = synth.matching_operation_specs(n_res=n_res)
    params = synth.make_params(specs, 31 + seed)
    if special:
        special(params)
    op = matching.MatchingOperation(number_of_residual_blocks=n_res, precision=precision)
    op.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    op = op.to(dev).eval()
    l = torch.from_numpy(synth.tensor((B, 64, H, W), 36 + seed)).to(dev)
    r = torch.from_numpy(synth.tensor((B, 64, H, W), 37 + seed)).to(dev)
    p = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    try:
        with torch.no_grad():
            t0 = time.time()
            out = op.match_all_disparities(l, r, D)
            torch.cuda.synchronize()
            dt = time.time() - t0
            ref = torch_port.matching(l, r, lambda x: torch_port.matching_operation(x, p, n_res), D - 1)
        err = (out - ref).abs()
        print(f'{tag:34s} {precision:7s} n_res={n_res} B={B} {H}x{W} D={D}: max-abs {err.max().item():.3e} '
              f'(ref scale {ref.abs().max().item():.2f}) mean-abs {err.mean().item():.2e}  [{dt*1e3:.1f} ms]', flush=True)
        if err.max().item() > 0.5 * ref.abs().max().item():
            bad = (err > 0.1 * ref.abs().max()).nonzero()
            print('   first bad indices (b, c, d, y, x):', bad[:6].tolist(), 'count', bad.shape[0])
        return err.max().item()
    except Exception as e:  # noqa
        print(f'{tag}: EXCEPTION {type(e).__name__}: {e}', flush=True)
        return None


def center_only(params):
    # conv0 = identity-ish on the centre tap for the first 8 channels; conv_last = centre tap pick
    for k, v in params.items():
        if k.endswith('.weight') and v.ndim == 4:
            v[:] = 0
    w0 = params['_matching_operation_modules.0.weight']
    for c in range(64):
        w0[c, c, 1, 1] = 1.0
    wl = params['_matching_operation_modules.1.weight']
    for c in range(8):
        wl[c, c, 1, 1] = 1.0


def right_center(params):
    center_only(params)
    w0 = params['_matching_operation_modules.0.weight']
    w0[:] = 0
    for c in range(64):
        w0[c, 64 + c, 1, 1] = 1.0


def shifted_tap(params):
    center_only(params)
    w0 = params['_matching_operation_modules.0.weight']
    w0[:] = 0
    for c in range(64):
        w0[c, c, 0, 2] = 1.0      # reads (y-1, x+1)


prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
run('identity centre tap (left)', 0, 1, 16, 8, 1, prec, special=center_only)
run('identity centre tap (right, D=3)', 0, 1, 16, 8, 3, prec, special=right_center)
run('single off-centre tap', 0, 1, 16, 8, 1, prec, special=shifted_tap)
run('random conv0+last, one tile', 0, 1, 16, 8, 1, prec)
run('random conv0+last, NT=3 ragged', 0, 2, 21, 37, 5, prec)
run('full op, one tile', 2, 1, 16, 8, 1, prec)
run('full op, 48x80 D=6', 2, 1, 48, 80, 6, prec)
run('full op, B=2 33x50 D=7 (ragged)', 2, 2, 33, 50, 7, prec)
for pr in ('fp16x2', 'bf16x3', 'bf16x2', 'bf16', 'fp16'):
    run('full op, 144x240 D=4', 2, 1, 144, 240, 4, pr)
