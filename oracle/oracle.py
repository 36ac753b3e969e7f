"""ctypes front end of the C oracle (oracle/pds_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of pds_oracle.c.  All functions take
and return contiguous float32 numpy arrays in PyTorch's NCHW / NCDHW layout.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libpds_oracle.so')
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, 'pds_oracle.c')
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    base = ['/usr/bin/gcc', '-O3', '-march=x86-64-v3', '-fPIC', '-std=c11',
            '-shared', '-o', _LIB_PATH, src, '-lm']
    try:
        subprocess.check_call(base[:2] + ['-fopenmp'] + base[2:],
                              stderr=subprocess.DEVNULL)
    except (subprocess.CalledProcessError, FileNotFoundError):
        base[0] = 'gcc' if not os.path.exists(base[0]) else base[0]
        subprocess.check_call(base)  # single-threaded oracle
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.pds_oracle_num_threads.restype = ctypes.c_int
        for name in ('pds_oracle_matching_operation', 'pds_oracle_regularization',
                     'pds_oracle_contraction_block', 'pds_oracle_expansion_block',
                     'pds_oracle_embedding'):
            getattr(_lib, name).restype = ctypes.c_long
    return _lib


def num_threads():
    return lib().pds_oracle_num_threads()


def set_num_threads(n):
    lib().pds_oracle_set_num_threads(int(n))


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f)


def conv3d(x, w, b, stride=(1, 1, 1), padding=(1, 1, 1)):
    x, xp = _c(x); w, wp = _c(w); b, bp = _c(b)
    N, Cin, D, H, W = x.shape
    Cout, _, kD, kH, kW = w.shape
    o = [(s + 2 * p - k) // st + 1 for s, p, k, st in
         zip((D, H, W), padding, (kD, kH, kW), stride)]
    out = np.empty((N, Cout, *o), np.float32)
    lib().pds_oracle_conv3d(xp, wp, bp, out.ctypes.data_as(_f), N, Cin, D, H, W,
                            Cout, kD, kH, kW, *stride, *padding)
    return out


def conv2d(x, w, b, stride=1):
    x, xp = _c(x); w, wp = _c(w); b, bp = _c(b)
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    oh, ow = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    out = np.empty((N, Cout, oh, ow), np.float32)
    lib().pds_oracle_conv2d(xp, wp, bp, out.ctypes.data_as(_f), N, Cin, H, W, Cout,
                            k, stride)
    return out


def conv_transpose3d(x, w, b, stride, padding):
    x, xp = _c(x); w, wp = _c(w); b, bp = _c(b)
    N, Cin, D, H, W = x.shape
    _, Cout, kD, kH, kW = w.shape
    o = [(s - 1) * st - 2 * p + k for s, p, k, st in
         zip((D, H, W), padding, (kD, kH, kW), stride)]
    out = np.empty((N, Cout, *o), np.float32)
    lib().pds_oracle_conv_transpose3d(xp, wp, bp, out.ctypes.data_as(_f), N, Cin, D,
                                      H, W, Cout, kD, kH, kW, *stride, *padding)
    return out


def instance_norm(x, gamma=None, beta=None):
    x = np.array(x, dtype=np.float32, order='C')
    N, C = x.shape[:2]
    S = int(np.prod(x.shape[2:]))
    g = _c(gamma)[1] if gamma is not None else None
    b = _c(beta)[1] if beta is not None else None
    lib().pds_oracle_instance_norm(x.ctypes.data_as(_f), g, b, N, C, ctypes.c_long(S))
    return x


def matching_concat(left, right, maximum_disparity):
    left, lp = _c(left); right, rp = _c(right)
    B, C, H, W = left.shape
    D = maximum_disparity + 1
    out = np.empty((B, D, 2 * C, H, W), np.float32)
    lib().pds_oracle_matching_concat(lp, rp, out.ctypes.data_as(_f), B, C, H, W, D)
    return out


def matching_operation(x, params_flat, f=64, csig=8, n_res=2):
    x, xp = _c(x); p, pp = _c(params_flat)
    N, Cin, H, W = x.shape
    out = np.empty((N, csig, H, W), np.float32)
    used = lib().pds_oracle_matching_operation(xp, pp, out.ctypes.data_as(_f), N, Cin,
                                               f, csig, n_res, H, W)
    assert used == p.size, (used, p.size)
    return out


def matching(left, right, params_flat, maximum_disparity, f=64, csig=8, n_res=2):
    left, lp = _c(left); right, rp = _c(right); p, pp = _c(params_flat)
    B, C, H, W = left.shape
    D = maximum_disparity + 1
    out = np.empty((B, csig, D, H, W), np.float32)
    lib().pds_oracle_matching(lp, rp, pp, out.ctypes.data_as(_f), B, C, f, csig,
                              n_res, H, W, D)
    return out


def regularization(signatures, shortcut, params_flat):
    s, sp = _c(signatures); sc, scp = _c(shortcut); p, pp = _c(params_flat)
    B, F, D, H, W = s.shape
    out = np.empty((B, 2 * D, 4 * H, 4 * W), np.float32)
    used = lib().pds_oracle_regularization(sp, scp, pp, out.ctypes.data_as(_f), B, F,
                                           D, H, W)
    assert used == p.size, (used, p.size)
    return out


def contraction_block(x, params_flat):
    x, xp = _c(x); p, pp = _c(params_flat)
    N, C, D, H, W = x.shape
    o = (N, 2 * C, (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1)
    down, smooth = np.empty(o, np.float32), np.empty(o, np.float32)
    used = lib().pds_oracle_contraction_block(xp, pp, down.ctypes.data_as(_f),
                                              smooth.ctypes.data_as(_f), N, C, D, H, W)
    assert used == p.size
    return down, smooth


def expansion_block(x, skip, params_flat):
    x, xp = _c(x); k, kp = _c(skip); p, pp = _c(params_flat)
    N, C, D, H, W = x.shape
    out = np.empty((N, C // 2, 2 * D, 2 * H, 2 * W), np.float32)
    used = lib().pds_oracle_expansion_block(xp, kp, pp, out.ctypes.data_as(_f), N, C,
                                            D, H, W)
    assert used == p.size
    return out


def embedding(image, params_flat, f=64, fs=8, n_res=2):
    x, xp = _c(image); p, pp = _c(params_flat)
    N, C, H, W = x.shape
    h4, w4 = ((H - 1) // 2 + 1 - 1) // 2 + 1, ((W - 1) // 2 + 1 - 1) // 2 + 1
    desc = np.empty((N, f, h4, w4), np.float32)
    short = np.empty((N, fs, h4, w4), np.float32)
    used = lib().pds_oracle_embedding(xp, pp, desc.ctypes.data_as(_f),
                                      short.ctypes.data_as(_f), N, C, f, fs, n_res, H, W)
    assert used == p.size
    return desc, short


def subpixel_map(similarities, half_support_window=4, disparity_step=2):
    """Returns (disparity float32 (B,H,W), argmax int64 (B,H,W))."""
    s, sp = _c(similarities)
    B, D, H, W = s.shape
    disp = np.empty((B, H, W), np.float32)
    idx = np.empty((B, H, W), np.int64)
    lib().pds_oracle_subpixel_map(sp, disp.ctypes.data_as(_f), idx.ctypes.data_as(_i64),
                                  B, D, H, W, half_support_window, disparity_step)
    return disp, idx


def pad(x, minimum_size=64):
    x, xp = _c(x)
    N, C, H, W = x.shape
    hp = -(-H // minimum_size) * minimum_size
    wp = -(-W // minimum_size) * minimum_size
    out = np.empty((N, C, hp, wp), np.float32)
    lib().pds_oracle_pad(xp, out.ctypes.data_as(_f), N, C, H, W, minimum_size)
    return out


def network_forward(left, right, p_embedding, p_matching, p_regularization,
                    maximum_disparity, return_cost=False):
    l, lp = _c(left); r, rp = _c(right)
    pe, pep = _c(p_embedding); pm, pmp = _c(p_matching); pr, prp = _c(p_regularization)
    B, _, H, W = l.shape
    disp = np.empty((B, H, W), np.float32)
    cost, cp = None, None
    if return_cost:
        hp, wp = -(-H // 64) * 64, -(-W // 64) * 64
        cost = np.empty((B, (maximum_disparity + 1) // 2, hp, wp), np.float32)
        cp = cost.ctypes.data_as(_f)
    rc = lib().pds_oracle_network_forward(lp, rp, pep, pmp, prp, disp.ctypes.data_as(_f),
                                          cp, B, H, W, maximum_disparity)
    if rc != 0:
        raise ValueError('"maximum_disparity" + 1 should be multiple of 64')
    return (disp, cost) if return_cost else disp


# ---- f4 (training): adjoints, restated in numpy (small cases only; float64 inside) -----------------
def matching_concat_backward(grad_volume):
    """Adjoint of matching_concat (what autograd accumulates through the reference's per-disparity
    pad / slice / cat nodes, matching.py:53-60): grad_volume (B, D, 2C, H, W) ->
    grad_left[b,c,y,x] = sum_d g[b,d,c,y,x];  grad_right[b,c,y,x] = sum_{d: x+d<W} g[b,d,C+c,y,x+d]."""
    g = np.asarray(grad_volume, dtype=np.float64)
    B, D, C2, H, W = g.shape
    C = C2 // 2
    grad_left = g[:, :, :C].sum(axis=1)
    grad_right = np.zeros((B, C, H, W))
    for d in range(min(D, W)):
        grad_right[..., :W - d] += g[:, d, C:, :, d:]
    return grad_left.astype(np.float32), grad_right.astype(np.float32)


def leaky_instance_norm(x, gamma, beta, eps=1e-5, slope=0.1):
    """InstanceNorm(affine)(LeakyReLU(x)) (network_blocks.py:47-58) and what its backward needs."""
    x = np.asarray(x, dtype=np.float64)
    z = np.where(x > 0, x, slope * x)
    axes = tuple(range(2, x.ndim))
    mean = z.mean(axis=axes, keepdims=True)
    var = z.var(axis=axes, keepdims=True)               # biased
    rstd = 1.0 / np.sqrt(var + eps)
    shape = (1, -1) + (1,) * (x.ndim - 2)
    zhat = (z - mean) * rstd
    y = zhat * np.asarray(gamma, np.float64).reshape(shape) + np.asarray(beta, np.float64).reshape(shape)
    return y, (zhat, rstd)


def leaky_instance_norm_backward(x, grad_y, gamma, eps=1e-5, slope=0.1):
    """-> grad_x, grad_gamma, grad_beta:  dz = gamma * rstd * (dy - mean(dy) - zhat * mean(dy * zhat)),
    dx = dz * LeakyReLU'(x) (slope at x <= 0), d gamma = sum dy * zhat, d beta = sum dy."""
    x = np.asarray(x, dtype=np.float64)
    dy = np.asarray(grad_y, dtype=np.float64)
    _, (zhat, rstd) = leaky_instance_norm(x, gamma, np.zeros_like(np.asarray(gamma, np.float64)), eps, slope)
    axes = tuple(range(2, x.ndim))
    shape = (1, -1) + (1,) * (x.ndim - 2)
    g = np.asarray(gamma, np.float64).reshape(shape)
    m1 = dy.mean(axis=axes, keepdims=True)
    m2 = (dy * zhat).mean(axis=axes, keepdims=True)
    dz = g * rstd * (dy - m1 - zhat * m2)
    dx = dz * np.where(x > 0, 1.0, slope)
    red = (0,) + axes
    return dx, (dy * zhat).sum(axis=red), dy.sum(axis=red)
