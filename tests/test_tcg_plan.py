"""Host logic of the generic tcgen05 convolution engine (csrc/conv_tcg_plan.cu), checked on the
CPU: the planner's program for a layer (TMA boxes with zero fill, K=16 MMA entries addressed by
start offset / LBO / SBO, weight gather table) is EMULATED in numpy and compared with the ATen
operator the reference layer calls (network_blocks.py:61-85, 106-131; embedding.py:33-38).
No kernel runs here; the GPU tests check the kernel that interprets the same program."""
import ctypes
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from practicaldeepstereo_nips2018_b200 import _capi

CONV3_S1, CONV3_S2, TCONV4_S2, CONV5_S2, TCONV4_S2M, CONV3_S1X4 = 0, 1, 2, 3, 4, 5


def describe(kind, nd, cin, cout, Z, Y, X, S=2):
    fn = ctypes.CDLL(_capi.LIB_PATH).pds_tcg_plan_describe
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 8 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    need = fn(kind, nd, cin, cout, Z, Y, X, S, None, 0)
    assert need > 0, f'planner refused the layer (status {-need})'
    buf = (ctypes.c_int * need)()
    assert fn(kind, nd, cin, cout, Z, Y, X, S, buf, need) == need
    a = np.frombuffer(buf, dtype=np.int32).copy()
    names = ('N nacc ntx ntz ncls nph P GZ GY GX OZ OY OX IZ IY IX BX BY BZ PB units_per_item '
             'resident stages box_bytes stage_bytes wres_bytes w_total_bytes nu nb ne max_boxes merged xg').split()
    p = dict(zip(names, a[:40].tolist()))
    o = 40
    p['units'] = a[o:o + p['nu'] * 6].reshape(-1, 6); o += p['nu'] * 6
    p['boxes'] = a[o:o + p['nb'] * 4].reshape(-1, 4); o += p['nb'] * 4
    p['entries'] = a[o:o + p['ne'] * 2].reshape(-1, 2).astype(np.int64) & 0xffffffff; o += p['ne'] * 2
    p['wsrc'] = a[o:o + p['ne'] * 8].reshape(-1, 2, 4); o += p['ne'] * 8
    p['tile_off'] = a[o:o + p['nacc']]
    return p


def emulate(p, kind, nd, x, w, S=2):
    """x (n, Cin, Z, Y, X) float64, w in PyTorch layout -> (n, Cout, OZ, OY, OX) as the kernel
    computes it from the plan (single exact term; the split only changes rounding)."""
    n, cin = x.shape[:2]
    cout = w.shape[1] if kind in (TCONV4_S2, TCONV4_S2M) else w.shape[0]
    merged = kind == TCONV4_S2M
    N, P, nph = p['N'], p['P'], p['nph']
    IZ, IY, IX, BX, BY, BZ, PB = (p[k] for k in ('IZ', 'IY', 'IX', 'BX', 'BY', 'BZ', 'PB'))
    box16 = p['box_bytes'] // 16
    # input planes [n][phase][P][IZ][IY][IX][8]
    x4 = kind == CONV3_S1X4
    if x4:       # four x-phase sub-volumes
        xin = np.stack([x[..., px::4] for px in range(4)], 1).reshape(n, 4, P, 8, IZ, IY, IX)
    elif nph == 1:
        xin = x.reshape(n, 1, P, 8, IZ, IY, IX)
    elif nd == 3:
        xin = np.stack([x[:, :, pz::2, py::2, px::2] for pz, py, px in itertools.product((0, 1), repeat=3)], 1)
        xin = xin.reshape(n, 8, P, 8, IZ, IY, IX)
    else:
        xin = np.stack([x[:, :, :, py::2, px::2] for py, px in itertools.product((0, 1), repeat=2)], 1)
        xin = xin.reshape(n, 4, P, 8, IZ, IY, IX)
    planes = np.moveaxis(xin, 3, -1).reshape(n, nph * P, IZ, IY, IX, 8)
    out = np.zeros((n, cout, p['OZ'], p['OY'], p['OX']))
    assert p['xg'] == (4 if x4 else 1) and p['OX'] == p['GX'] * p['xg'] * (2 if kind in (TCONV4_S2, TCONV4_S2M) else 1)
    rows = np.arange(128)
    row_off = (rows >> 3) * BX + (rows & 7)
    upi = p['units_per_item']
    ent16 = 2 * S * N
    ntx, ntz = p['ntx'], p['ntz']
    assert p['nacc'] == ntx * ntz
    for b in range(n):
        for z0, y0, x0 in itertools.product(range(0, p['GZ'], ntz), range(0, p['GY'], 16),
                                            range(0, p['GX'], 8 * ntx)):
            for cls in range(p['ncls']):
                D = np.zeros((p['nacc'], 128, N))
                for u in p['units'][cls * upi:(cls + 1) * upi]:
                    image = np.zeros((p['max_boxes'] * box16 + 4 * BZ * BY * BX, 8))  # slack: rows past a box
                    for j, (dx, dy, dz, plane) in enumerate(p['boxes'][u[2]:u[3]]):
                        box = np.zeros((PB, BZ, BY, BX, 8))
                        for pz, py, px in itertools.product(range(BZ), range(BY), range(BX)):
                            gz, gy, gx = z0 + dz + pz, y0 + dy + py, x0 + dx + px
                            if 0 <= gz < IZ and 0 <= gy < IY and 0 <= gx < IX:   # TMA zero fill otherwise
                                box[:, pz, py, px] = planes[b, plane:plane + PB, gz, gy, gx]
                        image[j * box16:j * box16 + PB * BZ * BY * BX] = box.reshape(-1, 8)
                    for e in range(u[0], u[1]):
                        a16, lbo = int(p['entries'][e, 0]) & 0xffff, int(p['entries'][e, 0]) >> 16
                        assert int(p['entries'][e, 1]) == (u[4] if p['resident'] else
                                                           p['max_boxes'] * S * box16) + (e - u[0]) * ent16
                        for h in range(2):
                            kz, ky, kx, g = p['wsrc'][e, h]
                            if g < 0:
                                continue
                            if merged:
                                # column = class * cout + channel; tap offset o = k - 1; class bit c uses
                                # offsets c (kernel index 1 - c) and c - 1 (kernel index 3 - c)
                                wm = np.zeros((8 * cout, 8))
                                for mc in range(8):
                                    ks = []
                                    for cb, off in zip(((mc >> 2) & 1, (mc >> 1) & 1, mc & 1),
                                                       (kz - 1, ky - 1, kx - 1)):
                                        t = cb - off
                                        ks.append(1 - cb + 2 * t if t in (0, 1) else None)
                                    if None not in ks:
                                        wm[mc * cout:(mc + 1) * cout] = w[8 * g:8 * g + 8, :, ks[0], ks[1], ks[2]].T
                            elif x4:
                                # column = position j of the voxel group * cout + channel; extended x
                                # tap kx reaches output j through kernel index kx - j
                                wm = np.zeros((4 * cout, 8))
                                for j in range(4):
                                    if 0 <= kx - j <= 2:
                                        wm[j * cout:(j + 1) * cout] = w[:, 8 * g:8 * g + 8, kz, ky, kx - j]
                            elif kind == TCONV4_S2:
                                wm = w[8 * g:8 * g + 8, :, kz, ky, kx].T          # (cout, 8)
                            else:
                                wm = w[:, 8 * g:8 * g + 8, kz, ky, kx]
                            for i in range(p['nacc']):
                                A = image[a16 + int(p['tile_off'][i]) + h * lbo + row_off]   # (128, 8)
                                D[i, :, :wm.shape[0]] += A @ wm.T
                cz, cy, cx = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
                m = 2 if kind in (TCONV4_S2, TCONV4_S2M) else 1
                for i in range(p['nacc']):
                    ix, iz = i % ntx, i // ntx
                    for r in range(128):
                        gz, gy, gx = z0 + iz, y0 + (r >> 3), x0 + 8 * ix + (r & 7)
                        if gz < p['GZ'] and gy < p['GY'] and gx < p['GX']:
                            if merged:
                                for mc in range(8):
                                    out[b, :, 2 * gz + ((mc >> 2) & 1), 2 * gy + ((mc >> 1) & 1),
                                        2 * gx + (mc & 1)] = D[i, r, mc * cout:(mc + 1) * cout]
                            elif x4:
                                for j in range(4):
                                    out[b, :, gz, gy, 4 * gx + j] = D[i, r, j * cout:(j + 1) * cout]
                            else:
                                out[b, :, m * gz + cz if nd == 3 else 0, m * gy + cy, m * gx + cx] = D[i, r, :cout]
    return out


def reference(kind, nd, x, w):
    xt, wt = torch.from_numpy(x), torch.from_numpy(w)
    if nd == 2:
        xt, wt = xt[:, :, 0], wt[:, :, 0]
    if kind in (CONV3_S1, CONV3_S1X4):
        y = (F.conv3d if nd == 3 else F.conv2d)(xt, wt, padding=1)
    elif kind == CONV3_S2:
        y = (F.conv3d if nd == 3 else F.conv2d)(xt, wt, padding=1, stride=2)
    elif kind == CONV5_S2:
        y = F.conv2d(xt, wt, padding=2, stride=2)
    else:      # TCONV4_S2 and its merged-class variant
        y = F.conv_transpose3d(xt, wt, padding=1, stride=2)
    y = y.numpy()
    return y[:, :, None] if nd == 2 else y


CASES = [
    # kind, nd, Cin, Cout, Z, Y, X
    (CONV3_S1, 3, 8, 8, 5, 18, 19),       # paired taps, partial tiles
    (CONV3_S1, 3, 16, 16, 3, 17, 9),
    (CONV3_S1, 3, 32, 32, 2, 6, 10),
    (CONV3_S1, 3, 64, 64, 3, 5, 9),       # weights streamed per (chunk, dz)
    (CONV3_S1, 3, 128, 128, 2, 3, 5),
    (CONV3_S2, 3, 8, 16, 6, 20, 18),      # phase-separated input, pairs across phases
    (CONV3_S2, 3, 16, 32, 4, 8, 20),
    (CONV3_S2, 3, 64, 128, 2, 6, 10),
    (TCONV4_S2, 3, 128, 64, 2, 3, 5),
    (TCONV4_S2, 3, 32, 16, 3, 6, 9),
    (TCONV4_S2, 3, 16, 8, 3, 18, 10),
    (TCONV4_S2, 3, 8, 4, 5, 17, 17),
    (TCONV4_S2M, 3, 8, 4, 5, 17, 17),     # parity classes merged along N
    (TCONV4_S2M, 3, 16, 8, 3, 18, 10),
    (CONV3_S1X4, 3, 8, 8, 5, 18, 20),     # four voxels per GEMM row, x-phase-separated input
    (CONV3_S1X4, 3, 8, 8, 3, 5, 76),      # several tiles along x, partial last tile
    (CONV3_S1X4, 3, 8, 4, 2, 17, 12),
    (CONV5_S2, 2, 64, 64, 1, 36, 20),
    (CONV3_S1, 2, 64, 8, 1, 20, 70),
    (CONV3_S1, 2, 64, 64, 1, 17, 12),
]


@pytest.mark.parametrize('kind,nd,cin,cout,Z,Y,X', CASES)
def test_plan_emulation_matches_aten(kind, nd, cin, cout, Z, Y, X):
    rng = np.random.RandomState(kind * 1000 + cin + cout + Z + Y + X)
    x = rng.randn(2 if cin <= 16 else 1, cin, Z, Y, X)
    k = {CONV3_S1: 3, CONV3_S2: 3, TCONV4_S2: 4, CONV5_S2: 5, TCONV4_S2M: 4, CONV3_S1X4: 3}[kind]
    kz = k if nd == 3 else 1
    w = rng.randn(*((cin, cout) if kind in (TCONV4_S2, TCONV4_S2M) else (cout, cin)), kz, k, k) / np.sqrt(cin * k * k * kz)
    p = describe(kind, nd, cin, cout, Z, Y, X)
    # structural invariants the kernel relies on
    assert p['nacc'] * 2 * p['N'] <= 512 and p['stages'] >= 2
    assert p['box_bytes'] % 128 == 0 and p['stage_bytes'] % 128 == 0
    assert (p['entries'][:, 0] >> 16).max() < (1 << 14)
    got = emulate(p, kind, nd, x, w)
    ref = reference(kind, nd, x, w)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, atol=1e-10)


def test_plan_full_size_layers_fit():
    """Every layer of the C2 / C3 / C4 hourglass and of the embedding plans within shared memory."""
    for (D, H, W) in [(48, 144, 240), (64, 144, 240), (48, 96, 320), (16, 16, 32)]:
        c, z, y, x = 8, D, H, W
        assert describe(CONV3_S1, 3, 8, 8, z, y, x)['stages'] >= 2
        assert describe(CONV3_S1X4, 3, 8, 8, z, y, x)['stages'] >= 2
        for _ in range(4):
            assert describe(CONV3_S2, 3, c, 2 * c, z, y, x)['stages'] >= 2
            c, z, y, x = 2 * c, z // 2, y // 2, x // 2
            assert describe(CONV3_S1, 3, c, c, z, y, x)['stages'] >= 2
        for _ in range(4):
            assert describe(TCONV4_S2M if c // 2 <= 8 else TCONV4_S2, 3, c, c // 2, z, y, x)['stages'] >= 2
            c, z, y, x = c // 2, z * 2, y * 2, x * 2
            assert describe(CONV3_S1, 3, c, c, z, y, x)['stages'] >= 2
        assert describe(TCONV4_S2M, 3, 8, 4, z, y, x)['stages'] >= 2
    assert describe(CONV5_S2, 2, 64, 64, 1, 288, 480)['stages'] >= 2
    assert describe(CONV3_S1, 2, 64, 8, 1, 144, 240)['stages'] >= 2


def test_plan_rejects_unsupported():
    fn = ctypes.CDLL(_capi.LIB_PATH).pds_tcg_plan_describe
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 8 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    assert fn(CONV3_S2, 3, 8, 16, 5, 8, 8, 2, None, 0) < 0      # odd extent under stride 2
    assert fn(CONV3_S1, 3, 12, 8, 4, 8, 8, 2, None, 0) < 0      # channels not a multiple of 8
    assert fn(TCONV4_S2, 2, 16, 8, 1, 8, 8, 2, None, 0) < 0     # 2-D transposed: not served
