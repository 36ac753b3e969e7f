"""Deterministic synthetic parameters and inputs shared by the golden-vector
generator (oracle/make_golden.py) and the tests.  TEST INFRASTRUCTURE ONLY.

Parameters are drawn with numpy's legacy ``RandomState`` (frozen bit stream),
in the reference modules' ``state_dict()`` order, so fixtures only need to
store outputs: the tests regenerate identical weights and inputs.

Key/shape lists restate the reference constructors:
  MatchingOperation  matching.py:69-95
  Regularization     regularization.py:74-92
  Embedding          embedding.py:14-44
and are asserted against the real ``state_dict()`` in make_golden.py.
"""
import numpy as np


def _block(prefix, cout, cin, kernel, transposed=False):
    wshape = (cin, cout) + kernel if transposed else (cout, cin) + kernel
    return [(prefix + '.0.weight', wshape), (prefix + '.0.bias', (cout,)),
            (prefix + '.2.weight', (cout,)), (prefix + '.2.bias', (cout,))]


def matching_operation_specs(cin=128, f=64, csig=8, n_res=2):
    p = '_matching_operation_modules'
    specs = [(f'{p}.0.weight', (f, cin, 3, 3)), (f'{p}.0.bias', (f,))]
    for r in range(n_res):
        for j in range(2):
            specs += _block(f'{p}.{1 + r}.convolutions.{j}', f, f, (3, 3))
    specs += [(f'{p}.{1 + n_res}.weight', (csig, f, 3, 3)),
              (f'{p}.{1 + n_res}.bias', (csig,))]
    return specs


def regularization_specs(f=8):
    k3, k4 = (3, 3, 3), (4, 4, 4)
    specs = _block('_smoothing', f, f, k3)
    for i, s in enumerate([1, 2, 4, 8]):
        c = f * s
        specs += _block(f'_contraction_blocks.{i}._downsampling_2x', 2 * c, c, k3)
        specs += _block(f'_contraction_blocks.{i}._smoothing', 2 * c, 2 * c, k3)
    for i, s in enumerate([16, 8, 4, 2]):
        c = f * s
        specs += _block(f'_expansion_blocks.{i}._upsampling_2x', c // 2, c, k4,
                        transposed=True)
        specs += _block(f'_expansion_blocks.{i}._smoothing', c // 2, c // 2, k3)
    specs += _block('_upsample_to_halfsize', f // 2, f, k4, transposed=True)
    specs += [('_upsample_to_fullsize.weight', (f // 2, 1, 3, 4, 4)),
              ('_upsample_to_fullsize.bias', (1,))]
    return specs


def contraction_block_specs(c):
    k3 = (3, 3, 3)
    return (_block('_downsampling_2x', 2 * c, c, k3) +
            _block('_smoothing', 2 * c, 2 * c, k3))


def expansion_block_specs(c):
    return (_block('_upsampling_2x', c // 2, c, (4, 4, 4), transposed=True) +
            _block('_smoothing', c // 2, c // 2, (3, 3, 3)))


def embedding_specs(cimg=3, f=64, fs=8, n_res=2):
    p = '_embedding_modules'
    specs = _block(f'{p}.1', f, cimg, (5, 5)) + _block(f'{p}.2', f, f, (5, 5))
    for r in range(n_res):
        for j in range(2):
            specs += _block(f'{p}.{3 + r}.convolutions.{j}', f, f, (3, 3))
    specs += _block('_shortcut', fs, f, (3, 3))
    return specs


def network_specs():
    """state_dict() of PdsNetwork.default(): embedding, matching, regularization."""
    return ([('_embedding.' + k, s) for k, s in embedding_specs()] +
            [('_matching._operation.' + k, s) for k, s in matching_operation_specs()] +
            [('_regularization.' + k, s) for k, s in regularization_specs()])


def make_params(specs, seed):
    """Ordered dict key -> float32 array.  Conv weights ~ N(0, 1/fan_in) scaled
    up 1.5x, biases ~ 0.1 N, InstanceNorm gamma ~ 1 + 0.2 N, beta ~ 0.2 N."""
    rng = np.random.RandomState(seed)
    out = {}
    for key, shape in specs:
        if len(shape) > 1:
            is_transposed = 'upsampl' in key or 'upsample' in key
            taps = int(np.prod(shape[2:]))
            fan_in = (shape[0] if is_transposed else shape[1]) * taps
            if is_transposed:
                fan_in = max(1, fan_in // (8 if len(shape) == 5 else 4))
            v = rng.standard_normal(shape) * (1.5 / np.sqrt(fan_in))
        elif key.endswith('.2.weight'):
            v = 1.0 + 0.2 * rng.standard_normal(shape)
        elif key.endswith('.2.bias'):
            v = 0.2 * rng.standard_normal(shape)
        else:
            v = 0.1 * rng.standard_normal(shape)
        out[key] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def flatten(params):
    return np.concatenate([v.ravel() for v in params.values()]).astype(np.float32)


def tensor(shape, seed, scale=1.0, uniform=False):
    rng = np.random.RandomState(seed)
    v = rng.random_sample(shape) if uniform else rng.standard_normal(shape)
    return np.ascontiguousarray(v * scale, dtype=np.float32)
