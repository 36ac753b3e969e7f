"""ctypes binding of the C-ABI declared in include/pds_b200.h.

This is the only bridge between the Python modules and the CUDA library.  There
is NO fallback: if libpds_b200.so is missing or fails to load, every kernel
entry point raises -- the product path never routes through PyTorch ops or the
CPU oracle.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpds_b200.so')
if os.environ.get('PDS_B200_LIB'):          # experiments: an alternative build of the same library
    LIB_PATH = os.environ['PDS_B200_LIB']

PDS_OK, PDS_ERR_INVALID_ARGUMENT, PDS_ERR_CUDA, PDS_ERR_WORKSPACE, PDS_ERR_UNSUPPORTED = range(5)
PDS_F32, PDS_BF16 = 0, 1
PRECISIONS = {'fp32': 0, 'bf16x3': 1, 'bf16x2': 2, 'bf16': 3, 'fp16x2': 4, 'fp16': 5}
# Arithmetic of the convolution stacks when a module is built without an explicit `precision`:
# the fp32-grade split-operand tensor-core mode.  'fp32' selects the CUDA-core
# FFMA kernels (10x slower); PDS_B200_PRECISION overrides the default for a whole process.
DEFAULT_PRECISION = os.environ.get('PDS_B200_PRECISION', 'fp16x2')
IMAGE_LAYOUTS = {'f32_nchw': 0, 'u8_nchw': 1, 'u8_nhwc': 2}   # enum pds_image_layout

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes): must list every symbol of include/pds_b200.h
SIGNATURES = {
    'pds_version': (_i, []),
    'pds_status_string': (ctypes.c_char_p, [_i]),
    'pds_last_error': (ctypes.c_char_p, []),
    'pds_launch_count': (ctypes.c_ulonglong, []),
    'pds_profiler_enable': (None, [_i]),
    'pds_profiler_reset': (None, []),
    'pds_profiler_read': (_i, [_i, ctypes.c_char_p, _i, ctypes.POINTER(ctypes.c_ulonglong),
                               ctypes.POINTER(ctypes.c_double)]),
    'pds_profiler_read_work': (_i, [_i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    'pds_matching_concat': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'pds_matching_stack': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'pds_matching_unstack': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'pds_matching_concat_backward': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'pds_instance_norm_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, _vp]),
    'pds_instance_norm_backward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, ctypes.c_longlong, ctypes.c_float, _vp]),
    'pds_matching_op_create': (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp]),
    'pds_matching_op_destroy': (None, [_vp]),
    'pds_matching_op_workspace_bytes': (_sz, [_vp, _i, _i, _i, _i]),
    'pds_matching_op_forward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    'pds_regularization_create': (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i, _vp]),
    'pds_regularization_destroy': (None, [_vp]),
    'pds_regularization_workspace_bytes': (_sz, [_vp, _i, _i, _i, _i]),
    'pds_regularization_forward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    'pds_regularization_forward_disparity': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i,
                                                  _vp, _sz, _vp]),
    'pds_contraction_block_workspace_bytes': (_sz, [_i, _i, _i, _i, _i]),
    'pds_contraction_block_forward': (_i, [ctypes.POINTER(_vp), _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    'pds_expansion_block_workspace_bytes': (_sz, [_i, _i, _i, _i, _i]),
    'pds_expansion_block_forward': (_i, [ctypes.POINTER(_vp), _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    'pds_embedding_create': (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp]),
    'pds_embedding_destroy': (None, [_vp]),
    'pds_embedding_workspace_bytes': (_sz, [_vp, _i, _i, _i]),
    'pds_embedding_forward': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    'pds_embedding_forward_images': (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _sz,
                                          _vp]),
    'pds_subpixel_map': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'pds_disparity_errors': (_i, [_vp, _vp, _vp, _vp, _sz, ctypes.c_float, _vp, _vp]),
    'pds_subpixel_cross_entropy_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _i, _vp]),
    'pds_subpixel_cross_entropy_backward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                                 ctypes.c_float, _i, _vp]),
}

_lib = None


class PdsLibraryError(RuntimeError):
    pass


def lib():
    """Loads libpds_b200.so (once).  Fails loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PdsLibraryError(
                f'{LIB_PATH} is missing: build it with `python -m '
                'practicaldeepstereo_nips2018_b200.build` (there is no fallback path)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if a symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status):
    if status == PDS_OK:
        return
    msg = lib().pds_last_error().decode() or lib().pds_status_string(status).decode()
    if status == PDS_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise RuntimeError(f'pds_b200: {msg}')


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float32:
        return PDS_F32
    if t.dtype == torch.bfloat16:
        return PDS_BF16
    raise TypeError(f'pds_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}')


def require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                'pds_b200 inference kernels need CUDA tensors; there is no CPU path '
                '(the CPU oracle under oracle/ is test infrastructure only)')


def ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


def pointer_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def launch_count():
    return int(lib().pds_launch_count())


def profiler_enable(on=True):
    lib().pds_profiler_enable(1 if on else 0)


def profiler_reset():
    lib().pds_profiler_reset()


def profiler_report():
    """{kernel class: (launches, total device ms, algorithmic FLOPs, algorithmic HBM bytes)}
    since the last reset."""
    out, i = {}, 0
    name = ctypes.create_string_buffer(128)
    n, ms = ctypes.c_ulonglong(), ctypes.c_double()
    fl, by = ctypes.c_double(), ctypes.c_double()
    while lib().pds_profiler_read(i, name, 128, ctypes.byref(n), ctypes.byref(ms)):
        lib().pds_profiler_read_work(i, ctypes.byref(fl), ctypes.byref(by))
        out[name.value.decode()] = (int(n.value), float(ms.value), float(fl.value),
                                    float(by.value))
        i += 1
    return out
