"""ctypes binding of libxunet_b200.so (include/xunet_b200.h).  There is NO fallback: if the CUDA
library is missing or fails to load, importing the compute path raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# XUNET_LIB=<path> loads another build of the same library (same-call A/B of a kernel change against the previous build)
LIB_PATH = Path(os.environ['XUNET_LIB']).resolve() if os.environ.get('XUNET_LIB') else Path(__file__).resolve().parent / 'libxunet_b200.so'
MAX_LEVELS = 8
DTYPE_F32, DTYPE_BF16 = 0, 1
RAYS = {'v3d130_ij': 0, 'opencv_uv': 1}


class XunetConfig(C.Structure):
    _fields_ = [('ch', C.c_int), ('n_levels', C.c_int), ('ch_mult', C.c_int * MAX_LEVELS), ('emb_ch', C.c_int),
                ('num_res_blocks', C.c_int), ('n_attn_resolutions', C.c_int),
                ('attn_resolutions', C.c_int * MAX_LEVELS), ('attn_heads', C.c_int), ('dropout', C.c_float),
                ('use_pos_emb', C.c_int), ('use_ref_pose_emb', C.c_int), ('ray_convention', C.c_int)]


class XunetBatch(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('x', 'z', 'logsnr', 'R1', 't1', 'R2', 't2', 'K', 'cond_mask', 'rays')]


# void fn(void* user, long long elem_offset, long long n_elems)  -- the data-parallel gradient-bucket hook
BUCKET_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_longlong, C.c_longlong)


# name -> (restype, argtypes): every symbol include/xunet_b200.h declares
SYMBOLS = {
    'xunet_last_error': (C.c_char_p, []),
    'xunet_version': (C.c_int, []),
    'xunet_create': (C.c_int, [C.POINTER(XunetConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'xunet_destroy': (None, [C.c_void_p]),
    'xunet_param_count': (C.c_longlong, [C.c_void_p]),
    'xunet_param_leaves': (C.c_int, [C.c_void_p]),
    'xunet_param_leaf': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                   C.POINTER(C.c_longlong * 5), C.POINTER(C.c_longlong)]),
    'xunet_workspace_bytes': (C.c_longlong, [C.c_void_p]),
    'xunet_tap_count': (C.c_int, [C.c_void_p]),
    'xunet_tap': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int * 4),
                            C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    'xunet_forward': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(XunetBatch), C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    'xunet_set_static_conditioning': (C.c_int, [C.c_void_p, C.c_int]),
    'xunet_backward': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(XunetBatch), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    'xunet_set_grad_bucket_callback': (C.c_int, [C.c_void_p, BUCKET_FN, C.c_void_p, C.c_longlong]),
    'xunet_grad_bucket_count': (C.c_int, [C.c_void_p]),
    'xunet_count_kernels': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(XunetBatch), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'xunet_adam_step': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong,
                                  C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    'xunet_sampler_update': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_ulonglong,
                                       C.c_void_p]),
    'xunet_sampler_step_table': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'xunet_forward_diffusion': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong,
                                          C.c_void_p]),
    'xunet_dropout_mask': (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_ulonglong, C.c_float, C.c_void_p]),
    'xunet_op_conv': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] +
                      [C.c_int] * 8 + [C.c_float, C.c_void_p]),
    'xunet_op_conv_dgrad': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 +
                            [C.c_float, C.c_int, C.c_void_p]),
    'xunet_op_conv_wgrad': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] +
                            [C.c_int] * 8 + [C.c_float, C.c_void_p]),
    'xunet_op_attention': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] +
                           [C.c_int] * 5 + [C.c_void_p]),
    'xunet_op_attention_bwd': (C.c_int, [C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p]),
}

_lib = None


def load():
    """Load the CUDA library; raises (never falls back) if it is absent or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -m novel_view_synthesis_3d_b200.build` '
                           f'(or __graft_entry__.build()); there is no CPU/PyTorch fallback for the X-UNet path')
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        raise RuntimeError(f'{what}: {load().xunet_last_error().decode()}')
