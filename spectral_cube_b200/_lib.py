"""
ctypes binding of libsc_b200.so (include/sc_b200.h).  This is the only place the product
talks to native code; there is NO CPU fallback: if the library is missing or has no CUDA
device to run on, calls raise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libsc_b200.so')

SC_MASK_MAX_NODES = 16
# sc_mask_kind
MASK_FINITE, MASK_CMP_SCALAR, MASK_CMP_ARRAY, MASK_BOOL, MASK_AND, MASK_OR, MASK_XOR, MASK_NOT = range(1, 9)
# sc_cmp_op
GT, GE, LT, LE, EQ, NE = range(6)
# sc_dtype
F32, F64, U8 = 0, 1, 2
# sc_op
OP_MOMENTS, OP_SPECTRAL_SMOOTH, OP_SPATIAL_SMOOTH, OP_SPECTRAL_INTERP, OP_REPROJECT, OP_SMOOTH_MOMENTS = range(1, 7)
WANT_M0, WANT_M1, WANT_M2 = 1, 2, 4


class MaskNode(C.Structure):
    _fields_ = [('kind', C.c_int32), ('op', C.c_int32), ('a', C.c_int32), ('b', C.c_int32),
                ('array_dtype', C.c_int32), ('reserved', C.c_int32),
                ('value', C.c_double),
                ('data', C.c_void_p), ('ds_c', C.c_int64), ('ds_y', C.c_int64),
                ('array', C.c_void_p), ('as_c', C.c_int64), ('as_y', C.c_int64), ('as_x', C.c_int64)]


class MaskDesc(C.Structure):
    _fields_ = [('n_nodes', C.c_int32), ('reserved', C.c_int32), ('nodes', MaskNode * SC_MASK_MAX_NODES)]


class LibraryError(RuntimeError):
    pass


class _NoDevice(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def on_device_of(tensor):
    """Context manager making the CUDA device that holds `tensor` the current one: the library launches on the
    calling thread's current device, on that device's current stream."""
    import torch
    if tensor is None or not getattr(tensor, 'is_cuda', False):
        return _NoDevice()
    return torch.cuda.device(tensor.device)


_lib = None

_i64, _i32, _dbl, _vp, _sz = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_size_t
_pd = C.POINTER(C.c_double)
_pmask = C.POINTER(MaskDesc)

# name -> (restype, argtypes); every symbol include/sc_b200.h declares
SIGNATURES = {
    'sc_last_error': (C.c_char_p, []),
    'sc_version': (_i32, []),
    'sc_workspace_bytes': (_sz, [_i32, _i64, _i64, _i64, _i64]),
    'sc_launch_count': (_i64, []),
    'sc_mask_include': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _vp, _vp]),
    'sc_fill_masked': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl, _vp, _vp]),
    'sc_moments_axis0': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _pd, _dbl, _dbl, _i32,
                                _vp, _vp, _vp, _vp, _sz, _vp]),
    'sc_moment_central_axis0': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _pd, _vp, _i32,
                                       _vp, _vp, _sz, _vp]),
    'sc_moments_spatial': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _i32, _pmask, _vp, _dbl, _i32,
                                  _vp, _vp, _vp, _vp]),
    'sc_pixel_offsets': (_i32, [_pd, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
    'sc_moments_axis0_host': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _pd, _dbl, _dbl, _i32,
                                     _vp, _vp, _vp, _sz, _i32]),
    'sc_spectral_smooth': (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                  _pd, _i32, _i32, _vp, _sz, _vp]),
    'sc_smooth_moments_axis0': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl, _pd, _i32, _i32,
                                       _pd, _dbl, _dbl, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    'sc_spatial_smooth_sep': (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                     _pd, _i32, _pd, _i32, _vp, _vp, _i32, _i32, _vp, _sz, _vp]),
    'sc_spatial_smooth_sep_ex': (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                        _pd, _i32, _pd, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _sz, _vp]),
    'sc_spatial_missing_sample': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _vp, _vp]),
    'sc_spatial_smooth_2d': (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                    _pd, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _sz, _vp]),
    'sc_scale': (_i32, [_vp, _i32, _i64, _i64, _i64, _i64, _i64, _dbl, _i32, _vp, _vp]),
    'sc_reshard_scatter': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, C.POINTER(C.c_uint64), _i32, _i32, C.POINTER(C.c_int64),
                                  _i64, _i64, _vp]),
    'sc_mosaic_accumulate': (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _vp]),
    'sc_mosaic_normalize': (_i32, [_vp, _vp, _i64, _i64, _i64, _vp]),
    'sc_pack_filled_rows': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl, _i64, _i64, _vp, _vp]),
    'sc_spectral_interp': (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                  _pd, _pd, _i32, _dbl, _i32, _i32, _i32, _vp, _sz, _vp]),
    'sc_spectral_interp_scatter': (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                                          _pd, _pd, _i32, _dbl, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    'sc_reproject': (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                            _vp, _vp, _i32, _vp]),
    'sc_reproject_ex': (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _pmask, _dbl,
                               _vp, _vp, _i32, _vp]),
    'sc_wcs_pixel_map': (_i32, [_pd, _pd, _i64, _i64, _vp, _vp, _vp]),
    'sc_fits_decode': (_i32, [_vp, _vp, _i64, _i32, _dbl, _dbl, _i32, C.c_longlong, _vp]),
    'sc_reduce_axis0': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _pmask, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'sc_reduce_spatial': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _i32, _pmask, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'sc_synth_cube': (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, C.c_uint64, _vp, _i32, _i32, _vp]),
    'sc_last_kernel_ms': (C.c_float, [_i32]),
    'sc_enable_kernel_timing': (None, [_i32]),
}


def load():
    """Load the shared library (no GPU needed for this) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            "%s is missing: build it with `python -m spectral_cube_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise LibraryError("libsc_b200 error %d: %s" % (rc, load().sc_last_error().decode()))


def as_double_array(seq):
    import numpy as np
    a = np.ascontiguousarray(seq, dtype=np.float64)
    return a, a.ctypes.data_as(_pd)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise LibraryError("spectral_cube_b200 needs a CUDA device (B200, sm_100a); none is "
                           "visible and there is no CPU fallback")
    return torch
