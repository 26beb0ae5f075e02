"""ctypes binding of libkatimager_b200.so (the C ABI in include/katimager_b200.h).

There is deliberately no fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
katsdpimager_b200/csrc``) importing any operation raises, and if no CUDA device
is usable the first runtime call raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint32, c_ulonglong, c_void_p)

LIB_NAME = 'libkatimager_b200.so'
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

_vp = c_void_p
_i = c_int
_i64 = c_int64
_d = c_double
_f = c_float
_sz = c_size_t

#: name -> argument types; every entry point returns int (0 = success) except
#: kib_version / kib_last_error.  Kept in the order of include/katimager_b200.h.
SIGNATURES = {
    'kib_device_count': [POINTER(c_int)],
    'kib_set_device': [_i],
    'kib_get_device': [POINTER(c_int)],
    'kib_device_name': [_i, c_char_p, _i],
    'kib_device_attr': [_i, _i, POINTER(c_int64)],
    'kib_mem_info': [POINTER(c_size_t), POINTER(c_size_t)],
    'kib_stream_create': [POINTER(c_void_p)],
    'kib_stream_destroy': [_vp],
    'kib_stream_sync': [_vp],
    'kib_stream_wait_event': [_vp, _vp],
    'kib_event_create': [POINTER(c_void_p)],
    'kib_event_record': [_vp, _vp],
    'kib_event_sync': [_vp],
    'kib_event_query': [_vp, POINTER(c_int)],
    'kib_event_elapsed_ms': [_vp, _vp, POINTER(c_float)],
    'kib_event_destroy': [_vp],
    'kib_malloc': [POINTER(c_void_p), _sz],
    'kib_free': [_vp],
    'kib_host_alloc': [POINTER(c_void_p), _sz],
    'kib_host_free': [_vp],
    'kib_host_register': [_vp, _sz],
    'kib_host_unregister': [_vp],
    'kib_memset_async': [_vp, _i, _sz, _vp],
    'kib_memcpy_h2d_async': [_vp, _vp, _sz, _vp],
    'kib_memcpy_d2h_async': [_vp, _vp, _sz, _vp],
    'kib_memcpy_d2d_async': [_vp, _vp, _sz, _vp],
    'kib_memcpy3d_async': [_vp, _sz, _sz, _vp, _sz, _sz, _sz, _sz, _sz, _i, _vp],
    'kib_fft_plan2d_create': [POINTER(c_void_p), _i, _i, _i, _i],
    'kib_fft_plan2d_real_create': [POINTER(c_void_p), _i, _i, _i, _i, _i, _i],
    'kib_fft_plan1d_create': [POINTER(c_void_p), _i, _i64, _i64, _i, _i],
    'kib_fft_plan2d_exec': [_vp, _vp, _vp, _i, _vp],
    'kib_fft_plan2d_destroy': [_vp],
    'kib_grid': [_vp, _i, _i64, _i, _i,
                 _vp, _i, _i64,
                 _vp, _vp, _vp,
                 _vp, _i, _i,
                 _i, _i, _i, _i,
                 _i64, _vp, _vp],
    'kib_degrid': [_vp, _i, _i64, _i, _i,
                   _vp, _vp, _vp, _vp,
                   _vp, _i, _i,
                   _i, _i, _i, _i,
                   _i64, _vp, _vp],
    'kib_grid_to_layer': [_vp, _i, _i, _vp, _i, _i, _i, _vp],
    'kib_layer_to_grid': [_vp, _i, _i, _vp, _i, _i, _i, _vp],
    'kib_layer_to_image': [_vp, _i, _vp, _i, _i, _vp, _d, _d, _d, _i, _vp],
    'kib_image_to_layer': [_vp, _i, _vp, _i, _i, _vp, _d, _d, _d, _i, _vp],
    'kib_grid_to_image': [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _d, _d, _d, _i, _vp],
    'kib_grid_to_image_supported': [_i, _i, _i],
    'kib_grid_to_image_columns': [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp],
    'kib_grid_to_image_fold_bytes': [_i, _i, POINTER(c_int64)],
    'kib_grid_to_image_columns_kernels': [_i],
    'kib_image_to_grid_rows': [_vp, _i, _i, _i, _vp, _i, _vp, _d, _d, _d, _vp, _i, _i, _vp],
    'kib_image_to_grid_columns': [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp],
    'kib_image_to_grid_sparse_supported': [_i, _i, _i],
    'kib_image_to_grid_rows_sparse': [_vp, _i, _i, _i, _vp, _i, _vp, _d, _d, _d, _vp, _i, _vp],
    'kib_image_to_grid_rows_classified': [_vp, _i, _i, _i, _vp, _i, _vp, _d, _d, _d, _vp, _i, _vp],
    'kib_image_to_grid_columns_sparse': [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp],
    'kib_grid_to_image_rows': [_vp, _i, _vp, _i, _i, _i, _vp, _d, _d, _d, _vp, _i, _i, _vp],
    'kib_column_occupancy': [_vp, c_int64, c_int64, _i, _i, _vp, _vp],
    'kib_row_presence': [_vp, _i, _i, _vp, _vp],
    'kib_clear_columns': [_vp, _i, c_int64, _i, _i, _vp, _i, _vp],
    'kib_grid_to_image_columns_occ': [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp],
    'kib_grid_to_image_rows_occ': [_vp, _i, _vp, _i, _i, _i, _vp, _d, _d, _d, _vp, _i, _vp, _i, _vp],
    'kib_image_to_grid_columns_occ': [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp],
    'kib_scale': [_vp, _i, _i64, _i, _i, _i, POINTER(c_double), _i, _vp],
    'kib_add_image': [_vp, _i, _i64, _vp, _i, _i64, _i, _i, _i, _i, _vp],
    'kib_apply_primary_beam': [_vp, _i, _i64, _vp, _i, _i, _i, _d, _d, _i, _vp],
    'kib_update_tiles': [_vp, _i, _i64, _i, _i, _i, _i, _i,
                         _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'kib_find_peak': [_vp, _i, _i64, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    'kib_subtract_psf': [_vp, _vp, _i, _i64, _i, _i, _i,
                         _vp, _i, _i64, _i, _i, _i, _i,
                         _vp, _i, _i, _d, _i, _vp],
    'kib_clean_minor_cycles': [_vp, _vp, _i, _i64, _i, _i, _i, _i, _i,
                               _vp, _i, _i64, _i, _i, _i, _i,
                               _vp, _vp, _i, _i, _i,
                               _vp, _vp, _vp,
                               _d, _d, _i,
                               _vp, _i, _vp, _vp, _i, _vp],
    'kib_clean_minor_cycles_launches': [_i, _i],
    'kib_psf_patch': [_vp, _i, _i64, _i, _i, _i, _i, _i, _i, _i, _d, _vp, _i, _vp],
    'kib_abs_histogram': [_vp, _i, _i64, _i, _i, _i, _i, c_uint32, _i, _i, _i, _vp, _i, _vp],
    'kib_abs_histogram_window': [_vp, _i, _i64, _i, _i, _i, _i, c_uint32, _i, _i, _i, _i, _vp, _vp,
                                 _i, _vp],
    'kib_rank': [_vp, _i, _i64, _i, _i, _i, _i, _d, _vp, _i, _vp],
    'kib_grid_weights': [_vp, _i, _i64, _i, _i, _vp, _vp, _i, _i64, _vp],
    'kib_mean_weight': [_vp, _i, _i, _i, _vp, _vp],
    'kib_density_weights': [_vp, _i, _i64, _i, _i, _i, _f, _f, _vp, _vp],
    'kib_fourier_beam': [_vp, _i, _d, _d, _d, _d, _i, _i, _i, _vp],
    'kib_fits_plane': [_vp, _vp, _i, _i64, _i, _i, _i, _i, _vp],
    'kib_fill': [_vp, _i, _i64, _i, _i, _i, _d, _i, _vp],
    'kib_unpack_records': [_vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp],
    'kib_predict': [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _f, _f, _vp],
    'kib_preprocess_scratch_bytes': [_i64, _i, _i, POINTER(c_int64)],
    'kib_preprocess': [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _d, _d, _i, _i, _i, _i64,
                       _vp, POINTER(c_int64), _vp, _i64, _vp],
    'kib_fp32_peak_kernel': [_vp, _i, _i, POINTER(c_double), _vp],
}

F32 = 0
F64 = 1


class KibError(RuntimeError):
    """A call into libkatimager_b200.so failed."""


_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            '{} has not been built; run `python -c "import __graft_entry__ as g; g.build()"` '
            'in the repository root. There is no CPU fallback.'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    lib.kib_version.restype = c_int
    lib.kib_version.argtypes = []
    lib.kib_last_error.restype = c_char_p
    lib.kib_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().kib_last_error()
        raise KibError('{} (code {})'.format(msg.decode('utf-8', 'replace') if msg else 'error', rc))


#: entry points that launch exactly one of this library's kernels
_ONE_KERNEL = frozenset([
    'kib_grid', 'kib_degrid', 'kib_grid_to_layer', 'kib_layer_to_grid', 'kib_layer_to_image',
    'kib_image_to_layer', 'kib_scale', 'kib_add_image', 'kib_apply_primary_beam',
    'kib_update_tiles', 'kib_find_peak', 'kib_subtract_psf', 'kib_psf_patch',
    'kib_abs_histogram', 'kib_abs_histogram_window', 'kib_rank', 'kib_grid_weights', 'kib_mean_weight',
    'kib_density_weights', 'kib_fill', 'kib_fits_plane', 'kib_fourier_beam', 'kib_predict', 'kib_fp32_peak_kernel',
    'kib_unpack_records', 'kib_grid_to_image_rows', 'kib_image_to_grid_rows',
    'kib_grid_to_image_rows_occ', 'kib_column_occupancy', 'kib_row_presence',
    'kib_clear_columns', 'kib_image_to_grid_rows_classified'])

#: number of hand-written kernels launched through this module (cuFFT and memset/memcpy
#: are not counted); bench.py reports the difference over its timed region
kernel_launches = 0


def call(name, *args):
    """Invoke entry point `name`, raising :class:`KibError` on failure."""
    global kernel_launches
    check(getattr(load(), name)(*args))
    if name in _ONE_KERNEL:
        kernel_launches += 1
    elif name in ('kib_grid_to_image_columns', 'kib_grid_to_image_columns_occ'):
        kernel_launches += load().kib_grid_to_image_columns_kernels(int(args[2]))
    elif name in ('kib_image_to_grid_columns', 'kib_image_to_grid_columns_occ'):
        kernel_launches += load().kib_grid_to_image_columns_kernels(int(args[5]))
    elif name == 'kib_image_to_grid_rows_sparse':
        kernel_launches += 2                 # row classification + transforms of non-empty rows
    elif name == 'kib_image_to_grid_columns_sparse':
        kernel_launches += 1
    elif name == 'kib_grid_to_image':
        kernel_launches += 1 + load().kib_grid_to_image_columns_kernels(int(args[8]))
    elif name == 'kib_preprocess':
        kernel_launches += 4                 # quantise, heads, merge, gather (+ cub scan / sort)
    elif name == 'kib_clean_minor_cycles':
        kernel_launches += load().kib_clean_minor_cycles_launches(int(args[26]), int(args[31]))


def grid_to_image_fold_bytes(size, grid_size):
    """Bytes of fold scratch kib_grid_to_image_columns needs."""
    nbytes = c_int64()
    call('kib_grid_to_image_fold_bytes', int(size), int(grid_size), ctypes.byref(nbytes))
    return int(nbytes.value)


def grid_to_image_supported(size, grid_size, dtype):
    """Whether the fused pruned transform (kib_grid_to_image) covers this case."""
    return bool(load().kib_grid_to_image_supported(int(size), int(grid_size), dtype_code(dtype)))


def dtype_code(dtype):
    import numpy as np
    dtype = np.dtype(dtype)
    if dtype in (np.float32, np.complex64):
        return F32
    if dtype in (np.float64, np.complex128):
        return F64
    raise TypeError('dtype {} is not supported'.format(dtype))
