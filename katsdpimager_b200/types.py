"""dtype helpers (reference katsdpimager/types.py:6-44)."""
import numpy as np


def real_to_complex(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return np.dtype(np.complex64)
    if dtype == np.float64:
        return np.dtype(np.complex128)
    raise ValueError('Unrecognised dtype {}'.format(dtype))


def complex_to_real(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.complex64:
        return np.dtype(np.float32)
    if dtype == np.complex128:
        return np.dtype(np.float64)
    raise ValueError('Unrecognised dtype {}'.format(dtype))
