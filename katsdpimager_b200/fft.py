"""cuFFT 2-D complex transforms (katsdpsigproc.fft surface used at reference
katsdpimager/image.py:585-600, 629, 698)."""
import ctypes
import enum
import weakref

import numpy as np

from . import _lib, accel
from .profiling import profile_device


class FftMode(enum.Enum):
    FORWARD = 0
    INVERSE = 1


def _destroy_plan(handle):
    try:
        _lib.load().kib_fft_plan2d_destroy(ctypes.c_void_p(handle))
    except Exception:
        pass


class FftTemplate:
    """A cuFFT plan for an unnormalised 2-D transform: complex-to-complex (the imaging
    transform, equal source and destination types and paddings, usually in place), or real
    to half-complex / half-complex to real (the restoring-beam convolution, reference
    beam.py:327-330; the complex side has shape height x (width // 2 + 1)).
    """

    def __init__(self, context, N, shape, dtype_src, dtype_dest,
                 padded_shape_src, padded_shape_dest, tuning=None):
        if N != 2 or len(shape) != 2:
            raise ValueError('only 2-D transforms are supported')
        dtype_src = np.dtype(dtype_src)
        dtype_dest = np.dtype(dtype_dest)
        self.context = context
        self.shape = tuple(shape)
        self.dtype_src = dtype_src
        self.dtype_dest = dtype_dest
        self.padded_shape_src = tuple(padded_shape_src)
        self.padded_shape_dest = tuple(padded_shape_dest)
        half = (shape[0], shape[1] // 2 + 1)
        self.shape_src = self.shape_dest = self.shape
        handle = ctypes.c_void_p()
        if dtype_src.kind == 'c' and dtype_dest.kind == 'c':
            if dtype_src != dtype_dest:
                raise ValueError('source and destination types must match')
            if tuple(padded_shape_src) != tuple(padded_shape_dest):
                raise ValueError('source and destination padding must match')
            if padded_shape_src[0] != shape[0]:
                raise ValueError('padding of the slow axis is not supported')
            _lib.call('kib_fft_plan2d_create', ctypes.byref(handle), shape[0], shape[1],
                      padded_shape_src[1], _lib.dtype_code(dtype_src))
        elif dtype_src.kind == 'f' and dtype_dest.kind == 'c':
            if padded_shape_src[0] != shape[0] or padded_shape_dest[0] != shape[0]:
                raise ValueError('padding of the slow axis is not supported')
            self.shape_dest = half
            _lib.call('kib_fft_plan2d_real_create', ctypes.byref(handle), shape[0], shape[1],
                      padded_shape_src[1], padded_shape_dest[1], 0, _lib.dtype_code(dtype_src))
        elif dtype_src.kind == 'c' and dtype_dest.kind == 'f':
            if padded_shape_src[0] != shape[0] or padded_shape_dest[0] != shape[0]:
                raise ValueError('padding of the slow axis is not supported')
            self.shape_src = half
            _lib.call('kib_fft_plan2d_real_create', ctypes.byref(handle), shape[0], shape[1],
                      padded_shape_dest[1], padded_shape_src[1], 1, _lib.dtype_code(dtype_dest))
        else:
            raise ValueError('unsupported combination of types')
        self.plan = handle
        self._finalizer = weakref.finalize(self, _destroy_plan, handle.value)

    def instantiate(self, command_queue, mode, allocator=None):
        return Fft(self, command_queue, mode, allocator)


class Fft(accel.Operation):
    """.. rubric:: Slots

    **src**, **dest** : complex arrays of the plan's shape (may be the same buffer)
    """

    def __init__(self, template, command_queue, mode, allocator=None):
        super().__init__(command_queue, allocator)
        self.template = template
        self.mode = mode
        src_dims = [accel.Dimension(s, min_padded_size=p, exact=(s == p))
                    for s, p in zip(template.shape_src, template.padded_shape_src)]
        dest_dims = [accel.Dimension(s, min_padded_size=p, exact=(s == p))
                     for s, p in zip(template.shape_dest, template.padded_shape_dest)]
        self.slots['src'] = accel.IOSlot(src_dims, template.dtype_src)
        self.slots['dest'] = accel.IOSlot(dest_dims, template.dtype_dest)

    def _run(self):
        src = self.buffer('src')
        dest = self.buffer('dest')
        if src.padded_shape != self.template.padded_shape_src \
                or dest.padded_shape != self.template.padded_shape_dest:
            raise ValueError('buffer padding does not match the FFT plan')
        with profile_device(self.command_queue, 'fft'):
            _lib.call('kib_fft_plan2d_exec', self.template.plan, src.ptr, dest.ptr,
                      self.mode.value, self.command_queue.stream)
