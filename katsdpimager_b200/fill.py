"""Fill an array with a constant (katsdpsigproc.fill.FillTemplate as used at
reference katsdpimager/weight.py:403, 480-484)."""
import numpy as np

from . import _lib, accel


class FillTemplate:
    def __init__(self, context, dtype, ctype=None, tuning=None):
        self.context = context
        self.dtype = np.dtype(dtype)

    def instantiate(self, command_queue, shape, allocator=None):
        return Fill(self, command_queue, shape, allocator)


class Fill(accel.Operation):
    """.. rubric:: Slots

    **data** : array of any shape (up to 3 dimensions)
    """

    def __init__(self, template, command_queue, shape, allocator=None):
        super().__init__(command_queue, allocator)
        if not 1 <= len(shape) <= 3:
            raise ValueError('Fill supports 1 to 3 dimensions')
        self.template = template
        self.shape = tuple(shape)
        self.slots['data'] = accel.IOSlot(shape, template.dtype)
        self.value = template.dtype.type(0)

    def set_value(self, value):
        self.value = self.template.dtype.type(value)

    def _run(self):
        data = self.buffer('data')
        shape = (1,) * (3 - len(data.shape)) + data.shape
        padded = (1,) * (3 - len(data.shape)) + data.padded_shape
        _lib.call('kib_fill', data.ptr, padded[2], padded[1] * padded[2],
                  shape[2], shape[1], shape[0], float(self.value),
                  _lib.dtype_code(data.dtype), self.command_queue.stream)
