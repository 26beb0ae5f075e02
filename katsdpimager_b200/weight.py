"""Imaging (density) weights on B200: natural, uniform and robust weighting.

Same surface as the reference's :mod:`katsdpimager.weight` (reference
weight.py:55-538; see its module docstring and [Bri95] for the definitions):
statistical weights are summed per UV cell without convolution, robust weighting
derives S^2 = (5 * 10^-R)^2 / (sum W^2 / sum W) from polarization 0, and each cell
becomes 1 / (a W + b) (0 where empty).  Kernels: csrc/kib_weight.cu.
"""
import enum

import numpy as np

from . import _lib, accel, fill
from .profiling import profile_device


class WeightType(enum.Enum):
    NATURAL = 0
    UNIFORM = 1
    ROBUST = 2


def _strides(array):
    return array.padded_shape[2], array.padded_shape[1] * array.padded_shape[2]


class GridWeightsTemplate:
    """Accumulate statistical weights onto a grid (reference weight.py:61-87)."""

    def __init__(self, context, num_polarizations, tuning=None):
        _lib.load()
        self.context = context
        self.num_polarizations = num_polarizations

    def instantiate(self, *args, **kwargs):
        return GridWeights(self, *args, **kwargs)


class GridWeights(accel.Operation):
    """.. rubric:: Slots

    **uv** : int16, max_vis x 4 (only the first two of each four are used; (0, 0) is
    the grid centre);  **weights** : float32, max_vis x polarizations;
    **grid** : float32, polarizations x height x width
    """

    def __init__(self, template, command_queue, grid_shape, max_vis, allocator=None):
        super().__init__(command_queue, allocator)
        self.template = template
        if grid_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        if grid_shape[1] % 2 or grid_shape[2] % 2:
            raise ValueError('Odd-sized grid not currently supported')
        self.max_vis = max_vis
        self.slots['grid'] = accel.IOSlot(grid_shape, np.float32)
        self.slots['uv'] = accel.IOSlot((max_vis, accel.Dimension(4, exact=True)), np.int16)
        self.slots['weights'] = accel.IOSlot(
            (max_vis, accel.Dimension(template.num_polarizations, exact=True)), np.float32)
        self._num_vis = 0

    @property
    def num_vis(self):
        return self._num_vis

    @num_vis.setter
    def num_vis(self, n):
        if n < 0 or n > self.max_vis:
            raise ValueError('Number of visibilities {} is out of range 0..{}'.format(
                n, self.max_vis))
        self._num_vis = n

    def _run(self):
        grid = self.buffer('grid')
        row_stride, pol_stride = _strides(grid)
        with profile_device(self.command_queue, 'grid_weights'):
            _lib.call('kib_grid_weights', grid.ptr, row_stride, pol_stride,
                      grid.shape[2], grid.shape[1], self.buffer('uv').ptr,
                      self.buffer('weights').ptr, grid.shape[0], self._num_vis,
                      self.command_queue.stream)

    def parameters(self):
        return {'num_polarizations': self.template.num_polarizations, 'max_vis': self.max_vis}


class DensityWeightsTemplate:
    """Statistical weight sums -> density weights (reference weight.py:186-214)."""

    def __init__(self, context, num_polarizations, tuning=None):
        _lib.load()
        self.context = context
        self.num_polarizations = num_polarizations

    def instantiate(self, *args, **kwargs):
        return DensityWeights(self, *args, **kwargs)


class DensityWeights(accel.Operation):
    """In-place W -> 1 / (a W + b); returns (rms, normalized rms) per equations 3.3 and
    3.5 of [Bri95] (reference weight.py:217-284).

    .. rubric:: Slots

    **grid** : float32, polarizations x height x width;  **sums** : float64[3] scratch
    """

    def __init__(self, template, command_queue, grid_shape, allocator=None):
        super().__init__(command_queue, allocator)
        self.template = template
        if grid_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        self.a = 1.0
        self.b = 0.0
        self.slots['grid'] = accel.IOSlot(grid_shape, np.float32)
        self.slots['sums'] = accel.IOSlot((3,), np.float64)
        self._sums_host = accel.HostArray((3,), np.float64, context=command_queue.context)

    def _run(self):
        grid = self.buffer('grid')
        sums = self.buffer('sums')
        sums.zero(self.command_queue)
        row_stride, pol_stride = _strides(grid)
        with profile_device(self.command_queue, 'density_weights'):
            _lib.call('kib_density_weights', grid.ptr, row_stride, pol_stride,
                      grid.shape[2], grid.shape[1], grid.shape[0],
                      float(self.a), float(self.b), sums.ptr, self.command_queue.stream)
        sums.get(self.command_queue, self._sums_host)
        sum_w, sum_dw, sum_d2w = (float(x) for x in self._sums_host)
        rms = np.sqrt(sum_d2w) / sum_dw
        return rms, rms * np.sqrt(sum_w)

    def parameters(self):
        return {'a': self.a, 'b': self.b,
                'num_polarizations': self.template.num_polarizations}


class MeanWeightTemplate:
    """Mean weight of equation 3.17 of [Bri95] (reference weight.py:296-323)."""

    def __init__(self, context, tuning=None):
        _lib.load()
        self.context = context

    def instantiate(self, *args, **kwargs):
        return MeanWeight(self, *args, **kwargs)


class MeanWeight(accel.Operation):
    """Returns sum W^2 / sum W over the cells of polarization 0.

    .. rubric:: Slots

    **grid** : float32, polarizations x height x width;  **sums** : float64[2] scratch
    """

    def __init__(self, template, command_queue, grid_shape, allocator=None):
        super().__init__(command_queue, allocator)
        self.template = template
        self.slots['grid'] = accel.IOSlot(grid_shape, np.float32)
        self.slots['sums'] = accel.IOSlot((2,), np.float64)
        self._sums_host = accel.HostArray((2,), np.float64, context=command_queue.context)

    def _run(self):
        grid = self.buffer('grid')
        sums = self.buffer('sums')
        self.command_queue.enqueue_zero_buffer(sums.buffer)
        with profile_device(self.command_queue, 'mean_weight'):
            _lib.call('kib_mean_weight', grid.ptr, grid.padded_shape[2],
                      grid.shape[2], grid.shape[1], sums.ptr, self.command_queue.stream)
        sums.get(self.command_queue, self._sums_host)
        return float(self._sums_host[1] / self._sums_host[0])


class WeightsTemplate:
    """Compound template for computing imaging weights (reference weight.py:379-416)."""

    def __init__(self, context, weight_type, num_polarizations,
                 grid_weights_tuning=None, mean_weight_tuning=None,
                 density_weights_tuning=None):
        self.context = context
        self.weight_type = weight_type
        self.grid_weights = None
        self.mean_weight = None
        self.density_weights = None
        self.fill = None
        if weight_type == WeightType.NATURAL:
            self.fill = fill.FillTemplate(context, np.float32, 'float')
        else:
            self.grid_weights = GridWeightsTemplate(context, num_polarizations,
                                                    tuning=grid_weights_tuning)
            if weight_type == WeightType.ROBUST:
                self.mean_weight = MeanWeightTemplate(context, tuning=mean_weight_tuning)
            self.density_weights = DensityWeightsTemplate(context, num_polarizations,
                                                          tuning=density_weights_tuning)

    def instantiate(self, *args, **kwargs):
        return Weights(self, *args, **kwargs)


class Weights(accel.OperationSequence):
    """Instantiation of :class:`WeightsTemplate` (reference weight.py:419-538):
    ``clear()``, then ``grid(N)`` per batch of N weights placed in the **uv** /
    **weights** slots, then ``finalize()`` -> (rms, normalized rms).

    The **uv** and **weights** slots are absent for natural weighting.
    """

    def __init__(self, template, command_queue, grid_shape, max_vis, allocator=None):
        self.template = template
        operations = []
        compounds = {'grid': []}
        self._grid_weights = self._fill = self._mean_weight = self._density_weights = None
        self.robustness = None
        if template.grid_weights is not None:
            self._grid_weights = template.grid_weights.instantiate(
                command_queue, grid_shape, max_vis, allocator)
            operations.append(('grid_weights', self._grid_weights))
            compounds['grid'].append('grid_weights:grid')
            compounds['uv'] = ['grid_weights:uv']
            compounds['weights'] = ['grid_weights:weights']
        if template.fill is not None:
            self._fill = template.fill.instantiate(command_queue, grid_shape, allocator)
            self._fill.set_value(1)
            operations.append(('fill', self._fill))
            compounds['grid'].append('fill:data')
        if template.mean_weight is not None:
            self._mean_weight = template.mean_weight.instantiate(
                command_queue, grid_shape, allocator)
            operations.append(('mean_weight', self._mean_weight))
            compounds['grid'].append('mean_weight:grid')
            self.robustness = 0.0
        if template.density_weights is not None:
            self._density_weights = template.density_weights.instantiate(
                command_queue, grid_shape, allocator)
            operations.append(('density_weights', self._density_weights))
            compounds['grid'].append('density_weights:grid')
        super().__init__(command_queue, operations, compounds, allocator=allocator)

    def _run(self):
        raise NotImplementedError('Weights should not be used as a callable')

    def clear(self):
        self.ensure_all_bound()
        if self._fill is None:      # natural weights are simply overwritten in finalize
            self.buffer('grid').zero(self.command_queue)

    def grid(self, N):
        self.ensure_all_bound()
        if self._grid_weights is not None:
            self._grid_weights.num_vis = N
            return self._grid_weights()

    def finalize(self):
        self.ensure_all_bound()
        if self._mean_weight is not None:
            mean_weight = self._mean_weight()
            self._density_weights.a = (5 * 10**(-self.robustness))**2 / mean_weight
            self._density_weights.b = 1.0
        if self._density_weights is not None:
            rms, normalized_rms = self._density_weights()
        else:
            rms, normalized_rms = None, 1.0
        if self._fill is not None:
            self._fill()
        return rms, normalized_rms
