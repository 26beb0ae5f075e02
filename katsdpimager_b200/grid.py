"""Convolutional gridding and degridding with W projection on B200.

Same public surface as the reference's :mod:`katsdpimager.grid` (``GridderTemplate``
/ ``Gridder``, ``DegridderTemplate`` / ``Degridder``, ``VisOperation``,
``ConvolutionKernel``; reference grid.py:344-1029) so that ``imaging.py`` can use it
unchanged; the device work is done by ``kib_grid`` / ``kib_degrid`` in
libkatimager_b200.so (csrc/kib_grid.cu, csrc/kib_degrid.cu).

The separable anti-aliasing x W convolution kernel is evaluated on the host in
double precision exactly as the reference does (grid.py:235-334, 358-389) and
uploaded once per channel as a ``complex64[w_planes][oversample][kernel_width]``
look-up table.  Unlike the reference no zero padding of the table is needed (the
sm_100a gridder assigns footprint cells to threads cyclically, not in aligned
bins), so ``pad`` is 0 and ``bin_size == kernel_width``.
"""
import math

import numpy as np

from . import _lib, accel
from .fast_math import expj2pi
from .profiling import profile_device, profile_function
from .types import real_to_complex  # noqa: F401  (re-exported for API parity)


def kaiser_bessel(x, width, beta):
    """Kaiser-Bessel window with support [-width/2, width/2] (grid.py:136-155)."""
    x = np.asarray(x, dtype=np.float64)
    inside = 1 - (2 * x / width) ** 2
    values = np.i0(beta * np.sqrt(np.clip(inside, 0, None))) / np.i0(beta)
    return np.where(inside >= 0, values, 0.0)


def kaiser_bessel_fourier(f, width, beta, out=None):
    """Continuous Fourier transform of :func:`kaiser_bessel` (grid.py:158-184):
    ``width / I0(beta) * sinc(sqrt((width f)^2 - (beta/pi)^2))`` with the usual
    analytic continuation (sinc of an imaginary argument) inside the main lobe."""
    f = np.asarray(f, dtype=np.float64)
    arg2 = (width * f) ** 2 - (beta / math.pi) ** 2
    # sinc(sqrt(t)) = sin(pi sqrt(t)) / (pi sqrt(t)); for t < 0 it is sinh(pi sqrt(-t)) / (pi sqrt(-t)).
    # np.sinc on a complex argument handles both branches.
    ans = width / np.i0(beta) * np.sinc(np.lib.scimath.sqrt(arg2)).real
    if out is not None:
        out[:] = ans
        return out
    return ans


def antialias_beta(antialias_width):
    """Kaiser-Bessel shape parameter: first null of the taper just outside the
    image (grid.py:373-378)."""
    return 1.2 * math.pi * math.sqrt(0.25 * antialias_width ** 2 - 1.0)


def antialias_w_kernel(cell_wavelengths, w, width, oversample, antialias_width,
                       image_oversample, beta, out=None):
    """Combined anti-aliasing and W-projection kernel (grid.py:235-334).

    The image-plane function ``aa(l) * exp(2 pi i (-w (-l^2/2 - 5 l^4/24) + shift l))``
    is sampled on ``width * oversample * image_oversample`` points, transformed to
    the UV plane, cropped to ``width * oversample`` samples about DC and split into
    `oversample` sub-pixel phases of `width` taps (sub-pixel index reversed, because
    it is the visibility's offset rather than the tap's).

    Returns complex64 array of shape ``(len(w), oversample, width)``.
    """
    w = np.atleast_1d(np.asarray(w, dtype=np.float64))
    taps = oversample * width
    if taps % 2:
        raise ValueError('oversample * width must be even')
    samples = taps * image_oversample
    step = 1.0 / (width * cell_wavelengths * image_oversample)
    l = (np.arange(samples) - samples // 2) * step
    l2 = l * l
    envelope = cell_wavelengths * kaiser_bessel_fourier(l * cell_wavelengths, antialias_width, beta)
    w_term = -0.5 * l2 - (5.0 / 24.0) * (l2 * l2)
    half_subpixel = -0.5 * cell_wavelengths / oversample
    phase = np.outer(-w, w_term) + half_subpixel * l
    image_plane = envelope * expj2pi(phase)
    uv_plane = np.fft.fft(np.fft.ifftshift(image_plane, axes=-1), axis=-1) * step
    centred = np.concatenate((uv_plane[:, -(taps // 2):], uv_plane[:, :taps // 2]), axis=-1)
    # centred[w, tap * oversample + s]  ->  lut[w, oversample - 1 - s, tap]
    lut = centred.reshape(len(w), width, oversample)[:, :, ::-1].transpose(0, 2, 1)
    if out is None:
        out = np.empty(lut.shape, np.complex64)
    out[:] = lut
    return out


def subpixel_coord(x, oversample):
    """Cell and sub-cell index of a coordinate in cells (grid.py:337-341)."""
    xs = int(math.floor(x * oversample))
    return xs // oversample, xs % oversample


class ConvolutionKernel:
    """Host copy of the per-channel kernel look-up table (grid.py:344-423)."""

    def __init__(self, image_parameters, grid_parameters, data=None):
        self.grid_parameters = grid_parameters
        fixed = grid_parameters.fixed
        shape = (grid_parameters.w_planes, fixed.oversample, fixed.kernel_width)
        self.data = np.empty(shape, np.complex64) if data is None else data
        cell_wavelengths = float(image_parameters.cell_size / image_parameters.wavelength)
        slice_spacing = float(fixed.max_w / (grid_parameters.w_slices * image_parameters.wavelength))
        plane_spacing = slice_spacing / grid_parameters.w_planes
        self.beta = antialias_beta(fixed.antialias_width)
        # planes are centred on w = +-(slice - plane) / 2 within the slice
        extreme = 0.5 * (slice_spacing - plane_spacing)
        ws = np.linspace(-extreme, extreme, grid_parameters.w_planes)
        antialias_w_kernel(cell_wavelengths, ws, fixed.kernel_width, fixed.oversample,
                           fixed.antialias_width, fixed.image_oversample, self.beta,
                           out=self.data)

    def taper(self, N, out=None):
        """Image-plane taper of the gridding kernel for an N-pixel image, including
        the sinc from piecewise-constant sub-pixel sampling (grid.py:404-423)."""
        x = np.arange(N) / N - 0.5
        fixed = self.grid_parameters.fixed
        values = kaiser_bessel_fourier(x, fixed.antialias_width, self.beta)
        values *= np.sinc(x / fixed.oversample)
        if out is not None:
            out[:] = values
            return out
        return values


class ConvolutionKernelDevice(ConvolutionKernel):
    """:class:`ConvolutionKernel` whose table also lives in device memory."""

    def __init__(self, context, image_parameters, grid_parameters, pad=0, allocator=None,
                 command_queue=None):
        if allocator is None:
            allocator = accel.DeviceAllocator(context)
        fixed = grid_parameters.fixed
        device = allocator.allocate(
            (grid_parameters.w_planes, fixed.oversample, fixed.kernel_width + 2 * pad),
            np.complex64)
        host = device.empty_like()
        host.fill(0)
        super().__init__(image_parameters, grid_parameters,
                         host[:, :, pad:pad + fixed.kernel_width])
        queue = command_queue if command_queue is not None else context.create_command_queue()
        device.set(queue, host)
        self.padded_data = device
        self.pad = pad

    @property
    def bin_size(self):
        return self.data.shape[-1] + self.pad


class GridderTemplate:
    """Gridding for one (precision, polarization count, kernel geometry).

    `tuning` is accepted for API compatibility; the kernel shape is selected inside
    the library from the kernel width and polarization count.
    """

    autotune_version = 0

    def __init__(self, context, fixed_image_parameters, fixed_grid_parameters, tuning=None):
        _lib.load()
        self.context = context
        self.fixed_image_parameters = fixed_image_parameters
        self.fixed_grid_parameters = fixed_grid_parameters
        self.kernel_pad = 0
        if not 1 <= len(fixed_image_parameters.polarizations) <= 4:
            raise ValueError('between 1 and 4 polarizations are supported')

    def instantiate(self, *args, **kwargs):
        return Gridder(self, *args, **kwargs)


class VisOperation(accel.Operation):
    """Operations that hold preprocessed visibilities in device buffers
    (grid.py:656-703).

    .. rubric:: Slots

    **uv** : int16, max_vis x 4 -- grid cell (u, v) then sub-cell (u, v)
    **w_plane** : int16, max_vis
    **vis** : complex64, max_vis x polarizations, pre-multiplied by statistical weights
    """

    def __init__(self, command_queue, num_polarizations, max_vis, allocator=None):
        super().__init__(command_queue, allocator)
        self.max_vis = max_vis
        self.slots['uv'] = accel.IOSlot((max_vis, accel.Dimension(4, exact=True)), np.int16)
        self.slots['w_plane'] = accel.IOSlot((max_vis,), np.int16)
        self.slots['vis'] = accel.IOSlot(
            (max_vis, accel.Dimension(num_polarizations, exact=True)), np.complex64)
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


class GridDegrid(VisOperation):
    """Common part of :class:`Gridder` and :class:`Degridder` (grid.py:706-773).

    Adds the **grid** slot (complex, polarizations x G x G with the DC cell at
    G/2) sized from the longest baseline exactly as the reference does, and owns
    the convolution kernel table.
    """

    def __init__(self, template, command_queue, array_parameters,
                 image_parameters, grid_parameters, max_vis, allocator=None):
        assert image_parameters.fixed == template.fixed_image_parameters
        assert grid_parameters.fixed == template.fixed_grid_parameters
        num_polarizations = len(image_parameters.fixed.polarizations)
        super().__init__(command_queue, num_polarizations, max_vis, allocator)
        self.convolve_kernel = ConvolutionKernelDevice(
            template.context, image_parameters, grid_parameters, template.kernel_pad,
            command_queue=command_queue)
        max_uv = float(array_parameters.longest_baseline / image_parameters.cell_size)
        kernel_size = self.convolve_kernel.padded_data.shape[-1]
        grid_pixels = 2 * (int(max_uv) + kernel_size // 2 + 1)
        if grid_pixels > image_parameters.pixels:
            raise ValueError('image_oversample is too small '
                             'to capture all visibilities in the UV plane')
        self.template = template
        self.image_parameters = image_parameters
        self.grid_parameters = grid_parameters
        self.slots['grid'] = accel.IOSlot(
            (num_polarizations, grid_pixels, grid_pixels), image_parameters.fixed.complex_dtype)
        # Count of visibilities skipped because their footprint left the grid
        self._rejected = accel.DeviceArray(template.context, (1,), np.int32)
        self._rejected.zero(command_queue)

    def parameters(self):
        return {'grid_parameters': self.grid_parameters,
                'image_parameters': self.image_parameters}

    def num_rejected(self):
        """Visibilities skipped so far for falling outside the grid (blocking)."""
        return int(self._rejected.get(self.command_queue)[0])

    def _lut_args(self):
        lut = self.convolve_kernel.padded_data
        fixed = self.grid_parameters.fixed
        return (lut.ptr, lut.padded_shape[2], self.convolve_kernel.pad,
                self.grid_parameters.w_planes, fixed.oversample, fixed.kernel_width)


class Gridder(GridDegrid):
    """Instantiation of :class:`GridderTemplate` (grid.py:776-867).

    Additional slot **weights_grid** (float32, same shape as **grid**): density
    weights looked up at each visibility's cell.
    """

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.slots['weights_grid'] = accel.IOSlot(self.slots['grid'].shape, np.float32)

    def _run(self):
        if self.num_vis == 0:
            return
        grid = self.buffer('grid')
        weights_grid = self.buffer('weights_grid')
        with profile_device(self.command_queue, 'grid'):
            _lib.call(
                'kib_grid',
                grid.ptr, grid.padded_shape[2], grid.padded_shape[1] * grid.padded_shape[2],
                grid.shape[2], _lib.dtype_code(grid.dtype),
                weights_grid.ptr, weights_grid.padded_shape[2],
                weights_grid.padded_shape[1] * weights_grid.padded_shape[2],
                self.buffer('uv').ptr, self.buffer('w_plane').ptr, self.buffer('vis').ptr,
                *self._lut_args(),
                len(self.image_parameters.fixed.polarizations),
                self.num_vis, self._rejected.ptr, self.command_queue.stream)


class DegridderTemplate:
    autotune_version = 0

    def __init__(self, context, fixed_image_parameters, fixed_grid_parameters, tuning=None):
        _lib.load()
        self.context = context
        self.fixed_image_parameters = fixed_image_parameters
        self.fixed_grid_parameters = fixed_grid_parameters
        self.kernel_pad = 0
        if not 1 <= len(fixed_image_parameters.polarizations) <= 4:
            raise ValueError('between 1 and 4 polarizations are supported')

    def instantiate(self, *args, **kwargs):
        return Degridder(self, *args, **kwargs)


class Degridder(GridDegrid):
    """Instantiation of :class:`DegridderTemplate` (grid.py:973-1029): subtracts the
    visibilities predicted from **grid** (scaled by **weights**) from **vis**.

    Additional slot **weights** (float32, max_vis x polarizations).
    """

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        num_polarizations = len(self.image_parameters.fixed.polarizations)
        self.slots['weights'] = accel.IOSlot(
            (self.max_vis, accel.Dimension(num_polarizations, exact=True)), np.float32)

    def _run(self):
        if self.num_vis == 0:
            return
        grid = self.buffer('grid')
        with profile_device(self.command_queue, 'degrid'):
            _lib.call(
                'kib_degrid',
                grid.ptr, grid.padded_shape[2], grid.padded_shape[1] * grid.padded_shape[2],
                grid.shape[2], _lib.dtype_code(grid.dtype),
                self.buffer('uv').ptr, self.buffer('w_plane').ptr,
                self.buffer('weights').ptr, self.buffer('vis').ptr,
                *self._lut_args(),
                len(self.image_parameters.fixed.polarizations),
                self.num_vis, self._rejected.ptr, self.command_queue.stream)


# keep the decorator importable under the reference's name
__all__ = ['kaiser_bessel', 'kaiser_bessel_fourier', 'antialias_w_kernel', 'subpixel_coord',
           'ConvolutionKernel', 'ConvolutionKernelDevice', 'GridderTemplate', 'VisOperation',
           'GridDegrid', 'Gridder', 'DegridderTemplate', 'Degridder', 'profile_function']
