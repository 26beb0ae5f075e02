"""Imaging parameter containers in plain SI floats.

The reference's :mod:`katsdpimager.parameters` (parameters.py:28-298) carries
astropy ``Quantity`` values; the hot path only ever evaluates ratios such as
``float(cell_size / wavelength)`` (grid.py:367-372, 754; imaging.py:90;
predict.py:140-148), so objects with the same attribute names holding floats
(metres, radians/direction cosines) are accepted by every operation here and by
the reference's own host classes.  The formulas are those of
SKA-TEL-SDP-0000003 as used by the reference.
"""
import math

import numpy as np


def is_smooth(x):
    """FFT-friendly size test (reference parameters.py:17-25): a multiple of 8
    whose only prime factors are 2, 3, 5, 7."""
    if x % 8:
        return False
    for prime in (2, 3, 5, 7):
        while x % prime == 0:
            x //= prime
    return x == 1


def next_smooth(x):
    while not is_smooth(x):
        x += 1
    return x


class ArrayParameters:
    def __init__(self, antenna_diameter, longest_baseline):
        self.antenna_diameter = float(antenna_diameter)
        self.longest_baseline = float(longest_baseline)


class FixedImageParameters:
    """Frequency-independent image properties (reference parameters.py:37-49)."""

    def __init__(self, polarizations, dtype):
        self.polarizations = list(polarizations)
        self.real_dtype = np.dtype(dtype)
        if self.real_dtype == np.float32:
            self.complex_dtype = np.dtype(np.complex64)
        elif self.real_dtype == np.float64:
            self.complex_dtype = np.dtype(np.complex128)
        else:
            raise ValueError('Unrecognised dtype {}'.format(dtype))

    def __eq__(self, other):
        return (isinstance(other, FixedImageParameters)
                and self.polarizations == other.polarizations
                and self.real_dtype == other.real_dtype)

    __hash__ = None


class ImageParameters:
    """Per-channel image geometry (reference parameters.py:52-115).

    `wavelength` in metres; `pixel_size` is the direction-cosine step per pixel.
    """

    def __init__(self, fixed, wavelength, pixels, pixel_size=None, array=None,
                 image_oversample=5.0, q_fov=1.0):
        self.fixed = fixed
        self.wavelength = float(wavelength)
        if pixel_size is None:
            if image_oversample < 3.0:
                raise ValueError('image_oversample is too small '
                                 'to capture all visibilities in the UV plane')
            uv_size = (2.0 / 3.0 * image_oversample) * array.longest_baseline
            pixel_size = self.wavelength / uv_size
        self.pixel_size = float(pixel_size)
        if pixels is None:
            cell = array.antenna_diameter * (math.pi / (7.6634 * q_fov))
            pixels = next_smooth(int(0.98 * (self.wavelength / cell) / self.pixel_size))
        elif not is_smooth(pixels):
            raise ValueError("Image size {} not supported - try {}".format(
                pixels, next_smooth(pixels)))
        self.pixels = int(pixels)
        self.image_size = self.pixel_size * self.pixels
        self.cell_size = self.wavelength / self.image_size


class FixedGridParameters:
    """Frequency-independent gridding parameters (reference parameters.py:208-238);
    `max_w` in metres."""

    def __init__(self, antialias_width, oversample, image_oversample, max_w, kernel_width,
                 degrid=False, beams=None):
        self.antialias_width = antialias_width
        self.oversample = int(oversample)
        self.image_oversample = int(image_oversample)
        self.max_w = float(max_w)
        self.kernel_width = int(kernel_width)
        self.degrid = degrid
        self.beams = beams


class GridParameters:
    def __init__(self, fixed, w_slices, w_planes):
        self.fixed = fixed
        self.w_slices = int(w_slices)
        self.w_planes = int(w_planes)


class WeightParameters:
    def __init__(self, weight_type, robustness=0.0):
        self.weight_type = weight_type
        self.robustness = robustness


class CleanParameters:
    def __init__(self, minor, loop_gain, major_gain, threshold, mode,
                 psf_cutoff, psf_limit, border):
        if psf_cutoff >= 1.0:
            raise ValueError('PSF cutoff must be less than 1')
        self.minor = minor
        self.loop_gain = loop_gain
        self.major_gain = major_gain
        self.threshold = threshold
        self.mode = mode
        self.psf_cutoff = psf_cutoff
        self.psf_limit = psf_limit
        self.border = border


def w_kernel_width(image_parameters, w, eps_w, antialias_width=0):
    """Support (in UV cells) of a W kernel, Eq 9 of SKA-TEL-SDP-0000003
    (reference parameters.py:135-158); `w` in metres."""
    fov = image_parameters.image_size
    wl = w / image_parameters.wavelength
    wk2 = 4 * fov**2 * ((wl * fov / 2)**2 + wl**1.5 * fov / (2 * math.pi * eps_w))
    return math.sqrt(wk2 + antialias_width**2)


def w_slices(image_parameters, max_w, eps_w, kernel_width, antialias_width=0):
    """Smallest number of W slices whose kernels fit `kernel_width`
    (reference parameters.py:161-178)."""
    half_w = 0.5 * max_w

    def support(slices):
        return w_kernel_width(image_parameters, half_w / (slices - 0.5), eps_w, antialias_width)

    hi = 1
    while support(hi) > kernel_width:
        hi *= 2
    lo = 0
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if support(mid) < kernel_width:
            hi = mid
        else:
            lo = mid
    return hi


def slice_mid_w(image_parameters, grid_parameters):
    """Central w (wavelengths) of every W slice (reference frontend.py:509-510)."""
    step = grid_parameters.fixed.max_w / image_parameters.wavelength / (grid_parameters.w_slices - 0.5)
    return np.arange(grid_parameters.w_slices) * step
