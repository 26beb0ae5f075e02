"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

numpy restatements of the reference's numpy host code and ctypes bindings for
the numba loops restated in oracle.c.  Each function cites the reference lines
it follows.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None

CLEAN_I = 0
CLEAN_SUMSQ = 1
_MEDIAN_TO_RMS = 1.4826022185056031   # clean.py:34
TILE = 32


def build(force=False):
    """Compile oracle.c with gcc if the shared object is missing or stale."""
    src = os.path.join(_HERE, 'oracle.c')
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B', '_build/liboracle.so'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# --------------------------------------------------------------------------------------
# fast_math.expj2pi (fast_math.py:7-16)
def expj2pi(x):
    x = np.asarray(x)
    y = 2 * np.pi * (x - np.rint(x))
    if x.dtype == np.float32:
        y = y.astype(np.float32)
        return (np.cos(y) + 1j * np.sin(y)).astype(np.complex64)
    return np.cos(y) + 1j * np.sin(y)


# --------------------------------------------------------------------------------------
# Convolution kernel generation (grid.py:158-184, 235-334, 358-423)
def kaiser_bessel_fourier(f, width, beta):
    alpha = beta / math.pi
    return width / np.i0(beta) * np.sinc(np.lib.scimath.sqrt((width * f)**2 - alpha * alpha)).real


def antialias_w_kernel(cell_wavelengths, w, width, oversample, antialias_width,
                       image_oversample, beta):
    out_pixels = oversample * width
    pixels = out_pixels * image_oversample
    uv_width = width * cell_wavelengths * image_oversample
    image_step = 1 / uv_width
    l = (np.arange(pixels) - (pixels // 2)) * image_step
    shift_by = -0.5 * cell_wavelengths / oversample
    aa_factor = cell_wavelengths * kaiser_bessel_fourier(l * cell_wavelengths, antialias_width, beta)
    l2 = l * l
    w_arg = np.outer(-np.asarray(w), -0.5 * l2 - 5.0 / 24.0 * l2 * l2)
    image_values = aa_factor * expj2pi(w_arg + shift_by * l)
    uv_values = np.fft.fft(np.fft.ifftshift(image_values, axes=-1), axis=-1) * image_step
    uv_values = np.concatenate(
        (uv_values[..., -(out_pixels // 2):], uv_values[..., :(out_pixels // 2)]), axis=-1)
    kernel = np.reshape(uv_values, np.shape(w) + (width, oversample))[..., ::-1]
    return np.ascontiguousarray(np.swapaxes(kernel, 1, 2)).astype(np.complex64)


def kernel_beta(antialias_width):
    return 1.2 * math.pi * math.sqrt(0.25 * antialias_width**2 - 1.0)


def convolution_kernel(image_parameters, grid_parameters):
    """LUT complex64[w_planes][oversample][kernel_width] (ConvolutionKernel.__init__)."""
    cell_wavelengths = float(image_parameters.cell_size / image_parameters.wavelength)
    w_slice_wavelengths = float(grid_parameters.fixed.max_w
                                / (grid_parameters.w_slices * image_parameters.wavelength))
    w_plane_wavelengths = w_slice_wavelengths / grid_parameters.w_planes
    max_w = (w_slice_wavelengths - w_plane_wavelengths) * 0.5
    ws = np.linspace(-max_w, max_w, grid_parameters.w_planes)
    return antialias_w_kernel(
        cell_wavelengths, ws, grid_parameters.fixed.kernel_width,
        grid_parameters.fixed.oversample, grid_parameters.fixed.antialias_width,
        grid_parameters.fixed.image_oversample, kernel_beta(grid_parameters.fixed.antialias_width))


def taper(grid_parameters, N, dtype=np.float64):
    """ConvolutionKernel.taper (grid.py:404-423)."""
    x = np.arange(N) / N - 0.5
    out = kaiser_bessel_fourier(x, grid_parameters.fixed.antialias_width,
                                kernel_beta(grid_parameters.fixed.antialias_width))
    out = out * np.sinc(x / grid_parameters.fixed.oversample)
    return out.astype(dtype)


# --------------------------------------------------------------------------------------
# Gridding / degridding (oracle.c)
def grid(lut, values, weights_grid, uv, sub_uv, w_plane, vis):
    """`_grid` (grid.py:1033-1052): accumulates into `values` (complex64/128
    [P][size][size]) in place."""
    lut = _c(lut, np.complex64)
    assert values.flags.c_contiguous and values.dtype in (np.complex64, np.complex128)
    P, size, size2 = values.shape
    assert size == size2
    weights_grid = _c(weights_grid, np.float32)
    assert weights_grid.shape == values.shape
    uv = _c(uv, np.int16)
    sub_uv = _c(sub_uv, np.int16)
    w_plane = _c(w_plane, np.int16)
    vis = _c(vis, np.complex64)
    n = len(w_plane)
    assert vis.shape == (n, P)
    fn = lib().kor_grid_f32 if values.dtype == np.complex64 else lib().kor_grid_f64
    fn(_ptr(lut), ctypes.c_int(lut.shape[1]), ctypes.c_int(lut.shape[2]), _ptr(values),
       ctypes.c_int(P), ctypes.c_int(size), _ptr(weights_grid), _ptr(uv), _ptr(sub_uv),
       _ptr(w_plane), _ptr(vis), ctypes.c_long(n))
    return values


def degrid(lut, values, uv, sub_uv, w_plane, weights, vis):
    """`_degrid` (grid.py:1139-1154): subtracts the prediction from `vis` in place."""
    lut = _c(lut, np.complex64)
    assert values.flags.c_contiguous and values.dtype in (np.complex64, np.complex128)
    P, size, _ = values.shape
    uv = _c(uv, np.int16)
    sub_uv = _c(sub_uv, np.int16)
    w_plane = _c(w_plane, np.int16)
    weights = _c(weights, np.float32)
    assert vis.flags.c_contiguous and vis.dtype == np.complex64
    n = len(w_plane)
    fn = lib().kor_degrid_f32 if values.dtype == np.complex64 else lib().kor_degrid_f64
    fn(_ptr(lut), ctypes.c_int(lut.shape[1]), ctypes.c_int(lut.shape[2]), _ptr(values),
       ctypes.c_int(P), ctypes.c_int(size), _ptr(uv), _ptr(sub_uv), _ptr(w_plane),
       _ptr(weights), _ptr(vis), ctypes.c_long(n))
    return vis


# --------------------------------------------------------------------------------------
# Grid <-> image (image.py:781-799, 836-848)
def _pad_grid(grid_values, pixels):
    """Centre a G x G device-style grid in a pixels x pixels host-style grid."""
    P, G, _ = grid_values.shape
    if G == pixels:
        return grid_values
    out = np.zeros((P, pixels, pixels), grid_values.dtype)
    pad = (pixels - G) // 2
    out[:, pad:pad + G, pad:pad + G] = grid_values
    return out


def grid_to_image(grid_values, image, kernel1d, lm_scale, lm_bias, w):
    """GridToImageHost.__call__: image += ... (in place).  `grid_values` may be
    smaller than the image (device grids are); it is zero-padded about its centre."""
    pixels = image.shape[-1]
    full = _pad_grid(grid_values, pixels)
    layer = np.fft.ifft2(np.fft.ifftshift(full, axes=(1, 2)), axes=(1, 2)).astype(full.dtype)
    scale = layer.shape[1] * layer.shape[2]
    lm = np.arange(pixels).astype(image.dtype) * lm_scale + lm_bias
    lm = np.fft.ifftshift(lm)
    lm2 = lm * lm
    n = np.sqrt(1 - (lm2[:, np.newaxis] + lm2[np.newaxis, :]))
    w_correct = expj2pi(w * (n - 1))
    layer *= w_correct
    out = layer.real.copy()
    out *= scale
    out *= n[np.newaxis, ...]
    out = np.fft.fftshift(out, axes=(1, 2))
    out /= np.outer(kernel1d, kernel1d)[np.newaxis, ...]
    image += out
    return image


def grid_to_image_threaded(grid_plane, image_plane, kernel1d, lm_scale, lm_bias, w, pool, threads):
    """:func:`grid_to_image` for ONE polarization plane with the work split over a thread
    pool (row blocks): the same arithmetic per pixel (the 2-D transform as two 1-D passes), so the
    result agrees to rounding (1e-7 of the peak, tests/test_oracle.py).  Used by bench.py's CPU legs so that the reference's single-threaded numpy
    path (GridToImageHost, image.py:781-799) is timed on all host cores.

    grid_plane : complex64 [G][G];  image_plane : float32 [N][N], accumulated into."""
    pixels = image_plane.shape[-1]
    full = _pad_grid(grid_plane[np.newaxis], pixels)[0]
    layer = np.fft.ifftshift(full)
    bounds = np.linspace(0, pixels, threads + 1).astype(int)
    blocks = [(bounds[i], bounds[i + 1]) for i in range(threads) if bounds[i] < bounds[i + 1]]

    def rows(block):
        a, b = block
        layer[a:b] = np.fft.ifft(layer[a:b], axis=1)

    def cols(block):
        a, b = block
        layer[:, a:b] = np.fft.ifft(layer[:, a:b], axis=0)

    list(pool.map(rows, blocks))
    list(pool.map(cols, blocks))
    scale = pixels * pixels
    lm = np.arange(pixels).astype(image_plane.dtype) * lm_scale + lm_bias
    lm = np.fft.ifftshift(lm)
    lm2 = lm * lm
    shifted = np.fft.ifftshift(np.arange(pixels))         # layer row r is image row shifted[r]
    k1 = np.asarray(kernel1d)

    def epilogue(block):
        a, b = block
        n = np.sqrt(1 - (lm2[a:b, np.newaxis] + lm2[np.newaxis, :]))
        part = layer[a:b] * expj2pi(w * (n - 1))
        out = part.real.copy()
        out *= scale
        out *= n
        out = np.fft.fftshift(out, axes=1)
        ys = shifted[a:b]
        out /= np.outer(k1[ys], k1)
        image_plane[ys] += out

    list(pool.map(epilogue, blocks))
    return image_plane


def image_to_grid(image, kernel1d, lm_scale, lm_bias, w, complex_dtype=np.complex64,
                  grid_size=None):
    """ImageToGridHost.__call__; returns the (optionally centre-cropped) grid."""
    pixels = image.shape[-1]
    lm = np.arange(pixels).astype(image.dtype) * lm_scale + lm_bias
    lm2 = lm * lm
    n = np.sqrt(1 - (lm2[:, np.newaxis] + lm2[np.newaxis, :]))[np.newaxis, ...]
    w_correct = expj2pi(-w * (n - 1))
    kernel = np.outer(kernel1d, kernel1d)[np.newaxis, ...]
    layer = (image / (kernel * n) * w_correct).astype(complex_dtype)
    full = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(layer, axes=(1, 2)), axes=(1, 2)),
                           axes=(1, 2)).astype(complex_dtype)
    if grid_size is not None and grid_size != pixels:
        pad = (pixels - grid_size) // 2
        full = np.ascontiguousarray(full[:, pad:pad + grid_size, pad:pad + grid_size])
    return full


# --------------------------------------------------------------------------------------
# CLEAN (clean.py:894-1075)
def psf_patch(psf, threshold, limit=None):
    """psf_patch_host (clean.py:894-935)."""
    if limit is not None:
        hlimit = (round(limit * min(psf.shape[1], psf.shape[2])) - 1) // 2
        mid_x = psf.shape[2] // 2
        mid_y = psf.shape[1] // 2
        min_x = max(0, mid_x - hlimit)
        min_y = max(0, mid_y - hlimit)
        max_x = min(psf.shape[2] - 1, mid_x + hlimit)
        max_y = min(psf.shape[1] - 1, mid_y + hlimit)
        psf = psf[:, min_y:max_y + 1, min_x:max_x + 1]
    nz = np.nonzero(np.abs(psf) >= threshold)
    if len(nz[0]) == 0:
        return (psf.shape[0], 1, 1)
    y_dist = np.max(np.abs(nz[1] - psf.shape[1] // 2))
    x_dist = np.max(np.abs(nz[2] - psf.shape[2] // 2))
    return (psf.shape[0], int(min(psf.shape[1], 2 * y_dist + 1)),
            int(min(psf.shape[2], 2 * x_dist + 1)))


def noise_est(image, border):
    """noise_est_host (clean.py:938-943)."""
    border_pixels = round(border * min(image.shape[1], image.shape[2]))
    image = image[:, border_pixels:-border_pixels, border_pixels:-border_pixels]
    return np.median(np.abs(image)) * _MEDIAN_TO_RMS


class CleanHost:
    """CleanHost (clean.py:971-1075) for float32 images, loops in oracle.c.

    ``__call__`` returns the position at which the component was subtracted; the
    position the reference returns (aliased to the tile's new peak, a quirk of
    clean.py:1063-1075) is available as ``reported_pos``."""

    def __init__(self, pixels, border, mode, loop_gain, image, psf, model):
        for a in (image, psf, model):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        self.image = image
        self.psf = psf
        self.model = model
        self.mode = mode
        self.loop_gain = loop_gain
        self.border_pixels = round(pixels * border)
        self.tiles_x = -(-(image.shape[2] - 2 * self.border_pixels) // TILE)
        self.tiles_y = -(-(image.shape[1] - 2 * self.border_pixels) // TILE)
        self.tile_max = np.zeros((self.tiles_y, self.tiles_x), np.float32)
        self.tile_pos = np.zeros((self.tiles_y, self.tiles_x, 2), np.int32)

    def reset(self):
        P, h, w = self.image.shape
        lib().kor_clean_reset(_ptr(self.image), P, h, w, self.border_pixels, self.mode,
                              _ptr(self.tile_max), _ptr(self.tile_pos), self.tiles_y, self.tiles_x)

    def __call__(self, psf_patch, threshold=0.0):
        P, h, w = self.image.shape
        peak_value = np.zeros(1, np.float32)
        peak_pos = np.zeros(2, np.int32)
        model_pixel = np.zeros(P, np.float32)
        reported = np.zeros(2, np.int32)
        done = lib().kor_clean_cycle(
            _ptr(self.image), _ptr(self.psf), _ptr(self.model), P, h, w,
            self.psf.shape[1], self.psf.shape[2], self.border_pixels, self.mode,
            ctypes.c_float(self.loop_gain), int(psf_patch[1]), int(psf_patch[2]),
            ctypes.c_float(threshold),
            _ptr(self.tile_max), _ptr(self.tile_pos), self.tiles_y, self.tiles_x,
            _ptr(peak_value), _ptr(peak_pos), _ptr(model_pixel), _ptr(reported))
        if not done:
            return None, None, None
        #: what the reference's CleanHost.__call__ returns (see oracle.c kor_clean_cycle)
        self.reported_pos = (int(reported[0]), int(reported[1]))
        return peak_value[0], (int(peak_pos[0]), int(peak_pos[1])), model_pixel


# --------------------------------------------------------------------------------------
# Weights (weight.py:541-605)
NATURAL, UNIFORM, ROBUST = 0, 1, 2


class WeightsHost:
    def __init__(self, weight_type, weights_grid):
        self.weight_type = getattr(weight_type, 'value', weight_type)
        self.robustness = 0.0
        self.weights_grid = weights_grid
        assert weights_grid.dtype == np.float32 and weights_grid.flags.c_contiguous

    def clear(self):
        if self.weight_type != NATURAL:
            self.weights_grid.fill(0)

    def grid(self, uv, weights):
        """WeightsHost.grid; unlike the reference (weight.py:570) `uv` is not modified."""
        uv = np.asarray(uv)
        assert uv.dtype == np.int16
        weights = _c(weights, np.float32)
        P, h, w = self.weights_grid.shape
        stride = uv.strides[0] // 2
        assert uv.strides[1] == 2
        lib().kor_grid_weights(_ptr(self.weights_grid), P, h, w, _ptr(uv), stride,
                               _ptr(weights), ctypes.c_long(len(uv)))

    def finalize(self):
        wg = self.weights_grid
        if self.weight_type == NATURAL:
            wg.fill(1)
            return None, 1.0
        elif self.weight_type == UNIFORM:
            sum_w = np.sum(wg[0])
            sum_dw = np.count_nonzero(wg[0])
            wg[wg == 0] = np.inf
            np.reciprocal(wg, out=wg)
            sum_d2w = np.sum(wg[0])
            rms = np.sqrt(sum_d2w) / sum_dw
            return rms, rms * np.sqrt(sum_w)
        elif self.weight_type == ROBUST:
            sum_sq = np.dot(wg[0].flat, wg[0].flat)
            total = np.sum(wg[0])
            mean_weight = sum_sq / total
            S2 = (5 * 10**(-self.robustness))**2 / mean_weight
            old0 = wg[0].copy()
            wg[wg == 0] = np.inf
            np.reciprocal(wg * S2 + 1, out=wg)
            sum_w = np.sum(old0)
            sum_dw = np.sum(wg[0] * old0)
            sum_d2w = np.sum(wg[0]**2 * old0)
            rms = np.sqrt(sum_d2w) / sum_dw
            return rms, rms * np.sqrt(sum_w)
        raise ValueError('Unknown weight_type {}'.format(self.weight_type))


# --------------------------------------------------------------------------------------
# Direct prediction (predict.py:122-149, 420-438)
def uvw_scale_bias(image_parameters, grid_parameters):
    ip, gp = image_parameters, grid_parameters
    uv_scale = float(ip.cell_size / gp.fixed.oversample / ip.wavelength)
    w_scale = float(gp.fixed.max_w / ((gp.w_slices - 0.5) * gp.w_planes) / ip.wavelength)
    w_bias = (0.5 - 0.5 * gp.w_planes) * w_scale
    return uv_scale, w_scale, w_bias


def predict(vis, uv, sub_uv, w_plane, weights, lmn, flux, oversample, uv_scale, w_scale, w_bias):
    """`_predict_host`: subtracts the direct-DFT prediction from `vis` in place."""
    assert vis.dtype == np.complex64 and vis.flags.c_contiguous
    n, P = vis.shape
    uv = _c(uv, np.int16)
    sub_uv = _c(sub_uv, np.int16)
    w_plane = _c(w_plane, np.int16)
    weights = _c(weights, np.float32)
    lmn = _c(lmn, np.float32)
    flux = _c(flux, np.float32)
    lib().kor_predict(_ptr(vis), _ptr(uv), _ptr(sub_uv), _ptr(w_plane), _ptr(weights),
                      _ptr(lmn), _ptr(flux), ctypes.c_long(n), len(lmn), P,
                      ctypes.c_float(oversample), ctypes.c_float(uv_scale),
                      ctypes.c_float(w_scale), ctypes.c_float(w_bias))
    return vis


# --------------------------------------------------------------------------------------
# Preprocessing (preprocess.cpp:313-513)
def preprocess_dtype(num_polarizations):
    """vis_t<P> including the w_slice field (preprocess.cpp:39-52)."""
    P = num_polarizations
    return np.dtype([('uv', 'i2', (2,)), ('sub_uv', 'i2', (2,)), ('w_plane', 'i2'),
                     ('w_slice', 'i2'), ('weights', 'f4', (P,)), ('vis', 'c8', (P,))])


def preprocess(uvw, weights, vis, mueller_stokes, num_polarizations, cell_size, max_w, w_slices,
               w_planes, oversample, feed_angle1=None, feed_angle2=None, mueller_circular=None,
               capacity=0):
    """visibility_collector<P>::add for one channel (see oracle.c kor_preprocess).  Returns
    (records ordered by W slice, counts per slice)."""
    uvw = _c(uvw, np.float32)
    weights = _c(weights, np.float32)
    vis = _c(vis, np.complex64)
    n, Q = vis.shape
    P = num_polarizations
    stokes = _c(mueller_stokes, np.complex64)
    circular = _c(mueller_circular, np.complex64) if mueller_circular is not None else None
    f1 = _c(feed_angle1, np.float32) if feed_angle1 is not None else None
    f2 = _c(feed_angle2, np.float32) if feed_angle2 is not None else None
    assert stokes.shape == ((P, Q) if f1 is None else (P, 4))
    out = np.zeros(max(n, 1), preprocess_dtype(P))
    counts = np.zeros(w_slices, np.int64)
    fn = lib().kor_preprocess
    fn.restype = ctypes.c_long
    total = fn(_ptr(uvw), _ptr(weights), _ptr(vis), ctypes.c_long(n), Q,
               _ptr(f1) if f1 is not None else None, _ptr(f2) if f2 is not None else None,
               _ptr(stokes), _ptr(circular) if circular is not None else None, P,
               ctypes.c_float(cell_size), ctypes.c_float(max_w), w_slices, w_planes, oversample,
               ctypes.c_long(capacity), _ptr(out), _ptr(counts))
    return out[:total].view(np.recarray), counts
