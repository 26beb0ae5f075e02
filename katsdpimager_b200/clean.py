"""Hogbom CLEAN minor cycles on B200.

Same surface as the reference's :mod:`katsdpimager.clean` (reference clean.py:37-891):
``PsfPatchTemplate``, ``NoiseEstTemplate``, ``CleanTemplate`` and their instantiations
with the same slots, plus the helpers ``metric_to_power`` / ``power_to_metric`` /
``noise_threshold_scale``.  Differences from the reference's device code, all in
csrc/kib_clean.cu:

* peak selection follows the HOST tie-break (first strict maximum, row-major within a
  tile, then ``np.argmax`` order over tiles), without FMA contraction, so the component
  list is bit-identical to ``CleanHost`` (reference clean.py:946-1075);
* a minor cycle is one kernel (subtract + tile update + next-peak search by the last
  block) and any number of cycles can be run back to back without host involvement:
  :meth:`Clean.run_cycles`.  :meth:`Clean.__call__` keeps the one-cycle-per-call API;
* the noise estimate is an exact median by radix selection (3-6 histogram passes)
  instead of ~32 rank passes with a host round trip each.
"""
import math

import numpy as np
import scipy.stats

from . import _lib, accel
from .profiling import profile_device

#: Use only Stokes I to find peaks
CLEAN_I = 0
#: Use the sum of squares of available Stokes components to find peaks
CLEAN_SUMSQ = 1

#: Scales median absolute value of a zero-mean Gaussian distribution to its standard deviation
_MEDIAN_TO_RMS = 1.4826022185056031

TILE = 32


def metric_to_power(mode, metric):
    """Convert a peak-finding metric to a value linear in flux density (clean.py:166-174)."""
    if mode == CLEAN_I:
        return metric
    elif mode == CLEAN_SUMSQ:
        return math.sqrt(metric)
    raise ValueError('Invalid mode {}'.format(mode))


def power_to_metric(mode, power):
    """Inverse of :func:`metric_to_power` (clean.py:177-184)."""
    if mode == CLEAN_I:
        return power
    elif mode == CLEAN_SUMSQ:
        return power * power
    raise ValueError('Invalid mode {}'.format(mode))


def noise_threshold_scale(mode, threshold, num_polarizations):
    """Sigma threshold adjusted for the chi-squared statistics of the SUMSQ metric
    (clean.py:187-203)."""
    if mode == CLEAN_I:
        return threshold
    elif mode == CLEAN_SUMSQ:
        p = 2 * scipy.stats.norm.sf(threshold)
        return np.sqrt(scipy.stats.chi2.isf(p, num_polarizations))
    raise ValueError('Invalid mode {}'.format(mode))


def _strides(array):
    """(row stride, plane stride) in elements of a polarizations x height x width array."""
    return array.padded_shape[2], array.padded_shape[1] * array.padded_shape[2]


class _Template:
    def __init__(self, context, dtype, num_polarizations, tuning=None):
        _lib.load()
        self.context = context
        self.dtype = np.dtype(dtype)
        self.num_polarizations = num_polarizations


# ---------------------------------------------------------------------------- PSF patch
class PsfPatchTemplate(_Template):
    """Bounding box of the PSF above a threshold (reference clean.py:37-69)."""

    def instantiate(self, *args, **kwargs):
        return PsfPatch(self, *args, **kwargs)


class PsfPatch(accel.Operation):
    """.. rubric:: Slots

    **psf** : real, polarizations x height x width, central value 1
    **bound** : int32[2], maximum |dx|, |dy| found (device scratch)
    """

    def __init__(self, template, command_queue, shape, allocator=None):
        if shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        super().__init__(command_queue, allocator)
        polarizations = accel.Dimension(template.num_polarizations, exact=True)
        self.slots['psf'] = accel.IOSlot([polarizations, shape[1], shape[2]], template.dtype)
        self.slots['bound'] = accel.IOSlot([accel.Dimension(2, exact=True)], np.int32)
        self._bound_host = accel.HostArray((2,), np.int32, context=command_queue.context)
        self.template = template

    def _run(self):
        raise NotImplementedError('use __call__(threshold, limit)')

    def __call__(self, threshold, limit=None, **kwargs):
        """Returns (polarizations, height, width) of the smallest centred box holding
        every pixel with |psf| >= threshold, examining only the central `limit`
        fraction of the PSF (clean.py:123-163)."""
        self.bind(**kwargs)
        self.ensure_all_bound()
        psf = self.buffer('psf')
        bound = self.buffer('bound')
        min_x, min_y = 0, 0
        max_x, max_y = psf.shape[2] - 1, psf.shape[1] - 1
        mid_x, mid_y = psf.shape[2] // 2, psf.shape[1] // 2
        if limit is not None:
            hlimit = (round(limit * min(psf.shape[1], psf.shape[2])) - 1) // 2
            min_x, min_y = max(min_x, mid_x - hlimit), max(min_y, mid_y - hlimit)
            max_x, max_y = min(max_x, mid_x + hlimit), min(max_y, mid_y + hlimit)
        row_stride, pol_stride = _strides(psf)
        with profile_device(self.command_queue, 'psf_patch'):
            _lib.call('kib_psf_patch', psf.ptr, row_stride, pol_stride, psf.shape[0],
                      min_x, min_y, max_x, max_y, mid_x, mid_y,
                      float(np.float32(threshold)) if psf.dtype == np.float32 else float(threshold),
                      bound.ptr, _lib.dtype_code(psf.dtype), self.command_queue.stream)
        bound.get(self.command_queue, self._bound_host)
        box = 2 * self._bound_host + 1
        return (self.template.num_polarizations,
                int(min(box[1], psf.shape[1])), int(min(box[0], psf.shape[2])))


# ----------------------------------------------------------------------- noise estimate
class NoiseEstTemplate(_Template):
    """Robust noise estimate: 1.4826 x median |pixel| inside the border
    (reference clean.py:206-244, host :938-943)."""

    def instantiate(self, *args, **kwargs):
        return NoiseEst(self, *args, **kwargs)


class NoiseEst(accel.Operation):
    """.. rubric:: Slots

    **dirty** : real, polarizations x height x width
    **rank** : uint32[2048] histogram scratch
    """

    #: (shift, bits) of the three radix passes: sign + exponent, 10 and 13 mantissa bits
    _DIGITS = ((23, 9), (13, 10), (0, 13))
    _WINDOW = 8                 # leading digits covered by a guessed pass (a factor of 256)

    def __init__(self, template, command_queue, image_shape, border, allocator=None):
        if image_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        if border >= 0.5:
            raise ValueError('Border must be less than half the image size')
        super().__init__(command_queue, allocator)
        self.template = template
        self.border_pixels = round(border * min(image_shape[1], image_shape[2]))
        self.slots['dirty'] = accel.IOSlot([
            accel.Dimension(template.num_polarizations, exact=True),
            image_shape[1], image_shape[2]], template.dtype)
        self.slots['rank'] = accel.IOSlot([accel.Dimension(8192 + 2, exact=True)], np.uint32)
        self._hist_host = accel.HostArray((8192 + 2,), np.uint32,
                                          context=command_queue.context)
        #: leading radix digit of the last median: the next estimate of a similar image starts
        #: from it and saves the (expensive) first pass
        self._guess = None

    def _run(self):
        raise NotImplementedError('use __call__()')

    def _histogram(self, prefix, prefix_bits, shift, bits):
        dirty = self.buffer('dirty')
        hist = self.buffer('rank')
        hist.zero(self.command_queue)
        row_stride, pol_stride = _strides(dirty)
        with profile_device(self.command_queue, 'abs_histogram'):
            if prefix_bits > 0:
                # few values match a prefix: plain shared-memory atomics, no warp aggregation
                _lib.call('kib_abs_histogram_window', dirty.ptr, row_stride, pol_stride,
                          dirty.shape[2], dirty.shape[1], dirty.shape[0], self.border_pixels,
                          prefix, 1, prefix_bits, shift, bits, hist.ptr,
                          (hist.ptr.value or 0) + 8192 * 4,
                          _lib.dtype_code(dirty.dtype), self.command_queue.stream)
            else:
                _lib.call('kib_abs_histogram', dirty.ptr, row_stride, pol_stride,
                          dirty.shape[2], dirty.shape[1], dirty.shape[0], self.border_pixels,
                          prefix, prefix_bits, shift, bits, hist.ptr,
                          _lib.dtype_code(dirty.dtype), self.command_queue.stream)
        hist.get(self.command_queue, self._hist_host)
        return self._hist_host[:1 << bits].astype(np.int64)

    def _select_from_guess(self, ranks):
        """Order statistics when all of them have a leading digit (exponent) within a factor of
        16 of the guess: one windowed pass (second digit for _WINDOW adjacent leading digits +
        count below) and the last pass; None if the guess was too far off."""
        (shift0, bits0), (shift1, bits1), (shift2, bits2) = self._DIGITS
        bins = 1 << bits1
        window = self._WINDOW
        first = min(max(self._guess - window // 2, 0), (1 << bits0) - window)
        dirty = self.buffer('dirty')
        hist = self.buffer('rank')
        hist.zero(self.command_queue)
        row_stride, pol_stride = _strides(dirty)
        below_ptr = (hist.ptr.value or 0) + 8192 * 4
        with profile_device(self.command_queue, 'abs_histogram'):
            _lib.call('kib_abs_histogram_window', dirty.ptr, row_stride, pol_stride,
                      dirty.shape[2], dirty.shape[1], dirty.shape[0], self.border_pixels,
                      first, window, bits0, shift1, bits1, hist.ptr, below_ptr,
                      _lib.dtype_code(dirty.dtype), self.command_queue.stream)
        hist.get(self.command_queue, self._hist_host)
        flat = self._hist_host[:window * bins].astype(np.int64)
        below = int(self._hist_host[8192:].view(np.uint64)[0])
        cumulative = np.cumsum(flat)
        for r in ranks:
            if r - below < 0 or r - below >= cumulative[-1]:
                return None
        groups = {}
        for r in ranks:
            idx = int(np.searchsorted(cumulative, r - below, side='right'))
            groups.setdefault(idx, []).append(r)
        results = {}
        for idx, rs in groups.items():
            prefix = ((first + idx // bins) << bits1) | (idx % bins)
            offset = below + (int(cumulative[idx - 1]) if idx > 0 else 0)
            last = self._histogram(prefix, 32 - shift2 - bits2, shift2, bits2)
            cum2 = np.cumsum(last)
            for r in rs:
                bucket = int(np.searchsorted(cum2, r - offset, side='right'))
                results[r] = (prefix << bits2) | bucket
        return [np.array(results[r], np.uint32).view(np.float32) for r in ranks]

    def _select(self, ranks):
        """Exact order statistics (0-based `ranks`, ascending) of |dirty| inside the
        border, as float32 bit patterns, by most-significant-digit radix selection."""
        if self._guess is not None:
            found = self._select_from_guess(ranks)
            if found is not None:
                self._guess = int(found[-1].view(np.uint32)) >> self._DIGITS[0][0]
                return found
        results = {}
        pending = [(0, 0, 0, list(ranks))]      # (pass index, prefix, rank offset, ranks)
        while pending:
            level, prefix, offset, wanted = pending.pop()
            shift, bits = self._DIGITS[level]
            prefix_bits = 32 - shift - bits
            hist = self._histogram(prefix, prefix_bits, shift, bits)
            cumulative = np.cumsum(hist)
            groups = {}
            for r in wanted:
                bucket = int(np.searchsorted(cumulative, r - offset, side='right'))
                groups.setdefault(bucket, []).append(r)
            for bucket, rs in groups.items():
                below = int(cumulative[bucket - 1]) if bucket > 0 else 0
                new_prefix = (prefix << bits) | bucket
                if level + 1 == len(self._DIGITS):
                    for r in rs:
                        results[r] = new_prefix
                else:
                    pending.append((level + 1, new_prefix, offset + below, rs))
        self._guess = results[ranks[-1]] >> self._DIGITS[0][0]
        return [np.array(results[r], np.uint32).view(np.float32) for r in ranks]

    def _binary_search(self):
        """Median by bisection on the value with a rank kernel (any precision);
        the algorithm of the reference's device code (clean.py:295-353)."""
        dirty = self.buffer('dirty')
        dtype = dirty.dtype
        itype = np.uint32 if dtype == np.float32 else np.uint64
        counter = accel.DeviceArray(self.command_queue.context, (1,), np.uint64)
        row_stride, pol_stride = _strides(dirty)
        median_rank = ((dirty.shape[1] - 2 * self.border_pixels)
                       * (dirty.shape[2] - 2 * self.border_pixels) * dirty.shape[0] // 2)
        low = dtype.type(0)
        high = dtype.type(np.inf)
        while high > np.finfo(dtype).tiny and high > low * 1.0001:
            ilow = low.view(itype)
            ihigh = high.view(itype)
            if ihigh - ilow == itype(1):
                break
            mid = (ilow + (ihigh - ilow) // itype(2)).view(dtype)
            counter.zero(self.command_queue)
            _lib.call('kib_rank', dirty.ptr, row_stride, pol_stride,
                      dirty.shape[2], dirty.shape[1], dirty.shape[0], self.border_pixels,
                      float(mid), counter.ptr, _lib.dtype_code(dtype), self.command_queue.stream)
            if int(counter.get(self.command_queue)[0]) < median_rank:
                low = mid
            else:
                high = mid
        return low * _MEDIAN_TO_RMS

    def __call__(self, **kwargs):
        self.bind(**kwargs)
        self.ensure_all_bound()
        dirty = self.buffer('dirty')
        if dirty.dtype != np.float32:
            return self._binary_search()
        count = ((dirty.shape[1] - 2 * self.border_pixels)
                 * (dirty.shape[2] - 2 * self.border_pixels) * dirty.shape[0])
        if count <= 0:
            return np.float32(np.nan)
        # np.median: middle element, or the mean of the two middle elements
        if count % 2:
            median = self._select([count // 2])[0]
        else:
            lo, hi = self._select([count // 2 - 1, count // 2])
            median = (lo + hi) / np.float32(2)
        return np.float32(median) * np.float32(_MEDIAN_TO_RMS)


# ------------------------------------------------------------------ individual CLEAN ops
class _UpdateTilesTemplate(_Template):
    """Per-tile peaks for the tiles intersecting a window (reference clean.py:356-395)."""

    def __init__(self, context, dtype, num_polarizations, mode, tuning=None):
        super().__init__(context, dtype, num_polarizations, tuning)
        if mode not in (CLEAN_I, CLEAN_SUMSQ):
            raise ValueError('Invalid mode {}'.format(mode))
        self.mode = mode
        self.tilex = TILE
        self.tiley = TILE

    def instantiate(self, *args, **kwargs):
        return _UpdateTiles(self, *args, **kwargs)


class _UpdateTiles(accel.Operation):
    """.. rubric:: Slots

    **dirty** : real, polarizations x height x width
    **tile_max** : real, tiles_y x tiles_x;  **tile_pos** : int32, tiles_y x tiles_x x 2 (row, col)
    """

    def __init__(self, template, command_queue, image_shape, border, allocator=None):
        if image_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        if border >= 0.5:
            raise ValueError('Border must be less than half the image size')
        super().__init__(command_queue, allocator)
        border_pixels = round(border * min(image_shape[1], image_shape[2]))
        num_tiles_x = accel.divup(image_shape[2] - 2 * border_pixels, template.tilex)
        num_tiles_y = accel.divup(image_shape[1] - 2 * border_pixels, template.tiley)
        image_dims = [accel.Dimension(template.num_polarizations, exact=True),
                      accel.Dimension(image_shape[1]), accel.Dimension(image_shape[2])]
        tiles_width = accel.Dimension(num_tiles_x)
        tiles_height = accel.Dimension(num_tiles_y)
        self.template = template
        self.border_pixels = border_pixels
        self.slots['dirty'] = accel.IOSlot(image_dims, template.dtype)
        self.slots['tile_max'] = accel.IOSlot([tiles_height, tiles_width], template.dtype)
        self.slots['tile_pos'] = accel.IOSlot(
            [tiles_height, tiles_width, accel.Dimension(2, exact=True)], np.int32)

    def _run(self):
        raise NotImplementedError('use __call__(x0, y0, x1, y1)')

    def __call__(self, x0, y0, x1, y1, **kwargs):
        """Update all tiles intersected by the pixel range [x0, x1) x [y0, y1)."""
        self.bind(**kwargs)
        self.ensure_all_bound()
        tile_max = self.buffer('tile_max')
        tile_pos = self.buffer('tile_pos')
        tx0 = max((x0 - self.border_pixels) // TILE, 0)
        ty0 = max((y0 - self.border_pixels) // TILE, 0)
        tx1 = min(accel.divup(x1 - self.border_pixels, TILE), tile_max.shape[1])
        ty1 = min(accel.divup(y1 - self.border_pixels, TILE), tile_max.shape[0])
        if tx0 < tx1 and ty0 < ty1:
            dirty = self.buffer('dirty')
            row_stride, pol_stride = _strides(dirty)
            assert tile_pos.padded_shape[1] == tile_max.padded_shape[1]
            with profile_device(self.command_queue, 'update_tiles'):
                _lib.call('kib_update_tiles', dirty.ptr, row_stride, pol_stride,
                          dirty.shape[2], dirty.shape[1], dirty.shape[0], self.border_pixels,
                          self.template.mode, tile_max.ptr, tile_pos.ptr,
                          tile_max.padded_shape[1], tx0, ty0, tx1, ty1,
                          _lib.dtype_code(dirty.dtype), self.command_queue.stream)


class _FindPeakTemplate(_Template):
    """Global peak from per-tile peaks (reference clean.py:483-513)."""

    def instantiate(self, *args, **kwargs):
        return _FindPeak(self, *args, **kwargs)


class _FindPeak(accel.Operation):
    """.. rubric:: Slots

    **dirty**, **tile_max**, **tile_pos** as :class:`_UpdateTiles`;
    **peak_value** : real[1];  **peak_pos** : int32[2] (row, col);
    **peak_pixel** : real[polarizations]
    """

    def __init__(self, template, command_queue, image_shape, tile_shape, allocator=None):
        if image_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        super().__init__(command_queue, allocator)
        self.template = template
        image_dims = [accel.Dimension(image_shape[0], exact=True),
                      accel.Dimension(image_shape[1]), accel.Dimension(image_shape[2])]
        tile_dims = [accel.Dimension(tile_shape[0]), accel.Dimension(tile_shape[1])]
        self.slots['dirty'] = accel.IOSlot(image_dims, template.dtype)
        self.slots['tile_max'] = accel.IOSlot(tile_dims, template.dtype)
        self.slots['tile_pos'] = accel.IOSlot(
            tile_dims + [accel.Dimension(2, exact=True)], np.int32)
        self.slots['peak_value'] = accel.IOSlot([1], template.dtype)
        self.slots['peak_pos'] = accel.IOSlot([2], np.int32)
        self.slots['peak_pixel'] = accel.IOSlot([template.num_polarizations], template.dtype)

    def _run(self):
        dirty = self.buffer('dirty')
        tile_max = self.buffer('tile_max')
        tile_pos = self.buffer('tile_pos')
        row_stride, pol_stride = _strides(dirty)
        assert tile_pos.padded_shape[1] == tile_max.padded_shape[1]
        with profile_device(self.command_queue, 'find_peak'):
            _lib.call('kib_find_peak', dirty.ptr, row_stride, pol_stride, dirty.shape[0],
                      tile_max.ptr, tile_pos.ptr, tile_max.padded_shape[1],
                      tile_max.shape[1], tile_max.shape[0],
                      self.buffer('peak_value').ptr, self.buffer('peak_pos').ptr,
                      self.buffer('peak_pixel').ptr,
                      _lib.dtype_code(dirty.dtype), self.command_queue.stream)


class _SubtractPsfTemplate(_Template):
    """Subtract a scaled PSF patch and update the model (reference clean.py:590-622)."""

    def instantiate(self, *args, **kwargs):
        return _SubtractPsf(self, *args, **kwargs)


class _SubtractPsf(accel.Operation):
    """.. rubric:: Slots

    **dirty**, **model** : real, polarizations x height x width (same padding);
    **psf** : real, polarizations x psf_height x psf_width, centre at (h // 2, w // 2);
    **peak_pixel** : real[polarizations]
    """

    def __init__(self, template, command_queue, loop_gain, image_shape, psf_shape, allocator=None):
        super().__init__(command_queue, allocator)
        if image_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        if psf_shape[0] != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        pol_dim = accel.Dimension(template.num_polarizations, exact=True)
        image_dims = [pol_dim, accel.Dimension(image_shape[1]), accel.Dimension(image_shape[2])]
        psf_dims = [pol_dim, accel.Dimension(psf_shape[1]), accel.Dimension(psf_shape[2])]
        self.slots['dirty'] = accel.IOSlot(image_dims, template.dtype)
        self.slots['model'] = accel.IOSlot(image_dims, template.dtype)
        self.slots['psf'] = accel.IOSlot(psf_dims, template.dtype)
        self.slots['peak_pixel'] = accel.IOSlot([pol_dim], template.dtype)
        self.loop_gain = loop_gain
        self.template = template

    def _run(self):
        raise NotImplementedError('use __call__(pos, psf_patch)')

    def __call__(self, pos, psf_patch, **kwargs):
        """Subtract the central `psf_patch` (polarizations, height, width) of the PSF,
        scaled by loop_gain * peak_pixel, centred at `pos` = (row, col)."""
        self.bind(**kwargs)
        self.ensure_all_bound()
        dirty = self.buffer('dirty')
        model = self.buffer('model')
        psf = self.buffer('psf')
        assert model.padded_shape == dirty.padded_shape
        row_stride, pol_stride = _strides(dirty)
        psf_row_stride, psf_pol_stride = _strides(psf)
        with profile_device(self.command_queue, 'subtract_psf'):
            _lib.call('kib_subtract_psf', dirty.ptr, model.ptr, row_stride, pol_stride,
                      dirty.shape[2], dirty.shape[1], dirty.shape[0],
                      psf.ptr, psf_row_stride, psf_pol_stride, psf.shape[2], psf.shape[1],
                      int(psf_patch[2]), int(psf_patch[1]),
                      self.buffer('peak_pixel').ptr, int(pos[0]), int(pos[1]),
                      float(self.loop_gain), _lib.dtype_code(dirty.dtype),
                      self.command_queue.stream)


# ------------------------------------------------------------------------ composite CLEAN
class CleanTemplate:
    """Composite template for the CLEAN minor cycles (reference clean.py:729-753).

    `tuning` may contain ``lookahead`` (default 1): with a value N > 1,
    :meth:`Clean.__call__` runs up to N cycles ahead on the device and hands the
    results out one call at a time.  The component sequence is unchanged, but the
    device buffers run ahead of the caller, so it is only safe for callers that (like
    ``frontend.process_channel``) keep calling until ``None`` or until
    ``clean_parameters.minor`` cycles have been made since :meth:`Clean.reset`.
    """

    def __init__(self, context, clean_parameters, dtype, num_polarizations, tuning=None):
        self.context = context
        self.clean_parameters = clean_parameters
        self.dtype = np.dtype(dtype)
        self.num_polarizations = num_polarizations
        self.lookahead = int((tuning or {}).get('lookahead', 1))
        self._update_tiles = _UpdateTilesTemplate(context, dtype, num_polarizations,
                                                  clean_parameters.mode)
        self._find_peak = _FindPeakTemplate(context, dtype, num_polarizations)
        self._subtract_psf = _SubtractPsfTemplate(context, dtype, num_polarizations)

    def instantiate(self, *args, **kwargs):
        return Clean(self, *args, **kwargs)


class Clean(accel.OperationSequence):
    """Instantiation of :class:`CleanTemplate` (reference clean.py:756-891).

    .. rubric:: Slots

    **dirty**, **model**, **psf** : real, polarizations x height x width
    **tile_max**, **tile_pos**, **peak_value**, **peak_pos**, **peak_pixel** : internal
    state, exposed so that the memory can be shared
    """

    def __init__(self, template, command_queue, image_parameters, allocator=None):
        if image_parameters.fixed.real_dtype != template.dtype:
            raise ValueError('dtype mismatch')
        image_shape = (len(image_parameters.fixed.polarizations),
                       image_parameters.pixels, image_parameters.pixels)
        self.template = template
        params = template.clean_parameters
        self._update_tiles = template._update_tiles.instantiate(
            command_queue, image_shape, params.border, allocator)
        tile_shape = self._update_tiles.slots['tile_max'].shape
        self._find_peak = template._find_peak.instantiate(
            command_queue, image_shape, tile_shape, allocator)
        self._subtract_psf = template._subtract_psf.instantiate(
            command_queue, params.loop_gain, image_shape, image_shape, allocator)
        ops = [('update_tiles', self._update_tiles),
               ('find_peak', self._find_peak),
               ('subtract_psf', self._subtract_psf)]
        compounds = {
            'dirty': ['update_tiles:dirty', 'find_peak:dirty', 'subtract_psf:dirty'],
            'model': ['subtract_psf:model'],
            'psf': ['subtract_psf:psf'],
            'tile_max': ['update_tiles:tile_max', 'find_peak:tile_max'],
            'tile_pos': ['update_tiles:tile_pos', 'find_peak:tile_pos'],
            'peak_value': ['find_peak:peak_value'],
            'peak_pos': ['find_peak:peak_pos'],
            'peak_pixel': ['find_peak:peak_pixel', 'subtract_psf:peak_pixel']
        }
        super().__init__(command_queue, ops, compounds, allocator=allocator)
        self._record_dtype = np.dtype([('pos', np.int32, (2,)), ('value', template.dtype),
                                       ('pixel', template.dtype, (template.num_polarizations,))])
        self._capacity = 0
        self._components = None
        self._components_host = None
        self._state = accel.DeviceArray(command_queue.context, (16,), np.int32)
        self._row_scratch = accel.DeviceArray(
            command_queue.context, (tile_shape[0] * (template.dtype.itemsize + 4),), np.uint8)
        self._state_host = accel.HostArray((16,), np.int32, context=command_queue.context)
        self._pending = []
        self._pending_key = None
        self._cycles_since_reset = 0

    def _run(self):
        raise NotImplementedError('use reset() and __call__(psf_patch, threshold)')

    def _ensure_capacity(self, n):
        if n > self._capacity:
            context = self.command_queue.context
            self._capacity = max(n, 2 * self._capacity, 64)
            self._components = accel.DeviceArray(context, (self._capacity,), self._record_dtype)
            self._components_host = accel.HostArray((self._capacity,), self._record_dtype,
                                                    context=context)

    def reset(self):
        """Call after populating the buffers but before the first minor cycle."""
        self.ensure_all_bound()
        dirty = self.buffer('dirty')
        self._update_tiles(0, 0, dirty.shape[2], dirty.shape[1])
        self._pending = []
        self._cycles_since_reset = 0

    def run_cycles(self, psf_patch, threshold=0.0, max_cycles=1):
        """Run up to `max_cycles` minor cycles on the device without host round trips.

        Returns ``(components, stopped)``: a structured array with fields ``pos``
        (row, col), ``value`` (peak metric before subtraction) and ``pixel``
        (loop_gain x dirty pixel = model increment) for every cycle executed, and
        whether the loop ended because the peak fell below `threshold`.
        """
        self.ensure_all_bound()
        if max_cycles <= 0:
            return np.empty(0, self._record_dtype), False
        self._ensure_capacity(max_cycles)
        queue = self.command_queue
        dirty = self.buffer('dirty')
        model = self.buffer('model')
        psf = self.buffer('psf')
        tile_max = self.buffer('tile_max')
        tile_pos = self.buffer('tile_pos')
        assert model.padded_shape == dirty.padded_shape
        assert tile_pos.padded_shape[1] == tile_max.padded_shape[1]
        row_stride, pol_stride = _strides(dirty)
        psf_row_stride, psf_pol_stride = _strides(psf)
        params = self.template.clean_parameters
        if dirty.dtype == np.float32 and not isinstance(threshold, np.floating):
            # numpy compares a float32 peak with a Python float in float32
            threshold = float(np.float32(threshold))
        self._find_peak()
        self._state.zero(queue)
        with profile_device(queue, 'clean_cycles'):
            _lib.call('kib_clean_minor_cycles', dirty.ptr, model.ptr, row_stride, pol_stride,
                      dirty.shape[2], dirty.shape[1], dirty.shape[0],
                      self._update_tiles.border_pixels, params.mode,
                      psf.ptr, psf_row_stride, psf_pol_stride, psf.shape[2], psf.shape[1],
                      int(psf_patch[2]), int(psf_patch[1]),
                      tile_max.ptr, tile_pos.ptr, tile_max.padded_shape[1],
                      tile_max.shape[1], tile_max.shape[0],
                      self.buffer('peak_value').ptr, self.buffer('peak_pos').ptr,
                      self.buffer('peak_pixel').ptr,
                      float(params.loop_gain), float(threshold), int(max_cycles),
                      self._components.ptr, self._record_dtype.itemsize, self._state.ptr,
                      self._row_scratch.ptr, _lib.dtype_code(dirty.dtype), queue.stream)
        self._state.get_async(queue, self._state_host)
        nbytes = max_cycles * self._record_dtype.itemsize
        _lib.call('kib_memcpy_d2h_async', self._components_host.ctypes.data,
                  self._components.ptr, nbytes, queue.stream)
        queue.finish()
        count = int(self._state_host[0])
        return np.array(self._components_host[:count]), bool(self._state_host[1])

    def __call__(self, psf_patch, threshold=0.0):
        """Run a single minor cycle (reference clean.py:848-891).

        Returns ``(peak_value, (row, col), model_pixel)``, or ``(None, None, None)``
        if the peak metric is below `threshold` (in which case nothing is changed).
        """
        key = (tuple(int(x) for x in psf_patch), float(threshold))
        if self._pending and key != self._pending_key:
            raise RuntimeError('psf_patch/threshold changed while look-ahead cycles are pending')
        if not self._pending:
            budget = getattr(self.template.clean_parameters, 'minor', None)
            ahead = self.template.lookahead
            if budget is not None:
                ahead = min(ahead, budget - self._cycles_since_reset)
            components, _ = self.run_cycles(psf_patch, threshold, max(1, ahead))
            self._pending = list(components)
            self._pending_key = key
            if not self._pending:
                return None, None, None
        record = self._pending.pop(0)
        self._cycles_since_reset += 1
        pos = (int(record['pos'][0]), int(record['pos'][1]))
        return record['value'], pos, np.array(record['pixel'])
