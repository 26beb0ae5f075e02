"""Direct (DFT) prediction of visibilities from point components on B200.

Same surface as the reference's :mod:`katsdpimager.predict` (reference
predict.py:30-416): ``PredictTemplate`` / ``Predict`` with slots ``uv, w_plane, vis,
weights, lmn, flux``, ``set_sky_model`` / ``set_sky_image`` / ``set_w``.  The kernel
(csrc/kib_predict.cu) evaluates, for the quantised UVW of every visibility,
``vis -= weights * sum_s flux_s exp(-2 pi i (l u + m v + (n - 1) w))``.
"""
import numpy as np

from . import _lib, accel, grid, polarization
from .profiling import profile_device, profile_function


def _extract_sky_model(image_parameters, grid_parameters, model, phase_centre):
    """(l, m, n-1) float32 and per-polarization flux float32 of a sky model object
    offering ``lmn(phase_centre)`` and ``flux_density(wavelength)`` (IQUV columns), with
    the sub-pixel sinc taper divided back out (reference predict.py:30-70)."""
    ip = image_parameters
    lmn = np.array(model.lmn(phase_centre), dtype=np.float64)
    lmn[:, 2] -= 1.0
    flux = np.array(model.flux_density(ip.wavelength), dtype=np.float64)
    taper_scale = float(ip.image_size * grid_parameters.fixed.oversample)
    flux *= np.prod(np.sinc(lmn[:, 0:2] / taper_scale), axis=1, keepdims=True)
    columns = [polarization.STOKES_IQUV.index(pol) for pol in ip.fixed.polarizations]
    return lmn.astype(np.float32), flux[:, columns].astype(np.float32)


def _extract_sky_image(image_parameters, grid_parameters, components):
    """Model for direct prediction from CLEAN components ``{(row, col): flux per
    polarization}`` (reference predict.py:73-119)."""
    dtype = image_parameters.fixed.real_dtype
    count = len(components)
    pixel_size = float(image_parameters.pixel_size)
    positions = np.array(list(components.keys()), dtype=np.float64).reshape(count, 2)
    centre = 0.5 * image_parameters.pixels
    l = (positions[:, 1] - centre) * pixel_size
    m = (positions[:, 0] - centre) * pixel_size
    lmn = np.empty((count, 3), np.float32)
    lmn[:, 0] = l
    lmn[:, 1] = m
    lmn[:, 2] = np.sqrt(1.0 - (np.square(l) + np.square(m))) - 1.0
    flux = np.empty((count, len(image_parameters.fixed.polarizations)), dtype)
    if count:
        flux[:] = list(components.values())
    taper_scale = float(image_parameters.image_size * grid_parameters.fixed.oversample)
    flux *= (np.sinc(l / taper_scale) * np.sinc(m / taper_scale))[:, np.newaxis]
    return lmn, flux


def _uvw_scale_bias(image_parameters, grid_parameters):
    """Scale and bias turning quantised UVW indices back into wavelengths
    (reference predict.py:122-149): ``uv = uv_scale (oversample g + s + 0.5)``,
    ``w = w_0 + w_scale w_plane + w_bias``."""
    ip, gp = image_parameters, grid_parameters
    uv_scale = float(ip.cell_size / gp.fixed.oversample / ip.wavelength)
    w_scale = float(gp.fixed.max_w / ((gp.w_slices - 0.5) * gp.w_planes) / ip.wavelength)
    w_bias = (0.5 - 0.5 * gp.w_planes) * w_scale
    return uv_scale, w_scale, w_bias


class PredictTemplate:
    autotune_version = 0

    def __init__(self, context, real_dtype, num_polarizations, tuning=None):
        _lib.load()
        self.context = context
        self.real_dtype = np.dtype(real_dtype)
        self.num_polarizations = num_polarizations
        if not 1 <= num_polarizations <= 4:
            raise ValueError('between 1 and 4 polarizations are supported')

    def instantiate(self, *args, **kwargs):
        return Predict(self, *args, **kwargs)


class Predict(grid.VisOperation):
    """Instantiation of :class:`PredictTemplate` (reference predict.py:252-416).

    .. rubric:: Slots

    In addition to those of :class:`~.grid.VisOperation`:
    **lmn** : float32, sources x 3 (l, m, n-1);  **flux** : float32, sources x polarizations;
    **weights** : float32, max_vis x polarizations
    """

    def __init__(self, template, command_queue, image_parameters, grid_parameters,
                 max_vis, max_sources, allocator=None):
        if len(image_parameters.fixed.polarizations) != template.num_polarizations:
            raise ValueError('Mismatch in number of polarizations')
        super().__init__(command_queue, template.num_polarizations, max_vis, allocator)
        self.template = template
        pol_dim = accel.Dimension(template.num_polarizations, exact=True)
        sources_dim = max(1, max_sources)
        self.slots['lmn'] = accel.IOSlot((sources_dim, accel.Dimension(3, exact=True)), np.float32)
        self.slots['flux'] = accel.IOSlot((sources_dim, pol_dim), np.float32)
        self.slots['weights'] = accel.IOSlot(
            (max_vis, accel.Dimension(template.num_polarizations, exact=True)), np.float32)
        self._num_sources = 0
        self.max_sources = max_sources
        self.image_parameters = image_parameters
        self.grid_parameters = grid_parameters
        self._w = 0.0
        context = command_queue.context
        self._host_lmn = accel.HostArray((sources_dim, 3), np.float32, context=context)
        self._host_flux = accel.HostArray((sources_dim, template.num_polarizations), np.float32,
                                          context=context)
        self._transfer_event = None

    def _upload(self, lmn, flux):
        count = len(lmn)
        if self._transfer_event is not None:
            self._transfer_event.wait()     # previous upload still reading the staging arrays
        self._host_lmn[:count] = lmn
        self._host_flux[:count] = flux
        self._num_sources = count
        self.ensure_bound('lmn')
        self.ensure_bound('flux')
        rows = np.s_[:count]
        self.buffer('lmn').set_region(self.command_queue, self._host_lmn, rows, rows,
                                      blocking=False)
        self.buffer('flux').set_region(self.command_queue, self._host_flux, rows, rows,
                                       blocking=False)
        self._transfer_event = self.command_queue.enqueue_marker()

    @profile_function()
    def set_sky_model(self, model, phase_centre):
        if len(model) > self.max_sources:
            raise ValueError('too many sources ({} > {})'.format(len(model), self.max_sources))
        self._upload(*_extract_sky_model(self.image_parameters, self.grid_parameters,
                                         model, phase_centre))

    @profile_function()
    def set_sky_image(self, components):
        lmn, flux = _extract_sky_image(self.image_parameters, self.grid_parameters, components)
        if len(lmn) > self.max_sources:
            raise ValueError('too many components ({} > {})'.format(len(lmn), self.max_sources))
        self._upload(lmn, flux)

    def set_sources(self, lmn, flux):
        """Set (l, m, n-1) and per-polarization fluxes directly."""
        if len(lmn) > self.max_sources:
            raise ValueError('too many sources ({} > {})'.format(len(lmn), self.max_sources))
        self._upload(np.asarray(lmn, np.float32), np.asarray(flux, np.float32))

    @property
    def num_sources(self):
        return self._num_sources

    def set_w(self, w):
        """Centre (wavelengths) of the W slice the quantised w_plane values belong to."""
        self._w = w

    def _run(self):
        if self.num_vis == 0 or self.num_sources == 0:
            return
        uv_scale, w_scale, w_bias = _uvw_scale_bias(self.image_parameters, self.grid_parameters)
        w_bias += self._w
        with profile_device(self.command_queue, 'predict'):
            _lib.call('kib_predict', self.buffer('vis').ptr, self.buffer('uv').ptr,
                      self.buffer('w_plane').ptr, self.buffer('weights').ptr,
                      self.buffer('lmn').ptr, self.buffer('flux').ptr,
                      self.num_vis, self.num_sources, self.template.num_polarizations,
                      self.grid_parameters.fixed.oversample,
                      float(np.float32(uv_scale)), float(np.float32(w_scale)),
                      float(np.float32(w_bias)), self.command_queue.stream)
