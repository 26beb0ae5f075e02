"""Restoring beam: Gaussian fit to the PSF core and convolution of the CLEAN model with it
(reference katsdpimager/beam.py).

Same surface as the reference: :class:`Beam`, :func:`fit_beam`, :func:`convolve_beam` (host),
``FourierBeamTemplate`` / ``ConvolveBeamTemplate`` with slots **image** and **fourier** and a
``beam`` attribute (beam.py:204-398), used by ``frontend.process_channel`` after the last major
cycle (frontend.py:623-641).  Differences:

* the reference fits with astropy's Levenberg-Marquardt fitter on an ``astropy`` model
  (beam.py:91-156); astropy is not a dependency here, so :class:`Gaussian2D` carries the three
  free parameters (amplitude and means fixed, as the reference fixes them) and
  ``scipy.optimize.least_squares(method='lm')`` -- the same MINPACK algorithm -- does the fit;
* :func:`restore` convolves every polarization of an :class:`~.imaging.Imaging` model without
  the per-plane staging copies of frontend.py:632-636 beyond one scratch plane.
"""
import math

import numpy as np

from . import _lib, accel, fft
from .profiling import profile_device, profile_function
from .types import real_to_complex


class Gaussian2D:
    """Elliptical Gaussian ``amplitude * exp(-(a x^2 + b x y + c y^2))`` centred on the origin,
    parameterised as astropy's ``models.Gaussian2D`` (standard deviations along the axes of
    the ellipse and the angle of the first axis)."""

    class _Value:
        def __init__(self, value):
            self.value = value

        def __float__(self):
            return float(self.value)

    def __init__(self, amplitude=1.0, x_stddev=1.0, y_stddev=1.0, theta=0.0):
        self.amplitude = float(amplitude)
        self.x_stddev = Gaussian2D._Value(float(x_stddev))
        self.y_stddev = Gaussian2D._Value(float(y_stddev))
        self.theta = Gaussian2D._Value(float(theta))

    @staticmethod
    def evaluate(x, y, amplitude, x_stddev, y_stddev, theta):
        cost2 = math.cos(theta) ** 2
        sint2 = math.sin(theta) ** 2
        sin2t = math.sin(2.0 * theta)
        xstd2 = x_stddev ** 2
        ystd2 = y_stddev ** 2
        a = 0.5 * (cost2 / xstd2 + sint2 / ystd2)
        b = 0.5 * (sin2t / xstd2 - sin2t / ystd2)
        c = 0.5 * (sint2 / xstd2 + cost2 / ystd2)
        return amplitude * np.exp(-(a * x * x + b * x * y + c * y * y))

    def __call__(self, x, y):
        return self.evaluate(np.asarray(x, np.float64), np.asarray(y, np.float64), self.amplitude,
                             self.x_stddev.value, self.y_stddev.value, self.theta.value)


class Beam:
    """Gaussian synthesised beam (reference beam.py:49-88): ``major`` / ``minor`` are full
    widths at half maximum in the units of the fit, ``theta`` (radians) is measured from the
    positive axis 0 of the PSF towards positive axis 1."""

    def __init__(self, model):
        self.model = model
        scale = math.sqrt(8 * math.log(2))
        self.major = model.x_stddev.value * scale
        self.minor = model.y_stddev.value * scale
        theta = model.theta.value
        if self.major < self.minor:
            self.minor, self.major = self.major, self.minor
            theta += math.pi / 2
        self.theta = theta % math.pi

    def __repr__(self):
        return 'Beam({0.major!r}, {0.minor!r}, {0.theta!r})'.format(self)


@profile_function()
def fit_beam(psf, step=1.0, threshold=0.01, init_threshold=0.5):
    """Fit a 2-D Gaussian to the central part of a PSF (reference beam.py:91-156): initial
    guess from the second moments of the pixels above `init_threshold` (corrected for the
    truncation), then a least-squares fit to every pixel above `threshold` with amplitude 1 and
    the centre fixed at the origin."""
    import scipy.optimize
    psf = np.asarray(psf, np.float64)

    def extract(level):
        mask = psf > level
        rows, cols = np.nonzero(mask)
        return psf[mask], (rows - psf.shape[0] // 2) * step, (cols - psf.shape[1] // 2) * step

    picked, x, y = extract(init_threshold)
    total = np.sum(picked)
    cov = np.empty((2, 2))
    cov[0, 0] = np.sum(picked * x ** 2) / total
    cov[0, 1] = cov[1, 0] = np.sum(picked * x * y) / total
    cov[1, 1] = np.sum(picked * y ** 2) / total
    # variance of a standard 2-D Gaussian truncated at radius R: 1 - (1 + R^2 / 2) exp(-R^2 / 2)
    r2 = -2 * math.log(init_threshold)
    cov /= 1 - (1 + 0.5 * r2) * math.exp(-0.5 * r2)
    eigvals, eigvecs = np.linalg.eigh(cov)
    eigvals = np.maximum(eigvals, 1e-12)
    # as astropy builds a model from a covariance matrix: first axis = largest eigenvalue
    order = np.argsort(eigvals)[::-1]
    x_std, y_std = np.sqrt(eigvals[order])
    vec = eigvecs[:, order[0]]
    theta = math.atan2(vec[1], vec[0])

    picked, x, y = extract(threshold)

    def residual(params):
        return Gaussian2D.evaluate(x, y, 1.0, params[0], params[1], params[2]) - picked

    result = scipy.optimize.least_squares(residual, [x_std, y_std, theta], method='lm')
    x_std, y_std, theta = result.x
    return Beam(Gaussian2D(1.0, abs(x_std), abs(y_std), theta))


def beam_covariance_sqrt(beam):
    """Square root M of the beam's covariance matrix (beam.py:159-168)."""
    model = beam.model
    c, s = math.cos(model.theta.value), math.sin(model.theta.value)
    q = np.array([[c, -s], [s, c]])
    d = np.diag([model.x_stddev.value, model.y_stddev.value])
    return q @ d @ q.T


def convolve_beam(model, beam, out=None):
    """Host restoration (reference beam.py:171-201): FFT convolution with the analytic
    transform of the beam."""
    if out is None:
        out = np.empty_like(model)
    model_ft = np.fft.fftn(model, axes=[1, 2])
    M = beam_covariance_sqrt(beam)
    amplitude = 2 * np.pi * beam.model.amplitude * np.abs(np.linalg.det(M))
    u = np.fft.fftfreq(model.shape[1])
    v = np.fft.fftfreq(model.shape[2])
    coords = np.stack(np.meshgrid(u, v, indexing='ij'), axis=-1)
    rotated = np.inner(coords, M)
    beam_ft = amplitude * np.exp(-2.0 * np.pi ** 2 * np.sum(rotated ** 2, axis=-1))
    out[:] = np.fft.ifftn(model_ft * beam_ft[np.newaxis, ...], axes=[1, 2]).real
    return out


class FourierBeamTemplate:
    """Multiply the half-complex transform of an image by the transform of a Gaussian beam
    (reference beam.py:204-238)."""

    def __init__(self, context, dtype, tuning=None):
        _lib.load()
        self.context = context
        self.dtype = np.dtype(dtype)

    def instantiate(self, *args, **kwargs):
        return FourierBeam(self, *args, **kwargs)


class FourierBeam(accel.Operation):
    """.. rubric:: Slots

    **data** : complex, height x (width // 2 + 1): real-to-complex transform of the image,
    multiplied in place.  ``beam`` must be set before the call.
    """

    def __init__(self, template, command_queue, image_shape, allocator=None):
        if len(image_shape) != 2:
            raise ValueError('image_shape must be 2D')
        super().__init__(command_queue, allocator)
        self.template = template
        self.image_shape = tuple(image_shape)
        self.slots['data'] = accel.IOSlot((image_shape[0], image_shape[1] // 2 + 1),
                                          real_to_complex(template.dtype))
        self.beam = None

    def _run(self):
        if self.beam is None:
            raise ValueError('Must set beam')
        M = beam_covariance_sqrt(self.beam)
        amplitude = 2 * np.pi * self.beam.model.amplitude * abs(np.linalg.det(M))
        # cuFFT does not normalise the inverse transform: folded into the amplitude
        amplitude /= self.image_shape[0] * self.image_shape[1]
        # integer frequencies -> cycles per pixel, folded into the matrix
        M = M @ np.diag([1.0 / self.image_shape[0], 1.0 / self.image_shape[1]])
        C = -2 * np.pi ** 2 * M.T @ M
        data = self.buffer('data')
        with profile_device(self.command_queue, 'fourier_beam'):
            _lib.call('kib_fourier_beam', data.ptr, data.padded_shape[1], float(amplitude),
                      float(C[0, 0]), float(2 * C[0, 1]), float(C[1, 1]),
                      data.shape[1], data.shape[0], _lib.dtype_code(self.template.dtype),
                      self.command_queue.stream)


class ConvolveBeamTemplate:
    """Convolution of one image plane with a Gaussian restoring beam by real FFTs
    (reference beam.py:304-338)."""

    def __init__(self, context, shape, dtype,
                 padded_shape_image=None, padded_shape_fourier=None, tuning=None):
        if padded_shape_image is None:
            padded_shape_image = tuple(shape)
        if padded_shape_fourier is None:
            padded_shape_fourier = tuple(shape[:-1]) + (shape[-1] // 2 + 1,)
        if len(shape) != 2 or len(padded_shape_image) != 2 or len(padded_shape_fourier) != 2:
            raise ValueError('wrong number of dimensions')
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        complex_dtype = real_to_complex(self.dtype)
        self.fft = fft.FftTemplate(context, 2, shape, dtype, complex_dtype,
                                   padded_shape_image, padded_shape_fourier)
        self.ifft = fft.FftTemplate(context, 2, shape, complex_dtype, dtype,
                                    padded_shape_fourier, padded_shape_image)
        self.fourier_beam = FourierBeamTemplate(context, dtype, tuning)

    def instantiate(self, *args, **kwargs):
        return ConvolveBeam(self, *args, **kwargs)


class ConvolveBeam(accel.OperationSequence):
    """.. rubric:: Slots

    **image** : real, height x width, convolved in place;  **fourier** : complex scratch,
    height x (width // 2 + 1).  ``beam`` must be set before the call (beam.py:341-398).
    """

    def __init__(self, template, command_queue, allocator=None):
        self._fft = template.fft.instantiate(command_queue, fft.FftMode.FORWARD, allocator)
        self._ifft = template.ifft.instantiate(command_queue, fft.FftMode.INVERSE, allocator)
        self._fourier_beam = template.fourier_beam.instantiate(
            command_queue, template.shape, allocator)
        operations = [('fft', self._fft), ('fourier_beam', self._fourier_beam),
                      ('ifft', self._ifft)]
        compounds = {'image': ['fft:src', 'ifft:dest'],
                     'fourier': ['fft:dest', 'ifft:src', 'fourier_beam:data']}
        super().__init__(command_queue, operations, compounds, allocator=allocator)

    @property
    def beam(self):
        return self._fourier_beam.beam

    @beam.setter
    def beam(self, value):
        self._fourier_beam.beam = value


def extract_psf(queue, psf, psf_patch, out=None):
    """Central `psf_patch` = (rows, cols) of polarization 0 of a device PSF on the host
    (reference frontend.extract_psf, frontend.py:152-173); `out` may supply the pinned
    destination."""
    y0 = (psf.shape[1] - psf_patch[0]) // 2
    x0 = (psf.shape[2] - psf_patch[1]) // 2
    if out is None:
        out = accel.HostArray((psf_patch[0], psf_patch[1]), psf.dtype, context=queue.context)
    psf.get_region(queue, out, np.s_[0, y0:y0 + psf_patch[0], x0:x0 + psf_patch[1]], np.s_[:, :])
    return out


class Restorer:
    """The restore step of ``frontend.process_channel`` for :func:`~.pipeline.process_channel`:
    ``extract`` copies the PSF core to the host right after the PSF patch is known (where the
    reference fits the beam, frontend.py:529-531), ``fit`` runs the host-side Gaussian fit --
    the pipeline calls it after enqueuing the first dirty pass, so the ~10 ms fit hides behind
    device work -- and calling the object convolves every polarization of the model with the
    beam (frontend.py:623-636).  One scratch plane, one half-complex plane and the pinned core
    buffer are allocated on first use and kept."""

    def __init__(self, context):
        self.context = context
        self._op = None
        self._shape = None
        self._core = None
        self._pending = False
        self.beam = None

    def extract(self, imager, psf_patch):
        queue = imager.command_queue
        psf = imager.buffer('psf')
        core_shape = (int(psf_patch[1]), int(psf_patch[2]))
        if self._core is None or self._core.shape != core_shape or self._core.dtype != psf.dtype:
            self._core = accel.HostArray(core_shape, psf.dtype, context=queue.context)
        extract_psf(queue, psf, core_shape, self._core)
        self._pending = True

    def fit(self):
        if self._pending:
            self.beam = fit_beam(self._core)
            self._pending = False
        return self.beam

    def __call__(self, imager, psf_patch):
        queue = imager.command_queue
        model = imager.buffer('model')
        if self.beam is None and not self._pending:
            self.extract(imager, psf_patch)
        self.fit()
        shape = tuple(model.shape[1:])
        if self._op is None or self._shape != shape or self._op.command_queue is not queue:
            template = ConvolveBeamTemplate(self.context, shape, model.dtype)
            self._op = template.instantiate(queue)
            self._op.ensure_all_bound()
            self._shape = shape
        self._op.beam = self.beam
        plane = self._op.buffer('image')
        for pol in range(model.shape[0]):
            model.copy_region(queue, plane, np.s_[pol], ())
            self._op()
            plane.copy_region(queue, model, (), np.s_[pol])
        beam, self.beam = self.beam, None        # the next channel has its own PSF
        self.last_beam = beam
        return beam
