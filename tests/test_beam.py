"""Restoring beam: the reference's own unit test (reference katsdpimager/test/test_beam.py:12-62:
FFT convolution with the analytic beam transform against a sampled-beam convolution,
rtol = atol = 1e-5), on the host here and on the device in tests/test_gpu_beam.py; plus the
Gaussian fit, which the reference leaves untested."""
import math

import numpy as np
import pytest
import scipy.signal

from katsdpimager_b200 import beam


def convolve_beam_reference(model, beam_):
    """Sampled-beam convolution without wrap-around (test_beam.py:12-31)."""
    hh = model.shape[1] // 2
    hw = model.shape[2] // 2
    m = np.arange(-hh + 1, hh)
    l = np.arange(-hw + 1, hw)
    beam_pixels = beam_.model(*np.meshgrid(m, l, indexing='ij'))
    out = np.empty_like(model)
    for pol in range(model.shape[0]):
        out[pol, ...] = scipy.signal.fftconvolve(model[pol], beam_pixels, 'same')
    return out


def reference_case():
    """test_beam.py:36-46."""
    b = beam.Beam(beam.Gaussian2D(amplitude=3.5, x_stddev=2.0, y_stddev=5.0, theta=1))
    model = np.zeros((4, 128, 128), np.float32)
    model[0, 32, 80] = 1.0
    model[0, 100, 40] = 2.0
    model[1, 50, 60] = 3.0
    model[2, 64, 64] = 4.0
    model[2, 80, 64] = 3.0
    return b, model, convolve_beam_reference(model, b)


def test_convolve_host():
    b, model, expected = reference_case()
    actual = beam.convolve_beam(model, b)
    np.testing.assert_allclose(expected, actual, rtol=1e-5, atol=1e-5)


def test_beam_swaps_axes():
    b = beam.Beam(beam.Gaussian2D(1.0, 2.0, 5.0, 0.25))
    scale = math.sqrt(8 * math.log(2))
    assert b.major == pytest.approx(5.0 * scale) and b.minor == pytest.approx(2.0 * scale)
    assert b.theta == pytest.approx(0.25 + math.pi / 2)


@pytest.mark.parametrize('x_std,y_std,theta', [(3.0, 3.0, 0.0), (4.0, 2.5, 0.4), (2.0, 6.0, 2.0)])
def test_fit_recovers_gaussian(x_std, y_std, theta):
    n = 65
    rows, cols = np.meshgrid(np.arange(n) - n // 2, np.arange(n) - n // 2, indexing='ij')
    truth = beam.Gaussian2D(1.0, x_std, y_std, theta)
    rs = np.random.RandomState(1)
    psf = truth(rows, cols) + 1e-4 * rs.standard_normal((n, n))
    fitted = beam.fit_beam(psf.astype(np.float32))
    expected = beam.Beam(truth)
    assert fitted.major == pytest.approx(expected.major, rel=2e-3)
    assert fitted.minor == pytest.approx(expected.minor, rel=2e-3)
    if abs(x_std - y_std) > 1e-6:
        delta = (fitted.theta - expected.theta + math.pi / 2) % math.pi - math.pi / 2
        assert abs(delta) < 5e-3
    # the fitted model reproduces the PSF where it matters
    assert np.abs(fitted.model(rows, cols) - truth(rows, cols)).max() < 2e-3
