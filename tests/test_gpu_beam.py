"""Device restoration (ConvolveBeam: cuFFT real transforms + kib_fourier_beam) against the
reference's unit-test expectation (reference test_beam.py:52-62, rtol = atol = 1e-5) and the
host implementation; the restore step inside the channel pipeline."""
import numpy as np
import pytest

from katsdpimager_b200 import beam, imaging, parameters as prm, pipeline, weight
from tests import cases
from tests.test_beam import reference_case

pytestmark = pytest.mark.gpu


def test_convolve_device(gpu):
    context, queue = gpu
    b, model, expected = reference_case()
    template = beam.ConvolveBeamTemplate(context, model.shape[1:], model.dtype)
    fn = template.instantiate(queue)
    fn.ensure_all_bound()
    assert fn.buffer('fourier').shape == (128, 65)
    for pol in range(model.shape[0]):
        fn.buffer('image').set(queue, model[pol])
        fn.beam = b
        fn()
        actual = fn.buffer('image').get(queue)
        np.testing.assert_allclose(expected[pol], actual, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        fresh = template.instantiate(queue)
        fresh.ensure_all_bound()
        fresh()                                  # beam not set


def test_convolve_device_large_matches_host(gpu):
    context, queue = gpu
    rs = np.random.RandomState(2)
    n = 2048
    model = np.zeros((1, n, n), np.float32)
    model[0, rs.randint(0, n, 300), rs.randint(0, n, 300)] = rs.uniform(0.1, 3.0, 300)
    b = beam.Beam(beam.Gaussian2D(1.0, 2.7, 1.9, 0.8))
    expected = beam.convolve_beam(model.astype(np.float64), b)
    fn = beam.ConvolveBeamTemplate(context, (n, n), np.float32).instantiate(queue)
    fn.ensure_all_bound()
    fn.buffer('image').set(queue, model[0])
    fn.beam = b
    fn()
    actual = fn.buffer('image').get(queue)
    assert np.abs(actual - expected[0]).max() <= 2e-6 * np.abs(expected).max()


def test_restore_in_pipeline(gpu):
    """process_channel with the restore step: final image = residual + model convolved with
    the beam fitted to the PSF (frontend.py:623-641)."""
    context, queue = gpu
    fx = cases.imaging_case()
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.UNIFORM)
    slices = [fx['reader']._data[0][w] for w in range(gp.w_slices)]
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    outs = []
    for restorer in (None, beam.Restorer(context)):
        imager = template.instantiate(queue, ip, gp, fx['vis_block'], 0, 2)
        imager.ensure_all_bound()
        vis = pipeline.ResidentVisibilities(queue, slices, len(ip.fixed.polarizations))
        pipeline.process_channel(imager, vis, ip, gp, cp, wp, 2, fx['vis_block'], restore=restorer)
        outs.append((imager.get_buffer('dirty'), imager.get_buffer('model')))
    (plain, model), (restored, restored_model) = outs
    b = restorer.last_beam
    assert 1.5 < b.minor <= b.major < 20
    expected_model = beam.convolve_beam(model.astype(np.float64), b)
    np.testing.assert_allclose(restored_model, expected_model, rtol=0,
                               atol=2e-5 * np.abs(expected_model).max())
    # residuals are the same in both runs up to gridding round-off
    residual_plain = plain - model
    residual_restored = restored - restored_model
    peak = np.abs(plain).max()
    assert np.sqrt(np.mean((residual_plain - residual_restored) ** 2)) < 1e-4 * peak
