"""Direct prediction against the reference's _predict_host (golden) and a float64 DFT."""
import numpy as np
import pytest

from katsdpimager_b200 import predict
from tests import cases
from tests.cases import load_golden

pytestmark = pytest.mark.gpu


def _truth(fx):
    u = (fx['uv'][:, 0].astype(np.float64) * fx['oversample'] + fx['sub_uv'][:, 0] + 0.5) * fx['uv_scale']
    v = (fx['uv'][:, 1].astype(np.float64) * fx['oversample'] + fx['sub_uv'][:, 1] + 0.5) * fx['uv_scale']
    w = fx['w_plane'].astype(np.float64) * fx['w_scale'] + fx['w_bias']
    lmn = fx['lmn'].astype(np.float64)
    phase = np.outer(u, lmn[:, 0]) + np.outer(v, lmn[:, 1]) + np.outer(w, lmn[:, 2])
    model = np.exp(-2j * np.pi * phase) @ fx['flux'].astype(np.float64)
    return fx['vis'] - fx['weights'] * model


@pytest.mark.parametrize('n_vis', [300, 2])
def test_predict(gpu, n_vis):
    context, queue = gpu
    fx = cases.predict_case(n_vis=n_vis)
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    pols = len(ip.fixed.polarizations)
    fn = predict.PredictTemplate(context, np.float32, pols).instantiate(
        queue, ip, gp, n_vis + 10, 16)
    fn.ensure_all_bound()
    fn.num_vis = n_vis
    fn.buffer('uv').set_region(queue, np.concatenate((fx['uv'], fx['sub_uv']), axis=1),
                               np.s_[:n_vis], np.s_[:])
    fn.buffer('w_plane').set_region(queue, fx['w_plane'], np.s_[:n_vis], np.s_[:])
    fn.buffer('vis').set_region(queue, fx['vis'], np.s_[:n_vis], np.s_[:])
    fn.buffer('weights').set_region(queue, fx['weights'], np.s_[:n_vis], np.s_[:])
    fn.set_sources(fx['lmn'], fx['flux'])
    fn.set_w(1.2)
    fn()
    actual = fn.buffer('vis').get(queue)[:n_vis]
    truth = _truth(fx)
    # Phases reach ~1000 turns, so single-precision coordinates limit both the device
    # and the host path to ~1e-3; the device (range-reduced) must not be worse.
    np.testing.assert_allclose(actual, truth, rtol=0, atol=2e-3)
    if n_vis == 300:
        golden = load_golden('predict_small')
        # the reference's own device-vs-host tolerance (test_predict.py:92)
        np.testing.assert_allclose(actual, golden['residual'], rtol=5e-4, atol=5e-4)
        assert np.abs(actual - truth).max() <= np.abs(golden['residual'] - truth).max() * 1.5


def test_too_many_sources(gpu):
    context, queue = gpu
    fx = cases.predict_case()
    fn = predict.PredictTemplate(context, np.float32, 3).instantiate(
        queue, fx['image_parameters'], fx['grid_parameters'], 10, 3)
    with pytest.raises(ValueError):
        fn.set_sources(fx['lmn'], fx['flux'])
    with pytest.raises(ValueError):
        predict.PredictTemplate(context, np.float32, 2).instantiate(
            queue, fx['image_parameters'], fx['grid_parameters'], 10, 3)
