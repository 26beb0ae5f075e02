"""Imaging weights: the reference's known-answer tests (reference
katsdpimager/test/test_weight.py) and WeightsHost golden vectors."""
import numpy as np
import pytest

from katsdpimager_b200 import weight
from tests import cases
from tests.cases import load_golden

pytestmark = pytest.mark.gpu


def _set_partial(queue, buffer, data):
    rs = np.random.RandomState(1)
    host = buffer.empty_like()
    host[:] = rs.uniform(size=host.shape) * 100
    host[tuple(np.s_[0:x] for x in data.shape)] = data
    buffer.set(queue, host)


def test_grid_weights_known_answer(gpu):
    context, queue = gpu
    grid_shape = (4, 100, 200)
    uv = np.array([[-10, 5, 0, 0], [23, 17, 0, 0], [-10, 5, 0, 0], [-10, 5, 0, 0],
                   [-10, 6, 0, 0], [-11, 5, 0, 0]], np.int16)
    weights = np.array([[1.0, 10.0, 100.0, 1000.0], [2.0, 20.0, 200.0, 2000.0],
                        [4.0, 40.0, 400.0, 4000.0], [8.0, 80.0, 800.0, 8000.0],
                        [16.0, 160.0, 1600.0, 16000.0], [32.0, 320.0, 3200.0, 32000.0]],
                       np.float32)
    fn = weight.GridWeightsTemplate(context, 4).instantiate(queue, grid_shape, 1000)
    fn.ensure_all_bound()
    _set_partial(queue, fn.buffer('uv'), uv)
    _set_partial(queue, fn.buffer('weights'), weights)
    fn.buffer('grid').zero(queue)
    fn.num_vis = len(uv)
    fn()
    expected = np.zeros(grid_shape, np.float32)
    for i in range(4):
        expected[i, 55, 90] = 13 * 10**i
        expected[i, 67, 123] = 2 * 10**i
        expected[i, 56, 90] = 16 * 10**i
        expected[i, 55, 89] = 32 * 10**i
    np.testing.assert_equal(expected, fn.buffer('grid').get(queue))
    with pytest.raises(ValueError):
        weight.GridWeightsTemplate(context, 4).instantiate(queue, (4, 99, 200), 10)


def test_density_weights(gpu):
    """reference test_weight.py TestDensityWeights."""
    context, queue = gpu
    rs = np.random.RandomState(1)
    grid_shape = (4, 50, 107)
    data = np.zeros(grid_shape, np.float32)
    expected = np.zeros(grid_shape, np.float32)
    sum_w = sum_dw = sum_d2w = 0.0
    for index in rs.choice(data.size, 100, replace=False):
        w = rs.uniform(low=0.1, high=2.0)
        d = 1.0 / (2.5 * w + 1.75)
        data.flat[index] = w
        expected.flat[index] = d
        if index < data[0].size:
            sum_w += w
            sum_dw += d * w
            sum_d2w += d**2 * w
    fn = weight.DensityWeightsTemplate(context, 4).instantiate(queue, grid_shape)
    fn.ensure_all_bound()
    fn.a = 2.5
    fn.b = 1.75
    fn.buffer('grid').set(queue, data)
    rms, normalized_rms = fn()
    np.testing.assert_allclose(expected, fn.buffer('grid').get(queue), 1e-5, 1e-5)
    np.testing.assert_allclose(np.sqrt(sum_d2w) / sum_dw, rms, 1e-6)
    np.testing.assert_allclose(np.sqrt(sum_d2w * sum_w) / sum_dw, normalized_rms, 1e-6)


def test_mean_weight(gpu):
    context, queue = gpu
    rs = np.random.RandomState(1)
    data = rs.uniform(size=(4, 50, 107)).astype(np.float32)
    fn = weight.MeanWeightTemplate(context).instantiate(queue, data.shape)
    fn.ensure_all_bound()
    fn.buffer('grid').set(queue, data)
    np.testing.assert_allclose(np.sum(data[0] * data[0]) / np.sum(data[0]), fn(), rtol=1e-5)


@pytest.mark.parametrize('name,weight_type', [('uniform', weight.WeightType.UNIFORM),
                                              ('robust', weight.WeightType.ROBUST),
                                              ('natural', weight.WeightType.NATURAL)])
def test_weights_golden(gpu, name, weight_type):
    """The compound Weights operation against the reference's WeightsHost."""
    context, queue = gpu
    fx = cases.weights_case()
    golden = load_golden('weights_' + name)
    fn = weight.WeightsTemplate(context, weight_type, 2).instantiate(queue, fx['shape'], 1000)
    fn.ensure_all_bound()
    if weight_type == weight.WeightType.ROBUST:
        fn.robustness = fx['robustness']
    fn.clear()
    if weight_type != weight.WeightType.NATURAL:
        n = len(fx['uv4'])
        for start in (0, 300):      # two batches
            stop = min(n, start + 300) if start == 0 else n
            count = stop - start
            fn.buffer('uv').set_region(queue, fx['uv4'][start:stop], np.s_[:count], np.s_[:])
            fn.buffer('weights').set_region(queue, fx['weights'][start:stop], np.s_[:count], np.s_[:])
            fn.grid(count)
    else:
        assert 'uv' not in fn.slots and 'weights' not in fn.slots
    rms, normalized_rms = fn.finalize()
    np.testing.assert_allclose(fn.buffer('grid').get(queue), golden['grid'], rtol=2e-6)
    if weight_type == weight.WeightType.NATURAL:
        assert rms is None and normalized_rms == 1.0
    else:
        np.testing.assert_allclose(rms, golden['rms'], rtol=1e-5)
        np.testing.assert_allclose(normalized_rms, golden['normalized_rms'], rtol=1e-5)
