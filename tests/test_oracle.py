"""The CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  These pin the oracle; the GPU tests then compare the
CUDA path with the oracle and with the same golden vectors."""
import numpy as np
import pytest

from tests import cases
from tests.cases import load_golden


def test_expj2pi(oracle):
    x = np.linspace(-20.3, 17.9, 1001)
    np.testing.assert_allclose(oracle.expj2pi(x), np.exp(2j * np.pi * x), atol=1e-13)
    x32 = x.astype(np.float32)
    assert oracle.expj2pi(x32).dtype == np.complex64


@pytest.mark.parametrize('name', ['test_grid', 'meerkat_k7'])
def test_convolution_kernel(oracle, name):
    ip, gp = cases.lut_cases()[name]
    golden = load_golden('lut_' + name)
    lut = oracle.convolution_kernel(ip, gp)
    assert lut.shape == golden['data'].shape
    np.testing.assert_allclose(lut, golden['data'], rtol=0, atol=2e-7)
    np.testing.assert_allclose(oracle.taper(gp, ip.pixels), golden['taper'], rtol=1e-12)


def test_grid_reference_fixture(oracle):
    """GridderHost / DegridderHost outputs on the reference's own unit-test inputs."""
    fx = cases.reference_grid_fixture()
    golden = load_golden('grid_reference_fixture')
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    lut = oracle.convolution_kernel(ip, gp)
    values = np.zeros((4, ip.pixels, ip.pixels), np.complex128)
    wgrid = np.zeros(values.shape, np.float32)
    cases.middle(wgrid, fx['weights_grid'].shape)[:] = fx['weights_grid']
    oracle.grid(lut, values, wgrid, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['vis'])
    np.testing.assert_allclose(values[:, ::3, :], golden['grid_rows'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(values.sum(axis=(1, 2)), golden['grid_sum'], rtol=1e-9)
    residual = fx['degrid_vis'].copy()
    oracle.degrid(lut, np.ascontiguousarray(fx['degrid_grid']), fx['uv'], fx['sub_uv'],
                  fx['w_plane'], fx['degrid_weights'], residual)
    np.testing.assert_allclose(residual, golden['residual'], rtol=1e-6, atol=1e-6)


def test_grid_reference_first_principles(oracle):
    """The expected-value computation of reference test_grid.py:91-112 (numpy, double)."""
    fx = cases.reference_grid_fixture()
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    lut = oracle.convolution_kernel(ip, gp)
    pixels = ip.pixels
    actual = np.zeros((4, pixels, pixels), np.complex128)
    wgrid = np.zeros(actual.shape, np.float32)
    cases.middle(wgrid, fx['weights_grid'].shape)[:] = fx['weights_grid']
    oracle.grid(lut, actual, wgrid, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['vis'])
    expected = np.zeros_like(actual)
    uv_bias = (lut.shape[-1] - 1) // 2 - pixels // 2
    wshape = fx['weights_grid'].shape
    for i in range(len(fx['uv'])):
        kernel = np.conj(np.outer(lut[fx['w_plane'][i], fx['sub_uv'][i, 1], :],
                                  lut[fx['w_plane'][i], fx['sub_uv'][i, 0], :]))
        u = fx['uv'][i, 0] - uv_bias
        v = fx['uv'][i, 1] - uv_bias
        wu = fx['uv'][i, 0] + wshape[2] // 2
        wv = fx['uv'][i, 1] + wshape[1] // 2
        for j in range(4):
            expected[j, v:v + kernel.shape[0], u:u + kernel.shape[1]] += \
                fx['vis'][i, j].astype(np.complex128) * fx['weights_grid'][j, wv, wu] * kernel
    np.testing.assert_allclose(expected, actual, 1e-5, 1e-8)


def test_grid_small(oracle):
    fx = cases.small_grid_case()
    golden = load_golden('grid_small')
    lut = oracle.convolution_kernel(fx['image_parameters'], fx['grid_parameters'])
    values = np.zeros(golden['grid'].shape, np.complex64)
    oracle.grid(lut, values, fx['weights_grid'], fx['uv'], fx['sub_uv'], fx['w_plane'], fx['vis'])
    scale = np.abs(golden['grid']).max()
    np.testing.assert_allclose(values, golden['grid'], rtol=0, atol=2e-6 * scale)
    residual = fx['vis'].copy()
    oracle.degrid(lut, golden['grid'], fx['uv'], fx['sub_uv'], fx['w_plane'], fx['weights'],
                  residual)
    np.testing.assert_allclose(residual, golden['residual'], rtol=0,
                               atol=2e-6 * np.abs(golden['residual']).max())


def test_image(oracle):
    fx = cases.image_case()
    golden = load_golden('image_small')
    image = np.zeros(fx['grid'].shape, np.float32)
    oracle.grid_to_image(fx['grid'], image, fx['kernel1d'], fx['lm_scale'], fx['lm_bias'], fx['w'])
    np.testing.assert_allclose(image, golden['image'], rtol=0,
                               atol=1e-6 * np.abs(golden['image']).max())
    # cropped (device-style) grid gives the same image
    size = fx['grid_size']
    crop = np.ascontiguousarray(cases.middle(fx['grid'], (2, size, size)))
    image2 = np.zeros_like(image)
    oracle.grid_to_image(crop, image2, fx['kernel1d'], fx['lm_scale'], fx['lm_bias'], fx['w'])
    np.testing.assert_array_equal(image, image2)
    back = oracle.image_to_grid(fx['model'], fx['kernel1d'], fx['lm_scale'], fx['lm_bias'], fx['w'])
    np.testing.assert_allclose(back, golden['grid_from_model'], rtol=0,
                               atol=1e-6 * np.abs(golden['grid_from_model']).max())


@pytest.mark.parametrize('name', ['clean_i', 'clean_sumsq'])
def test_clean(oracle, name):
    """Bit-exact CLEAN: component positions, values, fluxes, tiles and residual."""
    fx = cases.clean_case(name)
    golden = load_golden(name)
    cp = fx['clean_parameters']
    assert oracle.psf_patch(fx['psf'], cp.psf_cutoff, cp.psf_limit) == tuple(golden['patch'])
    assert oracle.noise_est(fx['dirty'], cp.border) == golden['noise']
    dirty = fx['dirty'].copy()
    model = np.zeros_like(dirty)
    cleaner = oracle.CleanHost(fx['image_parameters'].pixels, cp.border, cp.mode, cp.loop_gain,
                               dirty, fx['psf'], model)
    cleaner.reset()
    np.testing.assert_array_equal(cleaner.tile_max, golden['tile_max0'])
    np.testing.assert_array_equal(cleaner.tile_pos, golden['tile_pos0'])
    values, positions, reported, pixels = [], [], [], []
    for _ in range(fx['cycles']):
        value, pos, pixel = cleaner(fx['psf_patch'], fx['threshold'])
        if value is None:
            break
        values.append(value)
        positions.append(pos)
        reported.append(cleaner.reported_pos)
        pixels.append(pixel)
    assert len(values) == len(golden['values'])
    assert len(values) < fx['cycles']       # the threshold, not the cycle limit, ended it
    # The reference returns a position aliased to the tile's new peak (oracle.c
    # kor_clean_cycle); the true component positions are pinned through the model image.
    np.testing.assert_array_equal(np.array(reported, np.int32), golden['positions'])
    assert np.any(np.array(reported) != np.array(positions))
    expected_model = np.zeros_like(model)
    for pos, pixel in zip(positions, pixels):
        expected_model[:, pos[0], pos[1]] += pixel
    np.testing.assert_array_equal(model, expected_model)
    np.testing.assert_array_equal(model, golden['model'])
    np.testing.assert_array_equal(np.array(values, np.float32), golden['values'])
    np.testing.assert_array_equal(np.array(pixels, np.float32), golden['pixels'])
    np.testing.assert_array_equal(cleaner.tile_max, golden['tile_max'])
    np.testing.assert_array_equal(cleaner.tile_pos, golden['tile_pos'])
    np.testing.assert_array_equal(dirty, golden['residual'])


def test_psf_patch_known_answers(oracle):
    """Known answers of reference test_clean.py:11-37."""
    def fresh():
        psf = np.zeros((4, 206, 304), np.float32)
        psf[:, 103, 152] = 1.0
        return psf
    assert oracle.psf_patch(fresh(), 0.01) == (4, 1, 1)
    psf = fresh()
    psf[0, 0, 0] = 0.1
    assert oracle.psf_patch(psf, 0.01) == (4, 206, 304)
    psf = fresh()
    psf[3, 205, 303] = -0.2
    assert oracle.psf_patch(psf, 0.01) == (4, 205, 303)
    psf = fresh()
    psf[0, 0, 0] = 0.4
    psf[3, 205, 303] = 0.3
    psf[1, 110, 150] = 0.2
    assert oracle.psf_patch(psf, 0.01, limit=50 / 206) == (4, 15, 5)


@pytest.mark.parametrize('name', ['uniform', 'robust', 'natural'])
def test_weights(oracle, name):
    fx = cases.weights_case()
    golden = load_golden('weights_' + name)
    wgrid = np.zeros(fx['shape'], np.float32)
    weights = oracle.WeightsHost({'natural': 0, 'uniform': 1, 'robust': 2}[name], wgrid)
    weights.robustness = fx['robustness']
    weights.clear()
    uv = fx['uv'].copy()
    weights.grid(uv, fx['weights'])
    np.testing.assert_array_equal(uv, fx['uv'])
    rms, normalized_rms = weights.finalize()
    np.testing.assert_allclose(wgrid, golden['grid'], rtol=1e-6)
    if name == 'natural':
        assert rms is None and normalized_rms == 1.0
    else:
        np.testing.assert_allclose(rms, golden['rms'], rtol=1e-6)
        np.testing.assert_allclose(normalized_rms, golden['normalized_rms'], rtol=1e-6)


def test_weights_known_answers(oracle):
    """Known answers of reference test_weight.py:10-57."""
    uv = np.array([[-10, 5], [23, 17], [-10, 5], [-10, 5], [-10, 6], [-11, 5]], np.int16)
    w = np.array([[1, 10, 100, 1000], [2, 20, 200, 2000], [4, 40, 400, 4000],
                  [8, 80, 800, 8000], [16, 160, 1600, 16000], [32, 320, 3200, 32000]], np.float32)
    wgrid = np.zeros((4, 100, 200), np.float32)
    oracle.WeightsHost(1, wgrid).grid(uv, w)
    expected = np.zeros_like(wgrid)
    for i in range(4):
        expected[i, 55, 90] = 13 * 10**i
        expected[i, 67, 123] = 2 * 10**i
        expected[i, 56, 90] = 16 * 10**i
        expected[i, 55, 89] = 32 * 10**i
    np.testing.assert_array_equal(wgrid, expected)


def test_predict(oracle):
    fx = cases.predict_case()
    golden = load_golden('predict_small')
    np.testing.assert_allclose(
        oracle.uvw_scale_bias(fx['image_parameters'], fx['grid_parameters']),
        golden['scale_bias'], rtol=1e-12)
    vis = fx['vis'].copy()
    oracle.predict(vis, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['weights'], fx['lmn'],
                   fx['flux'], fx['oversample'], fx['uv_scale'], fx['w_scale'], fx['w_bias'])
    # single-precision phases of hundreds of turns: the reference's own tolerance for
    # this routine is rtol 5e-4 (test_predict.py:92)
    np.testing.assert_allclose(vis, golden['residual'], rtol=5e-4, atol=5e-4)


def test_grid_to_image_threaded_matches(oracle):
    """The thread-pool variant used by bench.py's CPU legs equals grid_to_image."""
    from concurrent.futures import ThreadPoolExecutor
    rs = np.random.RandomState(1)
    n, g = 256, 150
    grid = (rs.standard_normal((1, g, g)) + 1j * rs.standard_normal((1, g, g))).astype(np.complex64)
    taper = rs.uniform(1, 2, n).astype(np.float32)
    lm_scale = 0.2 / n
    lm_bias = -0.5 * n * lm_scale
    expected = np.zeros((1, n, n), np.float32)
    oracle.grid_to_image(grid, expected, taper, lm_scale, lm_bias, np.float64(31.5))
    actual = np.zeros((n, n), np.float32)
    with ThreadPoolExecutor(3) as pool:
        oracle.grid_to_image_threaded(grid[0], actual, taper, lm_scale, lm_bias,
                                      np.float64(31.5), pool, 3)
    assert np.abs(actual - expected[0]).max() <= 1e-6 * np.abs(expected).max()


def _reference_preprocess_case():
    """Inputs and expected records of the reference's own preprocessing test
    (reference katsdpimager/test/test_preprocess.py:76-136)."""
    uvw = np.array([[12.1, 2.3, 4.7], [12.102, 2.299, 4.6], [-5.2, -10.6, 7.2],
                    [-1.0, 2.0, 3.0]], np.float32)
    weights = np.array([
        [[1.3, 0.6, 1.2, 0.1], [1.1, 1.2, 1.3, 1.4], [0.5, 0.6, 0.7, 0.8], [1.0, 0.0, 1.0, 1.0]],
        [[0.2, 2.4, 1.2, 2.6], [2.8, 2.6, 2.4, 2.2], [1.6, 1.4, 1.2, 1.0], [2.0, 2.0, 0.0, 2.0]]],
        np.float32)
    vis = np.array([
        [[0.5 - 2.3j, 0.1 + 4.2j, 0.0 - 3j, 1.5 + 0j], [1.2 + 3.4j, 5.6 + 7.8j, 9.0 + 1.2j, 3.4 + 5.6j],
         [1.5 + 1.3j, 1.1 + 2.7j, 1.0 - 2j, 2.5 + 1j], [10.0, 10.0, 10.0, 10.0]],
        [[3.0 + 0j, 0.0 - 6j, 0.2 + 8.4j, 1.0 - 4.6j], [6.8 + 11.2j, 18.0 + 2.4j, 11.2 + 15.6j, 2.4 + 6.8j],
         [3.0 + 2j, 2.0 - 4j, 2.2 + 5.4j, 3.0 + 2.6j], [20.0, 20.0, 20.0, 20.0]]], np.complex64)
    expected = [
        dict(uv=[[96, 18], [-42, -85]], sub_uv=[[6, 3], [3, 1]], w_plane=[64, 65],
             weights=[[2.4, 1.8, 2.5, 1.5], [0.5, 0.6, 0.7, 0.8]],
             vis=[[1.97 + 0.75j, 6.78 + 11.88j, 11.7 - 2.04j, 4.91 + 7.84j],
                  [0.75 + 0.65j, 0.66 + 1.62j, 0.7 - 1.4j, 2.0 + 0.8j]]),
        dict(uv=[[387, 73], [387, 73], [-167, -340]], sub_uv=[[1, 4], [2, 4], [4, 6]],
             w_plane=[64, 64, 65],
             weights=[[0.2, 2.4, 1.2, 2.6], [2.8, 2.6, 2.4, 2.2], [1.6, 1.4, 1.2, 1.0]],
             vis=[[0.6 + 0.0j, 0.0 - 14.4j, 0.24 + 10.08j, 2.6 - 11.96j],
                  [19.04 + 31.36j, 46.8 + 6.24j, 26.88 + 37.44j, 5.28 + 14.96j],
                  [4.8 + 3.2j, 2.8 - 5.6j, 2.64 + 6.48j, 3.0 + 2.6j]])]
    # cell sizes of the test's two channels: wavelength / (pixel_size * pixels), pixel_size =
    # 1 / (4096 wavelength), pixels = 2048 -> 2 wavelength^2 ... in metres: 0.125, 0.03125
    cells = [0.25 / ((1.0 / (4096.0 * 0.25)) * 2048), 0.125 / ((1.0 / (4096.0 * 0.125)) * 2048)]
    return uvw, weights, vis, expected, cells


@pytest.mark.parametrize('feed_angles', [False, True])
def test_preprocess_reference_records(oracle, feed_angles):
    """oracle.preprocess (restatement of preprocess.cpp) against the reference's expected
    records, with the static Mueller matrix and through the parallactic-angle generator
    (zero feed angles, identity matrices, as the reference tests it)."""
    uvw, weights, vis, expected, cells = _reference_preprocess_case()
    identity = np.identity(4, np.complex64)
    for channel in range(2):
        kwargs = dict(feed_angle1=np.zeros(4, np.float32), feed_angle2=np.zeros(4, np.float32),
                      mueller_circular=identity) if feed_angles else {}
        records, counts = oracle.preprocess(uvw, weights[channel], vis[channel], identity, 4,
                                            cells[channel], 400.0, 1, 128, 8, capacity=64,
                                            **kwargs)
        want = expected[channel]
        assert list(counts) == [len(want['uv'])]
        np.testing.assert_array_equal(records.uv, want['uv'])
        np.testing.assert_array_equal(records.sub_uv, want['sub_uv'])
        np.testing.assert_array_equal(records.w_plane, want['w_plane'])
        np.testing.assert_allclose(records.weights, want['weights'])
        np.testing.assert_allclose(records.vis, want['vis'], rtol=1e-5)


def test_preprocess_matches_numpy_port(oracle):
    """...and against the numpy producer used by the benchmarks (identity Mueller matrix),
    on a random set with flagged samples, NaNs, negative w, several W slices and merging."""
    from katsdpimager_b200 import parameters as prm, preprocess
    rs = np.random.RandomState(5)
    n = 5000
    uvw = (rs.standard_normal((n, 3)) * [300.0, 300.0, 150.0]).astype(np.float32)
    uvw[100:140] = uvw[100]                    # a run of duplicates
    weights = rs.uniform(0.5, 1.5, (n, 2)).astype(np.float32)
    weights[rs.randint(0, n, 200), rs.randint(0, 2, 200)] = 0.0
    vis = (rs.standard_normal((n, 2)) + 1j * rs.standard_normal((n, 2))).astype(np.complex64)
    vis[rs.randint(0, n, 50), 0] = np.nan
    fixed = prm.FixedImageParameters([1, 2], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.2, pixels=2048, pixel_size=0.0001)
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, 600.0, 7), 5, 16)
    records, counts = oracle.preprocess(uvw, weights, vis, np.identity(2, np.complex64), 2,
                                        np.float32(ip.cell_size), 600.0, 5, 16, 8)
    q, w_slice = preprocess.quantise(uvw, weights, vis, ip, gp)
    q, w_slice = preprocess.compress(q, w_slice)
    slices = preprocess.bucket_by_slice(q, w_slice, 5)
    assert list(counts) == [len(s) for s in slices]
    expected = np.concatenate(slices).view(np.recarray)
    np.testing.assert_array_equal(records.uv, expected.uv)
    np.testing.assert_array_equal(records.sub_uv, expected.sub_uv)
    np.testing.assert_array_equal(records.w_plane, expected.w_plane)
    np.testing.assert_allclose(records.weights, expected.weights, rtol=3e-7)
    np.testing.assert_allclose(records.vis, expected.vis, rtol=1e-6, atol=1e-6)
