"""Device preprocessing (kib_preprocess) against the oracle's restatement of the reference's
preprocess.cpp -- itself pinned by the reference's expected records (tests/test_oracle.py) --
bit for bit where the arithmetic is IEEE (static Mueller matrix), to rounding where the feed
angle rotation uses device sincos."""
import numpy as np
import pytest

from katsdpimager_b200 import parameters as prm, pipeline, preprocess, simulate
from tests.test_oracle import _reference_preprocess_case

pytestmark = pytest.mark.gpu


def _params(pols, w_slices=5, w_planes=16, max_w=600.0, pixel_size=0.0001, wavelength=0.2,
            pixels=2048):
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=pixels, pixel_size=pixel_size)
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, max_w, 7), w_slices, w_planes)
    return ip, gp


def _device_records(queue, uvw, weights, vis, ip, gp, **kwargs):
    resident = pipeline.ResidentVisibilities.from_raw(queue, uvw, weights, vis, ip, gp, **kwargs)
    return resident, [resident.get(s) for s in range(gp.w_slices)]


def _compare(device_slices, records, counts, exact=True):
    assert [len(s) for s in device_slices] == list(counts)
    actual = np.concatenate(device_slices).view(np.recarray)
    np.testing.assert_array_equal(actual.uv, records.uv)
    np.testing.assert_array_equal(actual.sub_uv, records.sub_uv)
    np.testing.assert_array_equal(actual.w_plane, records.w_plane)
    if exact:
        np.testing.assert_array_equal(actual.weights, records.weights)
        np.testing.assert_array_equal(actual.vis, records.vis)
    else:
        np.testing.assert_allclose(actual.weights, records.weights, rtol=1e-5)
        np.testing.assert_allclose(actual.vis, records.vis, rtol=1e-4, atol=1e-5)


def test_reference_expected_records(gpu):
    """The reference's own test vectors (test/test_preprocess.py:76-136), both generators."""
    context, queue = gpu
    uvw, weights, vis, expected, cells = _reference_preprocess_case()
    identity = np.identity(4, np.complex64)
    for channel, wavelength in enumerate([0.25, 0.125]):
        ip, gp = _params(4, 1, 128, 400.0, 1.0 / (4096.0 * wavelength), wavelength)
        for kwargs in ({}, dict(feed_angle1=np.zeros(4, np.float32),
                                feed_angle2=np.zeros(4, np.float32), mueller_circular=identity)):
            _, slices = _device_records(queue, uvw, weights[channel], vis[channel], ip, gp,
                                        mueller_stokes=identity, capacity=64, **kwargs)
            want = expected[channel]
            actual = slices[0]
            np.testing.assert_array_equal(actual.uv, want['uv'])
            np.testing.assert_array_equal(actual.sub_uv, want['sub_uv'])
            np.testing.assert_array_equal(actual.w_plane, want['w_plane'])
            np.testing.assert_allclose(actual.weights, want['weights'])
            np.testing.assert_allclose(actual.vis, want['vis'], rtol=1e-5)


@pytest.mark.parametrize('pols,capacity', [(1, 0), (2, 0), (4, 0), (4, 4096), (3, 1000)])
def test_random_against_oracle(gpu, oracle, pols, capacity):
    """Flagged samples, NaNs, -0 weights, negative w, runs of duplicates that straddle buffer
    boundaries, several W slices: records identical to the host code's, bit for bit."""
    context, queue = gpu
    rs = np.random.RandomState(50 + pols)
    n = 200000
    uvw = (rs.standard_normal((n, 3)) * [300.0, 300.0, 150.0]).astype(np.float32)
    for start in rs.randint(0, n - 50, 300):                 # runs of duplicates
        uvw[start:start + rs.randint(2, 40)] = uvw[start]
    uvw[990:1010] = uvw[990]                                 # across the 1000-sample boundary
    weights = rs.uniform(0.5, 1.5, (n, pols)).astype(np.float32)
    weights[rs.randint(0, n, 5000), rs.randint(0, pols, 5000)] = 0.0
    weights[rs.randint(0, n, 50), 0] = -0.0
    vis = (rs.standard_normal((n, pols)) + 1j * rs.standard_normal((n, pols))).astype(np.complex64)
    vis[rs.randint(0, n, 500), rs.randint(0, pols, 500)] = np.nan
    vis[rs.randint(0, n, 100), 0] = np.inf
    ip, gp = _params(pols)
    mueller = (rs.standard_normal((pols, pols)) + 1j * rs.standard_normal((pols, pols))) \
        .astype(np.complex64)
    if pols > 1:
        mueller[0, -1] = 0.0                                 # exercises the MulZ products
    records, counts = oracle.preprocess(uvw, weights, vis, mueller, pols,
                                        np.float32(ip.cell_size), gp.fixed.max_w, gp.w_slices,
                                        gp.w_planes, gp.fixed.oversample, capacity=capacity)
    resident, slices = _device_records(queue, uvw, weights, vis, ip, gp, mueller_stokes=mueller,
                                       capacity=capacity)
    assert len(resident) == len(records) > n // 2
    _compare(slices, records, counts)


def test_feed_angles_and_stokes_conversion(gpu, oracle):
    """Linear feeds -> IQUV through the parallactic-angle generator (preprocess.cpp:167-182)."""
    context, queue = gpu
    rs = np.random.RandomState(9)
    n = 50000
    uvw = (rs.standard_normal((n, 3)) * [300.0, 300.0, 150.0]).astype(np.float32)
    weights = rs.uniform(0.5, 1.5, (n, 4)).astype(np.float32)
    vis = (rs.standard_normal((n, 4)) + 1j * rs.standard_normal((n, 4))).astype(np.complex64)
    feed1 = rs.uniform(-np.pi, np.pi, n).astype(np.float32)
    feed2 = rs.uniform(-np.pi, np.pi, n).astype(np.float32)
    stokes = np.array([[0.5, 0, 0, 0.5], [0, 0.5, 0.5, 0], [0, -0.5j, 0.5j, 0]], np.complex64)
    circular = (rs.standard_normal((4, 4)) + 1j * rs.standard_normal((4, 4))).astype(np.complex64)
    ip, gp = _params(3)
    records, counts = oracle.preprocess(uvw, weights, vis, stokes, 3, np.float32(ip.cell_size),
                                        gp.fixed.max_w, gp.w_slices, gp.w_planes,
                                        gp.fixed.oversample, feed_angle1=feed1, feed_angle2=feed2,
                                        mueller_circular=circular)
    _, slices = _device_records(queue, uvw, weights, vis, ip, gp, mueller_stokes=stokes,
                                feed_angle1=feed1, feed_angle2=feed2, mueller_circular=circular)
    _compare(slices, records, counts, exact=False)


def test_meerkat_channel_feeds_the_gridder(gpu, oracle):
    """A MeerKAT-shaped channel preprocessed on the device equals the numpy producer the
    benchmarks use (same slices, same coordinates, values to rounding)."""
    context, queue = gpu
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.2155, pixels=8192, array=array)
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, 7), 16, 16)
    uvw = simulate.uvw_tracks(600, dump_time=4.0)[::8].reshape(-1, 3).astype(np.float32)
    rs = np.random.RandomState(3)
    vis = (rs.standard_normal((len(uvw), 4)) + 1j * rs.standard_normal((len(uvw), 4))) \
        .astype(np.complex64)
    weights = rs.uniform(0.5, 1.5, (len(uvw), 4)).astype(np.float32)
    resident, slices = _device_records(queue, uvw, weights, vis, ip, gp)
    q, w_slice = preprocess.quantise(uvw, weights, vis, ip, gp)
    q, w_slice = preprocess.compress(q, w_slice)
    expected = preprocess.bucket_by_slice(q, w_slice, gp.w_slices)
    assert [len(s) for s in slices] == [len(s) for s in expected]
    for a, b in zip(slices, expected):
        np.testing.assert_array_equal(a.uv, b.uv)
        np.testing.assert_array_equal(a.sub_uv, b.sub_uv)
        np.testing.assert_array_equal(a.w_plane, b.w_plane)
        # the numpy producer keeps w instead of 1 / (1 / w) and sums runs pairwise: rounding only
        np.testing.assert_allclose(a.weights, b.weights, rtol=2e-6)
        np.testing.assert_allclose(a.vis, b.vis, rtol=1e-5, atol=1e-5)
