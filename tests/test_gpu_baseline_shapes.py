"""Parity at the shapes BASELINE.json names, not at toy sizes.

* grid / degrid at config 2 (8192^2, 4 pol, K = 7, 16 W planes per slice, G = 4940: the
  dense W slice of real MeerKAT tracks at the 4 s dump rate, > 1 M visibilities, which is
  where the staged gridder's run-length and wave-balancing heuristics switch on), config 4
  (16384^2, 4 pol, K = 32, 128 W planes, G ~ 9870: wide records, 32-bit offset limits) and
  config 1 (2048^2, 1 pol, K = 60, 1639 W planes: 12.6 MB kernel table) against the CPU oracle
  (restatement of reference grid.py:1033-1052 `_grid`, :1139-1154 `_degrid`);
* CLEAN at config 5 (4096^2, 1000 minor cycles, 255^2 and 1023^2 patches, I and SUMSQ)
  bit-exact against the oracle's CleanHost (reference clean.py:971-1075), plus a patch that
  needs several waves of blocks, one cycle per call and batched (the case of ADVICE r1).

Tolerances: grids -- float32 accumulation in a different order than the host loop; the error
against a float64 evaluation must not exceed 1e-5 of the largest cell, or 1.5 x the error the
host's own float32 loop makes against float64 when that is larger (dense slice: thousands of
accumulations per cell).  Model visibilities 1e-5 relative (north_star).  CLEAN bit exact.
"""
import types

import numpy as np
import pytest
import scipy.signal.windows

from katsdpimager_b200 import clean, grid, parameters as prm, preprocess, simulate
from tests import cases

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- inputs
def meerkat_slice(pixels, pols, kernel_width, w_slices, w_planes, baseline_step, dumps=3600,
                  want_slice=0, limit=None, seed=1, eps_w=None):
    """Quantised records of one W slice of a MeerKAT L-band channel: every
    `baseline_step`-th baseline of the 2016, `dumps` samples at 4 s, baseline-major
    (SURVEY.md section 8d recipe, the generator bench.py uses)."""
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.2155, pixels=pixels, array=array,
                             image_oversample=5.0)
    fixed_grid = prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, kernel_width,
                                         degrid=True)
    gp = prm.GridParameters(fixed_grid, w_slices, w_planes)
    uvw = simulate.uvw_tracks(dumps, dump_time=4.0)[::baseline_step].reshape(-1, 3)
    rs = np.random.RandomState(seed)
    vis = (rs.standard_normal((len(uvw), pols)) + 1j * rs.standard_normal((len(uvw), pols))) \
        .astype(np.complex64)
    weights = rs.uniform(0.5, 1.5, (len(uvw), pols)).astype(np.float32)
    records, w_slice = preprocess.quantise(uvw.astype(np.float32), weights, vis, ip, gp)
    slices = preprocess.bucket_by_slice(records, w_slice, w_slices)
    out = slices[want_slice]
    if limit is not None:
        out = out[:limit]
    return array, ip, gp, out


def _upload(fn, queue, records, with_weights=False):
    n = len(records)
    fn.num_vis = n
    fn.buffer('uv').set_region(
        queue, np.ascontiguousarray(np.concatenate((records.uv, records.sub_uv), axis=1)),
        np.s_[:n], np.s_[:])
    fn.buffer('w_plane').set_region(queue, np.ascontiguousarray(records.w_plane),
                                    np.s_[:n], np.s_[:])
    fn.buffer('vis').set_region(queue, np.ascontiguousarray(records.vis), np.s_[:n], np.s_[:])
    if with_weights:
        fn.buffer('weights').set_region(queue, np.ascontiguousarray(records.weights),
                                        np.s_[:n], np.s_[:])


def _grid_and_degrid(gpu, oracle, array, ip, gp, records, check_f64):
    context, queue = gpu
    pols = len(ip.fixed.polarizations)
    n = len(records)
    template = grid.GridderTemplate(context, ip.fixed, gp.fixed)
    fn = template.instantiate(queue, array, ip, gp, n)
    fn.ensure_all_bound()
    size = fn.buffer('grid').shape[-1]
    rs = np.random.RandomState(5)
    wgrid = rs.uniform(0.5, 1.5, (pols, size, size)).astype(np.float32)
    fn.buffer('grid').zero(queue)
    fn.buffer('weights_grid').set(queue, wgrid)
    _upload(fn, queue, records)
    fn()
    assert fn.num_rejected() == 0
    actual = fn.buffer('grid').get(queue)
    lut = oracle.convolution_kernel(ip, gp)
    uv = np.ascontiguousarray(records.uv)
    sub_uv = np.ascontiguousarray(records.sub_uv)
    w_plane = np.ascontiguousarray(records.w_plane)
    vis = np.ascontiguousarray(records.vis)
    expected = np.zeros(actual.shape, np.complex64)
    oracle.grid(lut, expected, wgrid, uv, sub_uv, w_plane, vis)
    peak = np.abs(expected).max()
    err = np.abs(actual - expected).max() / peak
    if check_f64:
        exact = np.zeros(actual.shape, np.complex128)
        oracle.grid(lut, exact, wgrid, uv, sub_uv, w_plane, vis)
        err_gpu = np.abs(actual - exact).max() / peak
        err_host = np.abs(expected - exact).max() / peak
        del exact
        assert err_gpu <= max(1e-5, 1.5 * err_host), (err_gpu, err_host)
        # and the two float32 results agree to the sum of their errors
        assert err <= max(1e-5, 2.5 * err_host)
    else:
        assert err <= 1e-5, err
    # nothing outside the footprints was touched: cells the host left at zero stay zero
    assert not actual[expected == 0].any()
    del actual

    # degridding from a random grid of the same shape (model visibilities: 1e-5 relative)
    dtemplate = grid.DegridderTemplate(context, ip.fixed, gp.fixed)
    dfn = dtemplate.instantiate(queue, array, ip, gp, n)
    dfn.bind(grid=fn.buffer('grid'))            # reuse the device allocation
    dfn.ensure_all_bound()
    assert dfn.buffer('grid').shape == expected.shape
    model = (rs.standard_normal(expected.shape) + 1j * rs.standard_normal(expected.shape)) \
        .astype(np.complex64)
    dfn.buffer('grid').set(queue, model)
    _upload(dfn, queue, records, with_weights=True)
    dfn()
    assert dfn.num_rejected() == 0
    residual = dfn.buffer('vis').get(queue)[:n]
    host_residual = vis.copy()
    oracle.degrid(lut, model, uv, sub_uv, w_plane, np.ascontiguousarray(records.weights),
                  host_residual)
    predicted = host_residual - vis
    scale = np.abs(predicted).max()
    assert np.abs(residual - host_residual).max() / scale < 1e-5
    return n, size


def test_config2_dense_slice(gpu, oracle):
    """8192^2, 4 pol, K = 7, 16 planes: slice 0 of 504 baselines x 3600 dumps."""
    array, ip, gp, records = meerkat_slice(8192, 4, 7, 16, 16, baseline_step=4)
    assert len(records) > 1000000
    n, size = _grid_and_degrid(gpu, oracle, array, ip, gp, records, check_f64=True)
    assert size == 2 * (int(array.longest_baseline / ip.cell_size) + 7 // 2 + 1)
    assert 4900 <= size <= 4960


def test_config2_sparse_slice(gpu, oracle):
    """A high-W slice of the same channel (short runs: sub-wave launch heuristics)."""
    array, ip, gp, records = meerkat_slice(8192, 4, 7, 16, 16, baseline_step=2, want_slice=4)
    assert 0 < len(records) < 200000
    _grid_and_degrid(gpu, oracle, array, ip, gp, records, check_f64=False)


def test_config4_wide_field(gpu, oracle):
    """16384^2, 4 pol, K = 32, 128 W planes: 3.1 GB grid, wide staged records."""
    array, ip, gp, records = meerkat_slice(16384, 4, 32, 4, 128, baseline_step=37,
                                           limit=120000)
    assert len(records) >= 100000
    n, size = _grid_and_degrid(gpu, oracle, array, ip, gp, records, check_f64=False)
    assert 9850 <= size <= 9900


def test_config1_wide_support(gpu, oracle):
    """2048^2, Stokes I, K = 60, 3 slices x 1639 planes (imager.py defaults on the
    simulate.py data set): 12.6 MB kernel table read through L1."""
    array, ip, gp, records = meerkat_slice(2048, 1, 60, 3, 1639, baseline_step=29,
                                           limit=120000)
    assert len(records) >= 100000
    _grid_and_degrid(gpu, oracle, array, ip, gp, records, check_f64=False)


# ------------------------------------------------------------------------------ CLEAN
def _clean_inputs(pixels, pols, seed, num_sources=60):
    """PSF with a sharp core, a broad pedestal and noisy sidelobes; dirty image = point
    sources (clustered so that patches overlap, some near the border) convolved with the PSF
    by FFT, plus noise."""
    rs = np.random.RandomState(seed)
    g1 = scipy.signal.windows.gaussian(pixels, 2.0)
    g2 = scipy.signal.windows.gaussian(pixels, pixels / 24.0)
    psf1 = np.outer(g1, g1) + 0.05 * np.outer(g2, g2) \
        + 0.002 * rs.standard_normal((pixels, pixels))
    psf1 /= psf1[pixels // 2, pixels // 2]
    sky = np.zeros((pixels, pixels))
    border = round(0.02 * pixels)
    ys = rs.randint(border, pixels - border, num_sources)
    xs = rs.randint(border, pixels - border, num_sources)
    ys[:8] = pixels // 2 + rs.randint(-40, 40, 8)
    xs[:8] = pixels // 2 + rs.randint(-40, 40, 8)
    ys[8:12] = border + rs.randint(0, 3, 4)              # on the border limit
    sky[ys, xs] = rs.uniform(0.5, 5.0, num_sources)
    conv = np.fft.irfft2(np.fft.rfft2(sky) * np.fft.rfft2(np.fft.ifftshift(psf1)),
                         s=(pixels, pixels))
    dirty = np.empty((pols, pixels, pixels), np.float32)
    for p in range(pols):
        frac = 1.0 if p == 0 else 0.3 * (-1) ** p
        dirty[p] = frac * conv + 0.01 * rs.standard_normal((pixels, pixels))
    psf = np.repeat(psf1[np.newaxis].astype(np.float32), pols, axis=0)
    psf[:, pixels // 2, pixels // 2] = 1.0
    return dirty, psf


def _make_clean(gpu, dirty, psf, mode, border, loop_gain=0.1):
    context, queue = gpu
    pols, pixels = dirty.shape[0], dirty.shape[1]
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
    ip = types.SimpleNamespace(fixed=fixed, pixels=pixels)
    cp = prm.CleanParameters(1000, loop_gain, 0.85, 5.0, mode, 0.01, 0.5, border)
    fn = clean.CleanTemplate(context, cp, np.float32, pols).instantiate(queue, ip)
    fn.ensure_all_bound()
    fn.buffer('dirty').set(queue, dirty)
    fn.buffer('psf').set(queue, psf)
    fn.buffer('model').zero(queue)
    fn.reset()
    return fn, queue


def _check_cycles(oracle, fn, queue, components, dirty, psf, mode, border, patch, loop_gain=0.1):
    host_dirty = dirty.copy()
    host_model = np.zeros_like(dirty)
    host = oracle.CleanHost(dirty.shape[1], border, mode, loop_gain, host_dirty, psf, host_model)
    host.reset()
    for i, record in enumerate(components):
        value, pos, pixel = host(patch, 0.0)
        assert (int(record['pos'][0]), int(record['pos'][1])) == pos, 'cycle {}'.format(i)
        assert record['value'] == value, 'cycle {}'.format(i)
        np.testing.assert_array_equal(record['pixel'], pixel)
    np.testing.assert_array_equal(fn.buffer('dirty').get(queue), host_dirty)
    np.testing.assert_array_equal(fn.buffer('model').get(queue), host_model)
    np.testing.assert_array_equal(fn.buffer('tile_max').get(queue), host.tile_max)
    np.testing.assert_array_equal(fn.buffer('tile_pos').get(queue), host.tile_pos)


@pytest.mark.parametrize('mode,pols,patch_size', [
    (clean.CLEAN_I, 1, 255), (clean.CLEAN_SUMSQ, 4, 255),
    (clean.CLEAN_I, 1, 1023), (clean.CLEAN_SUMSQ, 4, 1023)])
def test_config5_thousand_cycles(gpu, oracle, mode, pols, patch_size):
    """4096^2, border 0.02 (82 pixels, 123 x 123 tiles), 1000 minor cycles in one batch."""
    dirty, psf = _clean_inputs(4096, pols, 40 + pols)
    fn, queue = _make_clean(gpu, dirty, psf, mode, 0.02)
    patch = (pols, patch_size, patch_size)
    components, stopped = fn.run_cycles(patch, 0.0, 1000)
    assert len(components) == 1000 and not stopped
    _check_cycles(oracle, fn, queue, components, dirty, psf, mode, 0.02, patch)


@pytest.mark.parametrize('batched', [False, True])
def test_multi_wave_patch(gpu, oracle, batched):
    """A 2047^2 patch on 4096^2 (4225 blocks: several waves) -- every block of every cycle
    must do its share, also on the last cycle of a batch and with one cycle per call."""
    dirty, psf = _clean_inputs(4096, 2, 77, num_sources=20)
    fn, queue = _make_clean(gpu, dirty, psf, clean.CLEAN_SUMSQ, 0.02)
    patch = (2, 2047, 2047)
    cycles = 12
    if batched:
        components, stopped = fn.run_cycles(patch, 0.0, cycles)
        assert not stopped
    else:
        records = []
        for _ in range(cycles):
            value, pos, pixel = fn(patch, 0.0)
            records.append((pos, value, pixel))
        components = np.zeros(cycles, fn._record_dtype)
        for i, (pos, value, pixel) in enumerate(records):
            components[i]['pos'] = pos
            components[i]['value'] = value
            components[i]['pixel'] = pixel
    assert len(components) == cycles
    _check_cycles(oracle, fn, queue, components, dirty, psf, clean.CLEAN_SUMSQ, 0.02, patch)
