"""Gridder / degridder parity: the reference's own unit-test harness
(reference katsdpimager/test/test_grid.py), golden vectors from the reference host
classes, and the CPU oracle over a sweep of kernel widths and polarization counts."""
import types

import numpy as np
import pytest

from katsdpimager_b200 import grid, parameters as prm
from tests import cases
from tests.cases import load_golden

pytestmark = pytest.mark.gpu


def _load_vis(fn, queue, fx, vis):
    n = len(fx['uv'])
    fn.num_vis = n
    fn.buffer('uv').set_region(queue, np.concatenate((fx['uv'], fx['sub_uv']), axis=1),
                               np.s_[:n], np.s_[:])
    fn.buffer('w_plane').set_region(queue, fx['w_plane'], np.s_[:n], np.s_[:])
    fn.buffer('vis').set_region(queue, vis, np.s_[:n], np.s_[:])
    return n


def _run_gridder(gpu, fx, vis, weights_grid_full, max_vis=None):
    context, queue = gpu
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    template = grid.GridderTemplate(context, ip.fixed, gp.fixed)
    fn = template.instantiate(queue, fx['array_parameters'], ip, gp, max_vis or len(vis) + 280)
    fn.ensure_all_bound()
    fn.buffer('grid').zero(queue)
    wg = fn.buffer('weights_grid')
    host = wg.empty_like()
    host.fill(0)
    if weights_grid_full.shape[-1] <= host.shape[-1]:
        cases.middle(host, weights_grid_full.shape)[:] = weights_grid_full
    else:
        host[:] = cases.middle(weights_grid_full, host.shape)
    wg.set(queue, host)
    _load_vis(fn, queue, fx, vis)
    fn()
    assert fn.num_rejected() == 0
    return fn.buffer('grid').get(queue), np.array(host)


def _run_degridder(gpu, fx, grid_full, weights, vis):
    context, queue = gpu
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    template = grid.DegridderTemplate(context, ip.fixed, gp.fixed)
    fn = template.instantiate(queue, fx['array_parameters'], ip, gp, len(vis) + 280)
    fn.ensure_all_bound()
    buf = fn.buffer('grid')
    buf.set(queue, np.ascontiguousarray(cases.middle(grid_full, buf.shape)))
    n = _load_vis(fn, queue, fx, vis)
    fn.buffer('weights').set_region(queue, weights, np.s_[:n], np.s_[:])
    fn()
    assert fn.num_rejected() == 0
    return fn.buffer('vis').get(queue)[:n]


def test_reference_fixture_grid(gpu):
    """reference test_grid.py TestGridder.test: float64 grid, rtol 1e-5 / atol 1e-8
    against the first-principles numpy expectation (do_grid, test_grid.py:91-112)."""
    fx = cases.reference_grid_fixture()
    actual, _ = _run_gridder(gpu, fx, fx['vis'], fx['weights_grid'], max_vis=1280)
    lut = grid.ConvolutionKernel(fx['image_parameters'], fx['grid_parameters']).data
    expected = np.zeros_like(actual)
    pixels = actual.shape[-1]
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
    # and the reference's GridderHost output itself
    golden = load_golden('grid_reference_fixture')
    full = np.zeros((4, 256, 256), np.complex128)
    cases.middle(full, actual.shape)[:] = actual
    np.testing.assert_allclose(full[:, ::3, :], golden['grid_rows'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(full.sum(axis=(1, 2)), golden['grid_sum'], rtol=1e-6)


def test_reference_fixture_degrid(gpu):
    """reference test_grid.py TestDegridder.test (do_degrid, test_grid.py:114-135): rtol 1e-5."""
    fx = cases.reference_grid_fixture()
    actual = _run_degridder(gpu, fx, fx['degrid_grid'], fx['degrid_weights'], fx['degrid_vis'])
    golden = load_golden('grid_reference_fixture')
    np.testing.assert_allclose(actual, golden['residual'], rtol=1e-5)


def test_small_golden(gpu, oracle):
    """float32 grid, K = 7: golden output of the reference's GridderHost/DegridderHost."""
    fx = cases.small_grid_case()
    golden = load_golden('grid_small')
    actual, _ = _run_gridder(gpu, fx, fx['vis'], fx['weights_grid'])
    expected = cases.middle(golden['grid'], actual.shape)
    scale = np.abs(expected).max()
    # float32 accumulation in a different order: a few ulp of the largest cell
    np.testing.assert_allclose(actual, expected, rtol=0, atol=1e-5 * scale)
    # nothing was gridded outside the device grid
    outside = golden['grid'].copy()
    cases.middle(outside, actual.shape)[:] = 0
    assert not outside.any()
    residual = _run_degridder(gpu, fx, golden['grid'], fx['weights'], fx['vis'])
    # north_star: model visibilities within 1e-5 relative
    err = np.abs(residual - golden['residual']).max() / np.abs(golden['residual']).max()
    assert err < 1e-5


@pytest.mark.parametrize('kernel_width,pols', [(7, 1), (7, 4), (8, 3), (9, 2), (16, 4), (21, 1),
                                               (32, 4), (33, 2), (60, 1), (64, 4)])
def test_sweep_against_oracle(gpu, oracle, kernel_width, pols):
    pixels = 384 if kernel_width > 40 else 192
    fx = cases.small_grid_case(pixels=pixels, pols=pols, kernel_width=kernel_width, w_planes=3,
                               n_vis=1500, seed=100 + kernel_width + pols)
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    actual, wgrid = _run_gridder(gpu, fx, fx['vis'], fx['weights_grid'])
    lut = oracle.convolution_kernel(ip, gp)
    expected = np.zeros(actual.shape, np.complex64)
    oracle.grid(lut, expected, wgrid, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['vis'])
    np.testing.assert_allclose(actual, expected, rtol=0, atol=1e-5 * np.abs(expected).max())
    residual = _run_degridder(gpu, fx, expected, fx['weights'], fx['vis'])
    host_residual = fx['vis'].copy()
    oracle.degrid(lut, expected, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['weights'],
                  host_residual)
    err = np.abs(residual - host_residual).max() / np.abs(host_residual).max()
    assert err < 1e-5


def test_double_precision_small(gpu, oracle):
    fx = cases.small_grid_case(pols=3, kernel_width=12, dtype=np.float64, n_vis=1000)
    actual, wgrid = _run_gridder(gpu, fx, fx['vis'], fx['weights_grid'])
    assert actual.dtype == np.complex128
    lut = oracle.convolution_kernel(fx['image_parameters'], fx['grid_parameters'])
    expected = np.zeros(actual.shape, np.complex128)
    oracle.grid(lut, expected, wgrid, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['vis'])
    np.testing.assert_allclose(actual, expected, rtol=1e-5, atol=1e-8)


def test_linearity_large(gpu):
    """Size-independent property at a realistic chunk size: gridding is linear in the
    visibilities, and the grid total equals sum(sample * conj(kernel sums))."""
    context, queue = gpu
    fx = cases.small_grid_case(pixels=1024, pols=4, n_vis=200000, seed=9)
    a, _ = _run_gridder(gpu, fx, fx['vis'], fx['weights_grid'])
    b, _ = _run_gridder(gpu, fx, (2.5 * fx['vis']).astype(np.complex64), fx['weights_grid'])
    np.testing.assert_allclose(b, 2.5 * a, rtol=0, atol=2e-5 * np.abs(a).max())
    lut = grid.ConvolutionKernel(fx['image_parameters'], fx['grid_parameters']).data
    lut_sum = lut.astype(np.complex128).sum(axis=2)
    ksum = lut_sum[fx['w_plane'], fx['sub_uv'][:, 1]] * lut_sum[fx['w_plane'], fx['sub_uv'][:, 0]]
    size = a.shape[-1]
    wg = cases.middle(fx['weights_grid'], a.shape)
    w = wg[:, fx['uv'][:, 1] + size // 2, fx['uv'][:, 0] + size // 2].T
    expected_total = (fx['vis'].astype(np.complex128) * w * np.conj(ksum)[:, None]).sum(axis=0)
    np.testing.assert_allclose(a.astype(np.complex128).sum(axis=(1, 2)), expected_total,
                               rtol=2e-4)


def test_edge_cases(gpu):
    context, queue = gpu
    fx = cases.small_grid_case(n_vis=100)
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    template = grid.GridderTemplate(context, ip.fixed, gp.fixed)
    fn = template.instantiate(queue, fx['array_parameters'], ip, gp, 128)
    fn.ensure_all_bound()
    fn.buffer('grid').zero(queue)
    fn.buffer('weights_grid').set(queue, np.ones(fn.buffer('weights_grid').shape, np.float32))
    # empty launch
    fn.num_vis = 0
    fn()
    assert not fn.buffer('grid').get(queue).any()
    # num_vis range check (grid.py:697-703)
    with pytest.raises(ValueError):
        fn.num_vis = 129
    with pytest.raises(ValueError):
        fn.num_vis = -1
    # out-of-range coordinates are skipped and counted, in-range ones still gridded
    size = fn.buffer('grid').shape[-1]
    uv = np.zeros((4, 4), np.int16)
    uv[0, :2] = (size, 0)
    uv[1, :2] = (0, -size)
    uv[2, 2] = 8        # sub-pixel index out of range
    w_plane = np.array([0, 0, 0, gp.w_planes], np.int16)    # last one: bad w plane
    fn.num_vis = 4
    fn.buffer('uv').set_region(queue, uv, np.s_[:4], np.s_[:])
    fn.buffer('w_plane').set_region(queue, w_plane, np.s_[:4], np.s_[:])
    fn.buffer('vis').set_region(queue, np.ones((4, 2), np.complex64), np.s_[:4], np.s_[:])
    fn()
    assert fn.num_rejected() == 4
    assert not fn.buffer('grid').get(queue).any()
    # a baseline that does not fit the image is refused up front (grid.py:759-761)
    array = types.SimpleNamespace(longest_baseline=ip.cell_size * ip.pixels)
    with pytest.raises(ValueError):
        template.instantiate(queue, array, ip, gp, 128)
