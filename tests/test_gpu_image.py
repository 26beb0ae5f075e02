"""FFT-stage and image-arithmetic kernels: the reference's unit tests
(reference katsdpimager/test/test_image.py) and golden vectors from GridToImageHost /
ImageToGridHost."""
import math

import numpy as np
import pytest

from katsdpimager_b200 import accel, image
from tests import cases
from tests.cases import RandomState, load_golden

pytestmark = pytest.mark.gpu


def test_layer_to_image(gpu):
    """reference test_image.py TestLayerToImage (rtol 1e-4)."""
    context, queue = gpu
    slices, size = 3, 102
    shape = (slices, size, size)
    lm_scale = 0.1 / size
    lm_bias = -lm_scale * size / 3
    w = 12.3
    fn = image.LayerToImageTemplate(context, np.float32).instantiate(queue, shape, lm_scale, lm_bias)
    fn.set_w(w)
    fn.set_polarization(1)
    fn.ensure_all_bound()
    rs = RandomState(1)
    src = rs.complex_uniform(10.0, 100.0, shape[1:]).astype(np.complex64)
    kernel1d = rs.uniform(1.0, 2.0, size).astype(np.float32)
    fn.buffer('kernel1d').set(queue, kernel1d)
    fn.buffer('layer').set(queue, src)
    fn.buffer('image').zero(queue)
    lm = np.arange(size) * lm_scale + lm_bias
    lm2 = lm * lm
    n = np.sqrt(1 - lm2[np.newaxis, :, np.newaxis] - lm2[np.newaxis, np.newaxis, :])
    corrected = np.fft.fftshift(src) * np.exp(2j * math.pi * w * (n - 1))
    expected = np.zeros(shape, np.float32)
    expected[1] = corrected.real * n / np.outer(kernel1d, kernel1d)[np.newaxis, ...]
    fn()
    np.testing.assert_allclose(expected, fn.buffer('image').get(queue), 1e-4)
    # accumulates
    fn()
    np.testing.assert_allclose(2 * expected, fn.buffer('image').get(queue), 1e-4)
    with pytest.raises(IndexError):
        fn.set_polarization(3)
    with pytest.raises(ValueError):
        image.LayerToImageTemplate(context, np.float32).instantiate(queue, (1, 101, 101), 1, 0)


def test_image_to_layer_roundtrip(gpu):
    """reference test_image.py TestImageToLayer."""
    context, queue = gpu
    slices, size = 3, 102
    shape = (slices, size, size)
    lm_scale = 0.1 / size
    lm_bias = -lm_scale * size / 3
    w = 12.3
    i2l = image.ImageToLayerTemplate(context, np.float32).instantiate(queue, shape, lm_scale, lm_bias)
    i2l.ensure_all_bound()
    i2l.set_w(w)
    l2i = image.LayerToImageTemplate(context, np.float32).instantiate(queue, shape, lm_scale, lm_bias)
    l2i.set_w(w)
    l2i.bind(image=i2l.buffer('image'), layer=i2l.buffer('layer'), kernel1d=i2l.buffer('kernel1d'))
    rs = np.random.RandomState(1)
    expected = rs.uniform(10.0, 100.0, shape).astype(np.float32)
    kernel1d = rs.uniform(1.0, 2.0, size).astype(np.float32)
    kernel = np.outer(kernel1d, kernel1d)[np.newaxis, ...]
    i2l.buffer('image').set(queue, expected * kernel)
    i2l.buffer('kernel1d').set(queue, kernel1d)
    i2l.set_polarization(1)
    i2l()
    i2l.buffer('image').zero(queue)
    l2i.set_polarization(1)
    l2i()
    actual = l2i.buffer('image').get(queue) * kernel
    np.testing.assert_allclose(expected[1], actual[1], 1e-4)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_grid_to_image_and_back(gpu, oracle, dtype):
    """GridToImage / ImageToGrid against the reference's host classes (golden) and the
    oracle.  Bar: 1e-4 RMS relative to peak (north_star); observed ~1e-7."""
    context, queue = gpu
    fx = cases.image_case()
    golden = load_golden('image_small')
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    pols, pixels, _ = fx['grid'].shape
    size = fx['grid_size']
    template = image.GridImageTemplate(context, dtype)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    g2i = template.instantiate_grid_to_image(queue, (pols, size, size), fx['lm_scale'],
                                             fx['lm_bias'], plan)
    g2i.ensure_all_bound()
    g2i.buffer('grid').set(queue, np.ascontiguousarray(
        cases.middle(fx['grid'], (pols, size, size)).astype(cdtype)))
    g2i.buffer('kernel1d').set(queue, fx['kernel1d'].astype(dtype))
    g2i.buffer('image').zero(queue)
    g2i.set_w(fx['w'])
    g2i()
    actual = g2i.buffer('image').get(queue)
    rms = np.sqrt(np.mean((actual - golden['image']) ** 2)) / np.abs(golden['image']).max()
    assert rms < (5e-6 if dtype == np.float32 else 2e-5)   # golden is single precision
    i2g = template.instantiate_image_to_grid(queue, (pols, size, size), fx['lm_scale'],
                                             fx['lm_bias'], plan)
    i2g.bind(layer=g2i.buffer('layer'), kernel1d=g2i.buffer('kernel1d'))
    i2g.ensure_all_bound()
    i2g.buffer('image').set(queue, fx['model'].astype(dtype))
    i2g.set_w(fx['w'])
    i2g()
    back = i2g.buffer('grid').get(queue)
    expected = cases.middle(golden['grid_from_model'], back.shape)
    rms = np.sqrt(np.mean(np.abs(back - expected) ** 2)) / np.abs(expected).max()
    assert rms < (5e-6 if dtype == np.float32 else 2e-5)   # golden is single precision


def test_grid_to_image_w_precision(gpu, oracle):
    """Large W phases (thousands of turns) keep parity with the host computation."""
    context, queue = gpu
    fx = cases.image_case(pixels=128, grid_size=80, pols=1)
    pixels = 128
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    g2i = template.instantiate_grid_to_image(queue, (1, 80, 80), fx['lm_scale'], fx['lm_bias'], plan)
    g2i.ensure_all_bound()
    crop = np.ascontiguousarray(cases.middle(fx['grid'], (1, 80, 80)))
    g2i.buffer('grid').set(queue, crop)
    g2i.buffer('kernel1d').set(queue, fx['kernel1d'])
    w = np.float64(31234.56)
    g2i.buffer('image').zero(queue)
    g2i.set_w(w)
    g2i()
    expected = np.zeros((1, pixels, pixels), np.float32)
    oracle.grid_to_image(crop, expected, fx['kernel1d'], fx['lm_scale'], fx['lm_bias'], w)
    actual = g2i.buffer('image').get(queue)
    rms = np.sqrt(np.mean((actual - expected) ** 2)) / np.abs(expected).max()
    assert rms < 1e-5


def _fused_case(context, queue, pixels, grid_size, pols, seed, fused):
    rs = RandomState(seed)
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    g2i = template.instantiate_grid_to_image(queue, (pols, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    g2i.fused = fused
    g2i.ensure_all_bound()
    grid = rs.complex_normal(0.0j, 1.0, (pols, grid_size, grid_size)).astype(np.complex64)
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    g2i.buffer('grid').set(queue, grid)
    g2i.buffer('kernel1d').set(queue, kernel1d)
    return g2i, grid, kernel1d, lm_scale, lm_bias


@pytest.mark.parametrize('pixels,grid_size,pols', [(2048, 1230, 2), (2048, 2048, 1),
                                                   (2048, 1234, 1), (2048, 30, 1),
                                                   (4096, 2466, 1)])
def test_grid_to_image_fused_vs_oracle(gpu, oracle, pixels, grid_size, pols):
    """The pruned, fused transform (kib_grid_to_image) against the oracle's
    grid_to_image (numpy ifft2 in the reference's host arithmetic), including grids
    whose width is not a multiple of the 4-column block and a full-width grid.
    Bar: 1e-4 RMS relative to peak (north_star); observed ~3e-7."""
    context, queue = gpu
    from katsdpimager_b200 import _lib
    assert _lib.grid_to_image_supported(pixels, grid_size, np.float32)
    g2i, grid, kernel1d, lm_scale, lm_bias = _fused_case(
        context, queue, pixels, grid_size, pols, 11, True)
    before = _lib.kernel_launches
    expected = np.zeros((pols, pixels, pixels), np.float32)
    for w in (0.0, 57.25):
        g2i.set_w(w)
        g2i.buffer('image').zero(queue)
        g2i()
        actual = g2i.buffer('image').get(queue)
        expected[:] = 0
        oracle.grid_to_image(grid, expected, kernel1d, lm_scale, lm_bias, np.float64(w))
        rms = np.sqrt(np.mean((actual - expected) ** 2)) / np.abs(expected).max()
        assert rms < 2e-6
        assert np.abs(actual - expected).max() / np.abs(expected).max() < 2e-5
    per_plane = 1 + _lib.load().kib_grid_to_image_columns_kernels(pixels)
    assert _lib.kernel_launches - before == 2 * per_plane * pols
    # accumulates into the image
    g2i()
    np.testing.assert_allclose(g2i.buffer('image').get(queue), 2 * actual, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('pixels,grid_size', [(8192, 4940), (16384, 9864)])
def test_grid_to_image_fused_impulses(gpu, pixels, grid_size):
    """Full BASELINE sizes, analytic answer: a few non-zero grid cells give a sum of plane
    waves, image[y][x] = sum_k Re(a_k exp(2 pi i (u_k x + v_k y) / N)) n / (k1d[y] k1d[x])
    at w = 0 (x, y and u, v counted from the image / grid centre).  Checked on sampled rows in
    float64; independent of cuFFT and of the oracle."""
    context, queue = gpu
    rs = RandomState(31)
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    g2i = template.instantiate_grid_to_image(queue, (1, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    assert g2i.fused
    g2i.ensure_all_bound()
    half = grid_size // 2
    cells = [(0, 0), (half - 1, half - 1), (-half, -half), (37, -1201), (-half, half - 1),
             (1, 0), (0, -1), (-777, 2048)]
    amps = rs.complex_normal(0.0j, 1.0, len(cells)).astype(np.complex64)
    grid = np.zeros((1, grid_size, grid_size), np.complex64)
    for (v, u), a in zip(cells, amps):
        grid[0, v + half, u + half] += a
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    g2i.buffer('grid').set(queue, grid)
    g2i.buffer('kernel1d').set(queue, kernel1d)
    g2i.buffer('image').zero(queue)
    g2i.set_w(0.0)
    g2i()
    actual = g2i.buffer('image').get(queue)[0]
    rows = np.unique(np.concatenate(([0, 1, pixels // 2 - 1, pixels // 2, pixels - 1],
                                     rs.randint(0, pixels, 40))))
    x = np.arange(pixels, dtype=np.float64) - pixels // 2
    lm = (np.arange(pixels).astype(np.float32) * np.float32(lm_scale) + np.float32(lm_bias))
    lm2 = (lm * lm).astype(np.float64)
    worst = 0.0
    for yi in rows:
        y = float(yi - pixels // 2)
        value = np.zeros(pixels, np.complex128)
        for (v, u), a in zip(cells, amps):
            value += complex(a) * np.exp(2j * np.pi * (u * x + v * y) / pixels)
        n = np.sqrt(1.0 - (lm2[yi] + lm2))
        expected = value.real * n / (float(kernel1d[yi]) * kernel1d.astype(np.float64))
        worst = max(worst, float(np.abs(actual[yi] - expected).max()))
    assert worst < 2e-5 * np.abs(amps).sum()


def test_grid_to_image_fused_padded_buffers(gpu, oracle):
    """The fused transform honours row strides: image, layer (scratch) and grid buffers with
    padded rows give the same image as tightly packed ones."""
    context, queue = gpu
    pixels, grid_size, pols = 2048, 1226, 2
    g2i, grid, kernel1d, lm_scale, lm_bias = _fused_case(
        context, queue, pixels, grid_size, pols, 13, True)
    g2i.set_w(12.25)
    g2i.buffer('image').zero(queue)
    g2i()
    packed = g2i.buffer('image').get(queue)
    padded = {
        'image': accel.DeviceArray(context, (pols, pixels, pixels), np.float32,
                                   (pols, pixels, pixels + 16)),
        'layer': accel.DeviceArray(context, (pixels, pixels), np.complex64,
                                   (pixels, pixels + 16)),
        'grid': accel.DeviceArray(context, (pols, grid_size, grid_size), np.complex64,
                                  (pols, grid_size + 2, grid_size + 6)),
    }
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels + 16))
    other = template.instantiate_grid_to_image(queue, (pols, grid_size, grid_size),
                                               lm_scale, lm_bias, plan)
    other.bind(**padded)
    other.ensure_all_bound()
    other.buffer('grid').set(queue, grid)
    other.buffer('kernel1d').set(queue, kernel1d)
    other.buffer('image').zero(queue)
    other.set_w(12.25)
    other()
    np.testing.assert_array_equal(other.buffer('image').get(queue), packed)


@pytest.mark.parametrize('pixels,grid_size', [(4096, 2470), (8192, 4940), (16384, 1000),
                                              (16384, 9864)])
def test_grid_to_image_fused_vs_cufft(gpu, pixels, grid_size):
    """Full-size planes: the fused transform against the pad + cuFFT + layer_to_image
    sequence on the same random grid."""
    context, queue = gpu
    g2i, grid, kernel1d, lm_scale, lm_bias = _fused_case(
        context, queue, pixels, grid_size, 1, 12, True)
    g2i.set_w(133.5)
    g2i.buffer('image').zero(queue)
    g2i()
    fused = g2i.buffer('image').get(queue)
    g2i.fused = False
    g2i.buffer('image').zero(queue)
    g2i()
    plain = g2i.buffer('image').get(queue)
    peak = np.abs(plain).max()
    rms = np.sqrt(np.mean((fused.astype(np.float64) - plain) ** 2)) / peak
    assert rms < 2e-6
    assert np.abs(fused - plain).max() / peak < 2e-5


@pytest.mark.parametrize('pixels,grid_size,pols', [(2048, 1230, 2), (2048, 2048, 1),
                                                   (2048, 1234, 1), (2048, 30, 1),
                                                   (4096, 2466, 1)])
def test_image_to_grid_fused_vs_oracle(gpu, oracle, pixels, grid_size, pols):
    """The mirrored fused transform (kib_image_to_grid_rows / _columns) against the oracle's
    image_to_grid (reference ImageToGridHost).  Bar 1e-5 relative (north_star: model
    visibilities); observed ~3e-7 RMS of the largest grid value."""
    context, queue = gpu
    rs = RandomState(21)
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    i2g = template.instantiate_image_to_grid(queue, (pols, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    assert i2g.fused
    i2g.ensure_all_bound()
    model = rs.uniform(-1.0, 1.0, (pols, pixels, pixels)).astype(np.float32)
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    i2g.buffer('image').set(queue, model)
    i2g.buffer('kernel1d').set(queue, kernel1d)
    for w in (0.0, 57.25):
        i2g.set_w(w)
        i2g.buffer('grid').zero(queue)
        i2g()
        actual = i2g.buffer('grid').get(queue)
        expected = oracle.image_to_grid(model, kernel1d, lm_scale, lm_bias, np.float64(w),
                                        grid_size=grid_size)
        scale = np.abs(expected).max()
        assert np.sqrt(np.mean(np.abs(actual - expected) ** 2)) / scale < 2e-6
        assert np.abs(actual - expected).max() / scale < 2e-5


@pytest.mark.parametrize('pixels,grid_size', [(8192, 4940), (16384, 1000)])
def test_image_to_grid_fused_vs_cufft(gpu, pixels, grid_size):
    """Full-size planes: fused image -> grid against image_to_layer + cuFFT + layer_to_grid."""
    context, queue = gpu
    rs = RandomState(22)
    lm_scale = 0.2 / pixels
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    i2g = template.instantiate_image_to_grid(queue, (1, grid_size, grid_size),
                                             lm_scale, -lm_scale * pixels / 2, plan)
    i2g.ensure_all_bound()
    model = np.zeros((1, pixels, pixels), np.float32)
    ys = rs.randint(0, pixels, 500)
    xs = rs.randint(0, pixels, 500)
    model[0, ys, xs] = rs.uniform(0.5, 2.0, 500).astype(np.float32)
    i2g.buffer('image').set(queue, model)
    i2g.buffer('kernel1d').set(queue, rs.uniform(1.0, 2.0, pixels).astype(np.float32))
    i2g.set_w(133.5)
    i2g()
    fused = i2g.buffer('grid').get(queue)
    i2g.fused = False
    i2g()
    plain = i2g.buffer('grid').get(queue)
    scale = np.abs(plain).max()
    assert np.sqrt(np.mean(np.abs(fused - plain) ** 2)) / scale < 2e-6
    assert np.abs(fused - plain).max() / scale < 2e-5


@pytest.mark.parametrize('pixels,grid_size', [(8192, 4940), (16384, 9864)])
def test_image_to_grid_fused_impulses(gpu, pixels, grid_size):
    """Full BASELINE sizes, analytic answer: a few non-zero pixels give
    grid[v][u] = sum_k f_k / (k1d[y_k] k1d[x_k] n_k) exp(-2 pi i (u x_k + v y_k) / N) at w = 0
    (coordinates counted from the centres).  Checked on sampled grid rows in float64."""
    context, queue = gpu
    rs = RandomState(32)
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    i2g = template.instantiate_image_to_grid(queue, (1, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    assert i2g.fused
    i2g.ensure_all_bound()
    mid = pixels // 2
    pix = [(0, 0), (mid - 1, mid - 1), (-mid, -mid), (1234, -77), (-mid, mid - 1), (0, 1)]
    flux = rs.uniform(0.5, 2.0, len(pix))
    model = np.zeros((1, pixels, pixels), np.float32)
    for (y, x), f in zip(pix, flux):
        model[0, y + mid, x + mid] = f
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    i2g.buffer('image').set(queue, model)
    i2g.buffer('kernel1d').set(queue, kernel1d)
    i2g.set_w(0.0)
    i2g()
    actual = i2g.buffer('grid').get(queue)[0]
    half = grid_size // 2
    lm = (np.arange(pixels).astype(np.float32) * np.float32(lm_scale) + np.float32(lm_bias))
    lm2 = (lm * lm).astype(np.float64)
    u = np.arange(grid_size, dtype=np.float64) - half
    rows = np.unique(np.concatenate(([0, 1, half - 1, half, grid_size - 1],
                                     rs.randint(0, grid_size, 40))))
    worst = 0.0
    for gy in rows:
        v = float(gy - half)
        expected = np.zeros(grid_size, np.complex128)
        for (y, x), f in zip(pix, flux):
            yi, xi = y + mid, x + mid
            n = np.sqrt(1.0 - (lm2[yi] + lm2[xi]))
            amp = float(np.float32(f)) / (float(kernel1d[yi]) * float(kernel1d[xi]) * n)
            expected += amp * np.exp(-2j * np.pi * (u * x + v * y) / pixels)
        worst = max(worst, float(np.abs(actual[gy] - expected).max()))
    assert worst < 2e-5 * flux.sum()


def test_scale(gpu):
    context, queue = gpu
    shape = (4, 123, 234)
    rs = np.random.RandomState(1)
    fn = image.ScaleTemplate(context, np.float32, shape[0]).instantiate(queue, shape)
    fn.ensure_all_bound()
    src = rs.uniform(size=shape).astype(np.float32)
    scale_factor = np.array([1.2, 2.3, 3.4, -4.5], np.float32)
    fn.buffer('data').set(queue, src)
    fn.set_scale_factor(scale_factor)
    fn()
    np.testing.assert_allclose(src * scale_factor[:, np.newaxis, np.newaxis],
                               fn.buffer('data').get(queue))
    with pytest.raises(ValueError):
        image.ScaleTemplate(context, np.float32, 4).instantiate(queue, (3, 8, 8))


def test_add_image_different_padding(gpu):
    """reference test_image.py TestAddImage: src and dest padded differently."""
    context, queue = gpu
    shape = (4, 123, 234)
    rs = np.random.RandomState(1)
    fn = image.AddImageTemplate(context, np.float32, shape[0]).instantiate(queue, shape)
    for slot, padded in ((fn.slots['src'], (4, 128, 256)), (fn.slots['dest'], (4, 135, 240))):
        for i, size in enumerate(padded):
            slot.dimensions[i].link(accel.Dimension(slot.shape[i], min_padded_size=size))
    fn.ensure_all_bound()
    assert fn.buffer('src').padded_shape == (4, 128, 256)
    assert fn.buffer('dest').padded_shape == (4, 135, 240)
    src_host = rs.uniform(size=shape).astype(np.float32)
    dest_host = rs.uniform(size=shape).astype(np.float32)
    fn.buffer('src').set(queue, src_host)
    fn.buffer('dest').set(queue, dest_host)
    fn()
    np.testing.assert_allclose(src_host + dest_host, fn.buffer('dest').get(queue))


def test_apply_primary_beam(gpu):
    context, queue = gpu
    shape = (4, 123, 234)
    rs = np.random.RandomState(1)
    fn = image.ApplyPrimaryBeamTemplate(context, np.float32, shape[0]).instantiate(
        queue, shape, 0.2, 12345.0)
    fn.ensure_all_bound()
    src = rs.uniform(size=shape).astype(np.float32)
    beam = rs.uniform(size=shape[1:]).astype(np.float32)
    fn.buffer('data').set(queue, src)
    fn.buffer('beam_power').set(queue, beam)
    fn()
    np.testing.assert_allclose(np.where(beam < 0.2, 12345.0, src / beam),
                               fn.buffer('data').get(queue))
    # NaN replacement as used for the dirty image (imaging.py:130-131)
    fn.replacement = np.nan
    fn.buffer('data').set(queue, src)
    fn()
    out = fn.buffer('data').get(queue)
    assert np.all(np.isnan(out[:, beam < 0.2])) and not np.any(np.isnan(out[:, beam >= 0.2]))


def test_image_to_grid_sparse_model_rows(gpu, oracle):
    """All-zero image rows are answered without a transform (sparse_model, the CLEAN-model
    case): same grid as the dense route, also when the planes are occupied differently
    (one of them not at all) and for a -0.0 / NaN-free image."""
    context, queue = gpu
    rs = RandomState(41)
    pixels, grid_size, pols = 2048, 1230, 3
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    i2g = template.instantiate_image_to_grid(queue, (pols, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    i2g.ensure_all_bound()
    model = np.zeros((pols, pixels, pixels), np.float32)
    model[0, rs.randint(0, pixels, 40), rs.randint(0, pixels, 40)] = rs.uniform(0.5, 2.0, 40)
    model[2, rs.randint(0, pixels, 7), rs.randint(0, pixels, 7)] = rs.uniform(-2.0, -0.5, 7)
    model[2, 5, :] = -0.0
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    i2g.buffer('image').set(queue, model)
    i2g.buffer('kernel1d').set(queue, kernel1d)
    i2g.set_w(21.5)
    results = []
    for sparse in (True, False):
        i2g.sparse_model = sparse
        i2g.buffer('grid').set(queue, np.full((pols, grid_size, grid_size), 7 + 7j, np.complex64))
        i2g()
        results.append(i2g.buffer('grid').get(queue))
    scale = np.abs(results[1]).max()
    assert np.abs(results[0] - results[1]).max() <= 1e-6 * scale
    assert not results[0][1].any()                      # the empty plane gives an exactly zero grid
    expected = oracle.image_to_grid(model, kernel1d, lm_scale, lm_bias, np.float64(21.5),
                                    grid_size=grid_size)
    assert np.abs(results[0] - expected).max() / np.abs(expected).max() < 2e-5


def _occupancy_mask(occupied_groups, grid_size):
    words = np.zeros(image.occupancy_words(grid_size), np.uint32)
    for g in occupied_groups:
        words[g >> 5] |= np.uint32(1) << np.uint32(g & 31)
    return words


def test_column_occupancy_kernel(gpu):
    """kib_column_occupancy against numpy: footprints u0 .. u0 + K - 1 with the gridder's
    origin, out-of-range coordinates ignored, both from a packed uv array (stride 8) and from
    inside preprocessed records (stride 12 + 12 P)."""
    context, queue = gpu
    rs = RandomState(5)
    for grid_size, K in ((1230, 7), (4922, 7), (9864, 32), (512, 60)):
        n = 5000
        bias = (K - 1) // 2 - grid_size // 2
        limit = grid_size // 2
        u = rs.randint(-limit // 3, limit // 2, n).astype(np.int16)
        u[:20] = rs.randint(-limit - 40, -limit + 40, 20)          # around / beyond the left edge
        u[20:40] = rs.randint(limit - 70, limit + 40, 20)          # around / beyond the right edge
        expected = set()
        for x in u.astype(int):
            u0 = x - bias
            if u0 < 0 or u0 + K > grid_size:
                continue
            expected.update(range(u0 >> 3, ((u0 + K - 1) >> 3) + 1))
        want = _occupancy_mask(expected, grid_size)
        for stride in (8, 60):
            raw = np.zeros((n, stride), np.uint8)
            raw[:, :2] = u.view(np.uint8).reshape(n, 2)
            raw[:, 2:] = rs.randint(0, 255, (n, stride - 2))
            dev = accel.DeviceArray(context, raw.shape, np.uint8)
            dev.set(queue, raw)
            occ = image.column_occupancy(queue, dev, n, K, grid_size, stride)
            np.testing.assert_array_equal(occ.get(queue), want)


@pytest.mark.parametrize('pixels,grid_size,pols', [(2048, 1230, 2), (4096, 2470, 1),
                                                   (8192, 4922, 2), (16384, 9864, 1)])
def test_grid_to_image_occupancy(gpu, pixels, grid_size, pols):
    """A grid that is zero outside some groups of 8 columns, transformed with and without its
    occupancy mask: bit-identical images, and the masked route must not even read the other
    columns (they are filled with NaN for it)."""
    context, queue = gpu
    g2i, grid, kernel1d, lm_scale, lm_bias = _fused_case(
        context, queue, pixels, grid_size, pols, 21, True)
    rs = RandomState(22)
    groups = (grid_size + 7) // 8
    occupied = sorted(set(rs.randint(0, groups, groups // 3)) | {0, groups - 1})
    keep = np.zeros(grid_size, bool)
    for g in occupied:
        keep[8 * g:8 * g + 8] = True
    sparse = np.where(keep[np.newaxis, np.newaxis, :], grid, 0).astype(np.complex64)
    g2i.set_w(133.5)
    g2i.buffer('grid').set(queue, sparse)
    g2i.buffer('image').zero(queue)
    g2i()
    dense = g2i.buffer('image').get(queue)
    poisoned = np.where(keep[np.newaxis, np.newaxis, :], grid, np.nan + 1j * np.nan)
    g2i.buffer('grid').set(queue, poisoned.astype(np.complex64))
    # stale data in the half-transformed plane must not matter either
    layer = g2i.buffer('layer')
    layer.set(queue, np.full(layer.shape, np.nan, layer.dtype))
    occ = accel.DeviceArray(context, (image.occupancy_words(grid_size),), np.uint32)
    occ.set(queue, _occupancy_mask(occupied, grid_size))
    g2i.occupancy = occ
    g2i.buffer('image').zero(queue)
    g2i()
    masked = g2i.buffer('image').get(queue)
    assert np.isfinite(masked).all()
    np.testing.assert_array_equal(masked, dense)
    assert np.abs(dense).max() > 0


@pytest.mark.parametrize('pixels,grid_size,sparse_model', [(2048, 1230, True), (2048, 1230, False),
                                                           (8192, 4922, True), (16384, 9864, False)])
def test_image_to_grid_occupancy(gpu, pixels, grid_size, sparse_model):
    """image -> grid restricted to an occupancy mask: the occupied column groups are
    bit-identical to the dense transform; every other column is either left untouched or (when
    it shares a 16-column tile with an occupied group) computed all the same."""
    context, queue = gpu
    from katsdpimager_b200 import _lib
    rs = RandomState(31)
    pols = 2
    lm_scale = 0.2 / pixels
    lm_bias = -lm_scale * pixels / 2
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    i2g = template.instantiate_image_to_grid(queue, (pols, grid_size, grid_size),
                                             lm_scale, lm_bias, plan)
    i2g.ensure_all_bound()
    i2g.sparse_model = sparse_model
    if sparse_model and not _lib.load().kib_image_to_grid_sparse_supported(
            pixels, grid_size, _lib.dtype_code(np.dtype(np.complex64))):
        pytest.skip('no sparse route at this size')
    model = np.zeros((pols, pixels, pixels), np.float32)
    model[0, rs.randint(0, pixels, 60), rs.randint(0, pixels, 60)] = rs.uniform(0.5, 2.0, 60)
    model[1, rs.randint(0, pixels, 9), rs.randint(0, pixels, 9)] = rs.uniform(-2.0, -0.5, 9)
    i2g.buffer('image').set(queue, model)
    i2g.buffer('kernel1d').set(queue, rs.uniform(1.0, 2.0, pixels).astype(np.float32))
    i2g.set_w(21.5)
    i2g()
    dense = i2g.buffer('grid').get(queue)
    groups = (grid_size + 7) // 8
    occupied = sorted(set(rs.randint(0, groups, groups // 4)) | {0, groups - 1})
    keep = np.zeros(grid_size, bool)
    for g in occupied:
        keep[8 * g:8 * g + 8] = True
    occ = accel.DeviceArray(context, (image.occupancy_words(grid_size),), np.uint32)
    occ.set(queue, _occupancy_mask(occupied, grid_size))
    i2g.occupancy = occ
    sentinel = np.complex64(7 + 7j)
    i2g.buffer('grid').set(queue, np.full((pols, grid_size, grid_size), sentinel, np.complex64))
    i2g()
    masked = i2g.buffer('grid').get(queue)
    np.testing.assert_array_equal(masked[:, :, keep], dense[:, :, keep])
    rest = masked[:, :, ~keep]
    untouched = (rest == sentinel).all(axis=(0, 1))
    computed = (rest == dense[:, :, ~keep]).all(axis=(0, 1))
    assert (untouched | computed).all()
    assert untouched.sum() > 0.4 * rest.shape[2]
    assert np.abs(dense).max() > 0


def test_clear_columns(gpu):
    """kib_clear_columns zeroes exactly the occupied groups of 8 columns, in every
    polarization, also with a padded grid and a last group that is cut by the grid edge."""
    context, queue = gpu
    from katsdpimager_b200 import _lib
    rs = RandomState(7)
    for grid_size, padded in ((1230, 1232), (500, 500)):
        pols = 3
        groups = (grid_size + 7) // 8
        occupied = sorted(set(rs.randint(0, groups, groups // 2)) | {groups - 1})
        keep = np.zeros(grid_size, bool)
        for g in occupied:
            keep[8 * g:8 * g + 8] = True
        dev = accel.DeviceArray(context, (pols, grid_size, grid_size), np.complex64,
                                (pols, grid_size + 3, padded))
        ref = rs.complex_normal(0.0j, 1.0, dev.shape).astype(np.complex64)
        dev.set(queue, ref)
        occ = accel.DeviceArray(context, (image.occupancy_words(grid_size),), np.uint32)
        occ.set(queue, _occupancy_mask(occupied, grid_size))
        _lib.call('kib_clear_columns', dev.ptr, dev.padded_shape[2],
                  dev.padded_shape[1] * dev.padded_shape[2], grid_size, pols, occ.ptr,
                  _lib.dtype_code(dev.dtype), queue.stream)
        out = dev.get(queue)
        assert not out[:, :, keep].any()
        np.testing.assert_array_equal(out[:, :, ~keep], ref[:, :, ~keep])


@pytest.mark.parametrize('pixels,grid_size', [(2048, 1230), (8192, 4922)])
def test_grid_to_image_symmetric_factors(gpu, pixels, grid_size):
    """Quadrant factor table (factor modes 3 / 4: lm_bias = -N/2 lm_scale, symmetric taper)
    against the full factor plane, with and without the cache across calls: the images agree
    to the rounding of the direction cosines (l(N - x) = -l(x) holds to one ulp only)."""
    context, queue = gpu
    pols = 3
    g2i, grid, kernel1d, lm_scale, lm_bias = _fused_case(
        context, queue, pixels, grid_size, pols, 23, True)
    kernel1d[1:] = 0.5 * (kernel1d[1:] + kernel1d[1:][::-1])
    g2i.buffer('kernel1d').set(queue, kernel1d)
    g2i.set_w(133.5)
    g2i.buffer('image').zero(queue)
    g2i()
    full = g2i.buffer('image').get(queue)
    g2i.symmetric_factors = True
    g2i.factor_cache_planes = 2
    results = []
    for _ in range(2):                      # second call: every plane loads the cached quadrant
        g2i.buffer('image').zero(queue)
        g2i()
        results.append(g2i.buffer('image').get(queue))
    peak = np.abs(full).max()
    for img in results:
        # (single pixels at the image edge, where the taper division amplifies the rounding,
        # reach a few 1e-5 -- as between any two evaluations of the factor)
        assert np.abs(img - full).max() <= 1e-4 * peak
        assert np.sqrt(np.mean((img.astype(np.float64) - full) ** 2)) <= 1e-6 * peak
    # planes 1, 2 of the first call and all planes of the second read the same table
    np.testing.assert_array_equal(results[0][1:], results[1][1:])
