"""Host-side logic that needs neither a GPU nor the shared library's kernels: the
kernel look-up-table generator, parameter formulas, the preprocessing restatement,
slot/dimension negotiation of the accel layer and CLEAN bookkeeping."""
import numpy as np
import pytest

from katsdpimager_b200 import accel, clean, grid, parameters as prm, predict, preprocess, simulate
from tests import cases
from tests.cases import load_golden


def test_extract_sky_image():
    """reference test_predict.py test_extract_sky_image known answers + golden."""
    fx = cases.predict_case()
    golden = load_golden('predict_small')
    lmn, flux = predict._extract_sky_image(fx['image_parameters'], fx['grid_parameters'],
                                           fx['components'])
    np.testing.assert_allclose(lmn, golden['image_lmn'], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(flux, golden['image_flux'], rtol=1e-6)
    np.testing.assert_allclose(lmn[:, 0:2], [[2047e-5, -2048e-5], [-1536e-5, -1024e-5], [0, 0],
                                             [-2048e-5, 2047e-5]], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(
        predict._uvw_scale_bias(fx['image_parameters'], fx['grid_parameters']),
        golden['scale_bias'], rtol=1e-12)




@pytest.mark.parametrize('name', ['test_grid', 'meerkat_k7'])
def test_convolution_kernel_golden(name):
    """The product's LUT generator against the reference's ConvolutionKernel output."""
    ip, gp = cases.lut_cases()[name]
    golden = load_golden('lut_' + name)
    kernel = grid.ConvolutionKernel(ip, gp)
    assert kernel.data.shape == (gp.w_planes, gp.fixed.oversample, gp.fixed.kernel_width)
    np.testing.assert_allclose(kernel.data, golden['data'], rtol=0, atol=2e-7)
    np.testing.assert_allclose(kernel.taper(ip.pixels), golden['taper'], rtol=1e-12)
    assert kernel.beta == pytest.approx(float(golden['beta']))
    out = np.empty(ip.pixels, np.float32)
    assert kernel.taper(ip.pixels, out) is out
    np.testing.assert_allclose(out, golden['taper'], rtol=1e-6)


def test_kaiser_bessel_pair():
    """kaiser_bessel_fourier is the Fourier transform of kaiser_bessel."""
    width, beta = 7.0, grid.antialias_beta(7.0)
    x = np.linspace(-3.5, 3.5, 7001)
    window = grid.kaiser_bessel(x, width, beta)
    assert window[0] == pytest.approx(1 / np.i0(beta)) and window[3500] == pytest.approx(1.0)
    assert grid.kaiser_bessel(np.array([3.6]), width, beta)[0] == 0.0
    for f in (0.0, 0.05, 0.3, 0.8):
        numeric = np.trapezoid(window * np.cos(2 * np.pi * f * x), x)
        assert grid.kaiser_bessel_fourier(f, width, beta) == pytest.approx(numeric, abs=1e-6)
    assert grid.subpixel_coord(-1.3, 8) == (-2, 5)
    assert grid.subpixel_coord(2.99, 8) == (2, 7)


def test_parameters():
    assert prm.is_smooth(8192) and prm.is_smooth(5760) and not prm.is_smooth(8200)
    assert prm.next_smooth(8193) == 8232
    array = prm.ArrayParameters(13.5, 7697.58)
    fixed = prm.FixedImageParameters([1], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.2155, pixels=2048, array=array)
    # pixel_size = wavelength / (2/3 * 5 * longest baseline), cell = wavelength / image size
    assert ip.pixel_size == pytest.approx(0.2155 / (10 / 3 * 7697.58))
    assert ip.cell_size == pytest.approx(12.53, rel=1e-3)
    with pytest.raises(ValueError):
        prm.ImageParameters(fixed, wavelength=0.2, pixels=2050, pixel_size=1e-5)
    with pytest.raises(ValueError):
        prm.ImageParameters(fixed, wavelength=0.2, pixels=None, array=array, image_oversample=2)
    auto = prm.ImageParameters(fixed, wavelength=0.2155, pixels=None, array=array)
    assert prm.is_smooth(auto.pixels)
    slices = prm.w_slices(ip, 7697.58, 0.001, 60, 7.0)
    assert prm.w_kernel_width(ip, 0.5 * 7697.58 / (slices - 0.5), 0.001, 7.0) < 60
    if slices > 1:
        assert prm.w_kernel_width(ip, 0.5 * 7697.58 / (slices - 1.5), 0.001, 7.0) >= 60
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, 7697.58, 60), slices, 100)
    mid_w = prm.slice_mid_w(ip, gp)
    assert mid_w[0] == 0 and len(mid_w) == slices
    with pytest.raises(ValueError):
        prm.CleanParameters(10, 0.1, 0.85, 5.0, 0, 1.0, 0.5, 0.02)


def test_preprocess_golden_records():
    """Hand-computed records of the reference's preprocessing test
    (reference katsdpimager/test/test_preprocess.py:76-136, identity Mueller matrix)."""
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
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    fixed_grid = prm.FixedGridParameters(7.0, 8, 4, 400.0, 64)
    for channel, wavelength in enumerate([0.25, 0.125]):
        ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=2048,
                                 pixel_size=1.0 / (4096.0 * wavelength))
        gp = prm.GridParameters(fixed_grid, 1, 128)
        records, w_slice = preprocess.quantise(uvw, weights[channel], vis[channel], ip, gp)
        records, w_slice = preprocess.compress(records, w_slice)
        slices = preprocess.bucket_by_slice(records, w_slice, 1)
        actual = slices[0]
        want = expected[channel]
        np.testing.assert_array_equal(actual.uv, want['uv'])
        np.testing.assert_array_equal(actual.sub_uv, want['sub_uv'])
        np.testing.assert_array_equal(actual.w_plane, want['w_plane'])
        np.testing.assert_allclose(actual.weights, want['weights'], rtol=1e-6)
        np.testing.assert_allclose(actual.vis, want['vis'], rtol=1e-5)
    # record layout is the one Imaging.set_coordinates relies on (imaging.py:63-78)
    dtype = preprocess.make_dtype(4)
    assert dtype.fields['sub_uv'][1] == dtype.fields['uv'][1] + 4
    assert dtype.itemsize == 60
    # w < 0 flips the baseline and conjugates
    ip = prm.ImageParameters(fixed, wavelength=0.25, pixels=2048, pixel_size=1.0 / 1024)
    gp = prm.GridParameters(fixed_grid, 4, 16)
    rec, w_slice = preprocess.quantise(np.array([[3.0, -2.0, -50.0], [-3.0, 2.0, 50.0]], np.float32),
                                       np.ones((2, 4), np.float32),
                                       np.full((2, 4), 1 + 2j, np.complex64), ip, gp)
    np.testing.assert_array_equal(rec.uv[0], rec.uv[1])
    np.testing.assert_array_equal(rec.vis[0], np.conj(rec.vis[1]))
    assert w_slice[0] == w_slice[1]
    reader = preprocess.VisibilityReaderMem([preprocess.bucket_by_slice(rec, w_slice, 4)])
    assert sum(reader.len(0, s) for s in range(reader.num_w_slices(0))) == 2
    assert [len(c) for c in reader.iter_slice(0, int(w_slice[0]), 1)] == [1, 1]


def test_simulate():
    enu = simulate.meerkat_enu()
    assert enu.shape == (64, 3)
    assert simulate.longest_baseline() == pytest.approx(7697.58, abs=0.01)
    uvw = simulate.uvw_tracks(16, dump_time=4.0)
    assert uvw.shape == (2016, 16, 3)
    # rotation preserves baseline length
    np.testing.assert_allclose(np.linalg.norm(uvw, axis=2)[:, 0],
                               np.linalg.norm(simulate.baselines_enu(), axis=1), rtol=1e-12)
    # consecutive dumps move by much less than a cell of an 8192-pixel image (3.13 m)
    assert np.abs(np.diff(uvw, axis=1)).max() < 3.0
    lmn, flux = simulate.lsm_lmn_flux()
    assert lmn.shape == (8, 3) and flux.shape == (8, 4) and np.all(lmn[0] == 0)
    vis = simulate.dft_visibilities(np.array([[100.0, -50.0, 3.0]]), lmn[:2], flux[:2])
    expected = flux[0] + flux[1] * np.exp(-2j * np.pi * (np.array([100.0, -50.0, 3.0]) @ lmn[1]))
    np.testing.assert_allclose(vis[0], expected, rtol=1e-5)


# ------------------------------------------------------------------ accel layer (no device)
def test_dimension_linking():
    a = accel.Dimension(100)
    b = accel.Dimension(100, alignment=16)
    c = accel.Dimension(100, min_padded_size=130)
    a.link(b)
    b.link(c)
    assert a.required_padded_size() == 144 and c.required_padded_size() == 144
    assert a.valid(144) and a.valid(160) and not a.valid(130) and not a.valid(136)
    with pytest.raises(ValueError):
        accel.Dimension(100).link(accel.Dimension(101))
    with pytest.raises(ValueError):
        accel.Dimension(100, exact=True).link(accel.Dimension(100, alignment=16))
    exact = accel.Dimension(4, exact=True)
    assert exact.required_padded_size() == 4 and exact.valid(4) and not exact.valid(8)
    with pytest.raises(ValueError):
        a.link(accel.Dimension(100))        # padding already queried
    assert accel.Dimension(10, min_padded_round=8).required_padded_size() == 16
    assert accel.divup(10, 4) == 3 and accel.roundup(10, 4) == 12


def test_slots_and_sequences():
    class Op(accel.Operation):
        def __init__(self, dims, dtype=np.float32):
            self.command_queue = None
            self.allocator = None
            from collections import OrderedDict
            self.slots = OrderedDict()
            self.hidden_slots = OrderedDict()
            self.slots['data'] = accel.IOSlot(dims, dtype)

        def _run(self):
            pass

    class FakeBuffer:
        def __init__(self, shape, padded_shape, dtype=np.float32):
            self.shape, self.padded_shape, self.dtype = shape, padded_shape, np.dtype(dtype)

    op1 = Op((3, accel.Dimension(10, alignment=4), 20))
    op2 = Op((3, 10, accel.Dimension(20, min_padded_size=24)))
    op3 = Op((7,))
    # build the slot aliasing exactly as OperationSequence does, without a device
    compound = accel.CompoundIOSlot([op1.slots['data'], op2.slots['data']])
    assert compound.required_padded_shape() == (3, 12, 24)
    assert compound.required_bytes() == 3 * 12 * 24 * 4
    good = FakeBuffer((3, 10, 20), (3, 12, 24))
    compound.bind(good)
    assert op1.slots['data'].buffer is good and op2.slots['data'].buffer is good
    compound.bind(None)
    assert not op1.slots['data'].is_bound()
    with pytest.raises(ValueError):
        compound.bind(FakeBuffer((3, 10, 20), (3, 10, 24)))
    with pytest.raises(ValueError):
        compound.bind(FakeBuffer((3, 10, 21), (3, 12, 24)))
    with pytest.raises(TypeError):
        compound.bind(FakeBuffer((3, 10, 20), (3, 12, 24), np.float64))
    with pytest.raises(ValueError):
        accel.CompoundIOSlot([op1.slots['data'], op3.slots['data']])
    with pytest.raises(TypeError):
        accel.CompoundIOSlot([op3.slots['data'], Op((7,), np.int32).slots['data']])


def test_region_layout():
    axes = accel._normalise_region((4, 20, 30), np.s_[2, 3:10, ...])
    assert axes == [(2, 1, False), (3, 7, True), (0, 30, True)]
    offset, dims = accel._layout((4, 24, 32), 8, axes)
    assert offset == (2 * 24 * 32 + 3 * 32) * 8
    assert dims == [(7, 32 * 8), (30, 8)]
    assert accel._normalise_region((5,), np.s_[-2]) == [(3, 1, False)]
    assert accel._normalise_region((5, 6), np.s_[:, -4:]) == [(0, 5, True), (2, 4, True)]
    for bad in (np.s_[::2], np.s_[7], np.s_[0, 0]):
        with pytest.raises(IndexError):
            accel._normalise_region((5,), bad)


def test_clean_helpers():
    assert clean.metric_to_power(clean.CLEAN_I, 4.0) == 4.0
    assert clean.metric_to_power(clean.CLEAN_SUMSQ, 4.0) == 2.0
    assert clean.power_to_metric(clean.CLEAN_SUMSQ, 3.0) == 9.0
    assert clean.noise_threshold_scale(clean.CLEAN_I, 5.0, 4) == 5.0
    # chi-squared threshold with one degree of freedom reduces to the Gaussian one
    assert clean.noise_threshold_scale(clean.CLEAN_SUMSQ, 5.0, 1) == pytest.approx(5.0, rel=1e-6)
    assert clean.noise_threshold_scale(clean.CLEAN_SUMSQ, 5.0, 4) > 5.0
    for fn in (clean.metric_to_power, clean.power_to_metric):
        with pytest.raises(ValueError):
            fn(2, 1.0)


def test_column_occupancy_fractions_match_brute_force():
    """bench.column_occupancy_fractions (the byte accounting of the rooflines) against a
    cell-by-cell count of the footprints."""
    import bench
    rs = np.random.RandomState(11)
    grid_size = 1230
    K = bench.KERNEL_WIDTH
    bias = (K - 1) // 2 - grid_size // 2
    dtype = np.dtype([('uv', np.int16, (2,))])
    slices = []
    for n in (0, 1, 40, 3000):
        s = np.zeros(n, dtype).view(np.recarray)
        s.uv[:, 0] = rs.randint(-grid_size // 2 - 20, grid_size // 2 + 20, n)
        slices.append(s)
    got = bench.column_occupancy_fractions(slices, grid_size)
    groups = (grid_size + 7) // 8
    for s, fraction in zip(slices, got):
        mask = np.zeros(groups, bool)
        for u in s.uv[:, 0].astype(int):
            u0 = u - bias
            if u0 < 0 or u0 + K > grid_size:
                continue
            for c in range(u0, u0 + K):
                mask[c >> 3] = True
        assert fraction == pytest.approx(mask.sum() / groups)


def test_grid_clear_bookkeeping():
    """Imaging._clear_covers: a clear made for one occupancy mask only serves that very mask
    (same object, not refilled since); a whole-buffer clear serves everything."""
    from katsdpimager_b200.imaging import Imaging

    class Mask:
        generation = 0
    a, b = Mask(), Mask()
    assert Imaging._clear_covers(None, None) and Imaging._clear_covers(None, a)
    assert Imaging._clear_covers((a, 0), a)
    assert not Imaging._clear_covers((a, 0), b)
    assert not Imaging._clear_covers((a, 0), None)
    a.generation = 1                # refilled in place for another channel
    assert not Imaging._clear_covers((a, 0), a)


def test_make_dirty_call_sequence():
    """pipeline.make_dirty replays frontend.make_dirty (reference frontend.py:110-149): per
    non-empty W slice model_to_grid, clear_grid, (feed, predict, grid) per chunk, grid_to_image;
    with resident records every call carries the slice's occupancy mask, the clear also the
    next slice's (wrapping to the first: the next pass), and only the first model_to_grid of
    a pass looks for the model's empty rows."""
    from katsdpimager_b200 import pipeline

    class Buffer:
        shape = (4, 64, 64)

    class FakeImager:
        command_queue = None

        def __init__(self, with_kernel_width=True):
            self.calls = []
            if with_kernel_width:
                self.kernel_width = 7

        def buffer(self, name):
            return Buffer()

        def __getattr__(self, name):
            if name.startswith('_') or name == 'kernel_width':
                raise AttributeError(name)

            def record(*args, **kwargs):
                self.calls.append((name, args, kwargs))
            return record

    class FakeVis:
        counts = [5, 0, 3, 2]
        num_w_slices = 4

        def len(self, w_slice):
            return self.counts[w_slice]

        def chunks(self, w_slice, block):
            n = self.counts[w_slice]
            return [(s, min(block, n - s)) for s in range(0, n, block)]

        def occupancy(self, queue, w_slice, kernel_width, grid_size):
            assert (kernel_width, grid_size) == (7, 64)
            return 'mask%d' % w_slice

        def feed(self, imager, w_slice, start, count, field, with_weights):
            imager.calls.append(('feed', (w_slice, start, count, field, with_weights), {}))

    mid_w = [0.0, 1.0, 2.0, 3.0]
    imager = FakeImager()
    pipeline.make_dirty(imager, FakeVis(), 'vis', mid_w, 4, True, full_cycle=True)
    names = [c[0] for c in imager.calls]
    assert names == ['clear_dirty',
                     'model_to_grid', 'clear_grid', 'feed', 'predict', 'grid', 'feed', 'predict',
                     'grid', 'grid_to_image',
                     'model_to_grid', 'clear_grid', 'feed', 'predict', 'grid', 'grid_to_image',
                     'model_to_grid', 'clear_grid', 'feed', 'predict', 'grid', 'grid_to_image']
    by_name = {}
    for name, args, kwargs in imager.calls:
        by_name.setdefault(name, []).append((args, kwargs))
    assert [k for _, k in by_name['clear_grid']] == [
        {'occupancy': 'mask0', 'next_occupancy': 'mask2'},
        {'occupancy': 'mask2', 'next_occupancy': 'mask3'},
        {'occupancy': 'mask3', 'next_occupancy': 'mask0'}]
    assert by_name['model_to_grid'] == [
        ((0.0,), {'occupancy': 'mask0', 'model_unchanged': False}),
        ((2.0,), {'occupancy': 'mask2', 'model_unchanged': True}),
        ((3.0,), {'occupancy': 'mask3', 'model_unchanged': True})]
    assert by_name['grid_to_image'] == [((0.0,), {'occupancy': 'mask0'}),
                                        ((2.0,), {'occupancy': 'mask2'}),
                                        ((3.0,), {'occupancy': 'mask3'})]
    assert by_name['feed'][0][0] == (0, 0, 4, 'vis', True)
    assert by_name['feed'][1][0] == (0, 4, 1, 'vis', True)
    # PSF pass without degridding calls, occupancy switched off: the reference's plain calls
    imager = FakeImager()
    pipeline.make_dirty(imager, FakeVis(), 'weights', mid_w, 8, True, use_occupancy=False)
    assert [c[0] for c in imager.calls] == ['clear_dirty'] + ['clear_grid', 'feed', 'grid',
                                                              'grid_to_image'] * 3
    assert all(not kwargs for _, _, kwargs in imager.calls)
    # an imager without the extensions (the reference's own Imaging): plain calls as well
    imager = FakeImager(with_kernel_width=False)
    pipeline.make_dirty(imager, FakeVis(), 'vis', mid_w, 8, True, full_cycle=True)
    assert all(not kwargs for _, _, kwargs in imager.calls)
    assert [c[0] for c in imager.calls].count('model_to_grid') == 3
