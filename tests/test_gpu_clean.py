"""CLEAN kernels: the reference's unit tests (reference katsdpimager/test/test_clean.py),
bit-exact parity of the composite minor cycle with CleanHost (golden + oracle), and the
tie-break / edge cases the reference leaves untested."""
import numpy as np
import pytest
import scipy.signal.windows

from katsdpimager_b200 import clean, parameters as prm
from tests import cases
from tests.cases import load_golden

pytestmark = pytest.mark.gpu


class TestPsfPatch:
    """reference test_clean.py _TestPsfPatchBase known answers."""

    @pytest.fixture(autouse=True)
    def setup(self, gpu):
        context, queue = gpu
        self.fn = clean.PsfPatchTemplate(context, np.float32, 4).instantiate(queue, (4, 206, 304))
        self.fn.ensure_all_bound()
        self.psf = self.fn.buffer('psf')
        self.psf_host = self.psf.empty_like()
        self.psf_host.fill(0.0)
        self.psf_host[:, 103, 152] = 1.0

    def run(self, threshold=0.01, limit=None):
        self.psf.set(self.fn.command_queue, self.psf_host)
        return self.fn(threshold, limit)

    def test_peak_only(self):
        assert self.run() == (4, 1, 1)

    def test_low_corner(self):
        self.psf_host[0, 0, 0] = 0.1
        assert self.run() == (4, 206, 304)

    def test_high_corner(self):
        self.psf_host[3, 205, 303] = -0.2
        assert self.run() == (4, 205, 303)

    def test_1d(self):
        target = self.psf_host[1, 0, :152]
        target[:] = np.arange(152)
        threshold = 50.5
        box = self.run(threshold=threshold)
        hw = box[2] // 2
        assert sum(target[:-hw] >= threshold) == 0
        assert target[-hw] >= threshold

    def test_limit(self):
        self.psf_host[0, 0, 0] = 0.4
        self.psf_host[3, 205, 303] = 0.3
        self.psf_host[1, 110, 150] = 0.2
        assert self.run(limit=50 / 206) == (4, 15, 5)


def _zero(queue, buf):
    buf.set(queue, np.zeros(buf.shape, buf.dtype))


def test_update_tiles(gpu):
    """reference test_clean.py TestClean.test_update_tiles."""
    context, queue = gpu
    image_shape = (4, 567, 456)
    border_pixels = 65
    border = border_pixels / image_shape[2]
    rs = np.random.RandomState(seed=1)
    template = clean._UpdateTilesTemplate(context, np.float32, 4, clean.CLEAN_I)
    fn = template.instantiate(queue, image_shape, border)
    fn.ensure_all_bound()
    dirty = rs.standard_normal(image_shape).astype(np.float32)
    fn.buffer('dirty').set(queue, dirty)
    _zero(queue, fn.buffer('tile_max'))
    _zero(queue, fn.buffer('tile_pos'))
    fn(135, 161, 385, 450)
    tile_max = fn.buffer('tile_max').get(queue)
    tile_pos = fn.buffer('tile_pos').get(queue)
    num_tiles_y, num_tiles_x = tile_max.shape
    assert (num_tiles_y, num_tiles_x) == (14, 11)
    for y in range(num_tiles_y):
        for x in range(num_tiles_x):
            if x < 2 or x >= 10 or y < 3 or y >= 13:
                assert tile_max[y, x] == 0.0
                assert tuple(tile_pos[y, x]) == (0, 0)
            else:
                y0 = y * template.tiley + border_pixels
                x0 = x * template.tilex + border_pixels
                y1 = min(y0 + template.tiley, dirty.shape[1] - border_pixels)
                x1 = min(x0 + template.tilex, dirty.shape[2] - border_pixels)
                tile = np.abs(dirty[0, y0:y1, x0:x1])
                pos = np.unravel_index(np.argmax(tile), tile.shape)
                assert tile[pos] == tile_max[y, x]
                assert (pos[0] + y0, pos[1] + x0) == tuple(tile_pos[y, x])


def test_find_peak(gpu):
    """reference test_clean.py TestClean.test_find_peak."""
    context, queue = gpu
    image_shape = (4, 256, 256)
    tile_shape = (72, 67)
    fn = clean._FindPeakTemplate(context, np.float32, 4).instantiate(queue, image_shape, tile_shape)
    fn.ensure_all_bound()
    rs = np.random.RandomState(seed=1)
    dirty = rs.uniform(1.0, 2.0, image_shape).astype(np.float32)
    tile_max = rs.uniform(1.0, 2.0, tile_shape).astype(np.float32)
    tile_pos = np.array(
        [[[y, x] for x in range(tile_shape[1])] for y in range(tile_shape[0])], np.int32)
    fn.buffer('tile_max').set(queue, tile_max)
    fn.buffer('tile_pos').set(queue, tile_pos)
    fn.buffer('dirty').set(queue, dirty)
    fn()
    peak_value = fn.buffer('peak_value').get(queue)
    peak_pos = fn.buffer('peak_pos').get(queue)
    peak_pixel = fn.buffer('peak_pixel').get(queue)
    best = np.unravel_index(np.argmax(tile_max), tile_max.shape)
    assert tile_max[best] == peak_value[0]
    np.testing.assert_array_equal(tile_pos[best], peak_pos)
    np.testing.assert_array_equal(dirty[:, peak_pos[0], peak_pos[1]], peak_pixel)
    # ties: np.argmax takes the first maximum in row-major order
    tile_max[:] = 1.0
    for idx in [(70, 3), (12, 60), (12, 7), (40, 0)]:
        tile_max[idx] = 5.0
    fn.buffer('tile_max').set(queue, tile_max)
    fn()
    np.testing.assert_array_equal(fn.buffer('peak_pos').get(queue), [12, 7])


def test_subtract_psf(gpu):
    """reference test_clean.py TestClean.test_subtract_psf (patch clipped by the image)."""
    context, queue = gpu
    loop_gain = 0.25
    image_shape = (4, 200, 344)
    psf_patch = (4, 72, 130)
    pos = (170, 59)
    rs = np.random.RandomState(seed=1)
    dirty = rs.standard_normal(image_shape).astype(np.float32)
    psf = rs.standard_normal(psf_patch).astype(np.float32)
    expected = dirty.copy()
    peak_pixel = dirty[:, pos[0], pos[1]].copy()
    expected[:, 134:200, 0:124] -= \
        loop_gain * peak_pixel[:, np.newaxis, np.newaxis] * psf[:, :66, 6:]
    psf_full = np.ones(image_shape, np.float32)
    psf_full[:, 64:136, 107:237] = psf
    fn = clean._SubtractPsfTemplate(context, np.float32, 4).instantiate(
        queue, loop_gain, image_shape, image_shape)
    fn.ensure_all_bound()
    fn.buffer('dirty').set(queue, dirty)
    fn.buffer('psf').set(queue, psf_full)
    fn.buffer('peak_pixel').set(queue, peak_pixel)
    _zero(queue, fn.buffer('model'))
    fn(pos, psf_patch)
    np.testing.assert_allclose(expected, fn.buffer('dirty').get(queue), atol=1e-4)
    model = fn.buffer('model').get(queue)
    np.testing.assert_allclose(loop_gain * peak_pixel, model[:, pos[0], pos[1]])
    assert np.count_nonzero(model) == 4


def _noise(gpu, std, shape=(4, 400, 544), border_pixels=45):
    context, queue = gpu
    rs = np.random.RandomState(seed=1)
    border = border_pixels / shape[1]
    dirty = rs.standard_normal(shape).astype(np.float32)
    dirty[:, border_pixels:-border_pixels, border_pixels:-border_pixels] *= std
    dirty.flat[rs.choice(dirty.size, 1000, replace=False)] += 1e6
    fn = clean.NoiseEstTemplate(context, np.float32, shape[0]).instantiate(queue, shape, border)
    fn.ensure_all_bound()
    fn.buffer('dirty').set(queue, dirty)
    return fn(), dirty, border


def test_noise(gpu, oracle):
    """reference test_clean.py test_noise (rtol 1e-2) and exact agreement with the host
    median (noise_est_host, clean.py:938-943)."""
    estimated, dirty, border = _noise(gpu, 3.2)
    np.testing.assert_allclose(estimated, 3.2, rtol=1e-2)
    assert estimated == oracle.noise_est(dirty, border)
    # odd number of samples: a single middle element
    estimated, dirty, border = _noise(gpu, 1.7, shape=(1, 101, 103), border_pixels=10)
    assert estimated == oracle.noise_est(dirty, border)


def test_noise_zero(gpu):
    estimated, _, _ = _noise(gpu, 0.0)
    assert estimated == 0.0


def test_noise_double(gpu, oracle):
    context, queue = gpu
    rs = np.random.RandomState(5)
    shape = (2, 200, 220)
    dirty = rs.standard_normal(shape) * 2.5
    fn = clean.NoiseEstTemplate(context, np.float64, 2).instantiate(queue, shape, 0.1)
    fn.ensure_all_bound()
    fn.buffer('dirty').set(queue, dirty)
    np.testing.assert_allclose(fn(), oracle.noise_est(dirty, 0.1), rtol=2e-4)


def _make_clean(gpu, fx, lookahead=1):
    context, queue = gpu
    cp = fx['clean_parameters']
    pols = fx['dirty'].shape[0]
    template = clean.CleanTemplate(context, cp, np.float32, pols, tuning={'lookahead': lookahead})
    fn = template.instantiate(queue, fx['image_parameters'])
    fn.ensure_all_bound()
    fn.buffer('dirty').set(queue, fx['dirty'])
    fn.buffer('psf').set(queue, fx['psf'])
    fn.buffer('model').zero(queue)
    fn.reset()
    return fn, queue


def _check_against_golden(fn, queue, golden, positions, values, pixels, model_expected=True):
    np.testing.assert_array_equal(np.array(values, np.float32), golden['values'])
    np.testing.assert_array_equal(np.array(pixels, np.float32), golden['pixels'])
    np.testing.assert_array_equal(fn.buffer('dirty').get(queue), golden['residual'])
    np.testing.assert_array_equal(fn.buffer('model').get(queue), golden['model'])
    np.testing.assert_array_equal(fn.buffer('tile_max').get(queue), golden['tile_max'])
    np.testing.assert_array_equal(fn.buffer('tile_pos').get(queue), golden['tile_pos'])
    # component positions: the golden model image has a non-zero pixel exactly where
    # components were subtracted (the positions CleanHost *returns* are aliased, see
    # oracle.c kor_clean_cycle)
    expected_model = np.zeros_like(golden['model'])
    for pos, pixel in zip(positions, pixels):
        expected_model[:, pos[0], pos[1]] += pixel
    np.testing.assert_array_equal(expected_model, golden['model'])


@pytest.mark.parametrize('name', ['clean_i', 'clean_sumsq'])
def test_clean_single_cycles(gpu, name):
    """Clean.__call__ one cycle at a time (reference API): bit-exact with CleanHost."""
    fx = cases.clean_case(name)
    golden = load_golden(name)
    fn, queue = _make_clean(gpu, fx)
    np.testing.assert_array_equal(fn.buffer('tile_max').get(queue), golden['tile_max0'])
    np.testing.assert_array_equal(fn.buffer('tile_pos').get(queue), golden['tile_pos0'])
    values, positions, pixels = [], [], []
    for _ in range(fx['cycles']):
        value, pos, pixel = fn(fx['psf_patch'], fx['threshold'])
        if value is None:
            assert pos is None and pixel is None
            break
        values.append(value)
        positions.append(pos)
        pixels.append(pixel)
    assert len(values) == len(golden['values']) < fx['cycles']
    _check_against_golden(fn, queue, golden, positions, values, pixels)
    # a further call below the threshold changes nothing
    assert fn(fx['psf_patch'], fx['threshold']) == (None, None, None)
    np.testing.assert_array_equal(fn.buffer('dirty').get(queue), golden['residual'])


@pytest.mark.parametrize('name', ['clean_i', 'clean_sumsq'])
@pytest.mark.parametrize('batch', [7, 1000])
def test_clean_device_resident(gpu, oracle, name, batch):
    """Clean.run_cycles: many cycles per call, same component sequence."""
    fx = cases.clean_case(name)
    golden = load_golden(name)
    fn, queue = _make_clean(gpu, fx)
    values, positions, pixels = [], [], []
    stopped = False
    while not stopped and len(values) < fx['cycles']:
        components, stopped = fn.run_cycles(fx['psf_patch'], fx['threshold'], batch)
        values.extend(components['value'])
        positions.extend(tuple(p) for p in components['pos'])
        pixels.extend(components['pixel'])
    assert stopped
    _check_against_golden(fn, queue, golden, positions, values, pixels)
    # and the oracle's true positions, cycle by cycle
    dirty = fx['dirty'].copy()
    cp = fx['clean_parameters']
    host = oracle.CleanHost(fx['image_parameters'].pixels, cp.border, cp.mode, cp.loop_gain,
                            dirty, fx['psf'], np.zeros_like(dirty))
    host.reset()
    for pos in positions:
        assert host(fx['psf_patch'], fx['threshold'])[1] == tuple(int(x) for x in pos)


def test_clean_lookahead(gpu):
    """Look-ahead batching behind the one-cycle-per-call API."""
    fx = cases.clean_case('clean_i')
    golden = load_golden('clean_i')
    fn, queue = _make_clean(gpu, fx, lookahead=16)
    values, positions, pixels = [], [], []
    while True:
        value, pos, pixel = fn(fx['psf_patch'], fx['threshold'])
        if value is None:
            break
        values.append(value)
        positions.append(pos)
        pixels.append(pixel)
    _check_against_golden(fn, queue, golden, positions, values, pixels)


def test_clean_ties_and_empty_tiles(gpu, oracle):
    """Plateaus of equal values, all-zero tiles (the (x0, y0) quirk of clean.py:950),
    negative peaks and a patch clipped at two image edges."""
    pixels = 96
    dirty = np.zeros((1, pixels, pixels), np.float32)
    dirty[0, 40:44, 50:70] = 2.0           # plateau spanning two tiles
    dirty[0, 70, 5:90] = -2.0              # equal negative values in three tiles
    dirty[0, 6, 88] = 3.0                  # near the corner, inside the border limit
    psf = np.zeros((1, pixels, pixels), np.float32)
    psf[0, 40:57, 36:61] = 0.25
    psf[0, 48, 48] = 1.0
    fixed = prm.FixedImageParameters([1], np.float32)
    fx = dict(dirty=dirty, psf=psf,
              image_parameters=type('P', (), {'fixed': fixed, 'pixels': pixels})(),
              clean_parameters=prm.CleanParameters(100, 0.5, 0.85, 5.0, clean.CLEAN_I,
                                                   0.01, 0.5, 0.05))
    fn, queue = _make_clean(gpu, fx)
    host_dirty = dirty.copy()
    host = oracle.CleanHost(pixels, 0.05, clean.CLEAN_I, 0.5, host_dirty, psf,
                            np.zeros_like(dirty))
    host.reset()
    np.testing.assert_array_equal(fn.buffer('tile_max').get(queue), host.tile_max)
    np.testing.assert_array_equal(fn.buffer('tile_pos').get(queue), host.tile_pos)
    patch = (1, 17, 25)
    components, stopped = fn.run_cycles(patch, 0.0, 60)
    assert len(components) == 60 and not stopped
    for record in components:
        value, pos, pixel = host(patch, 0.0)
        assert tuple(record['pos']) == pos
        assert record['value'] == value
        np.testing.assert_array_equal(record['pixel'], pixel)
    np.testing.assert_array_equal(fn.buffer('dirty').get(queue), host_dirty)
    np.testing.assert_array_equal(fn.buffer('tile_max').get(queue), host.tile_max)
    np.testing.assert_array_equal(fn.buffer('tile_pos').get(queue), host.tile_pos)


def test_clean_large_patch(gpu, oracle):
    """A patch half the size of a 1024^2 image (psf_limit 0.5), 4 polarizations:
    idempotent stop, oracle agreement on a handful of cycles."""
    context, queue = gpu
    pixels = 1024
    rs = np.random.RandomState(3)
    g = scipy.signal.windows.gaussian(pixels, pixels / 8)
    psf1 = np.outer(g, g) + 0.01 * rs.standard_normal((pixels, pixels))
    psf1[pixels // 2, pixels // 2] = 1.0
    psf = np.repeat(psf1[np.newaxis].astype(np.float32), 4, axis=0)
    dirty = rs.standard_normal((4, pixels, pixels)).astype(np.float32)
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    fx = dict(dirty=dirty, psf=psf,
              image_parameters=type('P', (), {'fixed': fixed, 'pixels': pixels})(),
              clean_parameters=prm.CleanParameters(100, 0.1, 0.85, 5.0, clean.CLEAN_SUMSQ,
                                                   0.01, 0.5, 0.02))
    fn, _ = _make_clean(gpu, fx)
    patch = (4, 511, 511)
    components, stopped = fn.run_cycles(patch, 0.0, 5)
    host_dirty = dirty.copy()
    host = oracle.CleanHost(pixels, 0.02, clean.CLEAN_SUMSQ, 0.1, host_dirty, psf,
                            np.zeros_like(dirty))
    host.reset()
    for record in components:
        value, pos, pixel = host(patch, 0.0)
        assert tuple(record['pos']) == pos and record['value'] == value
    np.testing.assert_array_equal(fn.buffer('dirty').get(queue), host_dirty)


def test_noise_est_guessed_leading_digit(gpu, oracle):
    """Repeated estimates start from the leading radix digit of the previous median (one
    windowed pass instead of two): still the exact median, whether the guess is right (same
    or slightly scaled image), off by one bucket, or wrong (image scaled by 50: full select)."""
    context, queue = gpu
    rs = np.random.RandomState(8)
    pols, n = 2, 1024
    base = (rs.standard_normal((pols, n, n)) * 3e-3).astype(np.float32)
    base[0, 100:200, 100:300] += 0.5
    fn = clean.NoiseEstTemplate(context, np.float32, pols).instantiate(queue, (pols, n, n), 0.03)
    fn.ensure_all_bound()
    for scale in (1.0, 1.0, 1.02, 1.25, 0.7, 50.0, 50.0, 1e-3):
        image = (base * np.float32(scale)).astype(np.float32)
        fn.buffer('dirty').set(queue, image)
        assert fn() == oracle.noise_est(image, 0.03), scale
    # odd number of pixels: a single middle element
    fn = clean.NoiseEstTemplate(context, np.float32, 1).instantiate(queue, (1, 301, 301), 0.01)
    fn.ensure_all_bound()
    image = rs.standard_normal((1, 301, 301)).astype(np.float32)
    for _ in range(2):
        fn.buffer('dirty').set(queue, image)
        assert fn() == oracle.noise_est(image, 0.01)
