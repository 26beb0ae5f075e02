"""Seeded test inputs shared by the golden-vector generator (tests/golden/make_golden.py),
the oracle tests and the GPU parity tests.  Nothing here touches the GPU or the
reference tree; everything is deterministic in numpy's legacy RandomState.
"""
import os
import types
import warnings

import numpy as np

from katsdpimager_b200 import parameters as prm
from katsdpimager_b200 import preprocess, simulate

CLEAN_I, CLEAN_SUMSQ = 0, 1

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    """Golden vectors produced from the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


class RandomState(np.random.RandomState):
    """RandomState with the complex distributions of the reference's test utilities
    (reference katsdpimager/test/utils.py:6-27)."""

    def complex_normal(self, loc=0.0j, scale=1.0, size=None):
        return self.normal(np.real(loc), scale, size) + 1j * self.normal(np.imag(loc), scale, size)

    def complex_uniform(self, low=0.0, high=1.0, size=None):
        if not np.iscomplexobj(low):
            low = np.asarray(low) * (1 + 1j)
        if not np.iscomplexobj(high):
            high = np.asarray(high) * (1 + 1j)
        return self.uniform(np.real(low), np.real(high), size) \
            + 1j * self.uniform(np.imag(low), np.imag(high), size)


def middle(array, shape):
    """View of the central part of `array` with size `shape` (test_grid.py:13-21)."""
    index = []
    for a, s in zip(array.shape, shape):
        assert a >= s and (a - s) % 2 == 0
        pad = (a - s) // 2
        index.append(np.s_[pad:a - pad])
    return array[tuple(index)]


def _random_walk(rs, n_vis, cover, oversample, w_planes, jump=73):
    """Track that moves by at most one (sub-)cell per sample with occasional jumps
    (the generator of reference test_grid.py:67-86, bit for bit)."""
    uv = np.empty((n_vis, 2), np.int16)
    sub_uv = np.empty((n_vis, 2), np.int16)
    w_plane = np.empty((n_vis,), np.int16)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', DeprecationWarning)
        for i in range(n_vis):
            if i % jump == 0:
                uv[i, :] = rs.randint(0, cover, (2,))
                sub_uv[i, :] = rs.randint(0, oversample, (2,))
                w_plane[i] = rs.randint(0, w_planes)
            else:
                for j in range(2):
                    uv[i, j] = (uv[i - 1, j] + rs.random_integers(-1, 1)) % cover
                    sub_uv[i, j] = (sub_uv[i - 1, j] + rs.random_integers(-1, 1)) % oversample
                w_plane[i] = (w_plane[i - 1] + rs.random_integers(-1, 1)) % w_planes
    uv -= cover // 2
    return uv, sub_uv, w_plane


# ------------------------------------------------------------------------------ LUT
def lut_cases():
    fixed64 = prm.FixedImageParameters([1, 2, 3, 4], np.float64)
    fixed32 = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    out = {}
    # reference test_grid.py:25-60
    out['test_grid'] = (
        prm.ImageParameters(fixed64, wavelength=0.01, pixels=256, pixel_size=0.0001),
        prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, 5.0, 28), 1, 32))
    # MeerKAT L-band, 8192 pixels, 7 x 7 support (BASELINE config 2 geometry), 4 of 16 planes
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    out['meerkat_k7'] = (
        prm.ImageParameters(fixed32, wavelength=0.2155, pixels=8192, array=array),
        prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, 7), 16, 4))
    return out


# ----------------------------------------------------------------------------- grid
def reference_grid_fixture():
    """Inputs of the reference's gridding unit tests (test_grid.py:25-135)."""
    pixels, grid_cover, w_planes, oversample, n_vis, kernel_width = 256, 180, 32, 8, 1000, 28
    ip, gp = lut_cases()['test_grid']
    array = types.SimpleNamespace(longest_baseline=ip.cell_size * (grid_cover // 2))
    rs = np.random.RandomState(seed=1)
    uv, sub_uv, w_plane = _random_walk(rs, n_vis, grid_cover, oversample, w_planes)
    weights_grid = rs.uniform(size=(4, grid_cover, grid_cover)).astype(np.float32)
    rs = RandomState(seed=2)
    vis = rs.complex_uniform(-1, 1, size=(n_vis, 4)).astype(np.complex64)
    rs = RandomState(seed=2)
    degrid_grid = rs.complex_uniform(-1, 1, size=(4, pixels, pixels)).astype(np.complex128)
    degrid_vis = rs.complex_uniform(-1, 1, size=(n_vis, 4)).astype(np.complex64)
    degrid_weights = rs.uniform(0.5, 1.5, size=(n_vis, 4)).astype(np.float32)
    return dict(image_parameters=ip, grid_parameters=gp, array_parameters=array,
                uv=uv, sub_uv=sub_uv, w_plane=w_plane, weights_grid=weights_grid, vis=vis,
                degrid_grid=degrid_grid, degrid_vis=degrid_vis, degrid_weights=degrid_weights)


def small_grid_case(pixels=128, pols=2, kernel_width=7, w_planes=4, n_vis=3000, seed=5,
                    dtype=np.float32):
    """Small single-precision case with the MeerKAT kernel geometry (K = 7, S = 8)."""
    oversample = 8
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], dtype)
    ip = prm.ImageParameters(fixed, wavelength=0.2, pixels=pixels, pixel_size=0.002)
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, oversample, 4, 40.0, kernel_width), 2,
                            w_planes)
    cover = pixels - 2 * (kernel_width // 2 + 2)
    cover -= cover % 2
    array = types.SimpleNamespace(longest_baseline=ip.cell_size * (cover // 2))
    rs = RandomState(seed)
    uv, sub_uv, w_plane = _random_walk(rs, n_vis, cover, oversample, w_planes, jump=211)
    vis = rs.complex_normal(size=(n_vis, pols)).astype(np.complex64)
    weights = rs.uniform(0.5, 1.5, size=(n_vis, pols)).astype(np.float32)
    weights_grid = rs.uniform(0.5, 1.5, size=(pols, pixels, pixels)).astype(np.float32)
    return dict(image_parameters=ip, grid_parameters=gp, array_parameters=array,
                uv=uv, sub_uv=sub_uv, w_plane=w_plane, vis=vis, weights=weights,
                weights_grid=weights_grid)


# ---------------------------------------------------------------------------- image
def image_case(pixels=64, grid_size=40, pols=2, seed=7):
    rs = RandomState(seed)
    grid = np.zeros((pols, pixels, pixels), np.complex64)
    middle(grid, (pols, grid_size, grid_size))[:] = \
        rs.complex_normal(size=(pols, grid_size, grid_size)).astype(np.complex64)
    kernel1d = rs.uniform(1.0, 2.0, pixels).astype(np.float32)
    model = rs.uniform(-1.0, 1.0, (pols, pixels, pixels)).astype(np.float32)
    lm_scale = 0.3 / pixels
    return dict(grid=grid, grid_size=grid_size, kernel1d=kernel1d, model=model,
                lm_scale=lm_scale, lm_bias=-0.5 * pixels * lm_scale, w=np.float64(123.4))


# ---------------------------------------------------------------------------- clean
def _gaussian(n, sigma):
    x = np.arange(n) - n // 2
    return np.exp(-0.5 * (x / sigma) ** 2)


def clean_case(name, pixels=160):
    mode = CLEAN_I if name == 'clean_i' else CLEAN_SUMSQ
    pols = 1 if mode == CLEAN_I else 3
    rs = np.random.RandomState(11 + mode)
    # PSF: narrow core, broad pedestal, weak sidelobe noise; exactly 1 at the centre
    core = np.outer(_gaussian(pixels, 1.5), _gaussian(pixels, 2.5))
    pedestal = 0.08 * np.outer(_gaussian(pixels, 14.0), _gaussian(pixels, 9.0))
    psf1 = core + pedestal + 0.004 * rs.standard_normal((pixels, pixels))
    psf1 /= psf1[pixels // 2, pixels // 2]
    psf = np.repeat(psf1[np.newaxis], pols, axis=0).astype(np.float32)
    psf[:, pixels // 2, pixels // 2] = 1.0
    # Dirty image: point sources (some near the edges and the border) convolved by
    # shifting the PSF, plus noise
    dirty = 0.02 * rs.standard_normal((pols, pixels, pixels))
    sources = [(80, 80, 1.0), (30, 120, 0.8), (12, 15, 0.9), (150, 140, 0.7), (100, 9, 0.6),
               (81, 83, 0.5)]
    for y, x, flux in sources:
        shifted = np.roll(np.roll(psf1, y - pixels // 2, axis=0), x - pixels // 2, axis=1)
        for p in range(pols):
            dirty[p] += flux * (1.0 if p == 0 else 0.4 * (-1) ** p) * shifted
    dirty = dirty.astype(np.float32)
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
    ip = types.SimpleNamespace(fixed=fixed, pixels=pixels)
    cp = prm.CleanParameters(minor=200, loop_gain=0.1, major_gain=0.85, threshold=5.0, mode=mode,
                             psf_cutoff=0.01, psf_limit=0.5, border=0.05)
    threshold = 0.12 if mode == CLEAN_I else 0.12 ** 2
    return dict(image_parameters=ip, clean_parameters=cp, dirty=dirty, psf=psf,
                psf_patch=(pols, 41, 33), threshold=threshold, cycles=150)


# -------------------------------------------------------------------------- weights
def weights_case(seed=13):
    rs = np.random.RandomState(seed)
    shape = (2, 64, 96)
    n = 800
    uv = np.zeros((n, 4), np.int16)
    uv[:, 0] = rs.randint(-40, 40, n)
    uv[:, 1] = rs.randint(-28, 28, n)
    uv[:, 2:] = rs.randint(0, 8, (n, 2))
    # some repeated cells
    uv[100:200, :2] = uv[0:100, :2]
    weights = rs.uniform(0.5, 1.5, (n, 2)).astype(np.float32)
    return dict(shape=shape, uv=np.ascontiguousarray(uv[:, :2]), uv4=uv, weights=weights,
                robustness=0.5)


# -------------------------------------------------------------------------- predict
def predict_case(seed=17, n_vis=300, n_sources=7):
    pols = [1, 2, 4]
    fixed = prm.FixedImageParameters(pols, np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.2, pixels=4096, pixel_size=0.00001)
    gp = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, 5.0, 7), 10, 100)
    rs = RandomState(seed)
    uv = rs.randint(-2048, 2049, size=(n_vis, 2)).astype(np.int16)
    sub_uv = rs.randint(0, 8, size=(n_vis, 2)).astype(np.int16)
    w_plane = rs.randint(0, gp.w_planes, size=n_vis).astype(np.int16)
    weights = rs.uniform(size=(n_vis, len(pols))).astype(np.float32)
    vis = rs.complex_normal(size=(n_vis, len(pols))).astype(np.complex64)
    l = rs.uniform(-0.02, 0.02, n_sources)
    m = rs.uniform(-0.02, 0.02, n_sources)
    lmn = np.stack([l, m, np.sqrt(1 - l * l - m * m) - 1], axis=1).astype(np.float32)
    flux = rs.uniform(-1, 2, (n_sources, len(pols))).astype(np.float32)
    uv_scale = ip.cell_size / gp.fixed.oversample / ip.wavelength
    w_scale = gp.fixed.max_w / ((gp.w_slices - 0.5) * gp.w_planes) / ip.wavelength
    w_bias = (0.5 - 0.5 * gp.w_planes) * w_scale + 1.2
    components = {(0, 4095): np.array([4.0, 0.0, 0.0], np.float32),
                  (1024, 512): np.array([2.5, 1.5, 0.0], np.float32),
                  (2048, 2048): np.array([1.0, 2.0, 3.0], np.float32),
                  (4095, 0): np.array([5.0, 1.0, 2.0], np.float32)}
    return dict(image_parameters=ip, grid_parameters=gp, uv=uv, sub_uv=sub_uv, w_plane=w_plane,
                weights=weights, vis=vis, lmn=lmn, flux=flux, oversample=gp.fixed.oversample,
                uv_scale=uv_scale, w_scale=w_scale, w_bias=w_bias, components=components)


# -------------------------------------------------------------------------- imaging
def imaging_case(pixels=256, num_baselines=120, num_dumps=60, dump_time=240.0, pols=2,
                 degrid=True, seed=19):
    """A small simulated MeerKAT channel: the shortest-index `num_baselines` baselines,
    point sources inside the field, uniform weighting, W stacking with 2 slices."""
    enu = simulate.meerkat_enu()
    uvw = simulate.uvw_tracks(num_dumps, dump_time, start_hour_angle=-2.0, enu=enu,
                              max_baselines=num_baselines)
    longest = float(np.max(np.linalg.norm(uvw[..., :2], axis=-1))) * 1.02
    wavelength = 0.21
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, longest)
    fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=pixels, array=array,
                             image_oversample=5.0)
    max_w = float(np.max(np.abs(uvw[..., 2]))) * 1.01
    fixed_grid = prm.FixedGridParameters(7.0, 8, 4, max_w, 9, degrid=degrid)
    gp = prm.GridParameters(fixed_grid, w_slices=2, w_planes=16)
    cp = prm.CleanParameters(minor=40, loop_gain=0.1, major_gain=0.85, threshold=5.0,
                             mode=CLEAN_I if pols == 1 else CLEAN_SUMSQ,
                             psf_cutoff=0.02, psf_limit=0.5, border=0.02)
    # sources at pixel offsets from the image centre (row, col, I flux)
    rs = np.random.RandomState(seed)
    src = [(20, -31, 2.0), (-40, 12, 1.5), (3, 5, 1.0), (60, 70, 0.8)]
    lmn = np.array([[x * ip.pixel_size, y * ip.pixel_size, 0.0] for y, x, _ in src])
    lmn[:, 2] = np.sqrt(1 - lmn[:, 0] ** 2 - lmn[:, 1] ** 2) - 1
    flux = np.zeros((len(src), pols))
    flux[:, 0] = [f for _, _, f in src]
    if pols > 1:
        flux[:, 1] = 0.3 * flux[:, 0] * np.array([1, -1, 0.5, 0])[:len(src)]
    uvw_flat = uvw.reshape(-1, 3)
    vis = simulate.dft_visibilities(uvw_flat / wavelength, lmn, flux)
    vis += (0.05 * (rs.standard_normal(vis.shape) + 1j * rs.standard_normal(vis.shape))) \
        .astype(np.complex64)
    weights = rs.uniform(0.5, 1.5, (len(uvw_flat), pols)).astype(np.float32)
    records, w_slice = preprocess.quantise(uvw_flat.astype(np.float32), weights, vis, ip, gp)
    records, w_slice = preprocess.compress(records, w_slice)
    slices = preprocess.bucket_by_slice(records, w_slice, gp.w_slices)
    reader = preprocess.VisibilityReaderMem([slices])
    return dict(image_parameters=ip, grid_parameters=gp, clean_parameters=cp,
                array_parameters=array, reader=reader, major=2, vis_block=4096)


def run_imaging(make_imager, fx, host_style=False, clean_batch=None):
    """Replay frontend.process_channel (reference frontend.py:494-585) on `fx`.

    `make_imager` returns an object with the Imaging / ImagingHost method set.  With
    `clean_batch` the device-resident minor-cycle loop is used (Imaging.clean_cycles).
    Returns a dict of numpy outputs.
    """
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    reader, vis_block, degrid = fx['reader'], fx['vis_block'], gp.fixed.degrid
    imager = make_imager()
    if not host_style:
        imager.ensure_all_bound()
    imager.clear_model()
    out = {}

    def get(name):
        return np.array(imager.get_buffer(name))

    def make_dirty(field, full_cycle):
        imager.clear_dirty()
        if full_cycle and not degrid:
            imager.model_to_predict()
        for w_slice in range(reader.num_w_slices(0)):
            if reader.len(0, w_slice) == 0:
                continue
            if full_cycle and degrid:
                imager.model_to_grid(mid_w[w_slice])
            imager.clear_grid()
            for chunk in reader.iter_slice(0, w_slice, vis_block):
                imager.num_vis = len(chunk.uv)
                imager.set_coordinates(chunk)
                imager.set_vis(chunk[field])
                if full_cycle:
                    imager.set_weights(chunk.weights)
                    imager.predict(mid_w[w_slice])
                imager.grid()
            imager.grid_to_image(mid_w[w_slice])

    # weights (frontend.make_weights)
    imager.clear_weights()
    for w_slice in range(reader.num_w_slices(0)):
        for chunk in reader.iter_slice(0, w_slice, vis_block):
            imager.grid_weights(np.array(chunk.uv), chunk.weights)
    rms, normalized_rms = imager.finalize_weights()
    out['weights_rms'] = np.array([rms, normalized_rms], np.float64)

    mid_w = prm.slice_mid_w(ip, gp)
    make_dirty('weights', False)
    dirty = get('dirty')
    psf_peak = dirty[..., dirty.shape[1] // 2, dirty.shape[2] // 2].copy()
    out['psf_peak'] = psf_peak
    scale = np.reciprocal(psf_peak)
    imager.scale_dirty(scale)
    imager.dirty_to_psf()
    psf_patch = imager.psf_patch()
    out['psf_patch'] = np.array(psf_patch)
    out['psf'] = get('psf')[0]

    values, noises = [], []
    for i in range(fx['major']):
        make_dirty('vis', i != 0)
        imager.scale_dirty(scale)
        if i == 0:
            out['dirty0'] = get('dirty')
        noise = imager.noise_est()
        noises.append(noise)
        imager.clean_reset()
        peak_value = imager.clean_cycle(psf_patch)
        values.append(peak_value)
        threshold = 0.15 * peak_value
        if clean_batch:
            remaining = cp.minor - 1
            while remaining > 0:
                vals, stopped = imager.clean_cycles(psf_patch, threshold,
                                                    min(clean_batch, remaining))
                values.extend(vals)
                remaining -= len(vals)
                if stopped:
                    values.append(-1.0)
                    break
        else:
            for j in range(cp.minor - 1):
                value = imager.clean_cycle(psf_patch, threshold)
                if value is None:
                    values.append(-1.0)      # marks the terminating cycle
                    break
                values.append(value)
    out['noise'] = np.array(noises, np.float32)
    out['values'] = np.array(values, np.float32)
    components = imager._model_components
    keys = sorted(components)
    out['component_pos'] = np.array(keys, np.int32).reshape(len(keys), 2)
    out['component_flux'] = np.array([components[k] for k in keys], np.float32)
    out['residual'] = get('dirty')
    out['model'] = get('model')
    return out
