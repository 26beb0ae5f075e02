"""End to end: the Imaging facade replaying frontend.process_channel's call sequence
against the reference's ImagingHost (golden, tests/golden/make_golden.py)."""
import numpy as np
import pytest

from katsdpimager_b200 import imaging, parameters as prm, weight
from tests import cases
from tests.cases import load_golden

pytestmark = pytest.mark.gpu


def _rms_rel(actual, expected):
    return np.sqrt(np.mean((actual - expected) ** 2)) / np.abs(expected).max()


@pytest.mark.parametrize('clean_batch', [None, 16])
def test_imaging_against_host(gpu, clean_batch):
    context, queue = gpu
    fx = cases.imaging_case()
    golden = load_golden('imaging_small')
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.UNIFORM)
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)

    def make():
        return template.instantiate(queue, ip, gp, fx['vis_block'], 0, fx['major'])

    out = cases.run_imaging(make, fx, clean_batch=clean_batch)
    np.testing.assert_allclose(out['weights_rms'], golden['weights_rms'], rtol=1e-5)
    np.testing.assert_allclose(out['psf_peak'], golden['psf_peak'], rtol=1e-5)
    np.testing.assert_array_equal(out['psf_patch'], golden['psf_patch'])
    # north_star gate: images within 1e-4 RMS relative to peak.  Observed ~1e-5: single
    # precision FFT round-off (cuFFT vs pocketfft) amplified by the taper division
    # towards the image edges, where the taper falls to ~1e-2.
    assert _rms_rel(out['psf'], golden['psf']) < 1e-4
    assert _rms_rel(out['dirty0'], golden['dirty0']) < 1e-4
    np.testing.assert_allclose(out['noise'], golden['noise'], rtol=1e-4)
    # CLEAN: same number of cycles, same component pixels (bit-exact indices), fluxes 1e-5
    assert len(out['values']) == len(golden['values'])
    np.testing.assert_allclose(out['values'], golden['values'], rtol=2e-4)
    np.testing.assert_array_equal(np.argwhere(out['model'] != 0), np.argwhere(golden['model'] != 0))
    np.testing.assert_allclose(out['model'], golden['model'],
                               rtol=0, atol=1e-5 * np.abs(golden['model']).max())
    assert _rms_rel(out['residual'], golden['residual']) < 1e-4
    # our component dictionary is keyed by the true positions
    model = np.zeros_like(out['model'])
    for pos, flux in zip(out['component_pos'], out['component_flux']):
        model[:, pos[0], pos[1]] += flux
    np.testing.assert_allclose(model, out['model'], rtol=1e-6)


def test_buffer_management(gpu):
    context, queue = gpu
    fx = cases.imaging_case(num_baselines=10, num_dumps=5)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.NATURAL)
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    imager = template.instantiate(queue, ip, gp, 256, 0, 2)
    imager.ensure_all_bound()
    # dirty_to_psf swaps the two buffers (imaging.py:370-373)
    dirty, psf = imager.buffer('dirty'), imager.buffer('psf')
    imager.dirty_to_psf()
    assert imager.buffer('dirty') is psf and imager.buffer('psf') is dirty
    assert imager._clean.buffer('dirty') is psf
    # shared slots really share memory
    assert imager._gridder.buffer('grid') is imager._grid_to_image.buffer('grid')
    assert imager._predict.buffer('grid') is imager._image_to_grid.buffer('grid')
    assert imager._gridder.buffer('weights_grid') is imager._weights.buffer('grid')
    imager.free_buffer('layer')
    assert imager.buffer('layer') is None
    imager.clear_weights()
    assert imager.finalize_weights() == (None, 1.0)
    assert np.all(imager.get_buffer('weights_grid') == 1.0)
    with pytest.raises(ValueError):
        imager.num_vis = 257
    imager.num_vis = 3
    with pytest.raises(ValueError):
        imager.set_vis(np.zeros((4, 2), np.complex64))


def test_record_upload_matches_field_upload(gpu):
    """Uploading whole records and splitting them on the device (kib_unpack_records) gives
    the same buffers as the reference's per-field host staging (imaging.py:269-314)."""
    context, queue = gpu
    fx = cases.imaging_case(num_baselines=30, num_dumps=20)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.NATURAL)
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    imager = template.instantiate(queue, ip, gp, 1024, 0, 2)
    imager.ensure_all_bound()
    chunk = next(fx['reader'].iter_slice(0, 0, 500))
    n = len(chunk)
    assert n > 100
    imager.num_vis = n

    def snapshot():
        return {name: imager.get_buffer(name)[:n].copy()
                for name in ('uv', 'w_plane', 'vis', 'weights')}

    # record path (contiguous record array)
    assert imager._records.matches(chunk)
    imager.set_coordinates(chunk)
    imager.set_vis(chunk.vis)
    imager.set_weights(chunk.weights)
    via_records = snapshot()
    np.testing.assert_array_equal(via_records['uv'][:, :2], chunk.uv)
    np.testing.assert_array_equal(via_records['uv'][:, 2:], chunk.sub_uv)
    np.testing.assert_array_equal(via_records['w_plane'], chunk.w_plane)
    np.testing.assert_array_equal(via_records['vis'], chunk.vis)
    np.testing.assert_array_equal(via_records['weights'], chunk.weights)
    # PSF pass: the weights are gridded as if they were visibilities
    imager.set_vis(chunk.weights)
    np.testing.assert_array_equal(imager.get_buffer('vis')[:n], chunk.weights.astype(np.complex64))
    # field path: detached copies are not recognised as fields of the uploaded block
    for name in ('uv', 'w_plane', 'vis', 'weights'):
        imager.buffer(name).zero(queue)
    strided = chunk[::2]
    assert not imager._records.matches(strided)
    imager.num_vis = len(strided)
    imager.set_coordinates(strided)
    imager.set_vis(np.ascontiguousarray(strided.vis))
    imager.set_weights(np.ascontiguousarray(strided.weights))
    m = len(strided)
    np.testing.assert_array_equal(imager.get_buffer('uv')[:m, :2], strided.uv)
    np.testing.assert_array_equal(imager.get_buffer('vis')[:m], strided.vis)
    np.testing.assert_array_equal(imager.get_buffer('weights')[:m], strided.weights)


def test_imaging_pipeline_matches_single_imager(gpu):
    """Two imagers on two command queues taking alternate channels (ImagingPipeline, as
    bench.py's e2e leg does) produce the dirty image a single imager produces; all work of
    a turn stays asynchronous until wait()."""
    context, queue = gpu
    fx = cases.imaging_case(num_baselines=30, num_dumps=20)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.NATURAL)
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    mid_w = prm.slice_mid_w(ip, gp)
    reader = fx['reader']

    def dirty_image(imager, scale):
        imager.clear_dirty()
        for w_slice in range(reader.num_w_slices(0)):
            if reader.len(0, w_slice) == 0:
                continue
            imager.clear_grid()
            for chunk in reader.iter_slice(0, w_slice, 1024):
                imager.num_vis = len(chunk)
                imager.set_coordinates(chunk)
                imager.set_vis(chunk.vis * np.complex64(scale))
                imager.grid()
            imager.grid_to_image(mid_w[w_slice])

    single = template.instantiate(queue, ip, gp, 1024, 0, 2)
    single.ensure_all_bound()
    single.clear_weights()
    single.finalize_weights()
    expected = {}
    for scale in (1.0, 2.0, 3.0):
        dirty_image(single, scale)
        expected[scale] = single.get_buffer('dirty').copy()
    assert np.abs(expected[1.0]).max() > 0

    pipeline = imaging.ImagingPipeline(template, 2, ip, gp, 1024, 0, 2)
    assert len(pipeline) == 2
    hosts = []
    for imager in pipeline.imagers:
        imager.clear_weights()
        imager.finalize_weights()
        hosts.append(imager.buffer('dirty').empty_like())
    got = {}
    pending = {}
    for scale in (1.0, 2.0, 3.0):
        slot, imager = pipeline.acquire()
        if slot in pending:                      # acquire() waited: the host copy is valid
            got[pending.pop(slot)] = hosts[slot].copy()
        dirty_image(imager, scale)
        imager.buffer('dirty').get_async(imager.command_queue, hosts[slot])
        pipeline.release(slot)
        pending[slot] = scale
    pipeline.finish()
    for slot, scale in pending.items():
        got[scale] = hosts[slot].copy()
    # The order of the gridder's reductions differs from run to run (3e-8 of the largest grid
    # cell); dividing by the taper amplifies that towards the image edges, so two runs of the
    # SAME imager differ by ~1e-4 of the peak at worst and ~1e-5 RMS (profiles/pipe_diag.py).
    for scale in (1.0, 2.0, 3.0):
        assert _rms_rel(got[scale], expected[scale]) < 1e-4
        assert np.abs(got[scale] - expected[scale]).max() < 5e-3 * np.abs(expected[scale]).max()
    assert _rms_rel(got[2.0], 2 * expected[1.0]) < 1e-4      # each turn saw its own data
    assert _rms_rel(got[3.0], 3 * expected[1.0]) < 1e-3
    assert _rms_rel(got[3.0], expected[1.0]) > 0.05
    with pytest.raises(ValueError):
        imaging.ImagingPipeline(template, 0, ip, gp, 1024, 0, 2)


def test_imaging_fused_routes_match_cufft_routes(gpu):
    """A whole process_channel replay (PSF, dirty image, CLEAN, model -> grid -> degrid,
    residual) at 2048^2, where the fused pruned transforms are taken, against the same replay
    with the pad + cuFFT + epilogue routes that the golden files pin at small sizes."""
    context, queue = gpu
    fx = cases.imaging_case(pixels=2048, num_baselines=60, num_dumps=40)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.UNIFORM)
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    outputs = {}
    for fused in (True, False):
        def make():
            imager = template.instantiate(queue, ip, gp, fx['vis_block'], 0, fx['major'])
            assert imager._grid_to_image.fused and imager._image_to_grid.fused
            imager._grid_to_image.fused = fused
            imager._image_to_grid.fused = fused
            return imager
        outputs[fused] = cases.run_imaging(make, fx, clean_batch=16)
    a, b = outputs[True], outputs[False]
    np.testing.assert_array_equal(a['psf_patch'], b['psf_patch'])
    np.testing.assert_allclose(a['psf_peak'], b['psf_peak'], rtol=1e-5)
    for name in ('psf', 'dirty0', 'residual'):
        assert _rms_rel(a[name], b[name]) < 2e-5, name
    np.testing.assert_allclose(a['noise'], b['noise'], rtol=1e-4)
    # the same components: CLEAN is bit exact on identical inputs, and the two routes differ by
    # rounding only, which may reorder near-equal peaks but not the set of pixels found
    assert len(a['values']) == len(b['values'])
    np.testing.assert_array_equal(np.argwhere(a['model'] != 0), np.argwhere(b['model'] != 0))
    np.testing.assert_allclose(a['model'], b['model'], rtol=0, atol=1e-4 * np.abs(b['model']).max())


@pytest.mark.parametrize('degrid', [True, False])
def test_pipeline_resident_matches_host_chunks(gpu, degrid):
    """pipeline.process_channel (reference frontend.py:494-641) with the records resident in
    HBM gives bit-identical images and the same CLEAN history as the same driver feeding
    every chunk from the host on every pass (the reference's data flow)."""
    from katsdpimager_b200 import imaging, pipeline, weight
    context, queue = gpu
    fx = cases.imaging_case(degrid=degrid)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.ROBUST, 0.0)
    slices = [fx['reader']._data[0][w] for w in range(gp.w_slices)]
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    results = []
    for resident in (True, False):
        imager = template.instantiate(queue, ip, gp, fx['vis_block'], 0, 3)
        imager.ensure_all_bound()
        if resident:
            vis = pipeline.ResidentVisibilities(queue, slices, len(ip.fixed.polarizations))
        else:
            vis = pipeline.HostVisibilities(slices)
        out = imager.buffer('dirty').empty_like()
        stats = pipeline.process_channel(imager, vis, ip, gp, cp, wp, 3, fx['vis_block'], out=out)
        queue.finish()
        results.append((stats, np.array(out), imager.get_buffer('model'),
                        dict(imager._model_components)))
    (s0, img0, model0, comp0), (s1, img1, model1, comp1) = results
    assert s0['passes'] == 4 and s0['major'] == 3 and s0['minor'] > 10
    for key in ('passes', 'major', 'minor', 'psf_patch_size', 'compressed_vis'):
        assert s0[key] == s1[key], key
    for key in ('noise', 'weights_noise', 'normalized_noise'):
        np.testing.assert_allclose(s0[key], s1[key], rtol=1e-4)
    # gridding accumulates with atomics in no fixed order, so two runs -- of either kind -- agree
    # to rounding, not bit for bit: same components; images within the north_star bar
    # (1e-4 RMS of the peak; observed 2e-6, worst pixels 1e-4 at the image edge where the taper
    # division amplifies the rounding noise)
    assert sorted(comp0) == sorted(comp1)
    peak = np.abs(img1).max()
    assert np.abs(model0 - model1).max() <= 1e-5 * np.abs(model1).max()
    assert np.sqrt(np.mean((img0 - img1) ** 2)) <= 1e-5 * peak
    assert np.abs(img0 - img1).max() <= 1e-3 * peak
    assert np.count_nonzero(model0) > 0


def test_pipeline_occupancy_matches_dense(gpu):
    """pipeline.process_channel at a size the fused transform covers (2048^2): with the column
    occupancy of every W slice (clears, grid -> image and image -> grid restricted to the
    occupied column groups, factor planes cached across the passes) and without it (whole grids
    cleared and transformed) -- same CLEAN history, images equal to the rounding of the
    gridder's atomics."""
    from katsdpimager_b200 import imaging, pipeline, weight
    context, queue = gpu
    fx = cases.imaging_case(pixels=2048, degrid=True)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.ROBUST, 0.0)
    slices = [fx['reader']._data[0][w] for w in range(gp.w_slices)]
    template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
    results = []
    for use_occupancy in (True, False):
        imager = template.instantiate(queue, ip, gp, fx['vis_block'], 0, 3)
        imager.ensure_all_bound()
        assert imager._grid_to_image.uses_occupancy
        # poison both grids: only what a pass clears may be relied upon
        poison = np.full(imager.buffer('grid').shape, np.nan, np.complex64)
        imager.buffer('grid').set(queue, poison)
        imager.buffer('degrid').set(queue, poison)
        vis = pipeline.ResidentVisibilities(queue, slices, len(ip.fixed.polarizations))
        out = imager.buffer('dirty').empty_like()
        stats = pipeline.process_channel(imager, vis, ip, gp, cp, wp, 3, fx['vis_block'], out=out,
                                         use_occupancy=use_occupancy)
        queue.finish()
        results.append((stats, np.array(out), imager.get_buffer('model'),
                        dict(imager._model_components)))
        del imager
    (s0, img0, model0, comp0), (s1, img1, model1, comp1) = results
    assert np.isfinite(img0).all() and np.isfinite(img1).all()
    assert s0['passes'] == 4 and s0['minor'] > 10
    for key in ('passes', 'major', 'minor', 'psf_patch_size', 'compressed_vis'):
        assert s0[key] == s1[key], key
    np.testing.assert_allclose(s0['noise'], s1['noise'], rtol=1e-4)
    assert sorted(comp0) == sorted(comp1)
    peak = np.abs(img1).max()
    assert np.abs(model0 - model1).max() <= 1e-5 * np.abs(model1).max()
    assert np.sqrt(np.mean((img0 - img1) ** 2)) <= 1e-5 * peak
    assert np.abs(img0 - img1).max() <= 1e-3 * peak
