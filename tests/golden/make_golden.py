#!/usr/bin/env python3
"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference; see oracle/ref_import.py for how
the reference's host classes are imported):

    python tests/golden/make_golden.py

Every file holds seeded inputs (or the recipe to regenerate them) and the outputs of
the reference's own ``--host`` classes (GridderHost, DegridderHost, GridToImageHost,
ImageToGridHost, CleanHost, psf_patch_host, noise_est_host, WeightsHost,
_predict_host, ImagingHost).  The oracle (oracle/) and the CUDA path are both tested
against these files, so parity is pinned to the reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import                         # noqa: E402
from katsdpimager_b200 import parameters as prm       # noqa: E402
from katsdpimager_b200 import preprocess, simulate    # noqa: E402
from tests import cases                               # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('{:28s} {:8.1f} KiB'.format(name + '.npz', os.path.getsize(path) / 1024))


def lut_cases(ref):
    for name, (ip, gp) in cases.lut_cases().items():
        kernel = ref.grid.ConvolutionKernel(ip, gp)
        save('lut_' + name, data=kernel.data, taper=kernel.taper(ip.pixels), beta=kernel.beta)


def grid_cases(ref):
    # 1. the reference's own unit-test fixture (test/test_grid.py:25-135)
    fx = cases.reference_grid_fixture()
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    gridder = ref.grid.GridderHost(ip, gp)
    gridder.clear()
    cases.middle(gridder.weights_grid, fx['weights_grid'].shape)[:] = fx['weights_grid']
    gridder.num_vis = len(fx['uv'])
    gridder.set_coordinates(fx['uv'], fx['sub_uv'], fx['w_plane'])
    gridder.set_vis(fx['vis'])
    gridder()
    degridder = ref.grid.DegridderHost(ip, gp)
    degridder.values[:] = fx['degrid_grid']
    degridder.num_vis = len(fx['uv'])
    degridder.set_coordinates(fx['uv'], fx['sub_uv'], fx['w_plane'])
    degridder.set_weights(fx['degrid_weights'])
    residual = fx['degrid_vis'].copy()
    degridder.set_vis(residual)
    degridder()
    # keep every third row of the (mostly smooth) grid to bound the fixture size
    save('grid_reference_fixture', grid_rows=gridder.values[:, ::3, :].astype(np.complex64),
         grid_sum=gridder.values.sum(axis=(1, 2)), residual=residual)

    # 2. MeerKAT-like small case, float32 grid, K=7
    fx = cases.small_grid_case()
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    gridder = ref.grid.GridderHost(ip, gp)
    gridder.clear()
    gridder.weights_grid[:] = fx['weights_grid']
    gridder.num_vis = len(fx['uv'])
    gridder.set_coordinates(fx['uv'], fx['sub_uv'], fx['w_plane'])
    gridder.set_vis(fx['vis'])
    gridder()
    degridder = ref.grid.DegridderHost(ip, gp)
    degridder.values[:] = gridder.values
    degridder.num_vis = len(fx['uv'])
    degridder.set_coordinates(fx['uv'], fx['sub_uv'], fx['w_plane'])
    degridder.set_weights(fx['weights'])
    residual = fx['vis'].copy()
    degridder.set_vis(residual)
    degridder()
    save('grid_small', grid=gridder.values, residual=residual)


def image_cases(ref):
    fx = cases.image_case()
    grid, kernel1d = fx['grid'], fx['kernel1d']
    layer = np.empty_like(grid)
    image = np.zeros(grid.shape, np.float32)
    g2i = ref.image.GridToImageHost(grid, layer, image, kernel1d, fx['lm_scale'], fx['lm_bias'])
    g2i.set_w(fx['w'])
    g2i.clear()
    g2i()
    back = np.empty_like(grid)
    i2g = ref.image.ImageToGridHost(back, layer, fx['model'], kernel1d, fx['lm_scale'],
                                    fx['lm_bias'])
    i2g.set_w(fx['w'])
    i2g()
    save('image_small', image=image, grid_from_model=back)


def clean_cases(ref):
    for name in ('clean_i', 'clean_sumsq'):
        fx = cases.clean_case(name)
        dirty = fx['dirty'].copy()
        model = np.zeros_like(dirty)
        ip = fx['image_parameters']
        cp = fx['clean_parameters']
        noise = ref.clean.noise_est_host(dirty, cp.border)
        patch = ref.clean.psf_patch_host(fx['psf'], cp.psf_cutoff, cp.psf_limit)
        cleaner = ref.clean.CleanHost(ip, cp, dirty, fx['psf'], model)
        cleaner.reset()
        tile_max0 = cleaner._tile_max.copy()
        tile_pos0 = cleaner._tile_pos.copy()
        values, positions, pixels = [], [], []
        for _ in range(fx['cycles']):
            value, pos, pixel = cleaner(fx['psf_patch'], fx['threshold'])
            if value is None:
                break
            values.append(value)
            positions.append(pos)
            pixels.append(pixel)
        save(name, noise=np.float32(noise), patch=np.array(patch),
             tile_max0=tile_max0, tile_pos0=tile_pos0,
             values=np.array(values, np.float32), positions=np.array(positions, np.int32),
             pixels=np.array(pixels, np.float32),
             tile_max=cleaner._tile_max, tile_pos=cleaner._tile_pos,
             residual=dirty, model=model)


def weight_cases(ref):
    for name, weight_type in (('uniform', ref.weight.WeightType.UNIFORM),
                              ('robust', ref.weight.WeightType.ROBUST),
                              ('natural', ref.weight.WeightType.NATURAL)):
        fx = cases.weights_case()
        wgrid = np.zeros(fx['shape'], np.float32)
        weights = ref.weight.WeightsHost(weight_type, wgrid)
        weights.robustness = fx['robustness']
        weights.clear()
        weights.grid(fx['uv'].copy(), fx['weights'])      # the reference mutates uv
        rms, normalized_rms = weights.finalize()
        save('weights_' + name, grid=wgrid,
             rms=np.float64(np.nan if rms is None else rms),
             normalized_rms=np.float64(normalized_rms))


def predict_cases(ref):
    fx = cases.predict_case()
    vis = fx['vis'].copy()
    ref.predict._predict_host(
        vis, fx['uv'], fx['sub_uv'], fx['w_plane'], fx['weights'], fx['lmn'], fx['flux'],
        np.float32(fx['oversample']), np.float32(fx['uv_scale']), np.float32(fx['w_scale']),
        np.float32(fx['w_bias']), np.zeros(vis.shape[1], np.complex64))
    ip, gp = fx['image_parameters'], fx['grid_parameters']
    scale_bias = ref.predict._uvw_scale_bias(ip, gp)
    lmn, flux = ref.predict._extract_sky_image(ip, gp, fx['components'])
    save('predict_small', residual=vis, scale_bias=np.array(scale_bias),
         image_lmn=lmn, image_flux=flux)


def imaging_case(ref):
    """End-to-end ImagingHost run replaying frontend.process_channel's call sequence
    (reference frontend.py:494-585) on a small simulated MeerKAT channel."""
    fx = cases.imaging_case()
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(ref.weight.WeightType.UNIFORM)
    out = cases.run_imaging(lambda: ref.imaging.ImagingHost(ip, wp, gp, cp), fx,
                            host_style=True)
    save('imaging_small', **out)


def main():
    ref = ref_import.load()
    lut_cases(ref)
    grid_cases(ref)
    image_cases(ref)
    clean_cases(ref)
    weight_cases(ref)
    predict_cases(ref)
    imaging_case(ref)


if __name__ == '__main__':
    main()
