#!/usr/bin/env python3
"""Per-row measurements of SURVEY.md section 8 (one JSON object per row on stdout).

Complements bench.py (which times the config-2 dirty-image step): degridding, CLEAN minor
cycles (config 5 geometry: 4096^2, 1000 cycles), direct prediction (config 3: 1000 sources),
imaging weights and the image-arithmetic kernels, each against its roofline.  Device times
are CUDA events on the launching stream after warm-up; the CPU figures are the oracle port
on a bounded sample (1 thread) where it finishes in seconds.

    python profiles/bench_rows.py > profiles/r01_rows.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                                   # noqa: E402
from katsdpimager_b200 import (accel, clean, grid, image, parameters as prm,    # noqa: E402
                               predict, weight)

HBM = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def timed(queue, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    queue.finish()
    a = queue.enqueue_marker()
    for _ in range(reps):
        fn()
    b = queue.enqueue_marker()
    b.wait()
    return b.time_since(a) / reps


def emit(**row):
    print(json.dumps(row), flush=True)


def main():
    import oracle
    oracle.host.build()
    context = accel.Context(0)
    queue = context.create_command_queue()
    fp32_peak = None

    # ------------------------------------------------------------ grid / degrid (config 2)
    array, ip, gp, slices = bench.make_channel(0, 3600)
    s = slices[0]
    n = len(s)
    fixed_grid = prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, 7, degrid=True)
    gp_d = prm.GridParameters(fixed_grid, gp.w_slices, gp.w_planes)
    flops = bench.flops_per_vis(7, 4)
    for name, template_cls in (('grid', grid.GridderTemplate), ('degrid', grid.DegridderTemplate)):
        template = template_cls(context, ip.fixed, fixed_grid)
        op = template.instantiate(queue, array, ip, gp_d, n)
        op.ensure_all_bound()
        from katsdpimager_b200.imaging import _uv_view
        op.buffer('uv').set_region(queue, np.ascontiguousarray(_uv_view(s)), np.s_[:n], np.s_[:n])
        op.buffer('w_plane').set_region(queue, np.ascontiguousarray(s.w_plane), np.s_[:n], np.s_[:n])
        op.buffer('vis').set_region(queue, np.ascontiguousarray(s.vis), np.s_[:n], np.s_[:n])
        op.num_vis = n
        if name == 'grid':
            op.buffer('weights_grid').set(queue, np.ones(op.buffer('weights_grid').shape, np.float32))
            op.buffer('grid').zero(queue)
        else:
            op.buffer('weights').set_region(queue, np.ascontiguousarray(s.weights), np.s_[:n], np.s_[:n])
            rs = np.random.RandomState(1)
            g = op.buffer('grid')
            host = g.empty_like()
            host[:] = (rs.standard_normal(g.shape) + 1j * rs.standard_normal(g.shape))
            g.set(queue, host)
        seconds = timed(queue, op)
        emit(row=name, config='config 2 slice 0: K=7, P=4, {} vis, G={}'.format(
                 n, op.buffer('grid').shape[-1]),
             vis_per_s=n / seconds, ms=seconds * 1e3, tflops=n * flops / seconds / 1e12,
             bound='fp32', flops_per_vis=flops)

    # ------------------------------------------------------------ FFT stage (config 2 plane)
    gi = image.GridImageTemplate(context, np.float32)
    plan = gi.make_fft_plan((ip.pixels, ip.pixels), (ip.pixels, ip.pixels))
    size = 2 * (int(array.longest_baseline / ip.cell_size) + 7 // 2 + 1)
    lm_scale = float(ip.pixel_size)
    lm_bias = -0.5 * ip.pixels * lm_scale
    g2i = gi.instantiate_grid_to_image(queue, (1, size, size), lm_scale, lm_bias, plan)
    i2g = gi.instantiate_image_to_grid(queue, (1, size, size), lm_scale, lm_bias, plan)
    g2i.ensure_all_bound()
    i2g.bind(layer=g2i.buffer('layer'), kernel1d=g2i.buffer('kernel1d'), image=g2i.buffer('image'),
             grid=g2i.buffer('grid'))
    taper = oracle.taper(gp, ip.pixels, np.float32)
    g2i.buffer('kernel1d').set(queue, taper)
    g2i.buffer('grid').zero(queue)
    g2i.buffer('image').zero(queue)
    g2i.set_w(1234.5)
    i2g.set_w(1234.5)
    N = ip.pixels
    seconds = timed(queue, g2i)
    emit(row='grid_to_image', config='8192^2, 1 polarization (fused pruned transform: column pass + row pass)',
         ms=seconds * 1e3, planes_per_s=1 / seconds,
         algorithmic_gb=(8 * size * size + 16 * N * size + 8 * N * N) / 1e9,
         note='cuFFT route (pad/shift + cuFFT + epilogue): 0.87 ms, 3.95 GB')
    seconds = timed(queue, i2g)
    emit(row='image_to_grid', config='8192^2, 1 polarization (fused pruned transform: row pass + tile transforms + unfold)',
         ms=seconds * 1e3, planes_per_s=1 / seconds)

    # ------------------------------------------------------------ CLEAN (config 5 geometry)
    for pols, mode in ((1, clean.CLEAN_I), (4, clean.CLEAN_SUMSQ)):
        pixels = 4096
        rs = np.random.RandomState(2)
        dirty = rs.standard_normal((pols, pixels, pixels)).astype(np.float32)
        x = np.arange(pixels) - pixels // 2
        g1 = np.exp(-0.5 * (x / 3.0) ** 2) + 0.05 * np.exp(-0.5 * (x / 60.0) ** 2)
        psf1 = np.outer(g1, g1).astype(np.float32)
        psf1 /= psf1[pixels // 2, pixels // 2]
        psf = np.repeat(psf1[np.newaxis], pols, axis=0)
        fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
        cp = prm.CleanParameters(1000, 0.1, 0.85, 5.0, mode, 0.01, 0.5, 0.02)
        ipc = type('P', (), {'fixed': fixed, 'pixels': pixels})()
        op = clean.CleanTemplate(context, cp, np.float32, pols).instantiate(queue, ipc)
        op.ensure_all_bound()
        op.buffer('psf').set(queue, psf)
        op.buffer('model').zero(queue)
        for patch_side in (255, 1023):
            patch = (pols, patch_side, patch_side)
            op.buffer('dirty').set(queue, dirty)
            t0 = time.monotonic()
            op.reset()
            queue.finish()
            reset_s = time.monotonic() - t0
            op.run_cycles(patch, 0.0, 100)          # warm-up
            a = queue.enqueue_marker()
            components, _ = op.run_cycles(patch, 0.0, 1000)
            b = queue.enqueue_marker()
            b.wait()
            seconds = b.time_since(a)
            bytes_per_cycle = 12.0 * patch_side * patch_side * pols
            emit(row='clean_minor_cycles', config='4096^2, {} pol, {}^2 patch, 1000 cycles'.format(
                     pols, patch_side),
                 cycles_per_s=len(components) / seconds, us_per_cycle=seconds / len(components) * 1e6,
                 reset_ms=reset_s * 1e3, subtract_gb_s=bytes_per_cycle / (seconds / len(components)) / 1e9,
                 hbm_peak_gb_s=HBM, bound='latency (small patch) / hbm (large patch)')
        # one-cycle-per-call API (one sync per cycle, reference semantics)
        op.buffer('dirty').set(queue, dirty)
        op.reset()
        t0 = time.monotonic()
        for _ in range(200):
            op((pols, 255, 255), 0.0)
        seconds = time.monotonic() - t0
        emit(row='clean_cycle_per_call', config='4096^2, {} pol, 255^2 patch'.format(pols),
             cycles_per_s=200 / seconds)
        # CPU oracle on the same arrays
        host_dirty = dirty.copy()
        host = oracle.CleanHost(pixels, cp.border, mode, cp.loop_gain, host_dirty, psf,
                                np.zeros_like(dirty))
        t0 = time.monotonic()
        host.reset()
        reset_cpu = time.monotonic() - t0
        t0 = time.monotonic()
        for _ in range(100):
            host((pols, 255, 255), 0.0)
        emit(row='clean_cpu_oracle', config='4096^2, {} pol, 255^2 patch, 1 thread'.format(pols),
             cycles_per_s=100 / (time.monotonic() - t0), reset_ms=reset_cpu * 1e3)
        # noise estimate and PSF patch
        ne = clean.NoiseEstTemplate(context, np.float32, pols).instantiate(
            queue, dirty.shape, cp.border)
        ne.bind(dirty=op.buffer('dirty'))
        ne.ensure_all_bound()
        t0 = time.monotonic()
        ne()
        emit(row='noise_est', config='4096^2, {} pol (exact median, radix select)'.format(pols),
             ms=(time.monotonic() - t0) * 1e3)
        del op, ne

    # ------------------------------------------------------------ predict (config 3)
    pixels = 4096
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    ip3 = prm.ImageParameters(fixed, wavelength=0.2155, pixels=pixels, array=array)
    gp3 = prm.GridParameters(prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, 7), 16, 16)
    nvis, nsrc = 1 << 20, 1000
    op = predict.PredictTemplate(context, np.float32, 4).instantiate(queue, ip3, gp3, nvis, nsrc)
    op.ensure_all_bound()
    rs = np.random.RandomState(3)
    for name in ('uv', 'w_plane', 'vis', 'weights'):
        buf = op.buffer(name)
        host = buf.empty_like()
        if name == 'uv':
            host[:, :2] = rs.randint(-1200, 1200, (nvis, 2))
            host[:, 2:] = rs.randint(0, 8, (nvis, 2))
        elif name == 'w_plane':
            host[:] = rs.randint(0, 16, nvis)
        else:
            host[:] = rs.uniform(size=host.shape)
        buf.set(queue, host)
    lmn, flux = __import__('katsdpimager_b200.simulate', fromlist=['x']).random_sources(
        nsrc, 0.4 * pixels * ip3.pixel_size, 4)
    op.set_sources(lmn, flux)
    op.num_vis = nvis
    op.set_w(100.0)
    seconds = timed(queue, op)
    emit(row='predict', config='2^20 vis x 1000 sources, 4 pol', ms=seconds * 1e3,
         vis_sources_per_s=nvis * nsrc / seconds, vis_per_s=nvis / seconds,
         bound='fp32 + sfu: 5 flop phase + 1 sincos + 4P flop per (vis, source)')

    # ------------------------------------------------------------ weights (config 4 flavour)
    shape = (4, size, size)
    wop = weight.WeightsTemplate(context, weight.WeightType.ROBUST, 4).instantiate(
        queue, shape, n)
    wop.ensure_all_bound()
    from katsdpimager_b200.imaging import _uv_view
    wop.buffer('uv').set_region(queue, np.ascontiguousarray(_uv_view(s)), np.s_[:n], np.s_[:n])
    wop.buffer('weights').set_region(queue, np.ascontiguousarray(s.weights), np.s_[:n], np.s_[:n])
    wop.clear()
    seconds = timed(queue, lambda: wop.grid(n))
    emit(row='grid_weights', config='{} vis, 4 pol, G={}'.format(n, size), ms=seconds * 1e3,
         vis_per_s=n / seconds, gb_s=n * (8 + 8 * 4) / seconds / 1e9, bound='hbm/l2 atomics')
    t0 = time.monotonic()
    wop.finalize()
    emit(row='finalize_weights', config='robust, 4 x {}^2 cells'.format(size),
         ms=(time.monotonic() - t0) * 1e3, gb=8 * 4 * size * size / 1e9 + 4 * size * size / 1e9)


if __name__ == '__main__':
    main()
