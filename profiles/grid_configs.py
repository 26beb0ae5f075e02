#!/usr/bin/env python3
"""Gridder / degridder throughput on the largest W slice of the other BASELINE configs
(C1: 2048^2, P=1, K=60, W=1639 planes; C4: 16384^2, P=4, K=32, W=128), which run the generic
(table in L1/L2) gridder path, next to config 2 (K=7, table in shared memory).
One JSON object per row on stdout."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                                   # noqa: E402
from katsdpimager_b200 import accel, grid, parameters as prm, preprocess, simulate   # noqa: E402
from katsdpimager_b200.imaging import _uv_view                  # noqa: E402

CONFIGS = {
    'C1': dict(pixels=2048, pols=[1], K=60, w_slices=3, w_planes=1639, dumps=600),
    'C2': dict(pixels=8192, pols=[1, 2, 3, 4], K=7, w_slices=16, w_planes=16, dumps=600),
    'C4': dict(pixels=16384, pols=[1, 2, 3, 4], K=32, w_slices=4, w_planes=128, dumps=300),
}


def main():
    names = sys.argv[1:] or list(CONFIGS)
    context = accel.Context(0)
    queue = context.create_command_queue()
    for name in names:
        cfg = CONFIGS[name]
        P = len(cfg['pols'])
        wavelength = 299792458.0 / 1284e6
        array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
        fixed = prm.FixedImageParameters(cfg['pols'], np.float32)
        ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=cfg['pixels'], array=array,
                                 image_oversample=5.0)
        fixed_grid = prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, cfg['K'],
                                             degrid=True)
        gp = prm.GridParameters(fixed_grid, cfg['w_slices'], cfg['w_planes'])
        uvw = simulate.uvw_tracks(cfg['dumps'], dump_time=4.0).reshape(-1, 3)
        rs = np.random.RandomState(5)
        vis = (rs.standard_normal((len(uvw), P)) + 1j * rs.standard_normal((len(uvw), P))
               ).astype(np.complex64)
        weights = rs.uniform(0.5, 1.5, (len(uvw), P)).astype(np.float32)
        records, w_slice = preprocess.quantise(uvw.astype(np.float32), weights, vis, ip, gp)
        slices = preprocess.bucket_by_slice(records, w_slice, cfg['w_slices'])
        s = max(slices, key=len)
        n = len(s)
        flops = cfg['K'] ** 2 * (8 * P + 6)
        for kind, cls in (('grid', grid.GridderTemplate), ('degrid', grid.DegridderTemplate)):
            op = cls(context, ip.fixed, fixed_grid).instantiate(queue, array, ip, gp, n)
            op.ensure_all_bound()
            op.buffer('uv').set_region(queue, np.ascontiguousarray(_uv_view(s)), np.s_[:n], np.s_[:n])
            op.buffer('w_plane').set_region(queue, np.ascontiguousarray(s.w_plane), np.s_[:n], np.s_[:n])
            op.buffer('vis').set_region(queue, np.ascontiguousarray(s.vis), np.s_[:n], np.s_[:n])
            if kind == 'grid':
                op.buffer('weights_grid').set(queue, np.ones(op.buffer('weights_grid').shape, np.float32))
            else:
                op.buffer('weights').set_region(queue, np.ascontiguousarray(s.weights),
                                                np.s_[:n], np.s_[:n])
            op.buffer('grid').zero(queue)
            op.num_vis = n
            op()
            queue.finish()
            a = queue.enqueue_marker()
            reps = 5
            for _ in range(reps):
                op()
            b = queue.enqueue_marker()
            b.wait()
            seconds = b.time_since(a) / reps
            print(json.dumps({'config': name, 'row': kind, 'K': cfg['K'], 'P': P,
                              'w_planes': cfg['w_planes'], 'vis': n, 'ms': seconds * 1e3,
                              'gvis_per_s': n / seconds / 1e9,
                              'tflops': n * flops / seconds / 1e12,
                              'flops_per_vis': flops}), flush=True)
            del op


if __name__ == '__main__':
    main()
