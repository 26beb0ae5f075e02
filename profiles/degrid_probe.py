#!/usr/bin/env python3
"""Degridder alone on one W slice of the bench channel (config 2 geometry): device time per
call and TFLOP/s; run under ncu to capture degrid_kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                                   # noqa: E402
from katsdpimager_b200 import accel, grid, parameters as prm    # noqa: E402
from katsdpimager_b200.imaging import _uv_view                  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    context = accel.Context(0)
    queue = context.create_command_queue()
    array, ip, gp, slices = bench.make_channel(0, 3600)
    s = [x for x in slices if len(x)][which]
    n = len(s)
    fixed_grid = prm.FixedGridParameters(7.0, 8, 4, array.longest_baseline, 7, degrid=True)
    gp_d = prm.GridParameters(fixed_grid, gp.w_slices, gp.w_planes)
    op = grid.DegridderTemplate(context, ip.fixed, fixed_grid).instantiate(queue, array, ip, gp_d, n)
    op.ensure_all_bound()
    op.buffer('uv').set_region(queue, np.ascontiguousarray(_uv_view(s)), np.s_[:n], np.s_[:n])
    op.buffer('w_plane').set_region(queue, np.ascontiguousarray(s.w_plane), np.s_[:n], np.s_[:n])
    op.buffer('vis').set_region(queue, np.ascontiguousarray(s.vis), np.s_[:n], np.s_[:n])
    op.buffer('weights').set_region(queue, np.ascontiguousarray(s.weights), np.s_[:n], np.s_[:n])
    rs = np.random.RandomState(1)
    g = op.buffer('grid')
    host = g.empty_like()
    host[:] = (rs.standard_normal(g.shape) + 1j * rs.standard_normal(g.shape))
    g.set(queue, host)
    op.num_vis = n
    out = {'vis': n}
    for route in (os.environ.get('DEGRID_ROUTES', 'thread0,thread,vec,hoist').split(',')):
        os.environ['KIB_DEGRID_ROUTE'] = route
        op()
        queue.finish()
        a = queue.enqueue_marker()
        for _ in range(reps):
            op()
        b = queue.enqueue_marker()
        b.wait()
        seconds = b.time_since(a) / reps
        out[route] = {'ms': seconds * 1e3, 'gvis_per_s': n / seconds / 1e9,
                      'tflops': n * bench.flops_per_vis(7, 4) / seconds / 1e12}
    print(json.dumps(out))

if __name__ == '__main__':
    main()
