#!/usr/bin/env python3
"""GPU time of kib_grid for every W slice of the bench channel, launched back to back behind a
long kernel so that host submission never starves the stream (pure device time per call)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                                   # noqa: E402
from katsdpimager_b200 import _lib, accel, grid, parameters as prm    # noqa: E402
from katsdpimager_b200.imaging import _uv_view                  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    only = int(sys.argv[2]) if len(sys.argv) > 2 else None     # index among non-empty slices
    context = accel.Context(0)
    queue = context.create_command_queue()
    array, ip, gp, slices = bench.make_channel(0, 3600)
    template = grid.GridderTemplate(context, ip.fixed, gp.fixed)
    nmax = max(len(s) for s in slices)
    op = template.instantiate(queue, array, ip, gp, nmax)
    op.ensure_all_bound()
    op.buffer('weights_grid').set(queue, np.ones(op.buffer('weights_grid').shape, np.float32))
    op.buffer('grid').zero(queue)
    sink = accel.DeviceArray(context, (1,), np.float32)
    flops = _lib.c_double()
    out = []
    slices = [s for s in slices if len(s)]
    if only is not None:
        slices = [slices[only]]
    for s in slices:
        n = len(s)
        op.buffer('uv').set_region(queue, np.ascontiguousarray(_uv_view(s)), np.s_[:n], np.s_[:n])
        op.buffer('w_plane').set_region(queue, np.ascontiguousarray(s.w_plane), np.s_[:n], np.s_[:n])
        op.buffer('vis').set_region(queue, np.ascontiguousarray(s.vis), np.s_[:n], np.s_[:n])
        op.num_vis = n
        op()
        queue.finish()
        # ~10 ms of FFMA work keeps the stream busy while the host queues the calls
        _lib.call('kib_fp32_peak_kernel', sink.ptr, context.device.num_sms * 8, 2048,
                  _lib.ctypes.byref(flops), queue.stream)
        a = queue.enqueue_marker()
        for _ in range(reps):
            op()
        b = queue.enqueue_marker()
        b.wait()
        out.append({'vis': n, 'us_per_call': b.time_since(a) / reps * 1e6})
        print(json.dumps(out[-1]), flush=True)


if __name__ == '__main__':
    main()
