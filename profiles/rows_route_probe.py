"""Row pass of the fused grid -> image transform: times the routes of kib_grid_to_image_rows
(KIB_ROWS_ROUTE = direct | tma | tmapf) with CUDA events and checks that they produce the
same image.

    python profiles/rows_route_probe.py [pixels grid_size pols reps]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '.')
from katsdpimager_b200 import accel, image, profiling          # noqa: E402


def main():
    args = sys.argv[1:]
    pixels = int(args[0]) if len(args) > 0 else 8192
    grid_size = int(args[1]) if len(args) > 1 else 4922
    pols = int(args[2]) if len(args) > 2 else 4
    reps = int(args[3]) if len(args) > 3 else 10
    context = accel.create_some_context()
    queue = context.create_command_queue()
    lm_scale = 0.2 / pixels
    template = image.GridImageTemplate(context, np.float32)
    plan = template.make_fft_plan((pixels, pixels), (pixels, pixels))
    g2i = template.instantiate_grid_to_image(queue, (pols, grid_size, grid_size), lm_scale,
                                             -lm_scale * pixels / 2, plan)
    g2i.ensure_all_bound()
    rs = np.random.RandomState(1)
    grid = (rs.standard_normal((pols, grid_size, grid_size))
            + 1j * rs.standard_normal((pols, grid_size, grid_size))).astype(np.complex64)
    g2i.buffer('grid').set(queue, grid)
    g2i.buffer('kernel1d').set(queue, rs.uniform(1.0, 2.0, pixels).astype(np.float32))
    g2i.set_w(133.5)
    out = {}
    ref = None
    for route in ('direct', 'directpf', 'tma', 'tmapf', 'direct', 'directpf'):
        os.environ['KIB_ROWS_ROUTE'] = route
        for _ in range(2):
            g2i()
        queue.finish()
        timer = profiling.DeviceTimer()
        profiling.set_timer(timer)
        for _ in range(reps):
            g2i.buffer('image').zero(queue)      # also evicts the previous call's lines from L2
            g2i()
        queue.finish()
        profiling.set_timer(None)
        secs = timer.device_seconds()
        res = {k: v[1] / v[0] * 1e3 for k, v in secs.items()}
        img = g2i.buffer('image').get(queue)
        if ref is None:
            ref = img
        res['max_abs_diff_vs_direct'] = float(np.max(np.abs(img - ref)))
        res['peak'] = float(np.max(np.abs(ref)))
        out.setdefault(route, []).append(res)
        print(route, json.dumps(res), flush=True)
    print(json.dumps({'pixels': pixels, 'grid_size': grid_size, 'pols': pols, 'ms': out}))


if __name__ == '__main__':
    main()
