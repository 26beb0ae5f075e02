"""Times the two launches of the all-polarization fused grid -> image transform
(kib_grid_to_image_planes_columns / _rows) with CUDA events, for a few settings of the fold
tile ring (KIB_COLUMNS_CG column groups per chunk, KIB_COLUMNS_RING chunks).

    python profiles/fused_planes_probe.py [pixels grid_size pols reps] [cg:ring ...]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, '.')
from katsdpimager_b200 import accel, image, profiling          # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if ':' not in a]
    settings = [a for a in sys.argv[1:] if ':' in a] or ['default']
    pixels = int(args[0]) if len(args) > 0 else 8192
    grid_size = int(args[1]) if len(args) > 1 else 4940
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
    for setting in settings:
        g2i.fused_route = 'planes'
        os.environ['KIB_COLUMNS_ROUTE'] = 'cluster'
        if setting == 'plane:plane':
            g2i.fused_route = 'plane'
        elif setting.startswith('cluster'):
            os.environ['KIB_COLUMNS_M'] = setting.split(':')[1]
        elif setting != 'default':
            cg, ring = setting.split(':')
            os.environ['KIB_COLUMNS_ROUTE'] = 'ring'
            os.environ['KIB_COLUMNS_CG'] = cg
            os.environ['KIB_COLUMNS_RING'] = ring
        g2i._scratch = None
        for _ in range(3):
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
        out[setting] = {k: v[1] / v[0] * 1e3 for k, v in secs.items()}
        if g2i._scratch is not None:
            out[setting]['scratch_mb'] = g2i._scratch.shape[0] / 1e6
        print(setting, json.dumps(out[setting]), flush=True)
    print(json.dumps({'pixels': pixels, 'grid_size': grid_size, 'pols': pols, 'ms': out}))


if __name__ == '__main__':
    main()
