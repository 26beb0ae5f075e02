"""CLEAN minor cycles per second on a 4096^2 image (BASELINE config 5) for both launch routes
(KIB_CLEAN_ROUTE = persistent cooperative kernel / pdl = one launch per cycle)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from katsdpimager_b200 import accel, clean, parameters as prm          # noqa: E402
from tests.test_gpu_baseline_shapes import _clean_inputs               # noqa: E402


def main():
    import types
    context = accel.create_some_context()
    queue = context.create_command_queue()
    pixels = 4096
    out = []
    for pols, mode in ((1, clean.CLEAN_I), (4, clean.CLEAN_SUMSQ)):
        dirty, psf = _clean_inputs(pixels, pols, 40 + pols)
        fixed = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
        ip = types.SimpleNamespace(fixed=fixed, pixels=pixels)
        cp = prm.CleanParameters(1000, 0.1, 0.85, 5.0, mode, 0.01, 0.5, 0.02)
        op = clean.CleanTemplate(context, cp, np.float32, pols).instantiate(queue, ip)
        op.ensure_all_bound()
        op.buffer('psf').set(queue, psf)
        for route in ('persistent', 'pdl'):
            os.environ['KIB_CLEAN_ROUTE'] = route
            for side in (255, 1023, 2047):
                patch = (pols, side, side)
                op.buffer('dirty').set(queue, dirty)
                op.buffer('model').zero(queue)
                op.reset()
                op.run_cycles(patch, 0.0, 50)
                queue.finish()
                start = queue.enqueue_marker()
                components, _ = op.run_cycles(patch, 0.0, 1000)
                stop = queue.enqueue_marker()
                stop.wait()
                seconds = stop.time_since(start)
                nbytes = 12.0 * side * side * pols
                row = dict(pols=pols, patch=side, route=route, cycles=len(components),
                           us_per_cycle=seconds / len(components) * 1e6,
                           cycles_per_s=len(components) / seconds,
                           patch_gb_s=nbytes * len(components) / seconds / 1e9)
                out.append(row)
                print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
