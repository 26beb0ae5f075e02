"""Runs the fused grid -> image transform on one 8192^2 plane (grid 4940^2) a few times;
used under ncu to capture columns_kernel / rows_kernel (see profiles/README.md)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from katsdpimager_b200 import accel, image          # noqa: E402


def main():
    pixels = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    grid_size = int(sys.argv[2]) if len(sys.argv) > 2 else 4940
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    pols = int(sys.argv[4]) if len(sys.argv) > 4 else 1
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
    g2i.buffer('image').zero(queue)
    g2i.set_w(133.5)
    for fused in (True, False):
        g2i.fused = fused
        for _ in range(reps):
            queue.finish()
            start = time.perf_counter()
            g2i()
            queue.finish()
            print('fused' if fused else 'cufft', (time.perf_counter() - start) * 1e3, 'ms')


if __name__ == '__main__':
    main()
