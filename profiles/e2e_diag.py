import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from katsdpimager_b200 import accel, imaging, profiling, weight, clean, parameters as prm
array, ip, gp, slices = bench.make_channel(0, 3600)
context = accel.Context(0); queue = context.create_command_queue()
cp = prm.CleanParameters(1000, 0.1, 0.85, 5.0, clean.CLEAN_SUMSQ, 0.01, 0.5, 0.02)
wp = prm.WeightParameters(weight.WeightType.NATURAL)
template = imaging.ImagingTemplate(context, array, ip.fixed, wp, gp.fixed, cp)
imager = template.instantiate(queue, ip, gp, 1 << 20, 0, 1)
imager.ensure_all_bound(); imager.clear_weights(); imager.finalize_weights()
mid_w = prm.slice_mid_w(ip, gp)
out = imager.buffer('dirty').empty_like()
def step():
    t = {}
    def tick(name, t0): t[name] = t.get(name, 0) + time.monotonic() - t0
    imager.clear_dirty()
    for w_slice, s in enumerate(slices):
        if len(s) == 0: continue
        imager.clear_grid()
        for start in range(0, len(s), 1 << 20):
            chunk = s[start:start + (1 << 20)]
            imager.num_vis = len(chunk)
            t0 = time.monotonic(); imager.set_coordinates(chunk); tick('coords', t0)
            t0 = time.monotonic(); imager.set_vis(chunk.vis); tick('vis', t0)
            t0 = time.monotonic(); imager.grid(); tick('grid', t0)
        t0 = time.monotonic(); imager.grid_to_image(mid_w[w_slice]); tick('g2i', t0)
    t0 = time.monotonic(); queue.finish(); tick('finish', t0)
    t0 = time.monotonic(); imager.buffer('dirty').get_async(queue, out); queue.finish(); tick('d2h', t0)
    return t
step()
t0 = time.monotonic(); t = step(); total = time.monotonic() - t0
print('total', total, {k: round(v*1e3, 1) for k, v in t.items()})
