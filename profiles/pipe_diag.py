"""Run-to-run reproducibility of the dirty image path (diagnostic)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from katsdpimager_b200 import accel, imaging, parameters as prm, weight
from tests import cases

context = accel.Context(0)
queue = context.create_command_queue()
fx = cases.imaging_case(num_baselines=30, num_dumps=20)
ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
wp = prm.WeightParameters(weight.WeightType.NATURAL)
template = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
mid_w = prm.slice_mid_w(ip, gp)
reader = fx['reader']
im = template.instantiate(queue, ip, gp, 1024, 0, 2)
im.ensure_all_bound()
im.clear_weights(); im.finalize_weights()

def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())

def one_slice(w_slice, block):
    im.clear_grid()
    for chunk in reader.iter_slice(0, w_slice, block):
        im.num_vis = len(chunk)
        im.set_coordinates(chunk)
        im.set_vis(chunk.vis)
        im.grid()
    return im.get_buffer('grid').copy()

for w_slice in range(reader.num_w_slices(0)):
    n = reader.len(0, w_slice)
    if n == 0:
        continue
    g1 = one_slice(w_slice, 1024)
    g2 = one_slice(w_slice, 1024)
    g3 = one_slice(w_slice, 300)
    print('slice', w_slice, 'vis', n, 'grid repeat', rel(g2, g1), 'other chunking', rel(g3, g1),
          'max', float(np.abs(g1).max()), 'nonzero', int((g1 != 0).sum()))
    im.clear_dirty()
    im.grid_to_image(mid_w[w_slice])
    d1 = im.get_buffer('dirty').copy()
    im.clear_dirty()
    im.grid_to_image(mid_w[w_slice])
    d2 = im.get_buffer('dirty').copy()
    print('   image repeat from the same grid', rel(d2, d1), 'peak', float(np.abs(d1).max()),
          'median', float(np.median(np.abs(d1))))
