"""Where does the GPU plane of bench.oracle_plane_check differ from the oracle, and with which
factor route (quadrant table / full plane / computed on the fly)?"""
import sys

import numpy as np

sys.path.insert(0, '.')
import bench                                                     # noqa: E402
from katsdpimager_b200 import accel, imaging, parameters as prm, weight   # noqa: E402


def main():
    context = accel.Context(0)
    queue = context.create_command_queue()
    array, ip, gp, slices = bench.make_channel(0, bench.DUMPS if hasattr(bench, 'DUMPS') else 3600)
    mid_w = prm.slice_mid_w(ip, gp)
    cp = bench.clean_parameters()
    wp = prm.WeightParameters(weight.WeightType.ROBUST, bench.ROBUSTNESS)
    template = imaging.ImagingTemplate(context, array, ip.fixed, wp, gp.fixed, cp)
    imager = template.instantiate(queue, ip, gp, bench.VIS_BLOCK, 0, bench.MAJOR)
    imager.ensure_all_bound()
    imager.buffer('weights_grid').set(queue, np.ones(imager.buffer('weights_grid').shape, np.float32))
    import oracle
    oracle.host.build()
    n, w_slice = min((len(s), i) for i, s in enumerate(slices) if len(s))
    records = slices[w_slice]
    wgrid = imager.get_buffer('weights_grid')
    size = wgrid.shape[-1]
    grid = np.zeros((bench.POLS, size, size), np.complex64)
    oracle.grid(oracle.convolution_kernel(ip, gp), grid, np.ascontiguousarray(wgrid),
                np.ascontiguousarray(records.uv), np.ascontiguousarray(records.sub_uv),
                np.ascontiguousarray(records.w_plane), np.ascontiguousarray(records.vis))
    expected = np.zeros((1, ip.pixels, ip.pixels), np.float32)
    oracle.grid_to_image(grid[:1], expected, oracle.taper(gp, ip.pixels, np.float32),
                         float(ip.pixel_size), -0.5 * ip.pixels * float(ip.pixel_size),
                         np.float64(mid_w[w_slice]))
    peak = float(np.abs(expected).max())
    vis = bench._resident(queue, [records if i == w_slice else records[:0]
                                  for i in range(len(slices))])
    g2i = imager._grid_to_image
    for name, symmetric, planes, calls in (('quadrant+cache', True, 4, 2), ('full+cache', False, 4, 2),
                                           ('full', False, 0, 1)):
        g2i.symmetric_factors = symmetric
        g2i.clear_factor_cache()
        g2i.factor_cache_planes = planes
        for call in range(calls):
            imager.clear_dirty()
            imager.clear_grid()
            imager.set_resident(vis, w_slice, 0, n, 'vis')
            imager.grid()
            imager.grid_to_image(mid_w[w_slice])
            actual = imager.get_buffer('dirty')
            for pol in (0,):
                diff = np.abs(actual[pol] - expected[0])
                y, x = np.unravel_index(np.argmax(diff), diff.shape)
                rows = np.sqrt(np.mean(diff.astype(np.float64) ** 2, axis=1))
                worst_rows = np.argsort(rows)[-3:][::-1]
                print(name, 'call', call, 'pol', pol, 'rms %.3g max %.3g at (y=%d, x=%d)' % (
                    np.sqrt(np.mean(diff.astype(np.float64) ** 2)) / peak, diff.max() / peak, y, x),
                    'worst rows', [(int(r), float('%.3g' % (rows[r] / peak))) for r in worst_rows],
                    flush=True)


if __name__ == '__main__':
    main()
