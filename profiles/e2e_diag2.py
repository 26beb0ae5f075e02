"""Where the end-to-end step spends its time (host clock, queue drained between phases)."""
import sys, time, os, tempfile
sys.path.insert(0, '.')
import numpy as np
import bench
from katsdpimager_b200 import accel, beam, imaging, io, pipeline, weight, parameters as prm

context = accel.Context(0); queue = context.create_command_queue()
array, ip, gp, slices = bench.make_channel(0, 3600)
cp = bench.clean_parameters(); wp = prm.WeightParameters(weight.WeightType.ROBUST, 0.0)
template = imaging.ImagingTemplate(context, array, ip.fixed, wp, gp.fixed, cp)
imager = template.instantiate(queue, ip, gp, bench.VIS_BLOCK, 0, bench.MAJOR); imager.ensure_all_bound()
pinned = []
for s in slices:
    h = accel.HostArray((len(s),), s.dtype, context=context); h[:] = s; pinned.append(h.view(np.recarray))
vis = pipeline.ResidentVisibilities(queue, pinned, 4); vis.wait()
restorer = beam.Restorer(context)
name = '/dev/shm/kib_diag_cube.fits'
io.FitsCube.create(name, 1, ip, 1e9, 1e6).close()
cube = io.FitsCube(name); cube.pin(0, 1)
def t(label, fn):
    queue.finish(); a = time.perf_counter(); r = fn(); queue.finish(); print('%-28s %8.1f ms' % (label, (time.perf_counter() - a) * 1e3), flush=True); return r
for rep in range(2):
    t('upload', lambda: vis.upload(pinned))
    t('process_channel(no restore)', lambda: pipeline.process_channel(imager, vis, ip, gp, cp, wp, bench.MAJOR, bench.VIS_BLOCK))
    patch = imager.psf_patch()
    core = t('extract_psf', lambda: beam.extract_psf(queue, imager.buffer('psf'), patch[1:]))
    t('fit_beam', lambda: beam.fit_beam(core))
    t('restorer', lambda: restorer(imager, patch))
    t('store_device', lambda: cube.store_device(0, imager.buffer('dirty'), queue))
    out = imager.buffer('dirty').empty_like()
    t('get_async pinned', lambda: imager.buffer('dirty').get_async(queue, out))
cube.close(); os.unlink(name)
