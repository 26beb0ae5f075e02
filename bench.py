#!/usr/bin/env python3
"""Benchmark of the imaging hot path (contract: see the task description / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dumps T]

One *step* = one synthetic MeerKAT channel at BASELINE config 2 (8192^2 image, 4
polarizations, 16 W slices, 7x7 support x 8 oversample, 7.26 M visibilities) IMAGED, i.e. the
reference's ``frontend.process_channel`` (frontend.py:465-641) replayed by
katsdpimager_b200/pipeline.py: robust weights -> PSF pass -> MAJOR major cycles, each a dirty
/ residual pass (W-stacked gridding + fused grid->image transform; from the second cycle on
with image->grid + degridding of the model), a noise estimate and up to MINOR Hogbom minor
cycles -> model added back -> final image.  With N > 1 every rank images its own channel on
its own GPU (weak scaling; no collective on the data path) and stores its plane in a shared
FITS cube.

`value`  = visibilities gridded per second over the whole step (all passes) with the
           visibility records already resident in HBM.
`e2e`    = the same through the public API from HOST memory: per step the channel's pinned
           records are uploaded once (H2D), the channel is imaged, and the final image is
           written in FITS order into the mapped cube file (D2H), all inside the timed region.
`roofline` = the kernel that takes most of the step (row pass of the fused transform, HBM
           roofline over its algorithmic bytes); `rooflines` lists the other kernels of the path.
`--impl reference` / `cpu_baseline` time the CPU oracle (port of the reference's --host
path) on all host threads on a bounded sample: the same fraction of every stage of the step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from katsdpimager_b200 import parameters as prm      # noqa: E402
from katsdpimager_b200 import preprocess, simulate   # noqa: E402

PIXELS = 8192
POLS = 4
W_SLICES = 16
W_PLANES = 16
KERNEL_WIDTH = 7
OVERSAMPLE = 8
NUM_CHANNELS = 64
VIS_BLOCK = int(os.environ.get('KIB_BENCH_VIS_BLOCK', 1 << 20))   # the reference's --vis-block default
MAJOR = 3
MINOR = 1000
ROBUSTNESS = 0.0
METRIC = 'gridded_visibilities_per_sec'
UNIT = 'vis/s'
#: CLEAN work of the step as the B200 run of this synthetic channel executes it (CLEAN is bit
#: exact, so the host path would run the same cycles): PSF patch side and minor cycles over the
#: MAJOR major cycles.  Only used to size the CPU sample; the GPU arm reports what it ran.
NOMINAL_PATCH = 1145
NOMINAL_MINOR = 541


def make_parameters(channel):
    """Image / grid parameters of L-band channel `channel` of 64 (856-1712 MHz)."""
    frequency = 856e6 + 856e6 * (channel + 0.5) / NUM_CHANNELS
    wavelength = 299792458.0 / frequency
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=PIXELS, array=array,
                             image_oversample=5.0)
    fixed_grid = prm.FixedGridParameters(7.0, OVERSAMPLE, 4, array.longest_baseline, KERNEL_WIDTH,
                                         degrid=True)
    gp = prm.GridParameters(fixed_grid, W_SLICES, W_PLANES)
    return array, ip, gp


def clean_parameters():
    """The reference's defaults (frontend.py:337-351) with --major 3 --minor 1000."""
    from katsdpimager_b200 import clean
    return prm.CleanParameters(minor=MINOR, loop_gain=0.1, major_gain=0.85, threshold=5.0,
                               mode=clean.CLEAN_SUMSQ, psf_cutoff=0.01, psf_limit=0.5,
                               border=0.02)


def make_channel(channel, dumps, seed=1):
    """Synthetic visibilities of one channel, quantised and bucketed per W slice
    exactly as the reference's preprocessor does."""
    array, ip, gp = make_parameters(channel)
    uvw = simulate.uvw_tracks(dumps, dump_time=4.0 * 3600 / dumps).reshape(-1, 3)
    lmn, flux = simulate.lsm_lmn_flux()
    # only sources inside this field of view contribute meaningfully; keep all, it is cheap
    vis = simulate.dft_visibilities(uvw / ip.wavelength, lmn, flux)
    rs = np.random.RandomState(seed + channel)
    vis += (rs.standard_normal(vis.shape) + 1j * rs.standard_normal(vis.shape)).astype(np.complex64)
    weights = rs.uniform(0.5, 1.5, (len(uvw), POLS)).astype(np.float32)
    records, w_slice = preprocess.quantise(uvw.astype(np.float32), weights, vis, ip, gp)
    slices = preprocess.bucket_by_slice(records, w_slice, W_SLICES)
    return array, ip, gp, slices


def flops_per_vis(kernel_width, pols):
    """Algorithmic work of (de)gridding one visibility: K^2 (8 P + 6)."""
    return kernel_width ** 2 * (8 * pols + 6)


def column_occupancy_fractions(slices, grid_size):
    """Fraction of the groups of 8 grid columns that the footprints of each W slice cover
    (host restatement of kib_column_occupancy, for the byte accounting of the rooflines)."""
    groups = (grid_size + 7) // 8
    bias = (KERNEL_WIDTH - 1) // 2 - grid_size // 2
    out = []
    for s in slices:
        if len(s) == 0:
            out.append(0.0)
            continue
        u0 = np.unique(np.asarray(s.uv[:, 0], np.int64)) - bias
        u0 = u0[(u0 >= 0) & (u0 + KERNEL_WIDTH <= grid_size)]
        mask = np.zeros(groups + 1, bool)
        mask[u0 >> 3] = True
        mask[(u0 + KERNEL_WIDTH - 1) >> 3] = True       # K <= 8: at most two groups
        out.append(float(mask[:groups].sum()) / groups)
    return out


def step_work(slices):
    """Units of work in one step, shared by both arms."""
    total_vis = sum(len(s) for s in slices)
    planes = POLS * sum(1 for s in slices if len(s))
    return {
        'vis': total_vis,
        'gridded_vis': (1 + MAJOR) * total_vis,               # PSF pass + one pass per major cycle
        'degridded_vis': (MAJOR - 1) * total_vis,
        'weighted_vis': total_vis,
        'grid_to_image_planes': (1 + MAJOR) * planes,
        'image_to_grid_planes': (MAJOR - 1) * planes,
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, device):
        self.samples = []           # (arrival time, fields)
        self.window = [None, None]  # wall-clock bounds of the timed region
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(device), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            fields = [f.strip() for f in line.split(',')]
            if len(fields) >= 7:
                self.samples.append((time.monotonic(), fields))

    def mark_start(self):
        self.window[0] = time.monotonic()

    def mark_stop(self):
        self.window[1] = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, reasons, sm_max, power = [], set(), None, []
        lo, hi = self.window
        inside = [f for t, f in self.samples
                  if (lo is None or t >= lo) and (hi is None or t <= hi + 0.06)]
        # nvidia-smi needs a few hundred ms per sample when several ranks query at once: if no
        # sample fell inside a short timed region, use those taken under the same load just
        # before it (warm-up steps) and say so
        note = None
        if not inside and self.samples:
            inside = [f for _, f in self.samples[-3:]]
            note = 'no sample inside the timed region; last samples of the warm-up used'
        for f in inside:
            try:
                clocks.append(float(f[0]))
                sm_max = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, value in zip(self.NAMES, f[3:7]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        result = {'sm_mhz': float(np.median(clocks)) if clocks else None, 'sm_max_mhz': sm_max,
                  'power_w_max': max(power) if power else None, 'samples': len(clocks),
                  'reasons': sorted(reasons)}
        if note:
            result['note'] = note
        return result


# --------------------------------------------------------------------------- distributed
class Ranks:
    """Barrier and max-over-ranks; torch.distributed is used only for this plumbing
    (gloo on CPU tensors: the data path has no collective)."""

    def __init__(self):
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('gloo', rank=self.rank, world_size=self.world)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, value):
        if self.dist is None:
            return value
        import torch
        t = torch.tensor([value], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum(self, value):
        if self.dist is None:
            return value
        import torch
        t = torch.tensor([value], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU baseline
class CpuSample:
    """A bounded sample of one step for the CPU oracle: the SAME fraction `phi` of every stage
    of the step (gridding, degridding, both transforms, CLEAN), each stage spread over all
    host threads the way independent channels would be.  One call of :meth:`run` is one
    reference-arm step; its wall time / phi estimates the channel time."""

    def __init__(self, ip, gp, slices, threads, phi):
        import oracle
        oracle.host.build()
        self.oracle = oracle
        self.ip, self.gp = ip, gp
        self.threads = threads
        self.phi = phi
        work = step_work(slices)
        self.work = work
        records = np.concatenate([s for s in slices if len(s)]).view(np.recarray)
        self.n_grid = max(threads, int(round(phi * work['gridded_vis'])))
        self.n_degrid = max(threads, int(round(phi * work['degridded_vis'])))
        self.n_g2i = max(1, int(round(phi * work['grid_to_image_planes'])))
        self.n_i2g = max(1, int(round(phi * work['image_to_grid_planes'])))
        self.n_clean = max(1, int(round(phi * NOMINAL_MINOR)))
        self.records = records[:max(self.n_grid, self.n_degrid)]
        self.lut = oracle.convolution_kernel(ip, gp)
        max_uv = int(simulate.longest_baseline() / ip.cell_size)
        self.size = 2 * (max_uv + KERNEL_WIDTH // 2 + 1)
        self.taper = oracle.taper(gp, ip.pixels, np.float32)
        self.lm_scale = float(ip.pixel_size)
        self.lm_bias = -0.5 * ip.pixels * self.lm_scale
        self.mid_w = prm.slice_mid_w(ip, gp)
        self.pool = ThreadPoolExecutor(threads)
        rs = np.random.RandomState(4)
        size = self.size
        self.wgrid = np.ones((POLS, size, size), np.float32)
        self.model_grid = (rs.standard_normal((POLS, size, size))
                           + 1j * rs.standard_normal((POLS, size, size))).astype(np.complex64)
        # CLEAN inputs: noise + a few sources, Gaussian PSF with a pedestal
        n = ip.pixels
        g = np.exp(-0.5 * ((np.arange(n) - n // 2) / 3.0) ** 2).astype(np.float32)
        p = np.exp(-0.5 * ((np.arange(n) - n // 2) / (n / 16.0)) ** 2).astype(np.float32)
        psf1 = np.outer(g, g) + np.float32(0.05) * np.outer(p, p)
        self.psf = np.ascontiguousarray(np.broadcast_to(psf1, (POLS, n, n)))
        self.dirty = (0.01 * rs.standard_normal((POLS, n, n))).astype(np.float32)
        for _ in range(30):
            y, x = rs.randint(n // 8, n - n // 8, 2)
            self.dirty[:, y, x] += rs.uniform(1, 5)
        self.model = np.zeros_like(self.dirty)
        self.description = (
            'oracle (C/numpy port of the reference --host path), {} threads, fraction {:.4g} of '
            'every stage of one channel: {} vis gridded + {} degridded (chunks over threads), '
            '{} grid->image + {} image->grid 8192^2 planes (row blocks over threads), {} CLEAN '
            'cycles with a {}^2 x {} patch (1 thread, sequential by nature)'.format(
                threads, phi, self.n_grid, self.n_degrid, self.n_g2i, self.n_i2g, self.n_clean,
                NOMINAL_PATCH, POLS))

    def _grid(self, n):
        oracle, r = self.oracle, self.records[:n]
        chunks = np.array_split(np.arange(n), self.threads)

        def work(idx):
            values = np.zeros((POLS, self.size, self.size), np.complex64)
            c = r[idx]
            oracle.grid(self.lut, values, self.wgrid, c.uv, c.sub_uv, c.w_plane, c.vis)
            return values
        return list(self.pool.map(work, chunks))

    def _degrid(self, n):
        oracle, r = self.oracle, self.records[:n]
        chunks = np.array_split(np.arange(n), self.threads)

        def work(idx):
            c = r[idx]
            vis = np.ascontiguousarray(c.vis)
            oracle.degrid(self.lut, self.model_grid, c.uv, c.sub_uv, c.w_plane, c.weights, vis)
            return float(vis[0, 0].real) if len(vis) else 0.0
        return list(self.pool.map(work, chunks))

    def run(self):
        """One step; returns the wall time of each stage."""
        oracle = self.oracle
        times = {}
        t0 = time.monotonic()
        grids = self._grid(self.n_grid)
        times['grid'] = time.monotonic() - t0
        t0 = time.monotonic()
        self._degrid(self.n_degrid)
        times['degrid'] = time.monotonic() - t0
        t0 = time.monotonic()
        image = np.zeros((self.ip.pixels, self.ip.pixels), np.float32)
        for k in range(self.n_g2i):
            oracle.host.grid_to_image_threaded(
                grids[k % len(grids)][k % POLS], image, self.taper, self.lm_scale, self.lm_bias,
                np.float64(self.mid_w[k % W_SLICES]), self.pool, self.threads)
        times['grid_to_image'] = time.monotonic() - t0
        t0 = time.monotonic()
        # image -> grid costs the same two 1-D passes and one elementwise pass per plane
        for k in range(self.n_i2g):
            oracle.host.grid_to_image_threaded(
                grids[k % len(grids)][k % POLS], image, self.taper, self.lm_scale, self.lm_bias,
                np.float64(-self.mid_w[k % W_SLICES]), self.pool, self.threads)
        times['image_to_grid'] = time.monotonic() - t0
        t0 = time.monotonic()
        cp = clean_parameters()
        host = oracle.CleanHost(self.ip.pixels, cp.border, cp.mode, cp.loop_gain, self.dirty,
                                self.psf, self.model)
        host.reset()
        for _ in range(self.n_clean):
            host((POLS, NOMINAL_PATCH, NOMINAL_PATCH), 0.0)
        times['clean'] = time.monotonic() - t0
        times['total'] = sum(times.values())
        return times

    def close(self):
        self.pool.shutdown()


def run_reference(args, ranks):
    """--impl reference: the reference's CPU (--host) algorithm, as restated in oracle/,
    with all host threads.  Every step really executes a bounded sample (fraction phi of a
    channel); `ms_per_step` is its measured wall time and `value` the rate it implies."""
    if ranks.rank != 0:
        return
    threads = os.cpu_count() or 1
    array, ip, gp, slices = make_channel(0, args.dumps)
    sample = CpuSample(ip, gp, slices, threads, args.cpu_fraction)
    times = []
    for step in range(args.warmup + args.steps):
        t = sample.run()
        if step >= args.warmup:
            times.append(t)
    sample.close()
    seconds = float(np.mean([t['total'] for t in times]))
    value = args.cpu_fraction * sample.work['gridded_vis'] / seconds
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': seconds * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(sample.work['vis'], args, 1),
        'step_is': 'fraction {:.4g} of one channel (see cpu_baseline.sample)'.format(
            args.cpu_fraction),
        'channels_imaged_per_sec': args.cpu_fraction / seconds,
        'stage_seconds_per_step': {k: float(np.mean([t[k] for t in times])) for k in times[0]},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': sample.description},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(total_vis, args, world):
    return {
        'workload': ('BASELINE configs[1]: MeerKAT-64 L-band 4-pol, 8192^2 image, 16 w-slices, '
                     '7x7 support x8 oversample; one channel imaged per GPU per step '
                     '(frontend.process_channel replay: robust weights, PSF, {} major cycles with '
                     'degridding and up to {} minor cycles each, final image)'.format(MAJOR, MINOR)),
        'pixels': PIXELS, 'polarizations': POLS, 'w_slices': W_SLICES, 'w_planes': W_PLANES,
        'kernel_width': KERNEL_WIDTH, 'oversample': OVERSAMPLE, 'major': MAJOR, 'minor': MINOR,
        'weighting': 'robust {}'.format(ROBUSTNESS), 'degrid': True,
        'vis_per_channel': int(total_vis), 'dumps': args.dumps, 'channels_per_step': world,
        'parallelism': 'channel-parallel x{}'.format(world),
        'l2': 'working set per step (grids 1.6 GB, images 3.2 GB, records 0.4 GB) exceeds '
              'the 126 MB L2, no explicit flush',
    }


# ------------------------------------------------------------------------------ GPU arm
def oracle_plane_check(imager, ip, gp, slices, mid_w, queue):
    """Parity at the benchmark's own size: the sparsest non-empty W slice gridded and
    transformed on the GPU against the CPU oracle (polarization 0 of the 8192^2 plane).
    Returns RMS and maximum difference relative to the peak."""
    import oracle
    oracle.host.build()
    sizes = [(len(s), i) for i, s in enumerate(slices) if len(s)]
    n, w_slice = min(sizes)
    records = slices[w_slice]
    vis = _resident(queue, [records if i == w_slice else records[:0] for i in range(len(slices))])
    imager.clear_dirty()
    imager.clear_grid()
    imager.set_resident(vis, w_slice, 0, n, 'vis')
    imager.grid()
    imager.grid_to_image(mid_w[w_slice])
    actual = imager.get_buffer('dirty')[0]
    wgrid = imager.get_buffer('weights_grid')
    size = wgrid.shape[-1]
    grid = np.zeros((POLS, size, size), np.complex64)
    oracle.grid(oracle.convolution_kernel(ip, gp), grid, np.ascontiguousarray(wgrid),
                np.ascontiguousarray(records.uv), np.ascontiguousarray(records.sub_uv),
                np.ascontiguousarray(records.w_plane), np.ascontiguousarray(records.vis))
    expected = np.zeros((1, ip.pixels, ip.pixels), np.float32)
    oracle.grid_to_image(grid[:1], expected, oracle.taper(gp, ip.pixels, np.float32),
                         float(ip.pixel_size), -0.5 * ip.pixels * float(ip.pixel_size),
                         np.float64(mid_w[w_slice]))
    peak = float(np.abs(expected).max())
    diff = actual - expected[0]
    return {'w_slice': int(w_slice), 'vis': int(n),
            'rms_rel_peak': float(np.sqrt(np.mean(diff.astype(np.float64) ** 2)) / peak),
            'max_rel_peak': float(np.abs(diff).max() / peak), 'bar': 1e-4}


def _resident(queue, slices):
    from katsdpimager_b200 import pipeline
    return pipeline.ResidentVisibilities(queue, slices, POLS)


def side_rows(context, queue, hbm_peak, fp32_peak):
    """Rows of SURVEY.md section 8 that the channel job does not exercise at their own BASELINE
    configuration: direct prediction (config 3: 1000 sources) and the CLEAN minor cycle
    (config 5: 4096^2, 1000 cycles), each timed with events on a few repetitions."""
    import types
    from katsdpimager_b200 import accel, clean, predict
    rows = {}
    # ---- predict, config 3: 2^20 visibilities x 1000 sources x 4 polarizations
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    ip3 = prm.ImageParameters(fixed, wavelength=0.2155, pixels=4096, array=array)
    gp3 = prm.GridParameters(prm.FixedGridParameters(7.0, OVERSAMPLE, 4, array.longest_baseline,
                                                     KERNEL_WIDTH), W_SLICES, W_PLANES)
    nvis, nsrc = 1 << 20, 1000
    op = predict.PredictTemplate(context, np.float32, POLS).instantiate(queue, ip3, gp3, nvis, nsrc)
    op.ensure_all_bound()
    rs = np.random.RandomState(3)
    lmn, flux = simulate.random_sources(nsrc, 0.4 * ip3.pixels * ip3.pixel_size, POLS)
    op.set_sources(lmn, flux)
    uv = np.stack([rs.randint(-1200, 1200, nvis), rs.randint(-1200, 1200, nvis),
                   rs.randint(0, OVERSAMPLE, nvis), rs.randint(0, OVERSAMPLE, nvis)], axis=1)
    op.buffer('uv').set(queue, uv.astype(np.int16))
    op.buffer('w_plane').set(queue, rs.randint(0, W_PLANES, nvis).astype(np.int16))
    op.buffer('weights').set(queue, rs.uniform(0.5, 1.5, (nvis, POLS)).astype(np.float32))
    op.buffer('vis').zero(queue)
    op.num_vis = nvis
    op.set_w(100.0)
    for _ in range(2):
        op()
    queue.finish()
    a = queue.enqueue_marker()
    for _ in range(5):
        op()
    b = queue.enqueue_marker()
    b.wait()
    seconds = b.time_since(a) / 5
    pairs = float(nvis) * nsrc
    flops = pairs * (5 + 8 * POLS + 16)       # phase 5, complex MAC per pol, ~16 for the sincos
    rows['predict'] = {
        'config': 'BASELINE configs[2]: 2^20 visibilities x 1000 sources x 4 polarizations',
        'kernel': 'predict_kernel (kib_predict.cu)', 'ms': seconds * 1e3,
        'vis_sources_per_sec': pairs / seconds, 'bound': 'fp32 + sfu',
        'achieved': flops / seconds / 1e12, 'peak': fp32_peak / 1e12, 'unit': 'TFLOP/s',
        'frac': flops / seconds / fp32_peak,
        'note': 'flops per (visibility, source): 5 for the phase, 8 P for the complex '
                'multiply-accumulates, the polynomial sincos counted as 16'}
    # ---- CLEAN, config 5: 4096^2, 1000 minor cycles
    from tests.test_gpu_baseline_shapes import _clean_inputs
    for pols, mode, side in ((1, clean.CLEAN_I, 255), (4, clean.CLEAN_SUMSQ, 255),
                             (4, clean.CLEAN_SUMSQ, 1023)):
        dirty, psf = _clean_inputs(4096, pols, 40 + pols)
        fx = prm.FixedImageParameters([1, 2, 3, 4][:pols], np.float32)
        ipc = types.SimpleNamespace(fixed=fx, pixels=4096)
        cpc = prm.CleanParameters(1000, 0.1, 0.85, 5.0, mode, 0.01, 0.5, 0.02)
        cl = clean.CleanTemplate(context, cpc, np.float32, pols).instantiate(queue, ipc)
        cl.ensure_all_bound()
        cl.buffer('dirty').set(queue, dirty)
        cl.buffer('psf').set(queue, psf)
        cl.buffer('model').zero(queue)
        cl.reset()
        patch = (pols, side, side)
        cl.run_cycles(patch, 0.0, 50)
        queue.finish()
        a = queue.enqueue_marker()
        components, _ = cl.run_cycles(patch, 0.0, 1000)
        b = queue.enqueue_marker()
        b.wait()
        seconds = b.time_since(a)
        nbytes = 12.0 * side * side * pols
        gbs = nbytes * len(components) / seconds / 1e9
        rows['clean_{}pol_{}'.format(pols, side)] = {
            'config': 'BASELINE configs[4]: 4096^2, {} pol, {}^2 patch, 1000 minor cycles in one '
                      'launch'.format(pols, side),
            'kernel': 'clean_persistent_kernel (kib_clean.cu)', 'cycles': len(components),
            'cycles_per_sec': len(components) / seconds,
            'us_per_cycle': seconds / len(components) * 1e6, 'bound': 'hbm (L2-resident patch)',
            'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak}
        del cl
    return rows


def run_gpu(args, ranks):
    from katsdpimager_b200 import _lib, accel, beam, imaging, io, pipeline, profiling, weight

    context = accel.Context(ranks.local_rank)
    queue = context.create_command_queue()
    # weak scaling with FIXED work per GPU: every rank images its own copy of channel 0 (other
    # channels of the band differ in PSF patch size and CLEAN depth, which would measure the
    # slowest channel instead of the scaling; profiles/r02_scaling.md has such a run too)
    array, ip, gp, slices = make_channel(0, args.dumps)
    work = step_work(slices)
    total_vis = work['vis']
    mid_w = prm.slice_mid_w(ip, gp)
    cp = clean_parameters()
    wp = prm.WeightParameters(weight.WeightType.ROBUST, ROBUSTNESS)
    template = imaging.ImagingTemplate(context, array, ip.fixed, wp, gp.fixed, cp)
    imager = template.instantiate(queue, ip, gp, VIS_BLOCK, 0, MAJOR)
    imager.ensure_all_bound()

    # e2e inputs live in pinned host memory (contract: H2D from pinned memory in the timed region)
    pinned_slices = []
    for s in slices:
        host = accel.HostArray((len(s),), s.dtype, context=context)
        host[:] = s
        pinned_slices.append(host.view(np.recarray))
    resident = _resident(queue, pinned_slices)
    resident.wait()

    restorer = beam.Restorer(context)

    def step_resident():
        return pipeline.process_channel(imager, resident, ip, gp, cp, wp, MAJOR, VIS_BLOCK,
                                        restore=restorer)

    # ---- FP32 roofline denominator: FFMA micro-benchmark on all SMs
    sink = accel.DeviceArray(context, (1,), np.float32)
    flops = _lib.c_double()
    blocks = context.device.num_sms * 8

    def ffma():
        _lib.call('kib_fp32_peak_kernel', sink.ptr, blocks, 64, _lib.ctypes.byref(flops),
                  queue.stream)
    ffma()
    queue.finish()
    peaks = []
    for _ in range(5):
        a = queue.enqueue_marker()
        ffma()
        b = queue.enqueue_marker()
        b.wait()
        peaks.append(flops.value / b.time_since(a))
    fp32_peak = max(peaks)

    # ---- timed region: K steps with resident inputs
    sampler = ClockSampler(ranks.local_rank)        # already streaming when the timed region starts
    stats = None
    for _ in range(args.warmup):
        stats = step_resident()
    queue.finish()
    ranks.barrier()
    sampler.mark_start()
    timer = profiling.DeviceTimer()
    profiling.set_timer(timer)
    launches0 = _lib.kernel_launches
    start = queue.enqueue_marker()
    for _ in range(args.steps):
        stats = step_resident()
    stop = queue.enqueue_marker()
    stop.wait()
    queue.finish()
    sampler.mark_stop()
    seconds = stop.time_since(start)
    launches = _lib.kernel_launches - launches0
    profiling.set_timer(None)
    ranks.barrier()
    clocks = sampler.stop()
    seconds = ranks.max(seconds)
    per_kernel = timer.device_seconds()
    step_seconds = seconds / args.steps
    value = work['gridded_vis'] * ranks.world / step_seconds
    resident_image = imager.get_buffer('dirty')

    # ---- e2e: pinned host records -> device once per step, image the channel, final image in
    # FITS order straight into this rank's plane of a cube file shared by all ranks
    cube_dir = os.environ.get('KIB_CUBE_DIR') or ('/dev/shm' if os.path.isdir('/dev/shm')
                                                  else tempfile.gettempdir())
    cube_name = os.path.join(cube_dir, 'kib_bench_cube_{}.fits'.format(
        os.environ.get('MASTER_PORT', str(os.getpid()))))
    if ranks.rank == 0:
        io.FitsCube.create(cube_name, ranks.world, ip, 299792458.0 / ip.wavelength,
                           856e6 / NUM_CHANNELS).close()
    ranks.barrier()
    cube = io.FitsCube(cube_name)
    cube.pin(ranks.rank, ranks.rank + 1)
    # Copies overlap compute: the records of step k + 1 travel (second stream, copy engine) while
    # step k is imaged, and the image of step k leaves (third stream) while step k + 1 runs.
    # Every step still uploads one channel's records from pinned memory and downloads one
    # image, inside the timed region; the last image is awaited before the clock stops.
    h2d_queue = context.create_command_queue()
    d2h_queue = context.create_command_queue()
    e2e_bufs = [_resident(h2d_queue, pinned_slices), _resident(h2d_queue, pinned_slices)]
    read_done = [None, None]
    e2e_state = {'step': 0, 'stored': None}

    def step_e2e():
        k = e2e_state['step']
        cur, nxt = e2e_bufs[k % 2], e2e_bufs[(k + 1) % 2]
        if read_done[(k + 1) % 2] is not None:      # the step that last read `nxt` must be over
            h2d_queue.enqueue_wait_for_events([read_done[(k + 1) % 2]])
        if not os.environ.get('KIB_E2E_NO_UPLOAD'):         # diagnostics only
            nxt.upload(pinned_slices)
        queue.enqueue_wait_for_events([cur._uploaded])
        pipeline.process_channel(imager, cur, ip, gp, cp, wp, MAJOR, VIS_BLOCK,
                                 restore=restorer)
        read_done[k % 2] = queue.enqueue_marker()
        if os.environ.get('KIB_E2E_NO_STORE'):              # diagnostics only
            e2e_state['stored'] = queue.enqueue_marker()
        else:
            e2e_state['stored'] = cube.store_device(ranks.rank, imager.buffer('dirty'), queue,
                                                    d2h_queue)
        e2e_state['step'] = k + 1

    step_e2e()
    queue.finish()
    d2h_queue.finish()
    h2d_queue.finish()
    ranks.barrier()
    t0 = queue.enqueue_marker()
    for _ in range(args.steps):
        step_e2e()
    queue.enqueue_wait_for_events([e2e_state['stored']])
    t1 = queue.enqueue_marker()
    t1.wait()
    queue.finish()
    h2d_queue.finish()
    d2h_queue.finish()
    e2e_seconds = ranks.max(t1.time_since(t0)) / args.steps
    h2d = e2e_bufs[0].h2d_bytes
    d2h = int(np.prod(cube.shape[1:])) * 4
    # the e2e image (read back from the cube file) against the resident-path image
    e2e_image = np.array(cube.plane(ranks.rank)).astype(np.float32)[:, :, ::-1]
    peak = float(np.abs(resident_image).max())
    diff = e2e_image - resident_image
    parity = {'e2e_vs_resident': {
        'rms_rel_peak': float(np.sqrt(np.mean(diff.astype(np.float64) ** 2)) / peak),
        'max_rel_peak': float(np.abs(diff).max() / peak), 'bar': 1e-4,
        'note': 'final image from the cube file (FITS order undone) against the image of the '
                'resident-records run; gridding uses atomics, so two runs agree to rounding'}}
    del e2e_image, diff
    cube.close()
    ranks.barrier()
    if ranks.rank == 0:
        try:
            os.unlink(cube_name)
        except OSError:
            pass

    # ---- rooflines
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_file):
        hbm_peak, hbm_source = json.load(open(peaks_file))['hbm_gbs'], 'MEASURED_PEAKS.json'
    else:
        hbm_peak, hbm_source = 6650.0, 'fallback (B200_PROFILING.md)'
    traffic_file = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    grid_size = imager.buffer('grid').shape[-1]
    G, N = float(grid_size), float(PIXELS)

    def hbm_roofline(name, kernel, nbytes, traffic_key, note):
        count, secs = per_kernel.get(name, (0, 0.0))
        if not count:
            return None
        gbs = nbytes / (secs / count) / 1e9
        return {'kernel': kernel, 'label': name, 'bound': 'hbm', 'achieved': gbs,
                'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                'traffic': traffic.get(traffic_key), 'traffic_note': traffic.get('note'),
                'peak_source': hbm_source,
                'bytes_per_launch': nbytes, 'avg_launch_ms': secs / count * 1e3,
                'launches_per_step': count / args.steps, 'ms_per_step': secs / args.steps * 1e3,
                'note': note}

    def fp32_roofline(name, kernel, vis_per_step, traffic_key, bytes_per_vis):
        count, secs = per_kernel.get(name, (0, 0.0))
        if not count:
            return None
        vis_per_launch = vis_per_step * args.steps / count
        tflops = vis_per_launch * flops_per_vis(KERNEL_WIDTH, POLS) / (secs / count) / 1e12
        per_vis = traffic.get(traffic_key)
        return {'kernel': kernel, 'label': name, 'bound': 'fp32', 'achieved': tflops,
                'peak': fp32_peak / 1e12, 'unit': 'TFLOP/s', 'frac': tflops / (fp32_peak / 1e12),
                'traffic': per_vis * vis_per_launch if per_vis else None,
                'peak_source': 'FFMA micro-benchmark run in this process (burst); nominal 74.5',
                'flops_per_vis': flops_per_vis(KERNEL_WIDTH, POLS),
                'algorithmic_bytes_per_vis': bytes_per_vis, 'vis_per_launch': vis_per_launch,
                'vis_per_sec': vis_per_launch / (secs / count),
                'avg_launch_ms': secs / count * 1e3, 'ms_per_step': secs / args.steps * 1e3}

    # Column occupancy (katsdpimager_b200.image.column_occupancy): the transforms only touch the
    # groups of 8 grid columns a W slice's footprints cover.  Every non-empty slice is
    # transformed equally often, so the launch-averaged occupied fraction is the plain mean.
    occupied = column_occupancy_fractions(slices, grid_size)
    occ = float(np.mean([f for f in occupied if f > 0.0]))
    rows_roofline = hbm_roofline(
        'grid_to_image_rows', 'rows_kernel<8192,512,16,16,2,MODE,masked> (kib_gridfft.cu)',
        8.0 * N * G * occ + 8.0 * N * N, 'fused_rows_dram_bytes_per_launch',
        'algorithmic bytes = 8 N G f (occupied columns of the half-transformed plane, mean '
        'occupied fraction f = {:.3f}) + 8 N^2 (image read + write); the kernel is bound by its '
        'N-point FFTs (issue slots / shared memory), not by these bytes'.format(occ))
    rooflines = {
        'grid_to_image_columns': hbm_roofline(
            'grid_to_image_columns', 'columns_cluster_kernel<8,1024,8> (kib_gridfft.cu)',
            (8.0 * G * G + 8.0 * N * G) * occ, 'fused_columns_dram_bytes_per_launch',
            'algorithmic bytes = (8 G^2 (grid plane) + 8 N G (half-transformed plane written)) '
            'x occupied fraction {:.3f}'.format(occ)),
        'image_to_grid_rows': hbm_roofline(
            'image_to_grid_rows', 'rows_fwd_kernel<8192,256,16,16,2> (kib_gridfft.cu)',
            4.0 * N * N + 8.0 * N * G, 'fused_rows_fwd_dram_bytes_per_launch',
            'algorithmic bytes = 4 N^2 (image read) + 8 N G (half-transformed plane written); '
            'all-zero model rows are answered without a transform, hence frac > 1'),
        'image_to_grid_columns': hbm_roofline(
            'image_to_grid_columns', 'columns_fwd_cluster_kernel<8,1024,8> (kib_gridfft.cu)',
            (8.0 * N * G + 8.0 * G * G) * occ, 'fused_columns_fwd_dram_bytes_per_launch',
            'algorithmic bytes = (8 N G (read) + 8 G^2 (grid plane written)) x occupied '
            'fraction {:.3f}'.format(occ)),
        'gridder': fp32_roofline(
            'grid', 'grid_stage_kernel<4> + grid_tma_kernel<float,4,7,1> (kib_grid.cu)',
            work['gridded_vis'], 'grid_dram_bytes_per_vis', 10 + 12 * POLS),
        'degridder': fp32_roofline(
            'degrid', 'degrid_kernel (kib_degrid.cu)', work['degridded_vis'],
            'degrid_dram_bytes_per_vis', 10 + 20 * POLS),
    }
    cells = float(grid_size) ** 2
    rooflines['grid_weights'] = hbm_roofline(
        'grid_weights', 'grid_weights_kernel (kib_weight.cu)',
        (8.0 + 8.0 * POLS) * min(VIS_BLOCK, total_vis), None,
        'per visibility: uv 8 B + weights 4 P B read, 4 P B of atomic adds (W1, robust weights); '
        'bytes for a full 1 Mi-visibility chunk')
    rooflines['density_weights'] = hbm_roofline(
        'density_weights', 'density_weights_kernel (kib_weight.cu)', 8.0 * POLS * cells, None,
        '8 P B per grid cell (read + write)')
    rooflines['noise_est'] = hbm_roofline(
        'abs_histogram', 'abs_histogram_kernel (kib_clean.cu): one radix digit of the exact median',
        4.0 * POLS * (PIXELS - 2 * round(cp.border * PIXELS)) ** 2, None,
        '4 B per pixel inside the border per pass; 3 passes per estimate')
    rooflines['scale'] = hbm_roofline(
        'scale', 'scale_kernel (kib_image.cu)', 8.0 * POLS * N * N, None, '8 B per pixel')
    clean_count, clean_secs = per_kernel.get('clean_cycles', (0, 0.0))
    patch = stats.get('psf_patch_size', (0, 0))
    if clean_count and stats.get('minor'):
        cycles = (stats['minor'] - stats['major']) * args.steps     # batched cycles only
        nbytes = 12.0 * patch[0] * patch[1] * POLS
        gbs = nbytes * cycles / clean_secs / 1e9
        rooflines['clean'] = {
            'kernel': 'clean_persistent_kernel<4, SUMSQ> (kib_clean.cu)', 'label': 'clean_cycles',
            'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': gbs / hbm_peak, 'traffic': None, 'peak_source': hbm_source,
            'bytes_per_cycle': nbytes, 'cycles_per_sec': cycles / clean_secs,
            'us_per_cycle': clean_secs / cycles * 1e6, 'patch': list(patch),
            'ms_per_step': clean_secs / args.steps * 1e3,
            'note': '12 B per patch pixel and polarization (psf read, dirty read + write); the '
                    'patch is L2-resident between cycles, the cycle is a chain of dependent '
                    'L2 round trips (DESIGN.md section 4.4)'}
    extra_rows = {}
    if ranks.world == 1:
        extra_rows = side_rows(context, queue, hbm_peak, fp32_peak)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': ranks.world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_seconds * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(total_vis, args, ranks.world),
        'clocks': clocks, 'gpu_launches': launches,
        'channels_imaged_per_sec': ranks.world / step_seconds,
        'e2e': {'value': work['gridded_vis'] * ranks.world / e2e_seconds, 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_seconds * 1e3, 'steps': args.steps,
                'channels_imaged_per_sec': ranks.world / e2e_seconds,
                'output': 'FITS-ordered planes written by the device into a cube file mapped by '
                          'all ranks ({})'.format(cube_dir),
                'overlap': 'records of step k+1 are uploaded and the image of step k is '
                           'downloaded on copy streams while the device images; one upload '
                           'and one download per step, the last download awaited inside the '
                           'timed region'},
        'column_occupancy': {'mean': occ, 'per_w_slice': occupied,
                             'note': 'fraction of the groups of 8 grid columns each W slice touches; '
                                     'the grid <-> image transforms skip the others'},
        'step_stats': {k: (list(v) if isinstance(v, tuple) else
                           (float(v) if isinstance(v, (np.floating, float)) else v))
                       for k, v in stats.items()},
        'work_per_step': work,
        'roofline': rows_roofline,
        'rooflines': {k: v for k, v in rooflines.items() if v is not None},
        'rows': extra_rows,
        'kernels_ms_per_step': {k: v[1] / args.steps * 1e3 for k, v in sorted(per_kernel.items())},
        'parity': parity,
    }
    if ranks.rank == 0:
        if ranks.world == 1 and not args.no_cpu:
            line['parity']['plane_vs_oracle'] = oracle_plane_check(imager, ip, gp, slices,
                                                                   mid_w, queue)
            threads = os.cpu_count() or 1
            sample = CpuSample(ip, gp, slices, threads, args.cpu_fraction)
            sample.run() if args.cpu_warm else None
            t = sample.run()
            sample.close()
            rate = args.cpu_fraction * work['gridded_vis'] / t['total']
            line['cpu_baseline'] = {
                'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                'sample': sample.description, 'sample_seconds': t['total'],
                'stage_seconds': t, 'channels_imaged_per_sec': args.cpu_fraction / t['total']}
        print(json.dumps(line), flush=True)


def main():
    parser = argparse.ArgumentParser(description=__doc__,
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=10)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', choices=['b200', 'reference'], default='b200')
    parser.add_argument('--dumps', type=int, default=3600,
                        help='time samples per baseline (2016 baselines; 3600 -> 7.26 Mvis)')
    parser.add_argument('--cpu-fraction', type=float, default=1.0 / 80,
                        help='fraction of one channel the CPU sample executes per step')
    parser.add_argument('--cpu-warm', action='store_true',
                        help='run the CPU sample twice in the cpu_baseline leg, report the second')
    parser.add_argument('--no-cpu', action='store_true',
                        help='skip the cpu_baseline leg and the oracle parity check')
    args = parser.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    ranks = Ranks()
    try:
        if args.impl == 'reference':
            run_reference(args, ranks)
        else:
            run_gpu(args, ranks)
    finally:
        ranks.close()


if __name__ == '__main__':
    main()
