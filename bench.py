#!/usr/bin/env python3
"""Benchmark of the imaging hot path (contract: see the task description / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dumps T]

One *step* = one dirty-image pass over one synthetic MeerKAT channel at BASELINE
config 2 (8192^2 image, 4 polarizations, 16 W slices, 7x7 support x 8 oversample):
for every W slice { clear grid, grid the slice's visibilities, grid -> image for the 4
polarizations (fused pruned transform: column pass + row pass with the taper / W-term
epilogue) }.  With N > 1 every rank images its own channel on its own GPU (weak scaling, no
collective on the data path).

`value`  = visibilities gridded per second over the whole step with inputs resident in HBM.
`e2e`    = the same metric through the Imaging facade with HOST buffers: per 1 Mi-visibility
           chunk set_coordinates/set_vis (H2D of pinned records) + grid, and a D2H read of
           the dirty image, all inside the timed region; a few imagers on their own command
           queues take channels in turn so that copies overlap kernels (ImagingPipeline).
`roofline` = the kernel that takes most of the step (row pass of the fused transform, HBM
           roofline over its algorithmic bytes); `roofline_columns`, `roofline_gridder` follow.
`--impl reference` times the CPU oracle (port of the reference's --host path) on a bounded
sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
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
VIS_BLOCK = 1 << 20
E2E_DEPTH = int(os.environ.get('KIB_E2E_DEPTH', '3'))   # imagers (command queues) in flight in the e2e leg
METRIC = 'gridded_visibilities_per_sec'
UNIT = 'vis/s'


def make_parameters(channel):
    """Image / grid parameters of L-band channel `channel` of 64 (856-1712 MHz)."""
    frequency = 856e6 + 856e6 * (channel + 0.5) / NUM_CHANNELS
    wavelength = 299792458.0 / frequency
    array = prm.ArrayParameters(simulate.DISH_DIAMETER, simulate.longest_baseline())
    fixed = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=wavelength, pixels=PIXELS, array=array,
                             image_oversample=5.0)
    fixed_grid = prm.FixedGridParameters(7.0, OVERSAMPLE, 4, array.longest_baseline, KERNEL_WIDTH)
    gp = prm.GridParameters(fixed_grid, W_SLICES, W_PLANES)
    return array, ip, gp


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
    """Algorithmic work of gridding one visibility: K^2 (8 P + 6)."""
    return kernel_width ** 2 * (8 * pols + 6)


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
def cpu_sample(ip, gp, slices, threads, vis_per_thread, fft_planes):
    """Time the CPU oracle on a bounded sample: `threads` independent chunks of
    `vis_per_thread` visibilities gridded concurrently (one grid per thread, as separate
    channels would be), and `fft_planes` (W slice, polarization) planes through
    ifft2 + epilogue concurrently.  Returns per-visibility and per-plane wall times."""
    import oracle
    oracle.host.build()
    lut = oracle.convolution_kernel(ip, gp)
    max_uv = int(simulate.longest_baseline() / ip.cell_size)
    size = 2 * (max_uv + KERNEL_WIDTH // 2 + 1)
    records = np.concatenate([s for s in slices if len(s)]).view(np.recarray)
    records = records[:threads * vis_per_thread]
    chunks = np.array_split(np.arange(len(records)), threads)
    wgrid = np.ones((POLS, size, size), np.float32)

    def grid_chunk(idx):
        r = records[idx]
        values = np.zeros((POLS, size, size), np.complex64)
        oracle.grid(lut, values, wgrid, r.uv, r.sub_uv, r.w_plane, r.vis)
        return values

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(grid_chunk, [c[:1000] for c in chunks]))      # warm-up (page faults)
        t0 = time.monotonic()
        grids = list(pool.map(grid_chunk, chunks))
        t_grid = time.monotonic() - t0
    taper = oracle.taper(gp, ip.pixels, np.float32)
    lm_scale = float(ip.pixel_size)
    lm_bias = -0.5 * ip.pixels * lm_scale
    mid_w = prm.slice_mid_w(ip, gp)

    def image_plane(k):
        image = np.zeros((1, ip.pixels, ip.pixels), np.float32)
        oracle.grid_to_image(grids[k % len(grids)][k % POLS:k % POLS + 1], image, taper,
                             lm_scale, lm_bias, mid_w[k % W_SLICES])
        return float(image[0, ip.pixels // 2, ip.pixels // 2])

    fft_threads = min(threads, fft_planes)
    with ThreadPoolExecutor(fft_threads) as pool:
        t0 = time.monotonic()
        list(pool.map(image_plane, range(fft_planes)))
        t_fft = time.monotonic() - t0
    return t_grid / len(records), t_fft / fft_planes, len(records), fft_threads


def cpu_channel_rate(ip, gp, slices, threads, vis_per_thread, fft_planes):
    total_vis = sum(len(s) for s in slices)
    planes = POLS * sum(1 for s in slices if len(s))
    per_vis, per_plane, sample_vis, fft_threads = cpu_sample(
        ip, gp, slices, threads, vis_per_thread, fft_planes)
    seconds = per_vis * total_vis + per_plane * planes
    sample = ('oracle (C/numpy port of the --host path): {} vis gridded on {} threads '
              '({:.3g} us/vis wall) + {} of {} (w-slice, pol) planes through ifft2+epilogue on '
              '{} threads ({:.3g} s/plane wall); channel time extrapolated linearly').format(
        sample_vis, threads, per_vis * 1e6, fft_planes, planes, fft_threads, per_plane)
    return total_vis / seconds, seconds, sample, {
        'grid_vis_per_s': 1.0 / per_vis, 'image_planes_per_s': 1.0 / per_plane}


def run_reference(args, ranks):
    """--impl reference: the reference's CPU (--host) algorithm, as restated in oracle/,
    with all host threads, on a bounded sample per step."""
    if ranks.rank != 0:
        return
    threads = os.cpu_count() or 1
    array, ip, gp, slices = make_channel(0, args.dumps)
    total_vis = sum(len(s) for s in slices)
    rates, times = [], []
    sample = ''
    for step in range(args.warmup + args.steps):
        small = step < args.warmup
        rate, seconds, sample, extra = cpu_channel_rate(
            ip, gp, slices, threads, 2000 if small else args.cpu_vis, 1 if small else
            max(1, min(threads, args.cpu_planes)))
        if not small:
            rates.append(rate)
            times.append(seconds)
    value = float(np.mean(rates))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': float(np.mean(times)) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(total_vis, args, 1),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(total_vis, args, world):
    return {
        'workload': ('BASELINE configs[1]: MeerKAT-64 L-band 4-pol, 8192^2 image, 16 w-slices, '
                     '7x7 support x8 oversample; one channel per GPU per step (dirty image: '
                     'grid all slices + fused pruned grid->image transform per plane)'),
        'pixels': PIXELS, 'polarizations': POLS, 'w_slices': W_SLICES, 'w_planes': W_PLANES,
        'kernel_width': KERNEL_WIDTH, 'oversample': OVERSAMPLE,
        'vis_per_channel': int(total_vis), 'dumps': args.dumps, 'channels_per_step': world,
        'parallelism': 'channel-parallel x{}'.format(world),
        'l2': 'working set per step (grid 0.78 GB, layer 0.54 GB, image 1.07 GB, vis) exceeds '
              'the 126 MB L2, no explicit flush',
    }


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args, ranks):
    from katsdpimager_b200 import _lib, accel, imaging, profiling, weight, clean

    context = accel.Context(ranks.local_rank)
    queue = context.create_command_queue()
    channel = ranks.rank * (NUM_CHANNELS // max(ranks.world, 1)) % NUM_CHANNELS
    array, ip, gp, slices = make_channel(channel, args.dumps)
    total_vis = sum(len(s) for s in slices)
    max_slice = max(len(s) for s in slices)
    max_vis = max(VIS_BLOCK, max_slice)
    mid_w = prm.slice_mid_w(ip, gp)
    cp = prm.CleanParameters(minor=1000, loop_gain=0.1, major_gain=0.85, threshold=5.0,
                             mode=clean.CLEAN_SUMSQ, psf_cutoff=0.01, psf_limit=0.5, border=0.02)
    wp = prm.WeightParameters(weight.WeightType.NATURAL)
    template = imaging.ImagingTemplate(context, array, ip.fixed, wp, gp.fixed, cp)
    imager = template.instantiate(queue, ip, gp, max_vis, 0, 1)
    imager.ensure_all_bound()
    imager.clear_weights()
    imager.finalize_weights()          # natural weights: fill with ones

    # ---- device-resident inputs: one (uv, w_plane, vis) buffer set per W slice
    resident = []
    for s in slices:
        n = len(s)
        if n == 0:
            resident.append(None)
            continue
        bufs = {}
        for name, data in (('uv', imaging._uv_view(s)), ('w_plane', s.w_plane), ('vis', s.vis)):
            slot = imager.slots[name]
            dev = accel.DeviceArray(context, slot.shape, slot.dtype, slot.required_padded_shape())
            dev.set_region(queue, np.ascontiguousarray(data), np.s_[:n], np.s_[:n])
            bufs[name] = dev
        resident.append((n, bufs))
    staging = {name: imager.buffer(name) for name in ('uv', 'w_plane', 'vis')}
    # e2e inputs live in pinned host memory (contract: H2D from pinned memory in the timed region)
    pinned_slices = []
    for s in slices:
        host = accel.HostArray((len(s),), s.dtype, context=context)
        host[:] = s
        pinned_slices.append(host.view(np.recarray))

    def step_resident():
        imager.clear_dirty()
        for w_slice, entry in enumerate(resident):
            if entry is None:
                continue
            n, bufs = entry
            imager.clear_grid()
            imager.bind(**bufs)
            imager.num_vis = n
            imager.grid()
            imager.grid_to_image(mid_w[w_slice])

    def step_e2e(im, out):
        """One channel through the Imaging facade from pinned HOST records to a pinned HOST
        dirty image; everything is enqueued on the imager's own queue, nothing waits."""
        im.clear_dirty()
        for w_slice, s in enumerate(pinned_slices):
            if len(s) == 0:
                continue
            im.clear_grid()
            for start in range(0, len(s), VIS_BLOCK):
                chunk = s[start:start + VIS_BLOCK]
                im.num_vis = len(chunk)
                im.set_coordinates(chunk)
                im.set_vis(chunk.vis)
                im.grid()
            im.grid_to_image(mid_w[w_slice])
        im.buffer('dirty').get_async(im.command_queue, out)

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
    for _ in range(args.warmup):
        step_resident()
    queue.finish()
    ranks.barrier()
    sampler.mark_start()
    timer = profiling.DeviceTimer()
    profiling.set_timer(timer)
    launches0 = _lib.kernel_launches
    start = queue.enqueue_marker()
    for _ in range(args.steps):
        step_resident()
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
    grid_calls = [stop_.time_since(start_) for start_, stop_ in timer.records.get('grid', [])]
    slices_per_step = max(1, len(grid_calls) // args.steps)
    grid_ms_per_slice = [1e3 * float(np.mean(grid_calls[i::slices_per_step]))
                         for i in range(slices_per_step)]
    step_seconds = seconds / args.steps
    value = total_vis * ranks.world / step_seconds

    # ---- e2e: host buffers through the Imaging facade.  E2E_DEPTH imagers on their own command
    # queues take channels in turn (imaging.ImagingPipeline), so the record upload and image
    # download of one step overlap the kernels of the next; every step still copies all of
    # its inputs from pinned host memory and its dirty image back to the host.
    queue.finish()
    pipeline = imaging.ImagingPipeline(template, E2E_DEPTH, ip, gp, VIS_BLOCK, 0, 1)
    dirty_host = []
    for im in pipeline.imagers:
        im.clear_weights()
        im.finalize_weights()
        dirty_host.append(im.buffer('dirty').empty_like())
    for _ in range(len(pipeline)):      # warm-up (first touch, scratch allocation)
        slot, im = pipeline.acquire()
        step_e2e(im, dirty_host[slot])
        pipeline.release(slot)
    pipeline.finish()
    ranks.barrier()
    e2e_steps = args.steps
    t0 = pipeline.queues[0].enqueue_marker()
    ends = []
    for _ in range(e2e_steps):
        slot, im = pipeline.acquire()
        step_e2e(im, dirty_host[slot])
        ends.append(pipeline.release(slot))
    pipeline.finish()
    e2e_seconds = ranks.max(max(e.time_since(t0) for e in ends[-len(pipeline):])) / e2e_steps
    e2e_check = float(dirty_host[0][0, PIXELS // 2, PIXELS // 2])
    h2d = total_vis * slices[0].dtype.itemsize      # whole 60-byte records are uploaded
    d2h = dirty_host[0].nbytes

    # ---- rooflines.  `roofline` is the kernel that takes most of the step, the row pass of the
    # fused grid -> image transform (HBM roofline over its ALGORITHMIC bytes: the half-transformed
    # plane read once, the image plane read and written once); the column pass and the gridder
    # (FP32 pipe, the north star's named kernel) follow as extra objects.
    grid_count, grid_seconds = per_kernel.get('grid', (0, 0.0))
    grid_launch = grid_seconds / max(grid_count, 1)
    vis_per_launch = total_vis * args.steps / max(grid_count, 1)
    achieved = vis_per_launch * flops_per_vis(KERNEL_WIDTH, POLS) / grid_launch / 1e12
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_file):
        hbm_peak, hbm_source = json.load(open(peaks_file))['hbm_gbs'], 'MEASURED_PEAKS.json'
    else:
        hbm_peak, hbm_source = 6650.0, 'fallback (B200_PROFILING.md)'
    traffic_file = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    grid_traffic = (traffic['grid_dram_bytes_per_vis'] * vis_per_launch
                    if 'grid_dram_bytes_per_vis' in traffic else None)
    grid_size = imager.buffer('grid').shape[-1]

    def hbm_roofline(name, kernel, nbytes, traffic_key, note):
        count, seconds = per_kernel.get(name, (0, 0.0))
        if not count:
            return None
        gbs = nbytes / (seconds / count) / 1e9
        return {'kernel': kernel, 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak,
                'unit': 'GB/s', 'frac': gbs / hbm_peak, 'traffic': traffic.get(traffic_key),
                'peak_source': hbm_source, 'bytes_per_launch': nbytes,
                'avg_launch_ms': seconds / count * 1e3, 'launches_per_step': count / args.steps,
                'note': note}

    rows_roofline = hbm_roofline(
        'grid_to_image_rows', 'rows_kernel<8192,256,16,16,2> (kib_gridfft.cu)',
        8.0 * PIXELS * grid_size + 8.0 * PIXELS * PIXELS, 'fused_rows_dram_bytes_per_launch',
        'algorithmic bytes = 8 N G (half-transformed plane) + 8 N^2 (image read + write); the '
        'kernel is issue-bound (77 % of issue slots busy, profiles/r01_fused_fft_ncu_summary.csv)')
    columns_roofline = hbm_roofline(
        'grid_to_image_columns', 'columns_kernel (kib_gridfft.cu)',
        8.0 * grid_size * grid_size + 8.0 * PIXELS * grid_size,
        'fused_columns_dram_bytes_per_launch',
        'algorithmic bytes = 8 G^2 (grid plane) + 8 N G (half-transformed plane written)')
    epilogue_roofline = hbm_roofline(
        'layer_to_image', 'layer_to_image_x2_kernel (kib_image.cu; cuFFT route only)',
        16.0 * PIXELS * PIXELS, 'layer_to_image_dram_bytes_per_launch',
        '16 B per pixel: layer read, image read + write')
    gridder_roofline = {
        'kernel': 'grid_stage_kernel<4> + grid_tma_kernel<float,4,7,1> (kib_grid.cu)',
        'bound': 'fp32',
        'achieved': achieved, 'peak': fp32_peak / 1e12, 'unit': 'TFLOP/s',
        'frac': achieved / (fp32_peak / 1e12), 'traffic': grid_traffic,
        'traffic_note': 'DRAM bytes per launch from the committed ncu capture '
                        '(profiles/r01_traffic.json), scaled to the average launch; '
                        'algorithmic bytes are 58 B/vis, the staged records add 96 B/vis',
        'peak_source': 'FFMA micro-benchmark run in this process (burst); nominal 74.5',
        'flops_per_vis': flops_per_vis(KERNEL_WIDTH, POLS),
        'vis_per_launch': vis_per_launch, 'avg_launch_ms': grid_launch * 1e3}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': ranks.world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_seconds * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(total_vis, args, ranks.world),
        'clocks': clocks, 'gpu_launches': launches,
        'e2e': {'value': total_vis * ranks.world / e2e_seconds, 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_seconds * 1e3, 'steps': e2e_steps,
                'queues_in_flight': E2E_DEPTH, 'centre_pixel': e2e_check},
        'roofline': rows_roofline if rows_roofline is not None else epilogue_roofline,
        'roofline_columns': columns_roofline,
        'roofline_gridder': gridder_roofline,
        'kernels_ms_per_step': {k: v[1] / args.steps * 1e3 for k, v in sorted(per_kernel.items())},
        'channels_per_sec': ranks.world / step_seconds,
        'grid_ms_per_slice': grid_ms_per_slice,
        'vis_per_slice': [len(s) for s in slices if len(s)],
    }
    if ranks.rank == 0:
        if ranks.world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rate, _, sample, extra = cpu_channel_rate(ip, gp, slices, threads, args.cpu_vis,
                                                      max(1, min(threads, args.cpu_planes)))
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': threads,
                                    'kind': 'port', 'sample': sample, **extra}
        print(json.dumps(line), flush=True)


def main():
    parser = argparse.ArgumentParser(description=__doc__,
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=20)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', choices=['b200', 'reference'], default='b200')
    parser.add_argument('--dumps', type=int, default=3600,
                        help='time samples per baseline (2016 baselines; 3600 -> 7.26 Mvis)')
    parser.add_argument('--cpu-vis', type=int, default=150000,
                        help='visibilities per host thread in the CPU sample')
    parser.add_argument('--cpu-planes', type=int, default=4,
                        help='image planes in the CPU sample')
    parser.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
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
