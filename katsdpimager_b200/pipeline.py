"""Imaging one channel with the visibilities resident in HBM.

Host-side mirror of the reference's per-channel driver ``frontend.process_channel``
(reference frontend.py:465-658) and its helpers ``make_weights`` (:86-107) and
``make_dirty`` (:110-149), minus everything that is I/O (FITS writer, progress bars,
telstate statistics) -- same call sequence on the :class:`~.imaging.Imaging` facade, same
thresholds, same skipping of empty W slices.

What differs is where the visibilities live.  The reference re-reads every chunk from the
host store and uploads it again on every pass (weights, PSF, every major cycle:
frontend.py:128-138, imaging.py:269-314); on B200 that PCIe traffic costs more than all the
kernels.  Here :class:`ResidentVisibilities` uploads the preprocessed records of a channel
ONCE (array of structures, exactly as the preprocessor emits them, reference
preprocess.cpp:39-52) and every pass unpacks the fields it needs from device memory
(``kib_unpack_records``).  The records are never modified -- prediction subtracts from the
unpacked copy, as in the reference -- so any number of passes can replay them.

The minor cycles of a major cycle run as one device-resident batch
(:meth:`~.imaging.Imaging.clean_cycles`); the sequence of components is identical to calling
``clean_cycle`` in a loop as the reference does (frontend.py:577-582).
"""
import numpy as np

from . import _lib, accel, clean, weight
from .profiling import profile_device


#: bytes per host-to-device copy of an upload
UPLOAD_PIECE = 16 << 20


class ResidentVisibilities:
    """Preprocessed visibility records of one channel in HBM, W slices back to back.

    `slices` is a sequence of contiguous record arrays (fields ``uv, sub_uv, w_plane,
    weights, vis``; one entry per W slice, possibly empty).  Uploads are enqueued on
    `command_queue`; pinned arrays (:class:`~.accel.HostArray`) are copied without staging.
    :meth:`from_raw` builds the same object from raw correlator output, preprocessed on the
    device.
    """

    def __init__(self, command_queue, slices, num_polarizations):
        self.command_queue = command_queue
        self.num_polarizations = num_polarizations
        self.record_bytes = 12 + 12 * num_polarizations
        self._set_counts([len(s) for s in slices])
        self.buffer = accel.DeviceArray(command_queue.context,
                                        (max(1, len(self)) * self.record_bytes,), np.uint8)
        self.h2d_bytes = 0
        self._keepalive = []
        self._uploaded = None
        self.upload(slices)

    def _set_counts(self, counts):
        self.counts = [int(c) for c in counts]
        self.offsets = [int(o) * self.record_bytes
                        for o in np.concatenate(([0], np.cumsum(self.counts)[:-1]))]

    @classmethod
    def from_raw(cls, command_queue, uvw, weights, vis, image_parameters, grid_parameters,
                 mueller_stokes=None, feed_angle1=None, feed_angle2=None, mueller_circular=None,
                 capacity=0):
        """Preprocess one channel on the device (``kib_preprocess``: the whole of the
        reference's ``VisibilityCollector.add``, preprocess.cpp:401-513 + :335-397).

        uvw : (N, 3) float32 metres;  weights : (N, Q) float32;  vis : (N, Q) complex64, the
        raw correlation products; `mueller_stokes` (P x Q, or P x 4 with feed angles) maps them
        to the image's polarizations (identity by default).  The raw arrays are uploaded
        once; the records are produced in HBM and never visit the host.
        """
        ip, gp = image_parameters, grid_parameters
        P = len(ip.fixed.polarizations)
        uvw = np.ascontiguousarray(uvw, np.float32)
        weights = np.ascontiguousarray(weights, np.float32)
        vis = np.ascontiguousarray(vis, np.complex64)
        n, Q = vis.shape
        if mueller_stokes is None:
            if P != Q:
                raise ValueError('mueller_stokes is required when P != Q')
            mueller_stokes = np.identity(P, np.complex64)
        context = command_queue.context
        self = cls.__new__(cls)
        self.command_queue = command_queue
        self.num_polarizations = P
        self.record_bytes = 12 + 12 * P
        self.buffer = accel.DeviceArray(context, (max(1, n) * self.record_bytes,), np.uint8)
        self._keepalive = []
        self.h2d_bytes = 0

        def upload(array):
            if array is None:
                return None
            array = np.ascontiguousarray(array)
            dev = accel.DeviceArray(context, array.shape, array.dtype)
            dev.set(command_queue, array)
            self.h2d_bytes += array.nbytes
            return dev
        inputs = [upload(a) for a in (
            uvw, weights, vis,
            None if feed_angle1 is None else np.asarray(feed_angle1, np.float32),
            None if feed_angle2 is None else np.asarray(feed_angle2, np.float32),
            np.asarray(mueller_stokes, np.complex64),
            None if mueller_circular is None else np.asarray(mueller_circular, np.complex64))]
        nbytes = _lib.c_int64()
        _lib.call('kib_preprocess_scratch_bytes', n, P, gp.w_slices, _lib.ctypes.byref(nbytes))
        scratch = accel.DeviceArray(context, (max(1, nbytes.value),), np.uint8)
        counts = (_lib.c_int64 * gp.w_slices)()
        with profile_device(command_queue, 'preprocess'):
            _lib.call('kib_preprocess', *[d.ptr if d is not None else None for d in inputs[:3]],
                      n, Q, *[d.ptr if d is not None else None for d in inputs[3:]], P,
                      float(np.float32(ip.cell_size)), float(np.float32(gp.fixed.max_w)),
                      gp.w_slices, gp.w_planes, gp.fixed.oversample, int(capacity),
                      self.buffer.ptr, counts, scratch.ptr, scratch.shape[0],
                      command_queue.stream)
        self._set_counts(list(counts))
        self._uploaded = command_queue.enqueue_marker()
        return self

    def upload(self, slices):
        """(Re-)upload the records; `slices` must have the lengths given at construction."""
        if [len(s) for s in slices] != self.counts:
            raise ValueError('slice lengths differ from those this object was built for')
        queue = self.command_queue
        keepalive = []
        self.h2d_bytes = 0
        self.__dict__.pop('_occupancy_valid', None)     # new records, new footprints
        base = self.buffer.ptr.value or 0
        for records, offset in zip(slices, self.offsets):
            if len(records) == 0:
                continue
            if records.dtype.itemsize != self.record_bytes or not records.flags.c_contiguous:
                raise TypeError('records must be contiguous {}-byte structures'.format(
                    self.record_bytes))
            nbytes = len(records) * self.record_bytes
            raw = records.view(np.uint8).reshape(-1)
            if not accel.is_pinned(records):
                pinned = accel.HostArray((nbytes,), np.uint8, context=queue.context)
                pinned[:] = raw
                raw = pinned
            # in pieces, so that small copies of other streams are not stuck behind a 280 MB one
            for piece in range(0, nbytes, UPLOAD_PIECE):
                _lib.call('kib_memcpy_h2d_async', base + offset + piece, raw.ctypes.data + piece,
                          min(UPLOAD_PIECE, nbytes - piece), queue.stream)
            keepalive.append(raw)
            self.h2d_bytes += nbytes
        if self._uploaded is not None and self._keepalive:
            self._uploaded.wait()           # staging of the previous upload may go now
        self._keepalive = keepalive
        self._uploaded = queue.enqueue_marker()

    @property
    def num_w_slices(self):
        return len(self.counts)

    def __len__(self):
        return sum(self.counts)

    def len(self, w_slice):
        return self.counts[w_slice]

    def wait(self):
        """Block until the uploads have completed (the host arrays may then be reused)."""
        self._uploaded.wait()
        self._keepalive = []

    def chunks(self, w_slice, block_size):
        """(start, count) pairs covering the slice in chunks of at most `block_size`."""
        n = self.counts[w_slice]
        for start in range(0, n, block_size):
            yield start, min(block_size, n - start)

    def get(self, w_slice):
        """Records of a W slice as a host record array (blocking; for tests)."""
        from . import preprocess
        n = self.counts[w_slice]
        host = np.empty(n * self.record_bytes, np.uint8)
        if n:
            queue = self.command_queue
            pinned = accel.HostArray(host.shape, np.uint8, context=queue.context)
            _lib.call('kib_memcpy_d2h_async', pinned.ctypes.data,
                      (self.buffer.ptr.value or 0) + self.offsets[w_slice], host.nbytes,
                      queue.stream)
            queue.finish()
            host[:] = pinned
        return host.view(preprocess.make_dtype(self.num_polarizations)).view(np.recarray)

    def unpack(self, queue, w_slice, start, count, uv=None, w_plane=None, weights=None,
               vis=None, vis_from_weights=False):
        """Split records [start, start + count) of a slice into per-field device buffers."""
        def ptr(buffer):
            return buffer.ptr if buffer is not None else None
        base = (self.buffer.ptr.value or 0) + self.offsets[w_slice] + start * self.record_bytes
        with profile_device(queue, 'unpack_records'):
            _lib.call('kib_unpack_records', base, self.record_bytes, count,
                      self.num_polarizations, ptr(uv), ptr(w_plane), ptr(weights), ptr(vis),
                      int(vis_from_weights), queue.stream)

    def occupancy(self, queue, w_slice, kernel_width, grid_size):
        """Column occupancy of a W slice (:func:`.image.column_occupancy`): computed on the
        device from the resident records the first time it is asked for after an upload, then
        kept.  The mask buffers themselves live as long as this object (allocating or freeing
        device memory in the middle of a channel stalls the device)."""
        from . import image
        cache = self.__dict__.setdefault('_occupancy', {})
        valid = self.__dict__.setdefault('_occupancy_valid', set())
        key = (w_slice, kernel_width, grid_size)
        if key not in valid:
            base = (self.buffer.ptr.value or 0) + self.offsets[w_slice]
            out = cache.get(key)
            if out is not None:
                out.zero(queue)
                out.generation = getattr(out, 'generation', 0) + 1
            with profile_device(queue, 'column_occupancy'):
                cache[key] = image.column_occupancy(queue, base, self.counts[w_slice],
                                                    kernel_width, grid_size, self.record_bytes,
                                                    out=out)
            valid.add(key)
        return cache[key]

    def feed(self, imager, w_slice, start, count, field, with_weights):
        """Make records [start, start + count) the imager's current chunk."""
        imager.set_resident(self, w_slice, start, count, field, with_weights)

    def feed_weights(self, imager, w_slice, start, count):
        imager.grid_weights_resident(self, w_slice, start, count)


class HostVisibilities:
    """The same interface over host record arrays, chunk by chunk through the reference's
    calls (``set_coordinates`` / ``set_vis`` / ``set_weights`` / ``grid_weights``): every
    pass uploads every chunk again, as reference frontend.py:128-138 does."""

    def __init__(self, slices):
        self.slices = [s.view(np.recarray) for s in slices]
        self.counts = [len(s) for s in slices]

    @property
    def num_w_slices(self):
        return len(self.counts)

    def __len__(self):
        return sum(self.counts)

    def len(self, w_slice):
        return self.counts[w_slice]

    chunks = ResidentVisibilities.chunks

    def feed(self, imager, w_slice, start, count, field, with_weights):
        chunk = self.slices[w_slice][start:start + count]
        imager.num_vis = count
        imager.set_coordinates(chunk)
        imager.set_vis(chunk[field])
        if with_weights:
            imager.set_weights(chunk.weights)

    def feed_weights(self, imager, w_slice, start, count):
        chunk = self.slices[w_slice][start:start + count]
        imager.grid_weights(np.array(chunk.uv), chunk.weights)


def make_weights(imager, vis, weight_type, vis_block):
    """reference frontend.make_weights (frontend.py:86-107)."""
    imager.clear_weights()
    if weight_type != weight.WeightType.NATURAL:
        for w_slice in range(vis.num_w_slices):
            for start, count in vis.chunks(w_slice, vis_block):
                vis.feed_weights(imager, w_slice, start, count)
    return imager.finalize_weights()


def make_dirty(imager, vis, field, mid_w, vis_block, degrid, full_cycle=False,
               use_occupancy=True):
    """reference frontend.make_dirty (frontend.py:110-149) over resident records.

    `field` is ``'weights'`` (PSF: the weights are gridded as visibilities,
    frontend.py:511) or ``'vis'``."""
    imager.clear_dirty()
    if full_cycle and not degrid:
        imager.model_to_predict()
    # which columns of the grid each slice touches (resident records only): the clears and the
    # transforms skip the others -- same images, see image.column_occupancy
    w_slices = [w for w in range(vis.num_w_slices) if vis.len(w)]
    masks = {}
    if use_occupancy and hasattr(vis, 'occupancy') and hasattr(imager, 'kernel_width'):
        for w_slice in w_slices:
            masks[w_slice] = vis.occupancy(imager.command_queue, w_slice, imager.kernel_width,
                                           imager.buffer('grid').shape[-1])
    # the masks are made for the gridder's grid; the degridder's has the same geometry unless
    # somebody bound another one
    same_degrid_grid = False
    if masks and full_cycle and degrid:
        same_degrid_grid = (imager.buffer('degrid') is not None
                            and imager.buffer('degrid').shape == imager.buffer('grid').shape)
    for i, w_slice in enumerate(w_slices):
        occupancy = masks.get(w_slice)
        if full_cycle and degrid:
            if occupancy is not None and same_degrid_grid:
                # (the model only changes between passes)
                imager.model_to_grid(mid_w[w_slice], occupancy=occupancy, model_unchanged=i > 0)
            else:
                imager.model_to_grid(mid_w[w_slice])
        if occupancy is not None:
            # the grid cleared ahead of time for the next clear: the next slice, or the first
            # slice of the next pass
            imager.clear_grid(occupancy=occupancy,
                              next_occupancy=masks[w_slices[(i + 1) % len(w_slices)]])
        else:
            imager.clear_grid()
        for start, count in vis.chunks(w_slice, vis_block):
            vis.feed(imager, w_slice, start, count, field, full_cycle)
            if full_cycle:
                imager.predict(mid_w[w_slice])
            imager.grid()
        if occupancy is not None:
            imager.grid_to_image(mid_w[w_slice], occupancy=occupancy)
        else:
            imager.grid_to_image(mid_w[w_slice])


def process_channel(imager, vis, image_parameters, grid_parameters, clean_parameters,
                    weight_parameters, major, vis_block, restore=None, out=None,
                    use_occupancy=True):
    """reference frontend.process_channel (frontend.py:494-641) for one channel whose
    visibilities are resident.

    `use_occupancy`: let the grid <-> image transforms skip the grid columns no visibility of
    the W slice touches (:meth:`ResidentVisibilities.occupancy`; same images).

    `restore`, if given, is called as ``restore(imager, psf_patch)`` after the last major
    cycle and before the model is added back (the restoring-beam convolution,
    frontend.py:623-636).  If `out` (pinned host array, polarizations x N x N) is given the
    final image is copied into it asynchronously; call ``imager.command_queue.finish()``
    before reading it.

    Returns a dict of statistics (the quantities frontend.py:646-658 hands to the writer).
    """
    ip, gp, cp = image_parameters, grid_parameters, clean_parameters
    degrid = bool(gp.fixed.degrid)
    queue = imager.command_queue
    num_pols = len(ip.fixed.polarizations)
    stats = {'compressed_vis': len(vis), 'passes': 0}
    if len(vis) == 0:
        stats['skipped'] = 'no data'
        return stats
    if hasattr(imager, 'new_channel'):
        # factor planes of the non-empty W slices are computed in the first pass and reused by
        # the others (0.5 GB each at 8192^2; never across channels)
        imager.new_channel(sum(1 for w in range(vis.num_w_slices) if vis.len(w)))
    imager.clear_model()
    weights_noise, normalized_noise = make_weights(imager, vis, weight_parameters.weight_type,
                                                   vis_block)
    stats['weights_noise'] = weights_noise
    stats['normalized_noise'] = normalized_noise

    # PSF (frontend.py:507-540)
    slice_w_step = float(gp.fixed.max_w / ip.wavelength / (gp.w_slices - 0.5))
    mid_w = np.arange(gp.w_slices) * slice_w_step
    make_dirty(imager, vis, 'weights', mid_w, vis_block, degrid, use_occupancy=use_occupancy)
    stats['passes'] += 1
    dirty = imager.buffer('dirty')
    # pinned scratch is kept on the imager: freeing page-locked memory synchronises the device
    # and costs ~0.1 s once large mappings are registered (profiles/e2e_diag2.py)
    psf_peak = getattr(imager, '_psf_peak_host', None)
    if psf_peak is None or psf_peak.shape != (dirty.shape[0],) or psf_peak.dtype != dirty.dtype:
        psf_peak = accel.HostArray((dirty.shape[0],), dirty.dtype, context=queue.context)
        imager._psf_peak_host = psf_peak
    dirty.get_region(queue, psf_peak,
                     np.s_[:, dirty.shape[1] // 2, dirty.shape[2] // 2], np.s_[:])
    if np.any(psf_peak == 0):
        stats['skipped'] = 'no usable data'
        return stats
    scale = np.reciprocal(psf_peak)
    imager.scale_dirty(scale)
    imager.dirty_to_psf()
    psf_patch = imager.psf_patch()
    stats['psf_patch_size'] = (psf_patch[2], psf_patch[1])
    if restore is not None and hasattr(restore, 'extract'):
        restore.extract(imager, psf_patch)      # PSF core to the host (frontend.py:529-531)

    # major cycles (frontend.py:549-585)
    stats['major'] = 0
    stats['minor'] = 0
    noise = None
    for i in range(major):
        make_dirty(imager, vis, 'vis', mid_w, vis_block, degrid, full_cycle=i != 0,
                   use_occupancy=use_occupancy)
        stats['passes'] += 1
        if i == 0 and restore is not None and hasattr(restore, 'fit'):
            restore.fit()                       # host-side beam fit while the device grids
        imager.scale_dirty(scale)
        stats['major'] += 1
        noise = imager.noise_est()
        imager.clean_reset()
        peak_value = imager.clean_cycle(psf_patch)
        peak_power = clean.metric_to_power(cp.mode, peak_value)
        noise_threshold = noise * clean.noise_threshold_scale(cp.mode, cp.threshold, num_pols)
        mgain_threshold = (1.0 - cp.major_gain) * peak_power
        threshold = max(noise_threshold, mgain_threshold)
        if peak_power <= threshold:
            break
        threshold_metric = clean.power_to_metric(cp.mode, threshold)
        # frontend.py:577-582 counts the terminating call as a minor cycle too
        values, stopped = imager.clean_cycles(psf_patch, threshold_metric, cp.minor - 1)
        stats['minor'] += len(values) + (1 if stopped else 0)
        if i == major - 1:
            noise = imager.noise_est()
    stats['noise'] = noise

    if restore is not None:
        restore(imager, psf_patch)
    imager.add_model_to_dirty()
    if out is not None:
        imager.buffer('dirty').get_async(queue, out)
    return stats
