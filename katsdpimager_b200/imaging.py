"""Per-channel imaging facade: every operation of the hot path wired to shared buffers.

Offers the API of the reference's :class:`katsdpimager.imaging.ImagingTemplate` /
:class:`~katsdpimager.imaging.Imaging` (reference imaging.py:11-419) -- the object
``frontend.process_channel`` drives (reference frontend.py:465-658) -- on top of the
sm_100a operations in this package.  The reference file itself can also be used
unchanged over these modules; this module additionally exposes
the batched minor-cycle entry point :meth:`Imaging.clean_cycles`.

Buffer sharing follows the reference (imaging.py:185-209): one ``uv`` / ``w_plane`` /
``vis`` / ``weights`` staging set feeds weighting, prediction and gridding; the UV
``grid`` feeds the FFT stage; ``dirty``, ``model`` and ``psf`` are shared by CLEAN,
scaling, primary-beam correction and the noise estimate.
"""
import numpy as np

from . import _lib, accel, clean, grid, image, predict, weight
from .profiling import profile_device, profile_function


class ImagingTemplate:
    """Holds one template per operation (reference imaging.py:11-51)."""

    @profile_function()
    def __init__(self, context, array_parameters, fixed_image_parameters,
                 weight_parameters, fixed_grid_parameters, clean_parameters, clean_tuning=None):
        self.context = context
        self.array_parameters = array_parameters
        self.fixed_image_parameters = fixed_image_parameters
        self.weight_parameters = weight_parameters
        self.fixed_grid_parameters = fixed_grid_parameters
        self.clean_parameters = clean_parameters
        dtype = fixed_image_parameters.real_dtype
        pols = len(fixed_image_parameters.polarizations)
        self.weights = weight.WeightsTemplate(context, weight_parameters.weight_type, pols)
        self.gridder = grid.GridderTemplate(context, fixed_image_parameters, fixed_grid_parameters)
        self.predict = predict.PredictTemplate(context, dtype, pols)
        self.grid_image = image.GridImageTemplate(context, dtype)
        self.psf_patch = clean.PsfPatchTemplate(context, dtype, pols)
        self.noise_est = clean.NoiseEstTemplate(context, dtype, pols)
        self.clean = clean.CleanTemplate(context, clean_parameters, dtype, pols,
                                         tuning=clean_tuning)
        self.scale = image.ScaleTemplate(context, dtype, pols)
        self.add_image = image.AddImageTemplate(context, dtype, pols)
        self.apply_primary_beam = image.ApplyPrimaryBeamTemplate(context, dtype, pols)
        self.degridder = None
        if fixed_grid_parameters.degrid:
            self.degridder = grid.DegridderTemplate(
                context, fixed_image_parameters, fixed_grid_parameters)

    def instantiate(self, *args, **kwargs):
        return Imaging(self, *args, **kwargs)


class _Staging:
    """Pinned host mirror of a per-visibility slot plus the event that marks the end of
    the last upload from it (it must not be overwritten before then)."""

    def __init__(self, context, slot):
        self.array = accel.HostArray(slot.shape, slot.dtype, slot.required_padded_shape(),
                                     context=context)
        self.transfer_event = None

    def wait(self):
        if self.transfer_event is not None:
            self.transfer_event.wait()
            self.transfer_event = None


def _uv_view(coords):
    """View the adjacent ``uv`` and ``sub_uv`` int16 pairs of a record array as N x 4."""
    uv_type, uv_offset = coords.dtype.fields['uv'][:2]
    sub_type, sub_offset = coords.dtype.fields['sub_uv'][:2]
    pair = np.dtype(('i2', (2,)))
    if uv_type != pair or sub_type != pair or sub_offset != uv_offset + pair.itemsize:
        raise TypeError('uv and sub_uv must be adjacent pairs of int16')
    quad = np.dtype(dict(names=['q'], formats=[('i2', (4,))], offsets=[uv_offset],
                         itemsize=coords.dtype.itemsize))
    return coords.view(quad)['q']


class _RecordBlock:
    """A block of preprocessed visibility records (array of structures, the layout of
    reference preprocess.cpp:39-52) uploaded as is and split into the per-field device
    buffers by ``kib_unpack_records``.

    The reference extracts each field on the host into pinned staging
    (``staging[:N] = chunk.field``, reference imaging.py:269-314); for strided record
    fields that numpy copy costs more than all the device work of a chunk, so when the
    caller hands over a contiguous record array the whole block makes one trip instead.
    """

    def __init__(self, command_queue, max_vis, num_polarizations):
        self.command_queue = command_queue
        self.num_polarizations = num_polarizations
        self.record_bytes = 12 + 12 * num_polarizations
        context = command_queue.context
        self.device = accel.DeviceArray(context, (max_vis * self.record_bytes,), np.uint8)
        self.staging = accel.HostArray((max_vis * self.record_bytes,), np.uint8, context=context)
        self.staging_event = None
        self.upload_event = None
        self.source = None          # (address, count) of the records now on the device

    def matches(self, records):
        """True if `records` is a contiguous record array with the expected layout."""
        dtype = records.dtype
        if dtype.names is None or records.ndim != 1 or dtype.itemsize != self.record_bytes \
                or not records.flags.c_contiguous:
            return False
        P = self.num_polarizations
        want = {'uv': (np.dtype(('i2', (2,))), 0), 'sub_uv': (np.dtype(('i2', (2,))), 4),
                'w_plane': (np.dtype('i2'), 8), 'weights': (np.dtype(('f4', (P,))), 12),
                'vis': (np.dtype(('c8', (P,))), 12 + 4 * P)}
        return all(name in dtype.fields and dtype.fields[name][:2] == want[name] for name in want)

    def field_of(self, array, name, count):
        """True if `array` is field `name` of the record block currently on the device."""
        if self.source is None or not isinstance(array, np.ndarray) or len(array) != count:
            return False
        address, n = self.source
        P = self.num_polarizations
        offset, dtype, inner = {'weights': (12, np.float32, (P,)),
                                'vis': (12 + 4 * P, np.complex64, (P,)),
                                'uv': (0, np.int16, (2,))}[name]
        return (n == count and array.dtype == dtype and array.shape[1:] == inner
                and array.ctypes.data == address + offset
                and (count <= 1 or array.strides[0] == self.record_bytes)
                and (array.ndim == 1 or array.strides[1] == array.dtype.itemsize))

    def upload(self, records):
        """Start the transfer of `records`.  Subsequent ``set_vis`` / ``set_weights`` calls
        that pass fields of the same array (as frontend.make_dirty does) reuse this copy,
        so the array must not be modified in between."""
        count = len(records)
        address = records.ctypes.data
        nbytes = count * self.record_bytes
        raw = records.view(np.uint8).reshape(-1)
        queue = self.command_queue
        if accel.is_pinned(records):
            src = raw
        else:
            if self.staging_event is not None:
                self.staging_event.wait()
            self.staging[:nbytes] = raw         # one contiguous copy
            src = self.staging
        _lib.call('kib_memcpy_h2d_async', self.device.ptr, src.ctypes.data, nbytes, queue.stream)
        #: marks the end of the transfer: until then a pinned `records` array is being read by
        #: the copy engine and must not be overwritten (:meth:`wait`)
        self.upload_event = queue.enqueue_marker()
        if src is self.staging:
            self.staging_event = self.upload_event
        self.source = (address, count)

    def wait(self):
        """Block until the last upload has completed (a pinned source may then be reused)."""
        if self.upload_event is not None:
            self.upload_event.wait()
            self.upload_event = None

    def unpack(self, count, uv=None, w_plane=None, weights=None, vis=None, vis_from_weights=False):
        def ptr(buffer):
            return buffer.ptr if buffer is not None else None
        with profile_device(self.command_queue, 'unpack_records'):
            _lib.call('kib_unpack_records', self.device.ptr, self.record_bytes, count,
                      self.num_polarizations, ptr(uv), ptr(w_plane), ptr(weights), ptr(vis),
                      int(vis_from_weights), self.command_queue.stream)

    def invalidate(self):
        self.source = None


#: compound slot -> member slots, as reference imaging.py:185-204
_SHARED_SLOTS = {
    'weights': ['weights:weights', 'predict:weights', 'continuum_predict:weights'],
    'weights_grid': ['weights:grid', 'gridder:weights_grid'],
    'uv': ['weights:uv', 'gridder:uv', 'predict:uv', 'continuum_predict:uv'],
    'w_plane': ['gridder:w_plane', 'predict:w_plane', 'continuum_predict:w_plane'],
    'vis': ['gridder:vis', 'predict:vis', 'continuum_predict:vis'],
    'grid': ['gridder:grid', 'grid_to_image:grid'],
    'layer': ['grid_to_image:layer'],
    'dirty': ['grid_to_image:image', 'noise_est:dirty', 'clean:dirty', 'scale:data',
              'add_image:dest', 'apply_primary_beam_dirty:data'],
    'model': ['clean:model', 'apply_primary_beam_model:data', 'add_image:src'],
    'psf': ['clean:psf', 'psf_patch:psf'],
    'tile_max': ['clean:tile_max'],
    'tile_pos': ['clean:tile_pos'],
    'peak_value': ['clean:peak_value'],
    'peak_pos': ['clean:peak_pos'],
    'peak_pixel': ['clean:peak_pixel'],
    'beam_power': ['apply_primary_beam_model:beam_power', 'apply_primary_beam_dirty:beam_power'],
}


class Imaging(accel.OperationSequence):
    """All device state for imaging one channel (reference imaging.py:81-419)."""

    @profile_function()
    def __init__(self, template, command_queue, image_parameters, grid_parameters,
                 max_vis, max_sources, major, allocator=None):
        assert image_parameters.fixed == template.fixed_image_parameters
        assert grid_parameters.fixed == template.fixed_grid_parameters
        self.template = template
        context = template.context
        pixels = image_parameters.pixels
        dtype = image_parameters.fixed.real_dtype
        lm_scale = float(image_parameters.pixel_size)
        lm_bias = -0.5 * pixels * lm_scale
        image_shape = (len(image_parameters.fixed.polarizations), pixels, pixels)
        fft_plan = template.grid_image.make_fft_plan(image_shape[1:], image_shape[1:])
        degrid = bool(grid_parameters.fixed.degrid)

        def instantiate(tmpl, *args):
            return tmpl.instantiate(command_queue, *args, allocator)

        self._gridder = instantiate(template.gridder, template.array_parameters,
                                    image_parameters, grid_parameters, max_vis)
        grid_shape = self._gridder.slots['grid'].shape
        self._continuum_predict = instantiate(template.predict, image_parameters, grid_parameters,
                                              max_vis, max_sources)
        self._weights = instantiate(template.weights, grid_shape, max_vis)
        self._weights.robustness = template.weight_parameters.robustness
        self._grid_to_image = template.grid_image.instantiate_grid_to_image(
            command_queue, grid_shape, lm_scale, lm_bias, fft_plan, allocator)
        # lm_bias = -pixels / 2 * lm_scale above and the taper of grid.ConvolutionKernel is an
        # even function sampled symmetrically, so one quadrant of the factor plane would do
        # (GridToImage.symmetric_factors: 2 % faster, a quarter of the memory).  Left off: a
        # mirrored pixel then gets the factor of its partner, whose direction cosine differs in
        # the last bit, and at w = 2e4 wavelengths that moves single pixels at the image edge
        # (where the taper division amplifies everything) by 3e-3 of the peak against the host
        # result instead of 2e-4 (profiles/r02_transform.md).
        self._grid_to_image.symmetric_factors = False
        self._psf_patch = instantiate(template.psf_patch, image_shape)
        self._noise_est = instantiate(template.noise_est, image_shape,
                                      template.clean_parameters.border)
        self._clean = instantiate(template.clean, image_parameters)
        self._scale = instantiate(template.scale, image_shape)
        self._add_image = instantiate(template.add_image, image_shape)
        # thresholds are set by apply_primary_beam()
        self._apply_primary_beam_model = instantiate(
            template.apply_primary_beam, image_shape, 0.0, 0.0)
        self._apply_primary_beam_dirty = instantiate(
            template.apply_primary_beam, image_shape, 0.0, np.nan)

        def upload_taper(kernel):
            taper = accel.DeviceArray(context, (pixels,), dtype)
            host = taper.empty_like()
            kernel.taper(pixels, host)
            taper.set(command_queue, host)
            return taper

        self._grid_to_image.bind(kernel1d=upload_taper(self._gridder.convolve_kernel))
        self._image_to_grid = None
        if degrid:
            self._predict = instantiate(template.degridder, template.array_parameters,
                                        image_parameters, grid_parameters, max_vis)
            self._image_to_grid = template.grid_image.instantiate_image_to_grid(
                command_queue, self._predict.slots['grid'].shape, lm_scale, lm_bias, fft_plan,
                allocator)
            self._image_to_grid.bind(kernel1d=upload_taper(self._predict.convolve_kernel))
        else:
            max_components = min(pixels**2, (major - 1) * template.clean_parameters.minor)
            self._predict = instantiate(template.predict, image_parameters, grid_parameters,
                                        max_vis, max_components)
        self._model_components = {}

        operations = [
            ('weights', self._weights),
            ('gridder', self._gridder),
            ('predict', self._predict),
            ('continuum_predict', self._continuum_predict),
            ('grid_to_image', self._grid_to_image),
            ('psf_patch', self._psf_patch),
            ('noise_est', self._noise_est),
            ('clean', self._clean),
            ('scale', self._scale),
            ('add_image', self._add_image),
            ('apply_primary_beam_model', self._apply_primary_beam_model),
            ('apply_primary_beam_dirty', self._apply_primary_beam_dirty)
        ]
        compounds = {name: list(members) for name, members in _SHARED_SLOTS.items()}
        if self._weights.template.grid_weights is None:
            # natural weighting: the weights operation has no uv / weights slots
            compounds['weights'].remove('weights:weights')
            compounds['uv'].remove('weights:uv')
        if degrid:
            operations.append(('image_to_grid', self._image_to_grid))
            compounds['degrid'] = ['predict:grid', 'image_to_grid:grid']
            compounds['layer'].append('image_to_grid:layer')
            compounds['model'].append('image_to_grid:image')
        super().__init__(command_queue, operations, compounds, allocator=allocator)
        # dirty_to_psf swaps the two buffers, so their padding must be interchangeable
        for a, b in zip(self.slots['dirty'].dimensions, self.slots['psf'].dimensions):
            a.link(b)
        self.host_buffer = {name: _Staging(context, self.slots[name])
                            for name in ('weights', 'uv', 'w_plane', 'vis') if name in self.slots}
        self._records = _RecordBlock(command_queue, max_vis, image_shape[0])

    def __call__(self, **kwargs):
        raise NotImplementedError()

    # ------------------------------------------------------------------ visibilities
    @property
    def num_vis(self):
        return self._gridder.num_vis

    @num_vis.setter
    def num_vis(self, value):
        self._gridder.num_vis = value
        self._predict.num_vis = value
        self._continuum_predict.num_vis = value

    def _set_buffer(self, name, N, data, extra_index=()):
        """Stage `data` (N rows) in pinned memory and start its upload."""
        if len(data) != N:
            raise ValueError('Lengths do not match')
        staging = self.host_buffer[name]
        staging.wait()
        index = (np.s_[:N],) + extra_index
        staging.array[index] = data
        self.buffer(name).set_region(self.command_queue, staging.array, index, index,
                                     blocking=False)
        staging.transfer_event = self.command_queue.enqueue_marker()

    @profile_function()
    def set_coordinates(self, coords):
        """Upload UVW coordinates from a record array with fields ``uv``, ``sub_uv``
        (adjacent int16 pairs) and ``w_plane``."""
        if len(coords) != self.num_vis:
            raise ValueError('Lengths do not match')
        if self._records.matches(coords):
            # whole records in one transfer, fields split on the device
            self._records.upload(coords)
            self._records.unpack(self.num_vis, uv=self.buffer('uv'), w_plane=self.buffer('w_plane'))
            return
        self._records.invalidate()
        self._set_buffer('uv', self.num_vis, _uv_view(coords))
        self._set_buffer('w_plane', self.num_vis, coords['w_plane'])

    @profile_function()
    def set_vis(self, vis):
        N = self.num_vis
        if self._records.field_of(vis, 'vis', N):
            self._records.unpack(N, vis=self.buffer('vis'))
        elif self._records.field_of(vis, 'weights', N):
            # the PSF is made by gridding the weights (reference frontend.py:511)
            self._records.unpack(N, vis=self.buffer('vis'), vis_from_weights=True)
        else:
            self._set_buffer('vis', N, vis)

    @profile_function()
    def set_weights(self, weights):
        """Set statistical weights for prediction"""
        N = self.num_vis
        if self._records.field_of(weights, 'weights', N):
            self._records.unpack(N, weights=self.buffer('weights'))
        else:
            self._set_buffer('weights', N, weights)

    # ------------------------------------------------ visibilities resident in HBM
    @profile_function()
    def set_resident(self, resident, w_slice, start, count, field='vis', with_weights=False):
        """Equivalent of ``num_vis = count; set_coordinates(chunk); set_vis(chunk[field])``
        (and ``set_weights(chunk.weights)`` with `with_weights`) for records
        [start, start + count) of W slice `w_slice` of a
        :class:`~.pipeline.ResidentVisibilities`: the fields are unpacked from device memory,
        nothing crosses PCIe."""
        if field not in ('vis', 'weights'):
            raise ValueError('field must be vis or weights')
        self.num_vis = count
        self._records.invalidate()
        resident.unpack(self.command_queue, w_slice, start, count,
                        uv=self.buffer('uv'), w_plane=self.buffer('w_plane'),
                        weights=self.buffer('weights') if with_weights else None,
                        vis=self.buffer('vis'), vis_from_weights=field == 'weights')

    @profile_function()
    def grid_weights_resident(self, resident, w_slice, start, count):
        """:meth:`grid_weights` for resident records."""
        if count > self._gridder.max_vis:
            raise ValueError('chunk is larger than max_vis')
        self._records.invalidate()
        resident.unpack(self.command_queue, w_slice, start, count,
                        uv=self.buffer('uv'), weights=self.buffer('weights'))
        self._weights.grid(count)

    # ----------------------------------------------------------------------- weights
    @profile_function()
    def clear_weights(self):
        self._weights.clear()

    @profile_function()
    def grid_weights(self, uv, weights):
        self._set_buffer('uv', len(uv), uv, (np.s_[:2],))
        self._set_buffer('weights', len(uv), weights)
        self._weights.grid(len(uv))

    @profile_function()
    def finalize_weights(self):
        return self._weights.finalize()

    # ------------------------------------------------------------ gridding / prediction
    def _zero(self, name):
        with profile_device(self.command_queue, 'clear_' + name):
            self.buffer(name).zero(self.command_queue)

    #: Zero the grid of the next W slice on a second stream while the current slice is still
    #: being gridded / transformed (two grid buffers that swap at every clear_grid; the memset
    #: is HBM-bound, the transform kernels are not).  False = the reference's in-order clear.
    double_buffer_grid = True

    @profile_function()
    def clear_grid(self, occupancy=None, next_occupancy=None):
        """Zero the grid (reference imaging.py:253-255).

        Not in the reference: with `occupancy` (the column occupancy of everything that will be
        gridded before the next clear, :func:`.image.column_occupancy`) only those columns are
        guaranteed to be zero afterwards -- the caller reads the grid through
        ``grid_to_image(w, occupancy=...)`` only.  `next_occupancy` is the same for the clear
        after this one, which the second grid buffer receives ahead of time."""
        current = self.buffer('grid')
        if occupancy is not None and not self._grid_to_image.uses_occupancy:
            occupancy = next_occupancy = None       # the dense transform reads every column
        if not self.double_buffer_grid or current is None:
            self._clear_grid_buffer(self.command_queue, current, occupancy)
            return
        queue = self.command_queue
        spare = getattr(self, '_grid_spare', None)
        if spare is not None and (spare.shape != current.shape or spare.dtype != current.dtype
                                  or spare.padded_shape != current.padded_shape):
            spare = None                    # the caller bound a different grid buffer
        if spare is None:
            # first use: clear the bound grid in order, start clearing a second one on the side
            self._side_queue = queue.context.create_command_queue()
            self._grid_spare = accel.DeviceArray(queue.context, current.shape, current.dtype,
                                                 current.padded_shape)
            self._clear_grid_buffer(queue, current, occupancy)
            self._spare_cleared_for = self._clear_grid_buffer(self._side_queue, self._grid_spare,
                                                              next_occupancy)
            self._spare_zeroed = self._side_queue.enqueue_marker()
            return
        # everything enqueued so far (the transform of the previous slice) is done with
        # `current` when this marker completes; only then may the side stream wipe it
        done = queue.enqueue_marker()
        queue.enqueue_wait_for_events([self._spare_zeroed])
        self.bind(grid=spare)
        if not self._clear_covers(self._spare_cleared_for, occupancy):
            # the spare was prepared for other columns (first slice of a pass, new channel)
            self._clear_grid_buffer(queue, spare, occupancy)
        self._grid_spare = current
        self._side_queue.enqueue_wait_for_events([done])
        self._spare_cleared_for = self._clear_grid_buffer(self._side_queue, current,
                                                          next_occupancy)
        self._spare_zeroed = self._side_queue.enqueue_marker()

    @staticmethod
    def _clear_covers(cleared_for, occupancy):
        """Does a clear made for `cleared_for` (None = everything) leave the columns of
        `occupancy` zero?"""
        if cleared_for is None:
            return True
        if occupancy is None:
            return False
        return cleared_for[0] is occupancy and cleared_for[1] == getattr(occupancy, 'generation', 0)

    def _clear_grid_buffer(self, queue, buffer, occupancy):
        """Zero `buffer` (whole, or the column groups of `occupancy`) on `queue`; returns what
        :meth:`_clear_covers` compares."""
        with profile_device(queue, 'clear_grid'):
            pitch, plane = buffer.padded_shape[2], buffer.padded_shape[1] * buffer.padded_shape[2]
            if (occupancy is None or buffer.dtype != np.complex64 or pitch % 2 or plane % 2
                    or (buffer.ptr.value or 0) % 16):
                buffer.zero(queue)          # (kib_clear_columns wants 16-byte aligned pairs)
                return None
            _lib.call('kib_clear_columns', buffer.ptr, buffer.padded_shape[2],
                      buffer.padded_shape[1] * buffer.padded_shape[2], buffer.shape[2],
                      buffer.shape[0], occupancy.ptr, _lib.dtype_code(buffer.dtype), queue.stream)
        return (occupancy, getattr(occupancy, 'generation', 0))     # keeps the mask alive

    @profile_function()
    def clear_dirty(self):
        self._zero('dirty')

    @profile_function()
    def clear_model(self):
        self._zero('model')
        self._model_components.clear()

    @profile_function()
    def grid(self):
        self._gridder()

    @profile_function()
    def predict(self, w):
        if not self.template.fixed_grid_parameters.degrid:
            self._predict.set_w(w)
        self._predict()

    @profile_function()
    def continuum_predict(self, w):
        self._continuum_predict.set_w(w)
        self._continuum_predict()

    def set_sky_model(self, sky_model, phase_centre):
        self._continuum_predict.set_sky_model(sky_model, phase_centre)

    @profile_function()
    def grid_to_image(self, w, occupancy=None):
        """`occupancy` (not in the reference; :func:`.image.column_occupancy` of everything
        gridded since :meth:`clear_grid`) lets the transform skip the empty columns."""
        self._grid_to_image.set_w(w)
        self._grid_to_image.occupancy = occupancy
        self._grid_to_image()

    @profile_function()
    def model_to_grid(self, w, occupancy=None, model_unchanged=False):
        """`occupancy` (not in the reference): column occupancy of the visibilities that will be
        predicted from the grid; the other columns of the grid are not computed.
        `model_unchanged` (not in the reference): the model is what it was at the previous call
        (another W slice of the same pass), so its empty rows need not be looked for again."""
        if self._image_to_grid is None:
            raise RuntimeError('Can only use model_to_grid with degridding')
        self._image_to_grid.set_w(w)
        self._image_to_grid.occupancy = occupancy
        self._image_to_grid.image_unchanged = bool(model_unchanged)
        self._image_to_grid()

    @property
    def kernel_width(self):
        return self.template.fixed_grid_parameters.kernel_width

    def new_channel(self, w_slices=0):
        """Tell the imager that the next calls image another channel with `w_slices` W slices
        (not in the reference, which builds a new Imaging per channel): per-slice factor planes
        of the grid -> image transform are kept across the passes over this channel
        (:attr:`.image.GridToImage.factor_cache_planes`) and dropped here."""
        self._grid_to_image.clear_factor_cache()
        dirty = self.buffer('dirty')
        plane_bytes = (2 * dirty.dtype.itemsize * dirty.shape[1] * dirty.shape[2]
                       if dirty is not None else 1)
        # at most 24 GB of tables (fewer planes than W slices would only thrash: none then)
        fits = int(w_slices) * plane_bytes <= 24 << 30
        self._grid_to_image.factor_cache_planes = int(w_slices) if fits else 0
        if getattr(self, '_spare_zeroed', None) is not None:
            # the look-ahead clear of the second grid buffer reads an occupancy mask that the
            # new channel may be about to refill in place
            self.command_queue.enqueue_wait_for_events([self._spare_zeroed])

    @profile_function()
    def model_to_predict(self):
        if self.template.fixed_grid_parameters.degrid:
            raise RuntimeError('Can only use model_to_predict with direct prediction')
        self._predict.set_sky_image(self._model_components)

    # ------------------------------------------------------------------ image domain
    @profile_function()
    def scale_dirty(self, scale_factor):
        self._scale.set_scale_factor(scale_factor)
        self._scale()

    @profile_function()
    def add_model_to_dirty(self):
        self._add_image()

    @profile_function()
    def apply_primary_beam(self, threshold):
        """Applies primary beam power to both model and dirty images."""
        for op in (self._apply_primary_beam_model, self._apply_primary_beam_dirty):
            op.threshold = threshold
            op()

    @profile_function()
    def dirty_to_psf(self):
        """Make the current dirty image the PSF by exchanging the two buffers."""
        dirty = self.buffer('dirty')
        psf = self.buffer('psf')
        self.bind(dirty=psf, psf=dirty)

    @profile_function()
    def psf_patch(self):
        params = self.template.clean_parameters
        return self._psf_patch(params.psf_cutoff, params.psf_limit)

    @profile_function()
    def noise_est(self):
        return self._noise_est()

    # -------------------------------------------------------------------------- CLEAN
    @profile_function()
    def clean_reset(self):
        self._clean.reset()

    def _record_component(self, pos, pixel):
        if pos in self._model_components:
            self._model_components[pos] = self._model_components[pos] + pixel
        else:
            self._model_components[pos] = pixel

    @profile_function()
    def clean_cycle(self, psf_patch, threshold=0.0):
        peak_value, peak_pos, model_pixel = self._clean(psf_patch, threshold)
        if peak_pos is not None:
            self._record_component(peak_pos, model_pixel)
        return peak_value

    @profile_function()
    def clean_cycles(self, psf_patch, threshold=0.0, max_cycles=1):
        """Run up to `max_cycles` minor cycles without leaving the device.

        Equivalent to calling :meth:`clean_cycle` until it returns ``None`` or
        `max_cycles` calls have been made; returns ``(peak values, stopped)``.
        """
        components, stopped = self._clean.run_cycles(psf_patch, threshold, max_cycles)
        for record in components:
            self._record_component((int(record['pos'][0]), int(record['pos'][1])),
                                   np.array(record['pixel']))
        return components['value'], stopped

    # ------------------------------------------------------------------------ buffers
    @profile_function(labels=['name'])
    def get_buffer(self, name):
        """Get the contents of a buffer as a numpy array."""
        return self.buffer(name).get(self.command_queue)

    @profile_function(labels=['name'])
    def set_buffer(self, name, data):
        """Copy a numpy array to a buffer (blocking)."""
        self.buffer(name).set(self.command_queue, data)

    def free_buffer(self, name):
        """Release the device memory of a buffer that is no longer needed."""
        if name in self.slots:
            self.slots[name].bind(None)


class ImagingPipeline:
    """Several :class:`Imaging` instances of one template, each on its own command queue,
    handed out round-robin so that the host<->device copies of one channel overlap the
    kernels of another (the reference's frontend images channels strictly one after the
    other on one queue, reference frontend.py:583-692).

    Typical use, one channel per turn::

        slot, imager = pipeline.acquire()      # waits until that instance is idle
        ... imager.set_coordinates / grid / grid_to_image ...
        imager.buffer('dirty').get_async(imager.command_queue, host_image[slot])
        pipeline.release(slot)                 # marks the end of this channel's work
        ...
        pipeline.wait(slot)                    # host_image[slot] is now valid

    Everything an instance does stays ordered on its own queue; the only library state
    shared between queues is read-only (kernel tables) or keyed by stream (the gridder's
    staging scratch).
    """

    def __init__(self, template, depth, *args, **kwargs):
        if depth < 1:
            raise ValueError('depth must be at least 1')
        self.queues = [template.context.create_command_queue() for _ in range(depth)]
        self.imagers = [template.instantiate(queue, *args, **kwargs) for queue in self.queues]
        for imager in self.imagers:
            imager.ensure_all_bound()
        self._done = [None] * depth
        self._next = 0

    def __len__(self):
        return len(self.imagers)

    def acquire(self):
        """Next instance in turn, after its previously released work has completed."""
        slot = self._next
        self._next = (slot + 1) % len(self.imagers)
        self.wait(slot)
        return slot, self.imagers[slot]

    def release(self, slot):
        """Mark the end of the work enqueued on `slot` since :meth:`acquire`."""
        self._done[slot] = self.queues[slot].enqueue_marker()
        return self._done[slot]

    def wait(self, slot):
        if self._done[slot] is not None:
            self._done[slot].wait()
            self._done[slot] = None

    def finish(self):
        for slot in range(len(self.imagers)):
            self.wait(slot)
        for queue in self.queues:
            queue.finish()
