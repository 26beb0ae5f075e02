"""Image-domain operations and the grid <-> image conversion on B200.

Same surface as the reference's :mod:`katsdpimager.image` (reference image.py:15-740):
``LayerToImage`` / ``ImageToLayer``, ``Scale``, ``AddImage``, ``ApplyPrimaryBeam`` and the
compound ``GridToImage`` / ``ImageToGrid`` built by ``GridImageTemplate``.  The 2-D
transform itself is cuFFT (:mod:`.fft`); everything around it runs in the fused
kernels of csrc/kib_image.cu:

* grid -> layer: zero fill, zero padding and ifftshift of one polarization plane in a
  single pass (the reference issues a memset and four strided copies, image.py:659-671);
* layer -> image: fftshift, W-term rotation, n-term and taper division, accumulation.
"""
import numpy as np

from . import _lib, accel, fft
from .profiling import profile_device
from .types import real_to_complex


class _LayerImageTemplate:
    """Base of :class:`LayerToImageTemplate` and :class:`ImageToLayerTemplate`
    (reference image.py:15-86 documents the mathematics)."""

    def __init__(self, context, real_dtype, tuning=None):
        _lib.load()
        self.context = context
        self.real_dtype = np.dtype(real_dtype)


class _LayerImage(accel.Operation):
    """Conversion between a corner-centred complex *layer* and one polarization
    plane of a centred real *image* (reference image.py:89-180).

    .. rubric:: Slots

    **layer** : complex, height x width
    **image** : real, polarizations x height x width
    **kernel1d** : real, width -- image-plane taper of the gridding kernel
    """

    _entry = None

    def __init__(self, template, command_queue, shape, lm_scale, lm_bias, allocator=None):
        if len(shape) != 3 or shape[-1] != shape[-2]:
            raise ValueError('shape must be square, not {}'.format(shape))
        if shape[-1] % 2 != 0:
            raise ValueError('image size must be even, not {}'.format(shape[-1]))
        super().__init__(command_queue, allocator)
        self.template = template
        dims = [accel.Dimension(x) for x in shape]
        self.slots['layer'] = accel.IOSlot(dims[-2:], real_to_complex(template.real_dtype))
        self.slots['image'] = accel.IOSlot(dims, template.real_dtype)
        self.slots['kernel1d'] = accel.IOSlot((shape[-1],), template.real_dtype)
        self.lm_scale = lm_scale
        self.lm_bias = lm_bias
        self.w = 0
        self.polarization = 0

    def set_w(self, w):
        """Set the W coordinate of the layer (wavelengths)."""
        self.w = w

    def set_polarization(self, polarization):
        if polarization < 0 or polarization >= self.slots['image'].shape[0]:
            raise IndexError('polarization index out of range')
        self.polarization = polarization

    def _run(self):
        layer = self.buffer('layer')
        image = self.buffer('image')
        kernel1d = self.buffer('kernel1d')
        itemsize = image.dtype.itemsize
        plane = image.padded_shape[1] * image.padded_shape[2]
        image_ptr = (image.ptr.value or 0) + self.polarization * plane * itemsize
        size = image.shape[-1]
        with profile_device(self.command_queue, self._entry[4:]):
            if self._entry == 'kib_layer_to_image':
                _lib.call(self._entry, image_ptr, image.padded_shape[2],
                          layer.ptr, layer.padded_shape[1], size, kernel1d.ptr,
                          float(self.lm_scale), float(self.lm_bias), float(self.w),
                          _lib.dtype_code(image.dtype), self.command_queue.stream)
            else:
                _lib.call(self._entry, layer.ptr, layer.padded_shape[1],
                          image_ptr, image.padded_shape[2], size, kernel1d.ptr,
                          float(self.lm_scale), float(self.lm_bias), float(self.w),
                          _lib.dtype_code(image.dtype), self.command_queue.stream)


class LayerToImageTemplate(_LayerImageTemplate):
    def instantiate(self, *args, **kwargs):
        return LayerToImage(self, *args, **kwargs)


class LayerToImage(_LayerImage):
    """image[pol] += Re(fftshift(layer) * exp(2 pi i w (n - 1))) * n / taper (accumulates)."""
    _entry = 'kib_layer_to_image'


class ImageToLayerTemplate(_LayerImageTemplate):
    def instantiate(self, *args, **kwargs):
        return ImageToLayer(self, *args, **kwargs)


class ImageToLayer(_LayerImage):
    """layer = ifftshift(image[pol] / (taper * n) * exp(-2 pi i w (n - 1)))."""
    _entry = 'kib_image_to_layer'


class _ImageOpTemplate:
    def __init__(self, context, dtype, num_polarizations, tuning=None):
        _lib.load()
        self.context = context
        self.dtype = np.dtype(dtype)
        self.num_polarizations = num_polarizations


def _check_image_shape(template, shape):
    if len(shape) != 3:
        raise ValueError('Wrong number of dimensions in shape')
    if shape[0] != template.num_polarizations:
        raise ValueError('Mismatch in number of polarizations')


class ScaleTemplate(_ImageOpTemplate):
    """Scale an image by a constant per polarization (reference image.py:281-367)."""

    def instantiate(self, *args, **kwargs):
        return Scale(self, *args, **kwargs)


class Scale(accel.Operation):
    """.. rubric:: Slots

    **data** : real, polarizations x height x width, scaled in place
    """

    def __init__(self, template, command_queue, shape, allocator=None):
        super().__init__(command_queue, allocator)
        _check_image_shape(template, shape)
        self.template = template
        self.slots['data'] = accel.IOSlot(shape, template.dtype)
        self.scale_factor = np.zeros((shape[0],), template.dtype)

    def set_scale_factor(self, scale_factor):
        self.scale_factor[:] = scale_factor

    def _run(self):
        data = self.buffer('data')
        factors = np.ascontiguousarray(self.scale_factor, dtype=np.float64)
        with profile_device(self.command_queue, 'scale'):
            _lib.call('kib_scale', data.ptr, data.padded_shape[2],
                      data.padded_shape[1] * data.padded_shape[2],
                      data.shape[2], data.shape[1], data.shape[0],
                      factors.ctypes.data_as(_lib.POINTER(_lib.c_double)),
                      _lib.dtype_code(data.dtype), self.command_queue.stream)


class AddImageTemplate(_ImageOpTemplate):
    """Add one image to another (reference image.py:370-458)."""

    def instantiate(self, *args, **kwargs):
        return AddImage(self, *args, **kwargs)


class AddImage(accel.Operation):
    """.. rubric:: Slots

    **src** : real image to add
    **dest** : real image updated in place (may be padded differently from **src**)
    """

    def __init__(self, template, command_queue, shape, allocator=None):
        super().__init__(command_queue, allocator)
        _check_image_shape(template, shape)
        self.template = template
        self.slots['src'] = accel.IOSlot(shape, template.dtype)
        self.slots['dest'] = accel.IOSlot(shape, template.dtype)

    def _run(self):
        src = self.buffer('src')
        dest = self.buffer('dest')
        with profile_device(self.command_queue, 'add_image'):
            _lib.call('kib_add_image',
                      dest.ptr, dest.padded_shape[2], dest.padded_shape[1] * dest.padded_shape[2],
                      src.ptr, src.padded_shape[2], src.padded_shape[1] * src.padded_shape[2],
                      src.shape[2], src.shape[1], src.shape[0],
                      _lib.dtype_code(dest.dtype), self.command_queue.stream)


class ApplyPrimaryBeamTemplate(_ImageOpTemplate):
    """Divide an image by a polarization-independent primary beam
    (reference image.py:461-558)."""

    def instantiate(self, *args, **kwargs):
        return ApplyPrimaryBeam(self, *args, **kwargs)


class ApplyPrimaryBeam(accel.Operation):
    """.. rubric:: Slots

    **data** : real image, divided in place; pixels where the beam power is below
    `threshold` are set to `replacement`
    **beam_power** : real, height x width
    """

    def __init__(self, template, command_queue, shape, threshold, replacement, allocator=None):
        super().__init__(command_queue, allocator)
        _check_image_shape(template, shape)
        self.template = template
        y_dim = accel.Dimension(shape[1])
        x_dim = accel.Dimension(shape[2])
        self.slots['data'] = accel.IOSlot((shape[0], y_dim, x_dim), template.dtype)
        self.slots['beam_power'] = accel.IOSlot((y_dim, x_dim), template.dtype)
        self.threshold = threshold
        self.replacement = replacement

    def _run(self):
        data = self.buffer('data')
        beam_power = self.buffer('beam_power')
        assert beam_power.padded_shape == data.padded_shape[1:]
        with profile_device(self.command_queue, 'apply_primary_beam'):
            _lib.call('kib_apply_primary_beam', data.ptr, data.padded_shape[2],
                      data.padded_shape[1] * data.padded_shape[2], beam_power.ptr,
                      data.shape[2], data.shape[1], data.shape[0],
                      float(self.threshold), float(self.replacement),
                      _lib.dtype_code(data.dtype), self.command_queue.stream)


class GridImageTemplate:
    """Conversion between a centred complex UV grid and a centred real image via an
    unnormalised complex 2-D FFT (reference image.py:561-606)."""

    def __init__(self, context, real_dtype):
        self.context = context
        self.real_dtype = np.dtype(real_dtype)
        self.layer_to_image = LayerToImageTemplate(context, real_dtype)
        self.image_to_layer = ImageToLayerTemplate(context, real_dtype)

    def make_fft_plan(self, shape_layer, padded_shape_layer):
        complex_dtype = real_to_complex(self.real_dtype)
        return fft.FftTemplate(self.context, 2, shape_layer, complex_dtype, complex_dtype,
                               padded_shape_layer, padded_shape_layer)

    def instantiate_grid_to_image(self, *args, **kwargs):
        return GridToImage(self, *args, **kwargs)

    def instantiate_image_to_grid(self, *args, **kwargs):
        return ImageToGrid(self, *args, **kwargs)


def _check_grid(grid, layer):
    polarizations, height, width = grid.shape
    if height % 2 or width % 2 or height != width:
        raise ValueError('grid must be square with even size, not {}'.format(grid.shape))
    if width > layer.shape[-1]:
        raise ValueError('grid is larger than the image')
    return polarizations, width


def occupancy_words(grid_size):
    """Length (uint32 words) of the column-occupancy mask of a grid (kib_column_occupancy)."""
    return (grid_size // 8 + 31) // 32 + 1


def column_occupancy(command_queue, uv, num_vis, kernel_width, grid_size, stride_bytes=8,
                     out=None):
    """Mark the groups of 8 grid columns that the footprints of `num_vis` visibilities cover.

    `uv` is a device pointer (or DeviceArray) to the first int16 u coordinate, consecutive
    visibilities `stride_bytes` apart.  ORs into `out` (a zeroed uint32 DeviceArray of
    :func:`occupancy_words` words is made if None) and returns it."""
    if out is None:
        out = accel.DeviceArray(command_queue.context, (occupancy_words(grid_size),), np.uint32)
        out.zero(command_queue)
    ptr = uv.ptr if hasattr(uv, 'ptr') else uv
    _lib.call('kib_column_occupancy', ptr, int(stride_bytes), int(num_vis), int(kernel_width),
              int(grid_size), out.ptr, command_queue.stream)
    return out


def _check_occupancy(occupancy, grid_size):
    if occupancy is None:
        return None
    if occupancy.dtype != np.uint32 or occupancy.shape[0] < occupancy_words(grid_size):
        raise ValueError('occupancy must hold {} uint32 words'.format(occupancy_words(grid_size)))
    return occupancy


class GridToImage(accel.OperationSequence):
    """grid -> image for every polarization (reference image.py:609-673).

    .. rubric:: Slots

    **grid** : complex, polarizations x G x G;  **layer** : complex N x N scratch;
    **image** : real, polarizations x N x N (accumulated into);  **kernel1d** : real N
    """

    def __init__(self, template, command_queue, shape_grid,
                 lm_scale, lm_bias, fft_plan, allocator=None):
        self._ifft = fft_plan.instantiate(command_queue, fft.FftMode.INVERSE, allocator)
        shape_image = (shape_grid[0],) + tuple(fft_plan.shape)
        self._layer_to_image = template.layer_to_image.instantiate(
            command_queue, shape_image, lm_scale, lm_bias, allocator)
        operations = [('ifft', self._ifft), ('layer_to_image', self._layer_to_image)]
        compounds = {
            'layer': ['ifft:src', 'ifft:dest', 'layer_to_image:layer'],
            'image': ['layer_to_image:image'],
            'kernel1d': ['layer_to_image:kernel1d']
        }
        super().__init__(command_queue, operations, compounds, allocator=allocator)
        self.slots['grid'] = accel.IOSlot(shape_grid, fft_plan.dtype_src)
        #: use the fused pruned transform (kib_grid_to_image) when the library supports
        #: the size (single precision, power-of-two images of 2048..16384 pixels)
        self.fused = True
        #: column occupancy of the grid (:func:`column_occupancy`), or None.  When set, the
        #: caller promises that the grid is zero outside the occupied column groups; the fused
        #: transform then skips them (bit-identical image).
        self.occupancy = None
        #: Keep the per-pixel factor plane (W rotation, n, taper: N x N complex, 0.5 GB at
        #: 8192^2) of up to this many values of w: the W slices of a channel are transformed
        #: once per pass (PSF, every major cycle) with the same factors, which the first
        #: polarization of the first pass then computes and stores for all later launches.
        #: 0 = one plane, recomputed by the first polarization of every call (round 1).
        self.factor_cache_planes = 0
        #: The caller's promise that lm_bias = -N / 2 * lm_scale and that kernel1d is symmetric
        #: (kernel1d[i] == kernel1d[N - i]), as for the reference's taper: the factor plane is
        #: then symmetric about the image centre and only one quadrant is kept (a quarter of
        #: the bytes every launch has to read).
        self.symmetric_factors = False
        self._factors = None
        self._factor_cache = {}
        self._fold = None

    def set_w(self, w):
        self._layer_to_image.set_w(w)

    @property
    def uses_occupancy(self):
        """Whether :attr:`occupancy` is honoured (the fused transform is available for the bound
        shapes).  If not, every column of the grid is read and must be cleared."""
        grid = self.buffer('grid')
        layer = self.buffer('layer')
        if not self.fused or grid is None or layer is None:
            return False
        return bool(_lib.grid_to_image_supported(layer.shape[1], grid.shape[-1], grid.dtype))

    def clear_factor_cache(self):
        """Forget the cached factor planes (a new channel: new pixel size, new slice centres).
        The buffers are kept for reuse."""
        self._factor_pool = getattr(self, '_factor_pool', []) + list(self._factor_cache.values())
        self._factor_cache = {}

    def _factor_plane(self, n, dtype, key):
        """(plane, already computed?) for the factors identified by `key`; `n` is the side of
        the table (the image's, or half of it + 1 for symmetric factors)."""
        context = self.command_queue.context
        if self.factor_cache_planes <= 0:
            if self._factors is None or self._factors.shape != (n, n):
                self._factors = accel.DeviceArray(context, (n, n), dtype)
            return self._factors, False
        plane = self._factor_cache.get(key)
        if plane is not None:
            return plane, True
        pool = getattr(self, '_factor_pool', [])
        while len(self._factor_cache) >= self.factor_cache_planes:
            pool.append(self._factor_cache.pop(next(iter(self._factor_cache))))   # oldest first
        plane = None
        while pool and plane is None:
            candidate = pool.pop()
            if candidate.shape == (n, n) and candidate.dtype == dtype:
                plane = candidate
        self._factor_pool = pool
        if plane is None:
            plane = accel.DeviceArray(context, (n, n), dtype)
        self._factor_cache[key] = plane
        return plane, False

    def _run_fused(self, grid, layer, polarizations, size, plane_bytes):
        """Pruned transform with the layer_to_image arithmetic fused in
        (csrc/kib_gridfft.cu); the layer buffer only holds the half-transformed columns."""
        op = self._layer_to_image
        image = self.buffer('image')
        kernel1d = self.buffer('kernel1d')
        image_plane = image.padded_shape[1] * image.padded_shape[2] * image.dtype.itemsize
        dtype = _lib.dtype_code(grid.dtype)
        stream = self.command_queue.stream
        n = layer.shape[1]
        factors = None
        have_factors = False
        symmetric = False
        if polarizations > 1 or self.factor_cache_planes > 0:
            # the W rotation / n / taper factor of every pixel is computed for the first
            # polarization, kept, and reused by the others (and by later passes over the same
            # W slice when the cache is on)
            symmetric = bool(self.symmetric_factors) and abs(
                float(op.lm_bias) + 0.5 * n * float(op.lm_scale)) <= 1e-6 * abs(float(op.lm_scale))
            side = n // 2 + 1 if symmetric else n
            key = (float(op.w), float(op.lm_scale), float(op.lm_bias), kernel1d.ptr.value, side)
            plane, have_factors = self._factor_plane(side, grid.dtype, key)
            factors = plane.ptr
        fold_bytes = _lib.grid_to_image_fold_bytes(n, size)
        if self._fold is None or self._fold.shape[0] < fold_bytes:
            self._fold = accel.DeviceArray(self.command_queue.context, (fold_bytes,), np.uint8)
        occ = _check_occupancy(self.occupancy, size)
        presence = None
        if occ is not None:
            # the row pass's view of the mask, made once per mask and image size
            # (`generation` counts the times the owner refilled the mask in place)
            tables = occ.__dict__.setdefault('_row_presence', {})
            generation = getattr(occ, 'generation', 0)
            presence, made_for = tables.get((n, size), (None, None))
            if presence is None:
                presence = accel.DeviceArray(self.command_queue.context, (n // 16,), np.uint16)
            if made_for != generation:
                _lib.call('kib_row_presence', occ.ptr, size, n, presence.ptr, stream)
                tables[(n, size)] = (presence, generation)
        for pol in range(polarizations):
            mode = 0 if factors is None else (1 if pol == 0 and not have_factors else 2)
            if mode and symmetric:
                mode += 2                   # quadrant table
            grid_plane = (grid.ptr.value or 0) + pol * plane_bytes
            image_ptr = (image.ptr.value or 0) + pol * image_plane
            if occ is not None:
                with profile_device(self.command_queue, 'grid_to_image_columns'):
                    _lib.call('kib_grid_to_image_columns_occ', layer.ptr, layer.padded_shape[1],
                              layer.shape[1], grid_plane, grid.padded_shape[2], size,
                              self._fold.ptr, occ.ptr, dtype, stream)
                with profile_device(self.command_queue, 'grid_to_image_rows'):
                    _lib.call('kib_grid_to_image_rows_occ', image_ptr, image.padded_shape[2],
                              layer.ptr, layer.padded_shape[1], size, layer.shape[1],
                              kernel1d.ptr, float(op.lm_scale), float(op.lm_bias), float(op.w),
                              factors, mode, presence.ptr, dtype, stream)
                continue
            with profile_device(self.command_queue, 'grid_to_image_columns'):
                _lib.call('kib_grid_to_image_columns', layer.ptr, layer.padded_shape[1],
                          layer.shape[1], grid_plane, grid.padded_shape[2], size,
                          self._fold.ptr, dtype, stream)
            with profile_device(self.command_queue, 'grid_to_image_rows'):
                _lib.call('kib_grid_to_image_rows', image_ptr, image.padded_shape[2],
                          layer.ptr, layer.padded_shape[1], size, layer.shape[1], kernel1d.ptr,
                          float(op.lm_scale), float(op.lm_bias), float(op.w),
                          factors, mode, dtype, stream)

    def _run(self):
        grid = self.buffer('grid')
        layer = self.buffer('layer')
        polarizations, size = _check_grid(grid, layer)
        plane_bytes = grid.padded_shape[1] * grid.padded_shape[2] * grid.dtype.itemsize
        if self.fused and _lib.grid_to_image_supported(layer.shape[1], size, grid.dtype):
            self._run_fused(grid, layer, polarizations, size, plane_bytes)
            return
        for pol in range(polarizations):
            with profile_device(self.command_queue, 'grid_to_layer'):
                _lib.call('kib_grid_to_layer', layer.ptr, layer.padded_shape[1], layer.shape[1],
                          (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2], size,
                          _lib.dtype_code(grid.dtype), self.command_queue.stream)
            self._layer_to_image.set_polarization(pol)
            super()._run()


class ImageToGrid(accel.OperationSequence):
    """image -> grid for every polarization (reference image.py:676-740); slots as
    :class:`GridToImage` with **grid** as the output."""

    def __init__(self, template, command_queue, shape_grid,
                 lm_scale, lm_bias, fft_plan, allocator=None):
        shape_image = (shape_grid[0],) + tuple(fft_plan.shape)
        self._fft = fft_plan.instantiate(command_queue, fft.FftMode.FORWARD, allocator)
        self._image_to_layer = template.image_to_layer.instantiate(
            command_queue, shape_image, lm_scale, lm_bias, allocator)
        operations = [('image_to_layer', self._image_to_layer), ('fft', self._fft)]
        compounds = {
            'layer': ['fft:src', 'fft:dest', 'image_to_layer:layer'],
            'image': ['image_to_layer:image'],
            'kernel1d': ['image_to_layer:kernel1d']
        }
        super().__init__(command_queue, operations, compounds, allocator=allocator)
        self.slots['grid'] = accel.IOSlot(shape_grid, fft_plan.dtype_dest)
        #: use the fused pruned transform when the library supports the size
        self.fused = True
        #: the image is a CLEAN model (zero except for a few hundred pixels): rows that are
        #: entirely zero are not transformed.  Always exact; only the speed depends on it
        #: (a dense image is transformed faster with False: shared per-pixel factors).
        self.sparse_model = True
        #: column occupancy (:func:`column_occupancy`) of the visibilities that will be
        #: degridded from the grid, or None.  When set, only the occupied column groups of the
        #: grid are computed; the others keep whatever they held.
        self.occupancy = None
        #: The caller's promise that the image (the CLEAN model) has not changed since the
        #: previous call of this operation: the sparse route then reuses its classification of
        #: the rows instead of reading the whole plane again (the model is transformed once per
        #: W slice between two batches of CLEAN cycles).
        self.image_unchanged = False
        self._classified = None
        self._factors = None
        self._fold = None
        self._row_info = None

    def set_w(self, w):
        self._image_to_layer.set_w(w)

    def _run_fused(self, grid, layer, polarizations, size, plane_bytes):
        """Mirror image of :meth:`GridToImage._run_fused` (csrc/kib_gridfft.cu)."""
        op = self._image_to_layer
        image = self.buffer('image')
        kernel1d = self.buffer('kernel1d')
        image_plane = image.padded_shape[1] * image.padded_shape[2] * image.dtype.itemsize
        dtype = _lib.dtype_code(grid.dtype)
        stream = self.command_queue.stream
        n = layer.shape[1]
        context = self.command_queue.context
        occ = _check_occupancy(self.occupancy, size)
        if self.sparse_model and _lib.load().kib_image_to_grid_sparse_supported(n, size, dtype):
            # only rows with a non-zero pixel are transformed; the column pass never reads the
            # others (taken as zero)
            words = 2 * n + 1
            if self._row_info is None or self._row_info.shape[0] < polarizations * words:
                self._row_info = accel.DeviceArray(context, (polarizations * words,), np.int32)
                self._classified = None
            state = (image.ptr.value, n, polarizations)
            reuse = bool(self.image_unchanged) and self._classified == state
            self._classified = state
            for pol in range(polarizations):
                row_info = (self._row_info.ptr.value or 0) + pol * words * 4
                with profile_device(self.command_queue, 'image_to_grid_rows'):
                    _lib.call('kib_image_to_grid_rows_classified' if reuse
                              else 'kib_image_to_grid_rows_sparse', layer.ptr,
                              layer.padded_shape[1],
                              size, n, (image.ptr.value or 0) + pol * image_plane,
                              image.padded_shape[2], kernel1d.ptr, float(op.lm_scale),
                              float(op.lm_bias), float(op.w), row_info, dtype, stream)
                with profile_device(self.command_queue, 'image_to_grid_columns'):
                    if occ is not None:
                        _lib.call('kib_image_to_grid_columns_occ',
                                  (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2],
                                  size, layer.ptr, layer.padded_shape[1], n, None,
                                  row_info, occ.ptr, dtype, stream)
                    else:
                        _lib.call('kib_image_to_grid_columns_sparse',
                                  (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2],
                                  size, layer.ptr, layer.padded_shape[1], n, row_info,
                                  dtype, stream)
            return
        factors = None
        if polarizations > 1 and not self.sparse_model:
            if self._factors is None or self._factors.shape != (n, n):
                self._factors = accel.DeviceArray(context, (n, n), grid.dtype)
            factors = self._factors.ptr
        fold_bytes = _lib.grid_to_image_fold_bytes(n, size)
        if self._fold is None or self._fold.shape[0] < fold_bytes:
            self._fold = accel.DeviceArray(context, (fold_bytes,), np.uint8)
        for pol in range(polarizations):
            if self.sparse_model:
                mode = 3        # factor on the fly, all-zero image rows skipped
            else:
                mode = 0 if factors is None else (1 if pol == 0 else 2)
            with profile_device(self.command_queue, 'image_to_grid_rows'):
                _lib.call('kib_image_to_grid_rows', layer.ptr, layer.padded_shape[1], size, n,
                          (image.ptr.value or 0) + pol * image_plane, image.padded_shape[2],
                          kernel1d.ptr, float(op.lm_scale), float(op.lm_bias), float(op.w),
                          factors, mode, dtype, stream)
            with profile_device(self.command_queue, 'image_to_grid_columns'):
                if occ is not None:
                    _lib.call('kib_image_to_grid_columns_occ',
                              (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2],
                              size, layer.ptr, layer.padded_shape[1], n, self._fold.ptr, None,
                              occ.ptr, dtype, stream)
                else:
                    _lib.call('kib_image_to_grid_columns',
                              (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2],
                              size, layer.ptr, layer.padded_shape[1], n, self._fold.ptr, dtype,
                              stream)

    def _run(self):
        grid = self.buffer('grid')
        layer = self.buffer('layer')
        polarizations, size = _check_grid(grid, layer)
        plane_bytes = grid.padded_shape[1] * grid.padded_shape[2] * grid.dtype.itemsize
        if self.fused and _lib.grid_to_image_supported(layer.shape[1], size, grid.dtype):
            self._run_fused(grid, layer, polarizations, size, plane_bytes)
            return
        for pol in range(polarizations):
            self._image_to_layer.set_polarization(pol)
            super()._run()
            with profile_device(self.command_queue, 'layer_to_grid'):
                _lib.call('kib_layer_to_grid',
                          (grid.ptr.value or 0) + pol * plane_bytes, grid.padded_shape[2], size,
                          layer.ptr, layer.padded_shape[1], layer.shape[1],
                          _lib.dtype_code(grid.dtype), self.command_queue.stream)
