"""CUDA-only operation runtime with the ``katsdpsigproc.accel`` surface.

The reference builds every imaging operation on katsdpsigproc (un-vendored,
``setup.py:43``): ``Operation`` objects expose named *slots* that are bound to
``DeviceArray`` buffers, ``OperationSequence`` glues operations by aliasing
slots, and ``Dimension`` objects negotiate padding between aliased slots
(SURVEY.md section 8b.2 lists the exact calls made by the reference).  This
module provides that surface over the C ABI of libkatimager_b200.so through
ctypes.  There is one backend (CUDA on the current device) and no kernel JIT:
operations call the ``kib_*`` launchers directly on their queue's stream.
"""
import ctypes
import weakref
from collections import OrderedDict

import numpy as np

from . import _lib


def divup(x, y):
    """Divide x by y and round the result upwards."""
    return (x + y - 1) // y


def roundup(x, y):
    """Round x up to the next multiple of y."""
    return divup(x, y) * y


# ------------------------------------------------------------------ context / queue
class Device:
    def __init__(self, index):
        self.index = index
        buf = ctypes.create_string_buffer(256)
        _lib.call('kib_device_name', index, buf, 256)
        self.name = buf.value.decode()
        self.simd_group_size = self._attr(5)
        self.num_sms = self._attr(0)
        self.compute_capability = self._attr(4)
        self.is_cuda = True
        self.is_gpu = True

    def _attr(self, attr):
        value = ctypes.c_int64()
        _lib.call('kib_device_attr', self.index, attr, ctypes.byref(value))
        return int(value.value)


class Event:
    """Marker in a command queue (katsdpsigproc.abc.AbstractEvent)."""

    def __init__(self, stream):
        handle = ctypes.c_void_p()
        _lib.call('kib_event_create', ctypes.byref(handle))
        self._handle = handle
        self._finalizer = weakref.finalize(self, _destroy_event, handle.value)
        _lib.call('kib_event_record', handle, stream)

    def wait(self):
        _lib.call('kib_event_sync', self._handle)

    def time_since(self, prior):
        """Seconds elapsed between `prior` and this event (both must have completed)."""
        ms = ctypes.c_float()
        _lib.call('kib_event_elapsed_ms', prior._handle, self._handle, ctypes.byref(ms))
        return ms.value * 1e-3

    def time_till(self, nxt):
        return nxt.time_since(self)


def _destroy_event(handle):
    try:
        _lib.load().kib_event_destroy(ctypes.c_void_p(handle))
    except Exception:       # interpreter shutdown
        pass


def _destroy_stream(handle):
    try:
        _lib.load().kib_stream_destroy(ctypes.c_void_p(handle))
    except Exception:
        pass


class CommandQueue:
    """In-order CUDA stream (katsdpsigproc.abc.AbstractCommandQueue)."""

    def __init__(self, context, profile=False):
        self.context = context
        handle = ctypes.c_void_p()
        _lib.call('kib_stream_create', ctypes.byref(handle))
        self.stream = handle
        self._finalizer = weakref.finalize(self, _destroy_stream, handle.value)

    def enqueue_marker(self):
        return Event(self.stream)

    def enqueue_wait_for_events(self, events):
        """Later work on this queue waits for `events` (of any queue) to complete."""
        for event in events:
            _lib.call('kib_stream_wait_event', self.stream, event._handle)

    def enqueue_zero_buffer(self, buffer):
        _lib.call('kib_memset_async', buffer.ptr, 0, buffer.nbytes, self.stream)

    def flush(self):
        pass

    def finish(self):
        _lib.call('kib_stream_sync', self.stream)


class Context:
    """The CUDA primary context of one device."""

    def __init__(self, device_index=0):
        _lib.load()
        count = ctypes.c_int()
        _lib.call('kib_device_count', ctypes.byref(count))
        if count.value <= 0:
            raise RuntimeError('no CUDA device available')
        if device_index >= count.value:
            raise RuntimeError('CUDA device {} requested but only {} present'.format(
                device_index, count.value))
        _lib.call('kib_set_device', device_index)
        self.device = Device(device_index)

    def make_current(self):
        _lib.call('kib_set_device', self.device.index)

    def __enter__(self):
        self.make_current()
        return self

    def __exit__(self, *exc):
        return False

    def create_command_queue(self, profile=False):
        return CommandQueue(self, profile)

    def create_tuning_command_queue(self):
        return CommandQueue(self, True)

    def allocate_raw(self, n_bytes):
        return RawBuffer(n_bytes)

    def allocate_pinned(self, shape, dtype):
        return HostArray(shape, dtype, context=self)

    def mem_info(self):
        free = ctypes.c_size_t()
        total = ctypes.c_size_t()
        _lib.call('kib_mem_info', ctypes.byref(free), ctypes.byref(total))
        return free.value, total.value


def create_some_context(interactive=False, device_filter=None):
    """Context on the device selected by ``KIB_DEVICE`` / ``LOCAL_RANK`` (default 0)."""
    import os
    index = int(os.environ.get('KIB_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    return Context(index)


# ------------------------------------------------------------------------- memory
def _free_device(ptr):
    try:
        _lib.load().kib_free(ctypes.c_void_p(ptr))
    except Exception:
        pass


def _free_host(ptr):
    try:
        _lib.load().kib_host_free(ctypes.c_void_p(ptr))
    except Exception:
        pass


class RawBuffer:
    """An untyped device allocation."""

    def __init__(self, n_bytes):
        handle = ctypes.c_void_p()
        _lib.call('kib_malloc', ctypes.byref(handle), int(n_bytes))
        self.ptr = handle
        self.nbytes = int(n_bytes)
        self._finalizer = weakref.finalize(self, _free_device, handle.value)

    def __int__(self):
        return self.ptr.value or 0


class HostArray(np.ndarray):
    """numpy array in pinned host memory, optionally a view of a larger padded
    allocation so that it can be copied to a :class:`DeviceArray` with the same
    padding in one transfer (katsdpsigproc.accel.HostArray)."""

    def __new__(cls, shape, dtype, padded_shape=None, context=None):
        shape = tuple(int(x) for x in np.atleast_1d(shape)) if np.ndim(shape) else (int(shape),)
        if padded_shape is None:
            padded_shape = shape
        padded_shape = tuple(int(x) for x in padded_shape)
        assert len(shape) == len(padded_shape)
        assert all(p >= s for s, p in zip(shape, padded_shape))
        dtype = np.dtype(dtype)
        n_bytes = int(np.prod(padded_shape, dtype=np.int64)) * dtype.itemsize
        handle = ctypes.c_void_p()
        _lib.call('kib_host_alloc', ctypes.byref(handle), max(n_bytes, 1))
        raw = (ctypes.c_char * max(n_bytes, 1)).from_address(handle.value)
        # Every numpy view keeps `raw` alive through its base chain, and `raw` keeps the
        # pinned allocation alive.
        raw._owner = _PinnedOwner(handle.value)
        base = np.frombuffer(raw, dtype=np.uint8, count=n_bytes).view(dtype).reshape(padded_shape)
        obj = base[tuple(slice(0, s) for s in shape)].view(cls)
        obj._padded = base
        obj.padded_shape = padded_shape
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            # views lose the association with the padded storage
            self._padded = None
            self.padded_shape = None

    @classmethod
    def safe(cls, obj):
        return isinstance(obj, cls) and getattr(obj, '_padded', None) is not None

    def padded_view(self):
        return self._padded


def is_pinned(array):
    """True if `array` is (a view of) a :class:`HostArray`, i.e. lives in page-locked memory
    that the copy engines can read without staging."""
    while array is not None:
        if isinstance(array, HostArray):
            return True
        if isinstance(array, ctypes.Array):
            return hasattr(array, '_owner')
        array = getattr(array, 'base', None)
    return False


class _PinnedOwner:
    def __init__(self, ptr):
        self._finalizer = weakref.finalize(self, _free_host, ptr)


def _normalise_region(shape, region):
    """Turn a numpy basic index into per-axis (start, size, keep) triples."""
    if not isinstance(region, tuple):
        region = (region,)
    if any(r is Ellipsis for r in region):
        pos = [i for i, r in enumerate(region) if r is Ellipsis]
        if len(pos) > 1:
            raise IndexError('only one ellipsis allowed')
        fill = len(shape) - (len(region) - 1)
        region = region[:pos[0]] + (slice(None),) * fill + region[pos[0] + 1:]
    if len(region) > len(shape):
        raise IndexError('too many indices')
    region = region + (slice(None),) * (len(shape) - len(region))
    out = []
    for r, n in zip(region, shape):
        if isinstance(r, slice):
            start, stop, step = r.indices(n)
            if step != 1:
                raise IndexError('only unit-stride slices are supported')
            out.append((start, max(0, stop - start), True))
        else:
            r = int(r)
            if r < 0:
                r += n
            if not 0 <= r < n:
                raise IndexError('index out of range')
            out.append((r, 1, False))
    return out


def _layout(padded_shape, itemsize, axes):
    """(byte offset, [(size, byte stride)] of kept axes) of a region of a C-ordered array."""
    strides = []
    s = itemsize
    for n in reversed(padded_shape):
        strides.append(s)
        s *= n
    strides.reverse()
    offset = 0
    dims = []
    for (start, size, keep), stride in zip(axes, strides):
        offset += start * stride
        if keep:
            dims.append((size, stride))
    return offset, dims


def _strided_copy(queue, dst_ptr, dst_dims, src_ptr, src_dims, itemsize, kind):
    """Copy between two strided regions, each described outermost-first as
    [(size, byte stride)], using as few pitched copies as possible."""
    if [d[0] for d in dst_dims] != [d[0] for d in src_dims]:
        raise ValueError('region shapes do not match: {} vs {}'.format(
            [d[0] for d in dst_dims], [d[0] for d in src_dims]))
    dims = [(n, ds, ss) for (n, ds), (_, ss) in zip(dst_dims, src_dims) if n != 1]
    if any(n == 0 for n, _, _ in dims):
        return
    # Fold innermost contiguous axes into the row width.
    width = itemsize
    while dims and dims[-1][1] == width and dims[-1][2] == width:
        width *= dims[-1][0]
        dims.pop()
    # Merge remaining axes whose strides nest exactly (innermost first).
    merged = []
    for n, ds, ss in reversed(dims):
        if merged and merged[-1][0] * merged[-1][1] == ds and merged[-1][0] * merged[-1][2] == ss:
            merged[-1] = (merged[-1][0] * n, merged[-1][1], merged[-1][2])
        else:
            merged.append((n, ds, ss))
    height, d_row, s_row = merged[0] if merged else (1, width, width)
    if len(merged) > 1:
        depth, d_plane, s_plane = merged[1]
    else:
        depth, d_plane, s_plane = 1, height * d_row, height * s_row
    outer = merged[2:]

    def recurse(level, d_off, s_off):
        if level < 0:
            _lib.call('kib_memcpy3d_async',
                      ctypes.c_void_p(dst_ptr + d_off), d_row, d_plane,
                      ctypes.c_void_p(src_ptr + s_off), s_row, s_plane,
                      width, height, depth, kind, queue.stream)
            return
        n, ds, ss = outer[level]
        for i in range(n):
            recurse(level - 1, d_off + i * ds, s_off + i * ss)

    recurse(len(outer) - 1, 0, 0)


class DeviceArray:
    """N-dimensional array in device memory with optional padding
    (katsdpsigproc.accel.DeviceArray)."""

    def __init__(self, context, shape, dtype, padded_shape=None, raw=None):
        self.context = context
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        if padded_shape is None:
            padded_shape = self.shape
        self.padded_shape = tuple(int(x) for x in padded_shape)
        assert len(self.shape) == len(self.padded_shape)
        assert all(p >= s for s, p in zip(self.shape, self.padded_shape))
        n_bytes = int(np.prod(self.padded_shape, dtype=np.int64)) * self.dtype.itemsize
        if raw is None:
            raw = RawBuffer(n_bytes)
        elif raw.nbytes < n_bytes:
            raise ValueError('raw buffer is too small')
        self.buffer = raw
        self.nbytes = n_bytes

    @property
    def ptr(self):
        return self.buffer.ptr

    @property
    def ndim(self):
        return len(self.shape)

    def _contiguous(self):
        return self.shape == self.padded_shape

    def empty_like(self):
        return HostArray(self.shape, self.dtype, self.padded_shape, context=self.context)

    def asarray_like(self, ary):
        if HostArray.safe(ary) and ary.shape == self.shape and ary.dtype == self.dtype \
                and ary.padded_shape == self.padded_shape:
            return ary
        tmp = self.empty_like()
        np.copyto(tmp, ary, casting='no')
        return tmp

    def _check_host(self, ary):
        if ary.shape != self.shape:
            raise ValueError('shape mismatch: host {} device {}'.format(ary.shape, self.shape))
        if ary.dtype != self.dtype:
            raise TypeError('dtype mismatch: host {} device {}'.format(ary.dtype, self.dtype))

    def set_async(self, command_queue, ary):
        ary = np.asarray(ary) if not isinstance(ary, np.ndarray) else ary
        self._check_host(ary)
        if HostArray.safe(ary) and ary.padded_shape == self.padded_shape:
            _lib.call('kib_memcpy_h2d_async', self.ptr, ary.padded_view().ctypes.data,
                      self.nbytes, command_queue.stream)
        else:
            whole = tuple(slice(None) for _ in self.shape)
            self.set_region(command_queue, ary, whole, whole, blocking=False)

    def set(self, command_queue, ary):
        self.set_async(command_queue, ary)
        command_queue.finish()

    def get_async(self, command_queue, ary=None):
        if ary is None:
            ary = self.empty_like()
        self._check_host(ary)
        if HostArray.safe(ary) and ary.padded_shape == self.padded_shape:
            _lib.call('kib_memcpy_d2h_async', ary.padded_view().ctypes.data, self.ptr,
                      self.nbytes, command_queue.stream)
        else:
            whole = tuple(slice(None) for _ in self.shape)
            self.get_region(command_queue, ary, whole, whole, blocking=False)
        return ary

    def get(self, command_queue, ary=None):
        ary = self.get_async(command_queue, ary)
        command_queue.finish()
        return ary

    def zero(self, command_queue):
        _lib.call('kib_memset_async', self.ptr, 0, self.nbytes, command_queue.stream)

    def _region(self, region):
        axes = _normalise_region(self.shape, region)
        return _layout(self.padded_shape, self.dtype.itemsize, axes)

    def set_region(self, command_queue, ary, device_region, ary_region, blocking=True):
        """Copy ``ary[ary_region]`` (host) to ``self[device_region]``."""
        view = np.asarray(ary)[ary_region]
        if view.dtype != self.dtype:
            raise TypeError('dtype mismatch: host {} device {}'.format(view.dtype, self.dtype))
        offset, dims = self._region(device_region)
        src_dims = list(zip(view.shape, view.strides))
        if any(s < 0 for _, s in src_dims):
            raise ValueError('negative strides are not supported')
        _strided_copy(command_queue, (self.ptr.value or 0) + offset, dims,
                      view.ctypes.data, src_dims, self.dtype.itemsize, 0)
        if blocking:
            command_queue.finish()

    def get_region(self, command_queue, ary, device_region, ary_region, blocking=True):
        """Copy ``self[device_region]`` to ``ary[ary_region]`` (host)."""
        view = ary[ary_region]
        if view.dtype != self.dtype:
            raise TypeError('dtype mismatch: host {} device {}'.format(view.dtype, self.dtype))
        offset, dims = self._region(device_region)
        dst_dims = list(zip(view.shape, view.strides))
        _strided_copy(command_queue, view.ctypes.data, dst_dims,
                      (self.ptr.value or 0) + offset, dims, self.dtype.itemsize, 1)
        if blocking:
            command_queue.finish()

    def copy_region(self, command_queue, dest, src_region, dest_region):
        """Device-to-device copy of ``self[src_region]`` to ``dest[dest_region]``."""
        if dest.dtype != self.dtype:
            raise TypeError('dtype mismatch')
        s_off, s_dims = self._region(src_region)
        d_off, d_dims = dest._region(dest_region)
        _strided_copy(command_queue, (dest.ptr.value or 0) + d_off, d_dims,
                      (self.ptr.value or 0) + s_off, s_dims, self.dtype.itemsize, 2)


class DeviceAllocator:
    """Allocates :class:`DeviceArray` objects from a context."""

    def __init__(self, context):
        self.context = context

    def allocate(self, shape, dtype, padded_shape=None, raw=None):
        return DeviceArray(self.context, shape, dtype, padded_shape, raw)

    def allocate_raw(self, n_bytes):
        return RawBuffer(n_bytes)


AbstractAllocator = DeviceAllocator


# ------------------------------------------------------------------ slots
class Dimension:
    """Size and padding requirements of one axis of a slot.  Linked dimensions
    (``link``) share their requirements, so aliased slots agree on padding."""

    def __init__(self, size, alignment=1, min_padded_round=None, min_padded_size=None,
                 align_dtype=None, exact=False):
        if min_padded_size is None:
            min_padded_size = size
        if min_padded_round is not None:
            min_padded_size = max(min_padded_size, roundup(size, min_padded_round))
        if align_dtype is not None:
            alignment = max(1, alignment // np.dtype(align_dtype).itemsize)
        self._root = self
        self._size = int(size)
        self._alignment = int(alignment)
        self._min_padded_size = int(min_padded_size)
        self._exact = bool(exact)
        self._frozen = False

    def _find(self):
        node = self
        while node._root is not node:
            node._root = node._root._root
            node = node._root
        return node

    @property
    def size(self):
        return self._find()._size

    @property
    def exact(self):
        return self._find()._exact

    @property
    def alignment(self):
        return self._find()._alignment

    def required_padded_size(self):
        root = self._find()
        root._frozen = True
        if root._exact:
            return root._size
        return roundup(root._min_padded_size, root._alignment)

    def valid(self, padded_size):
        root = self._find()
        if root._exact:
            return padded_size == root._size
        return padded_size >= root._min_padded_size and padded_size % root._alignment == 0

    def link(self, other):
        a, b = self._find(), other._find()
        if a is b:
            return
        if a._size != b._size:
            raise ValueError('linked dimensions must have the same size ({} vs {})'.format(
                a._size, b._size))
        if a._frozen or b._frozen:
            raise ValueError('cannot link dimensions after their padding has been queried')
        exact = a._exact or b._exact
        alignment = int(np.lcm(a._alignment, b._alignment))
        min_padded = max(a._min_padded_size, b._min_padded_size)
        if exact and (min_padded > a._size or a._size % alignment != 0):
            raise ValueError('exact dimension of size {} cannot satisfy padding requirements'.format(
                a._size))
        b._root = a
        a._exact = exact
        a._alignment = alignment
        a._min_padded_size = min_padded


class IOSlotBase:
    def __init__(self):
        self.buffer = None

    def is_bound(self):
        return self.buffer is not None


class IOSlot(IOSlotBase):
    """A named array argument of an operation."""

    def __init__(self, dimensions, dtype):
        super().__init__()
        self.dimensions = tuple(d if isinstance(d, Dimension) else Dimension(d) for d in dimensions)
        self.shape = tuple(d.size for d in self.dimensions)
        self.dtype = np.dtype(dtype)

    def required_padded_shape(self):
        return tuple(d.required_padded_size() for d in self.dimensions)

    def required_bytes(self):
        return int(np.prod(self.required_padded_shape(), dtype=np.int64)) * self.dtype.itemsize

    def is_compatible(self, buffer):
        return (buffer.shape == self.shape and buffer.dtype == self.dtype
                and all(d.valid(p) for d, p in zip(self.dimensions, buffer.padded_shape)))

    def validate(self, buffer):
        if buffer.shape != self.shape:
            raise ValueError('shape mismatch: slot {} buffer {}'.format(self.shape, buffer.shape))
        if buffer.dtype != self.dtype:
            raise TypeError('dtype mismatch: slot {} buffer {}'.format(self.dtype, buffer.dtype))
        for d, p in zip(self.dimensions, buffer.padded_shape):
            if not d.valid(p):
                raise ValueError('padded shape {} does not meet the slot requirements'.format(
                    buffer.padded_shape))

    def bind(self, buffer):
        if buffer is not None:
            self.validate(buffer)
        self.buffer = buffer

    def allocate(self, allocator, bind=True):
        buffer = allocator.allocate(self.shape, self.dtype, self.required_padded_shape())
        if bind:
            self.bind(buffer)
        return buffer

    def allocate_host(self, context):
        return HostArray(self.shape, self.dtype, self.required_padded_shape(), context=context)


class CompoundIOSlot(IOSlotBase):
    """Several slots (of several operations) that must share one buffer."""

    def __init__(self, children):
        super().__init__()
        self.children = list(children)
        if not self.children:
            raise ValueError('empty compound slot')
        first = self.children[0]
        self.shape = first.shape
        self.dtype = first.dtype
        self.dimensions = first.dimensions
        for child in self.children[1:]:
            if child.shape != self.shape:
                raise ValueError('aliased slots have different shapes: {} vs {}'.format(
                    child.shape, self.shape))
            if child.dtype != self.dtype:
                raise TypeError('aliased slots have different dtypes: {} vs {}'.format(
                    child.dtype, self.dtype))
            for a, b in zip(self.dimensions, child.dimensions):
                a.link(b)
        for child in self.children:
            if child.buffer is not None:
                self.buffer = child.buffer
        if self.buffer is not None:
            self.bind(self.buffer)

    required_padded_shape = IOSlot.required_padded_shape
    required_bytes = IOSlot.required_bytes
    is_compatible = IOSlot.is_compatible
    validate = IOSlot.validate
    allocate = IOSlot.allocate
    allocate_host = IOSlot.allocate_host

    def bind(self, buffer):
        if buffer is not None:
            self.validate(buffer)
        for child in self.children:
            child.bind(buffer)
        self.buffer = buffer


class Operation:
    """Base class of everything that runs on the device: owns slots, allocates
    unbound ones on demand and enqueues work on `command_queue` when called."""

    def __init__(self, command_queue, allocator=None):
        self.command_queue = command_queue
        self.slots = OrderedDict()
        self.hidden_slots = OrderedDict()
        if allocator is None:
            allocator = DeviceAllocator(command_queue.context)
        self.allocator = allocator

    def bind(self, **kwargs):
        for name, buffer in kwargs.items():
            self.slots[name].bind(buffer)

    def ensure_bound(self, name):
        slot = self.slots[name] if name in self.slots else self.hidden_slots[name]
        if not slot.is_bound():
            slot.allocate(self.allocator)

    def ensure_all_bound(self):
        for name in self.slots:
            self.ensure_bound(name)
        for name in self.hidden_slots:
            self.ensure_bound(name)

    def buffer(self, name):
        slot = self.slots[name] if name in self.slots else self.hidden_slots[name]
        return slot.buffer

    def parameters(self):
        return {}

    def _run(self):
        raise NotImplementedError()

    def __call__(self, **kwargs):
        self.bind(**kwargs)
        self.ensure_all_bound()
        return self._run()


class OperationSequence(Operation):
    """Runs several operations in order.  Child slots appear as ``op:slot``
    unless grouped by `compounds` (new name -> list of ``op:slot``), in which case
    the children share a buffer and padding requirements."""

    def __init__(self, command_queue, operations, compounds=None, aliases=None, allocator=None):
        super().__init__(command_queue, allocator)
        self.operations = OrderedDict(operations)
        if compounds is None:
            compounds = {}
        if aliases:
            raise NotImplementedError('aliases (overlapping scratch buffers) are not supported')
        children = OrderedDict()
        for op_name, op in self.operations.items():
            for slot_name, slot in op.slots.items():
                children['{}:{}'.format(op_name, slot_name)] = slot
            for slot_name, slot in op.hidden_slots.items():
                self.hidden_slots['{}:{}'.format(op_name, slot_name)] = slot
        for name, members in compounds.items():
            if not members:
                continue
            self.slots[name] = CompoundIOSlot([children.pop(m) for m in members])
        for name, slot in children.items():
            self.slots[name] = slot

    def _run(self):
        for op in self.operations.values():
            op()

    def parameters(self):
        out = {}
        for op_name, op in self.operations.items():
            for key, value in op.parameters().items():
                out['{}:{}'.format(op_name, key)] = value
        return out
