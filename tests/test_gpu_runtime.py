"""Runtime layer through the C ABI: memory, strided region copies, events, cuFFT."""
import numpy as np
import pytest

from katsdpimager_b200 import accel, fft

pytestmark = pytest.mark.gpu


def test_set_get_roundtrip(gpu):
    context, queue = gpu
    rs = np.random.RandomState(1)
    for shape, padded in [((5,), None), ((7, 9), (8, 16)), ((3, 10, 11), (3, 12, 16))]:
        dev = accel.DeviceArray(context, shape, np.float32, padded)
        host = rs.uniform(size=shape).astype(np.float32)
        dev.set(queue, host)
        np.testing.assert_array_equal(dev.get(queue), host)
        pinned = dev.empty_like()
        pinned[:] = host * 2
        dev.set_async(queue, pinned)
        out = dev.get_async(queue)
        queue.finish()
        np.testing.assert_array_equal(out, host * 2)


def test_regions(gpu):
    context, queue = gpu
    rs = np.random.RandomState(2)
    dev = accel.DeviceArray(context, (4, 20, 30), np.complex64, (4, 24, 32))
    ref = (rs.uniform(size=dev.shape) + 1j * rs.uniform(size=dev.shape)).astype(np.complex64)
    dev.set(queue, ref)
    # set_region with broadcasting-free slices, including a collapsed axis
    patch = (rs.uniform(size=(7, 9)) + 0j).astype(np.complex64)
    dev.set_region(queue, patch, np.s_[2, 3:10, 11:20], np.s_[:, :])
    ref[2, 3:10, 11:20] = patch
    np.testing.assert_array_equal(dev.get(queue), ref)
    # get_region into a strided host view
    out = np.zeros((4, 5, 6), np.complex64)
    dev.get_region(queue, out, np.s_[:, 15:20, -6:], np.s_[:, :, :])
    np.testing.assert_array_equal(out, ref[:, 15:20, -6:])
    # copy_region between differently padded device arrays, dropping the first axis
    other = accel.DeviceArray(context, (25, 40), np.complex64, (25, 48))
    other.zero(queue)
    dev.copy_region(queue, other, np.s_[1, 5:15, :30], np.s_[10:20, 10:40])
    expected = np.zeros((25, 40), np.complex64)
    expected[10:20, 10:40] = ref[1, 5:15, :30]
    np.testing.assert_array_equal(other.get(queue), expected)
    # row-prefix upload as used for visibility staging (imaging.py:269-291)
    vis = accel.DeviceArray(context, (100, 4), np.int16)
    vis.zero(queue)
    host = vis.empty_like()
    host[:] = rs.randint(-100, 100, host.shape)
    vis.set_region(queue, host, np.s_[:37, :2], np.s_[:37, :2], blocking=False)
    queue.finish()
    expected = np.zeros((100, 4), np.int16)
    expected[:37, :2] = host[:37, :2]
    np.testing.assert_array_equal(vis.get(queue), expected)


def test_small_reads_into_pinned_memory(gpu):
    """Device-to-host reads of up to 256 KB into page-locked memory are written by a kernel
    (kib_runtime.cu: small_d2h) instead of the copy engine: contiguous and strided, word- and
    byte-aligned, into pinned and into pageable destinations (the latter take the copy engine),
    and just around the size limit."""
    context, queue = gpu
    rs = np.random.RandomState(3)
    # contiguous, several sizes around the limit and not multiples of 4 bytes
    for nbytes in (1, 3, 4, 1000, 4099, 256 * 1024, 256 * 1024 + 4):
        ref = rs.randint(0, 256, nbytes).astype(np.uint8)
        dev = accel.DeviceArray(context, (nbytes,), np.uint8)
        dev.set(queue, ref)
        pinned = accel.HostArray((nbytes,), np.uint8, context=context)
        pinned[:] = 0
        dev.get(queue, pinned)
        np.testing.assert_array_equal(np.asarray(pinned), ref)
        pageable = np.zeros(nbytes, np.uint8)
        dev.get(queue, pageable)
        np.testing.assert_array_equal(pageable, ref)
    # strided on both sides (the PSF-peak read of pipeline.process_channel is one of these)
    dev = accel.DeviceArray(context, (4, 50, 60), np.float32, (4, 56, 64))
    ref = rs.uniform(size=dev.shape).astype(np.float32)
    dev.set(queue, ref)
    peak = accel.HostArray((4,), np.float32, context=context)
    dev.get_region(queue, peak, np.s_[:, 25, 30], np.s_[:])
    np.testing.assert_array_equal(np.asarray(peak), ref[:, 25, 30])
    out = accel.HostArray((4, 9, 16), np.float32, context=context)
    out[:] = -1
    dev.get_region(queue, out, np.s_[1:3, 7:13, 5:12], np.s_[2:4, 1:7, 3:10])
    expected = np.full((4, 9, 16), -1, np.float32)
    expected[2:4, 1:7, 3:10] = ref[1:3, 7:13, 5:12]
    np.testing.assert_array_equal(np.asarray(out), expected)
    # int16 columns: byte-granular path
    vis = accel.DeviceArray(context, (100, 3), np.int16)
    host = rs.randint(-100, 100, (100, 3)).astype(np.int16)
    vis.set(queue, host)
    col = accel.HostArray((100, 1), np.int16, context=context)
    vis.get_region(queue, col, np.s_[:, 1:2], np.s_[:, :])
    np.testing.assert_array_equal(np.asarray(col), host[:, 1:2])


def test_events(gpu):
    context, queue = gpu
    dev = accel.DeviceArray(context, (1 << 22,), np.float32)
    start = queue.enqueue_marker()
    for _ in range(10):
        dev.zero(queue)
    stop = queue.enqueue_marker()
    stop.wait()
    assert 0 < stop.time_since(start) < 1.0


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128])
def test_fft(gpu, dtype):
    context, queue = gpu
    shape = (96, 120)
    rs = np.random.RandomState(3)
    data = (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(dtype)
    template = fft.FftTemplate(context, 2, shape, dtype, dtype, shape, shape)
    tol = 1e-5 if dtype == np.complex64 else 1e-13
    for mode, expected in [(fft.FftMode.FORWARD, np.fft.fft2(data)),
                           (fft.FftMode.INVERSE, np.fft.ifft2(data) * data.size)]:
        op = template.instantiate(queue, mode)
        buf = accel.DeviceArray(context, shape, dtype)
        buf.set(queue, data)
        op.bind(src=buf, dest=buf)
        op()
        np.testing.assert_allclose(buf.get(queue), expected, rtol=0,
                                   atol=tol * np.abs(expected).max())
