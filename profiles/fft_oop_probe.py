"""cuFFT 8192^2 C2C: in place against out of place (timing, and whether the input survives)."""
import sys, ctypes, numpy as np
sys.path.insert(0, '.')
from katsdpimager_b200 import _lib, accel
ctx = accel.Context(0); q = ctx.create_command_queue()
N = 8192
A = accel.DeviceArray(ctx, (N, N), np.complex64); B = accel.DeviceArray(ctx, (N, N), np.complex64)
rs = np.random.RandomState(1)
x = (rs.standard_normal((N, N)) + 1j * rs.standard_normal((N, N))).astype(np.complex64)
x[2461:N - 2461, :] = 0
x[:, 2461:N - 2461] = 0
h = ctypes.c_void_p(); _lib.call('kib_fft_plan2d_create', ctypes.byref(h), N, N, N, 0)
def timeit(fn, reps=20):
    fn(); q.finish()
    a = q.enqueue_marker()
    for _ in range(reps): fn()
    b = q.enqueue_marker(); b.wait()
    return b.time_since(a) / reps * 1e6
A.set(q, x)
print('in place     us', timeit(lambda: _lib.call('kib_fft_plan2d_exec', h, A.ptr, A.ptr, 1, q.stream)))
A.set(q, x)
print('out of place us', timeit(lambda: _lib.call('kib_fft_plan2d_exec', h, A.ptr, B.ptr, 1, q.stream)))
back = A.get(q)
print('input preserved:', bool(np.array_equal(back, x)))
ref = B.get(q).copy()
A.set(q, x); _lib.call('kib_fft_plan2d_exec', h, A.ptr, A.ptr, 1, q.stream)
print('same result:', float(np.abs(A.get(q) - ref).max() / np.abs(ref).max()))
