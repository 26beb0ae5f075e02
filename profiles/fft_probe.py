import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
from katsdpimager_b200 import _lib, accel
ctx = accel.Context(0); q = ctx.create_command_queue()
N=8192; G=4922; half=G//2
A = accel.DeviceArray(ctx,(N,N),np.complex64); B = accel.DeviceArray(ctx,(N,N),np.complex64)
A.zero(q); B.zero(q)
def plan1d(n, stride, dist, batch):
    h=ctypes.c_void_p(); _lib.call('kib_fft_plan1d_create', ctypes.byref(h), n, stride, dist, batch, 0); return h
def plan2d():
    h=ctypes.c_void_p(); _lib.call('kib_fft_plan2d_create', ctypes.byref(h), N, N, N, 0); return h
def timeit(fn, reps=10):
    fn(); q.finish()
    a=q.enqueue_marker()
    for _ in range(reps): fn()
    b=q.enqueue_marker(); b.wait()
    return b.time_since(a)/reps*1e6
p2=plan2d(); prow=plan1d(N,1,N,half); pcol=plan1d(N,N,1,N); prow_all=plan1d(N,1,N,N)
print('2d inplace', timeit(lambda: _lib.call('kib_fft_plan2d_exec', p2, A.ptr, A.ptr, 1, q.stream)))
print('rows all', timeit(lambda: _lib.call('kib_fft_plan2d_exec', prow_all, A.ptr, A.ptr, 1, q.stream)))
print('rows band x2', timeit(lambda: (_lib.call('kib_fft_plan2d_exec', prow, A.ptr, A.ptr, 1, q.stream), _lib.call('kib_fft_plan2d_exec', prow, A.ptr.value + (N-half)*N*8, A.ptr.value + (N-half)*N*8, 1, q.stream))))
print('cols inplace', timeit(lambda: _lib.call('kib_fft_plan2d_exec', pcol, A.ptr, A.ptr, 1, q.stream)))
print('cols A->B', timeit(lambda: _lib.call('kib_fft_plan2d_exec', pcol, A.ptr, B.ptr, 1, q.stream)))
# correctness of rows+cols vs 2d on random data
rs=np.random.RandomState(1); x=(rs.standard_normal((N,N))+1j*rs.standard_normal((N,N))).astype(np.complex64)
x[half:N-half,:]=0
A.set(q,x); _lib.call('kib_fft_plan2d_exec', p2, A.ptr, A.ptr, 1, q.stream); ref=A.get(q).copy()
A.set(q,x)
_lib.call('kib_fft_plan2d_exec', prow, A.ptr, A.ptr, 1, q.stream); _lib.call('kib_fft_plan2d_exec', prow, A.ptr.value + (N-half)*N*8, A.ptr.value + (N-half)*N*8, 1, q.stream)
_lib.call('kib_fft_plan2d_exec', pcol, A.ptr, B.ptr, 1, q.stream)
out=B.get(q)
print('err', np.abs(out-ref).max()/np.abs(ref).max())
