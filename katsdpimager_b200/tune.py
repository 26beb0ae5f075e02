"""Tuning hooks (katsdpsigproc.tune surface used by the reference: ``autotuner``,
``autotune``, ``make_measure``).

The reference autotunes Mako template parameters at run time and caches them in
sqlite.  The sm_100a kernels choose their thread/accumulator shapes from the
kernel support and polarization count inside the C library (kib_grid.cu
`choose_config`), so there is nothing to tune from Python; the decorators are kept
so code written against the reference API still imports.
"""
import functools


def autotuner(test=None):
    def decorator(fn):
        @functools.wraps(fn)
        def wrapper(cls, context, *args, **kwargs):
            return dict(test or {})
        return wrapper
    return decorator


def make_measure(queue, fn):
    def measure(iters):
        start = queue.enqueue_marker()
        for _ in range(iters):
            fn()
        stop = queue.enqueue_marker()
        stop.wait()
        return stop.time_since(start) / iters
    return measure


def autotune(generate, time_limit=0.1, **kwargs):
    raise NotImplementedError('run-time autotuning is not used by the sm_100a kernels')
