"""Minimal timing hooks with the call signatures the reference's operations use
(reference katsdpimager/profiling.py:340-443: ``profile_function``, ``profile``,
``profile_device``).  The full profiler (flamegraphs, NVTX) is outside the hot
path; what is kept is the ability to bracket device work with events on the
launching stream, which bench.py uses for per-kernel times.
"""
import contextlib
import functools
import time
from collections import defaultdict


class DeviceTimer:
    """Collects (start, stop) event pairs per label."""

    def __init__(self):
        self.records = defaultdict(list)
        self.host = defaultdict(float)

    def device_seconds(self):
        """label -> (count, total seconds); blocks until the events complete."""
        out = {}
        for name, pairs in self.records.items():
            total = 0.0
            for start, stop in pairs:
                stop.wait()
                total += stop.time_since(start)
            out[name] = (len(pairs), total)
        return out

    def clear(self):
        self.records.clear()
        self.host.clear()


_active = None


def set_timer(timer):
    """Install (or with None remove) the active :class:`DeviceTimer`."""
    global _active
    _active = timer


@contextlib.contextmanager
def profile_device(queue, name, labels=None):
    if _active is None:
        yield
        return
    start = queue.enqueue_marker()
    try:
        yield
    finally:
        _active.records[name].append((start, queue.enqueue_marker()))


@contextlib.contextmanager
def profile(name, labels=None):
    if _active is None:
        yield
        return
    t0 = time.monotonic()
    try:
        yield
    finally:
        _active.host[name] += time.monotonic() - t0


def profile_function(name=None, labels=()):
    def decorator(func):
        label = name if name is not None else '{}.{}'.format(func.__module__, func.__qualname__)

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if _active is None:
                return func(*args, **kwargs)
            with profile(label):
                return func(*args, **kwargs)
        return wrapper
    return decorator
