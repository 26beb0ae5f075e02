"""Import the reference's own host (CPU) classes from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that the golden-vector
generator (tests/golden/make_golden.py) and the oracle-pinning tests can run
the *unmodified* reference ``--host`` implementation in the build container.
It is never imported by the product package, by ``-m gpu`` tests, by
``__graft_entry__.smoke()`` or by ``bench.py`` (``/root/reference`` does not
exist on the GPU box).

The reference depends on ``katsdpsigproc`` and ``astropy`` which are not
installed here; only import-time names are needed by the host classes, so we
register minimal stub modules before importing (recipe: SURVEY.md section 8c).
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('KATSDPIMAGER_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'katsdpimager'))


def _install_stubs():
    if 'katsdpsigproc' in sys.modules and getattr(
            sys.modules['katsdpsigproc'], '_kib_oracle_stub', False):
        return

    def divup(x, y):
        return (x + y - 1) // y

    def roundup(x, y):
        return divup(x, y) * y

    class _Placeholder:
        def __init__(self, *args, **kwargs):
            pass

    sigproc = types.ModuleType('katsdpsigproc')
    sigproc._kib_oracle_stub = True
    sigproc.__path__ = []
    accel = types.ModuleType('katsdpsigproc.accel')
    accel.divup = divup
    accel.roundup = roundup
    for name in ['Operation', 'OperationSequence', 'IOSlot', 'Dimension', 'DeviceArray',
                 'HostArray', 'DeviceAllocator', 'AbstractAllocator']:
        setattr(accel, name, type(name, (_Placeholder,), {}))
    tune = types.ModuleType('katsdpsigproc.tune')

    def autotuner(test=None):
        def decorator(fn):
            return fn
        return decorator
    tune.autotuner = autotuner
    abc = types.ModuleType('katsdpsigproc.abc')
    abc.AbstractEvent = type('AbstractEvent', (), {})
    abc.AbstractCommandQueue = type('AbstractCommandQueue', (), {})
    fft = types.ModuleType('katsdpsigproc.fft')
    fill = types.ModuleType('katsdpsigproc.fill')
    for mod in (accel, tune, abc, fft, fill):
        sys.modules[mod.__name__] = mod
        setattr(sigproc, mod.__name__.split('.')[-1], mod)
    sys.modules['katsdpsigproc'] = sigproc

    if 'astropy' not in sys.modules:
        astropy = types.ModuleType('astropy')
        astropy.__path__ = []
        units = types.ModuleType('astropy.units')
        astropy.units = units
        sys.modules['astropy'] = astropy
        sys.modules['astropy.units'] = units


_cache = {}


def load():
    """Returns a namespace with the reference modules grid, image, clean, weight,
    predict, imaging, fast_math (host classes usable; device classes are not)."""
    if 'ns' in _cache:
        return _cache['ns']
    if not available():
        raise ImportError('reference tree not present at ' + REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from katsdpimager import grid, image, clean, weight, predict, imaging, fast_math
    ns = types.SimpleNamespace(grid=grid, image=image, clean=clean, weight=weight,
                               predict=predict, imaging=imaging, fast_math=fast_math)
    _cache['ns'] = ns
    return ns
