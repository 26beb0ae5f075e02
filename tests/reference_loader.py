"""Load the reference's own katsdpimager/imaging.py UNMODIFIED with its imports redirected to
this package (the sys.modules redirect of INTEGRATION.md section 1).

The file is looked for in $KATSDPIMAGER_REFERENCE, /root/reference, and baseline/_ref (where
`__graft_entry__.build()` stages it when the reference tree is present: the GPU box has no
/root/reference).  Nothing of the reference is vendored in the repository."""
import contextlib
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get('KATSDPIMAGER_REFERENCE'), '/root/reference',
              os.path.join(ROOT, 'baseline', '_ref')]


def reference_imaging_path():
    for base in CANDIDATES:
        if base:
            path = os.path.join(base, 'katsdpimager', 'imaging.py')
            if os.path.exists(path):
                return path
    return None


@contextlib.contextmanager
def reference_imaging():
    """Context manager yielding the reference's imaging module running over this package."""
    path = reference_imaging_path()
    if path is None:
        raise FileNotFoundError('reference katsdpimager/imaging.py not found')

    def ours(key):
        return key == 'katsdpsigproc' or key.startswith('katsdpsigproc.') \
            or key == 'katsdpimager' or key.startswith('katsdpimager.')
    saved = {k: v for k, v in sys.modules.items() if ours(k)}
    try:
        sigproc = types.ModuleType('katsdpsigproc')
        sigproc.__path__ = []
        for name in ('accel', 'fft', 'fill', 'tune'):
            module = importlib.import_module('katsdpimager_b200.' + name)
            sys.modules['katsdpsigproc.' + name] = module
            setattr(sigproc, name, module)
        sys.modules['katsdpsigproc'] = sigproc
        package = types.ModuleType('katsdpimager')
        package.__path__ = [os.path.dirname(path)]
        sys.modules['katsdpimager'] = package
        for name in ('grid', 'image', 'clean', 'weight', 'predict', 'profiling'):
            sys.modules['katsdpimager.' + name] = importlib.import_module('katsdpimager_b200.' + name)
        spec = importlib.util.spec_from_file_location('katsdpimager.imaging', path)
        module = importlib.util.module_from_spec(spec)
        sys.modules['katsdpimager.imaging'] = module
        spec.loader.exec_module(module)
        yield module
    finally:
        for key in [k for k in sys.modules if ours(k)]:
            del sys.modules[key]
        sys.modules.update(saved)
