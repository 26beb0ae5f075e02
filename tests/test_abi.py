"""The C ABI: every function declared in include/katimager_b200.h is exported by
libkatimager_b200.so and bound in katsdpimager_b200/_lib.py with the same number of
arguments.  No compute entry point is called (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

from katsdpimager_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'katimager_b200.h')


def _declarations():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = {}
    for match in re.finditer(r'\b(int|const char \*)\s*(kib_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        args = match.group(3).strip()
        count = 0 if args in ('', 'void') else len(args.split(','))
        decls[match.group(2)] = count
    return decls


def test_header_parses():
    decls = _declarations()
    assert len(decls) >= 50
    for name in ('kib_grid', 'kib_degrid', 'kib_layer_to_image', 'kib_clean_minor_cycles',
                 'kib_fft_plan2d_exec', 'kib_predict', 'kib_density_weights'):
        assert name in decls


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if ' T ' in line}
    missing = sorted(set(_declarations()) - exported)
    assert not missing, 'declared but not exported: {}'.format(missing)


def test_bindings_match_header():
    decls = _declarations()
    bound = dict(_lib.SIGNATURES)
    assert set(bound) | {'kib_version', 'kib_last_error'} == set(decls)
    for name, argtypes in bound.items():
        assert len(argtypes) == decls[name], name


def test_library_loads_without_gpu():
    lib = _lib.load()
    assert lib.kib_version() == 1
    assert isinstance(lib.kib_last_error(), bytes)
    # argument validation happens before any CUDA call
    rc = lib.kib_grid(None, 0, 0, 255, 0, None, 0, 0, None, None, None, None, 7, 0,
                      1, 8, 7, 1, 1, None, None)
    assert rc != 0
    assert b'odd grid size' in lib.kib_last_error()
    with pytest.raises(_lib.KibError):
        _lib.call('kib_scale', None, 0, 0, 1, 1, 5, None, 0, None)


def test_no_fallback_in_product():
    """The product never imports the oracle, torch or any alternative backend."""
    package = os.path.join(ROOT, 'katsdpimager_b200')
    for name in os.listdir(package):
        if name.endswith('.py'):
            text = open(os.path.join(package, name)).read()
            forbidden_imports = ['import oracle', 'from oracle', 'import numba', 'import triton',
                                 'import pycuda']
            if name != 'distributed.py':    # torch.distributed gathers finished planes only
                forbidden_imports.append('import torch')
            for forbidden in forbidden_imports:
                assert forbidden not in text, '{} in {}'.format(forbidden, name)
