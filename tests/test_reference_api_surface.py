"""API-surface compatibility with the reference (runs only where /root/reference exists):
the reference's own imaging.py, imported UNCHANGED over this package's modules (the
redirect of INTEGRATION.md section 1), must find every name it uses."""
import ast
import importlib
import os
import sys
import types

import numpy as np
import pytest

REFERENCE = os.environ.get('KATSDPIMAGER_REFERENCE', '/root/reference')
IMAGING = os.path.join(REFERENCE, 'katsdpimager', 'imaging.py')

pytestmark = pytest.mark.skipif(not os.path.exists(IMAGING), reason='reference tree not present')


def _used_attributes(path, modules):
    """{module: {names}} for every `module.name` expression in the file."""
    tree = ast.parse(open(path).read())
    used = {m: set() for m in modules}
    for node in ast.walk(tree):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) \
                and node.value.id in used:
            used[node.value.id].add(node.attr)
    return used


def test_names_used_by_reference_imaging_exist():
    used = _used_attributes(IMAGING, ['grid', 'predict', 'weight', 'image', 'clean', 'accel'])
    assert used['accel'] >= {'OperationSequence', 'HostArray', 'DeviceArray'}
    for module, names in used.items():
        mine = importlib.import_module('katsdpimager_b200.' + module)
        # ImagingHost (the --host CPU path, same file) is deliberately not provided: the
        # product has no CPU fallback; its role in tests is played by oracle/.
        names = {n for n in names if not n.endswith('Host') and not n.endswith('_host')}
        missing = sorted(n for n in names if not hasattr(mine, n))
        assert not missing, '{} lacks {}'.format(module, missing)


def test_names_used_by_reference_operations_exist():
    """Every katsdpsigproc.accel / fft / fill / tune name the reference's operation modules
    use is provided by the shim."""
    for ref_module in ('grid', 'image', 'clean', 'weight', 'predict'):
        path = os.path.join(REFERENCE, 'katsdpimager', ref_module + '.py')
        used = _used_attributes(path, ['accel', 'fft', 'fill', 'tune'])
        for module, names in used.items():
            mine = importlib.import_module('katsdpimager_b200.' + module)
            # accel.build compiles Mako kernels; there is nothing to build here
            missing = sorted(n for n in names - {'build'} if not hasattr(mine, n))
            assert not missing, '{}.py uses {}.{}'.format(ref_module, module, missing)


def test_reference_imaging_template_constructs_over_this_package():
    from tests import reference_loader
    with reference_loader.reference_imaging() as ref_imaging:
        from katsdpimager_b200 import clean, parameters as prm, weight
        fixed_image = prm.FixedImageParameters([1, 2, 3, 4], np.float32)
        fixed_grid = prm.FixedGridParameters(7.0, 8, 4, 1000.0, 7, degrid=True)
        clean_p = prm.CleanParameters(100, 0.1, 0.85, 5.0, clean.CLEAN_SUMSQ, 0.01, 0.5, 0.02)
        template = ref_imaging.ImagingTemplate(
            None, prm.ArrayParameters(13.5, 1000.0), fixed_image,
            prm.WeightParameters(weight.WeightType.ROBUST, 0.5), fixed_grid, clean_p)
        assert isinstance(template.gridder, sys.modules['katsdpimager.grid'].GridderTemplate)
        assert template.degridder is not None and template.clean.num_polarizations == 4
        # the Imaging class of the reference subclasses our OperationSequence
        assert issubclass(ref_imaging.Imaging, sys.modules['katsdpsigproc.accel'].OperationSequence)
    assert 'katsdpimager.imaging' not in sys.modules
