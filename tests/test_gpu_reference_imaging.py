"""The drop-in boundary, executed: the reference's OWN ``katsdpimager/imaging.py`` --
unmodified, loaded from the reference tree (or the copy `__graft_entry__.build()` stages under
baseline/_ref for the GPU box) -- drives this package's operations through a complete
``frontend.process_channel`` replay on the GPU, and lands on the outputs of the reference's
``ImagingHost`` (golden fixture made from the unmodified reference, tests/golden).

Everything the reference facade does goes through the shim exactly as it would through
katsdpsigproc: ``accel.OperationSequence`` slot aliasing (imaging.py:185-215), pinned
``_HostBuffer`` staging with ``set_region(..., blocking=False)`` + ``enqueue_marker``
(imaging.py:54-78, 269-291), ``dirty_to_psf`` buffer swapping (:370-373), one ``Clean.__call__``
per minor cycle (:389-396)."""
import numpy as np
import pytest

from katsdpimager_b200 import parameters as prm, weight
from tests import cases, reference_loader
from tests.cases import load_golden

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(reference_loader.reference_imaging_path() is None,
                                 reason='reference katsdpimager/imaging.py not available')]


def _rms_rel(actual, expected):
    return np.sqrt(np.mean((actual - expected) ** 2)) / np.abs(expected).max()


@pytest.mark.parametrize('degrid', [True, False])
def test_reference_imaging_runs_unmodified(gpu, degrid):
    context, queue = gpu
    fx = cases.imaging_case(degrid=degrid)
    ip, gp, cp = fx['image_parameters'], fx['grid_parameters'], fx['clean_parameters']
    wp = prm.WeightParameters(weight.WeightType.UNIFORM)
    with reference_loader.reference_imaging() as ref:
        assert ref.__file__.endswith('katsdpimager/imaging.py')
        assert 'katsdpimager_b200' not in ref.__file__
        template = ref.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)

        def make():
            imager = template.instantiate(queue, ip, gp, fx['vis_block'], 0, fx['major'])
            assert type(imager).__module__ == 'katsdpimager.imaging'
            return imager

        out = cases.run_imaging(make, fx)
    if degrid:
        golden = load_golden('imaging_small')
        np.testing.assert_allclose(out['weights_rms'], golden['weights_rms'], rtol=1e-5)
        np.testing.assert_allclose(out['psf_peak'], golden['psf_peak'], rtol=1e-5)
        np.testing.assert_array_equal(out['psf_patch'], golden['psf_patch'])
        assert _rms_rel(out['psf'], golden['psf']) < 1e-4
        assert _rms_rel(out['dirty0'], golden['dirty0']) < 1e-4
        np.testing.assert_allclose(out['noise'], golden['noise'], rtol=1e-4)
        assert len(out['values']) == len(golden['values'])
        np.testing.assert_allclose(out['values'], golden['values'], rtol=2e-4)
        np.testing.assert_array_equal(np.argwhere(out['model'] != 0),
                                      np.argwhere(golden['model'] != 0))
        np.testing.assert_allclose(out['model'], golden['model'],
                                   rtol=0, atol=1e-5 * np.abs(golden['model']).max())
        assert _rms_rel(out['residual'], golden['residual']) < 1e-4
    else:
        # direct prediction (model_to_predict): compare with this package's own facade
        from katsdpimager_b200 import imaging
        own = imaging.ImagingTemplate(context, fx['array_parameters'], ip.fixed, wp, gp.fixed, cp)
        expected = cases.run_imaging(
            lambda: own.instantiate(queue, ip, gp, fx['vis_block'], 0, fx['major']), fx)
        assert len(out['values']) == len(expected['values'])
        np.testing.assert_array_equal(out['component_pos'], expected['component_pos'])
        assert _rms_rel(out['residual'], expected['residual']) < 1e-4      # north_star bar
