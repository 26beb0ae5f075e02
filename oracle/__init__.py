"""CPU oracle for the katsdpimager imaging hot path.

TEST INFRASTRUCTURE ONLY -- not part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this package, and there only as the checker or the
timed CPU baseline.  ``katsdpimager_b200`` never imports it.

The oracle restates the reference's ``--host`` classes:

=====================  ==========================================  ===========================
oracle                 reference (relative to reference root)      pinned by
=====================  ==========================================  ===========================
oracle.c kor_grid_*    katsdpimager/grid.py:1033 ``_grid``         tests/golden/grid_*.npz
oracle.c kor_degrid_*  katsdpimager/grid.py:1139 ``_degrid``       tests/golden/degrid_*.npz
host.grid_to_image     katsdpimager/image.py:781 GridToImageHost   tests/golden/image_*.npz
host.image_to_grid     katsdpimager/image.py:836 ImageToGridHost   tests/golden/image_*.npz
oracle.c kor_clean_*   katsdpimager/clean.py:946-1075 CleanHost    tests/golden/clean_*.npz
host.psf_patch         katsdpimager/clean.py:894 psf_patch_host    reference test known answers
host.noise_est         katsdpimager/clean.py:938 noise_est_host    tests/golden/clean_*.npz
host.WeightsHost       katsdpimager/weight.py:541 WeightsHost      tests/golden/weights_*.npz
oracle.c kor_predict   katsdpimager/predict.py:420 _predict_host   tests/golden/predict_*.npz
=====================  ==========================================  ===========================

The golden vectors were produced by running the *unmodified* reference classes
(imported from /root/reference under stub katsdpsigproc/astropy modules, see
``oracle/ref_import.py``) with ``tests/golden/make_golden.py``; parity is
therefore pinned, not merely restated.  ``oracle/_ref`` is unused: the reference
path is Python + numba, nothing to compile (its only native file,
``preprocess.cpp``, needs Eigen3, which is absent, and is outside this path).
"""
from . import host  # noqa: F401
from .host import *  # noqa: F401,F403
