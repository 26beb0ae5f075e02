"""Device-side FITS ordering (kib_fits_plane) straight into a page-locked mapping of the cube
file: same bytes as the host writer (reference io.py:191-200 semantics)."""
import numpy as np
import pytest

from katsdpimager_b200 import accel, io, parameters as prm

pytestmark = pytest.mark.gpu


def test_store_device_matches_host_store(gpu, tmp_path):
    context, queue = gpu
    pols, n = 3, 96
    fixed = prm.FixedImageParameters([1, 2, 3], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.21, pixels=n, pixel_size=1e-4)
    rs = np.random.RandomState(4)
    planes = rs.standard_normal((4, pols, n, n)).astype(np.float32)
    name = str(tmp_path / 'cube.fits')
    io.FitsCube.create(name, 4, ip, 856e6, 1e6).close()
    cube = io.FitsCube(name)
    cube.pin(1, 3)                                   # this "rank" owns channels 1 and 2
    image = accel.DeviceArray(context, (pols, n, n), np.float32, (pols, n, n + 8))
    for channel in (1, 2):
        image.set(queue, planes[channel])
        cube.store_device(channel, image, queue)
        queue.finish()
    cube.store(0, planes[0])                         # host path
    cube.store(3, planes[3])
    cube.close()
    header, data = io.read_fits(name)
    assert header['NAXIS4'] == 4
    np.testing.assert_array_equal(data, planes[:, :, :, ::-1])
