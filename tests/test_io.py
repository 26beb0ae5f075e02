"""FITS writer (katsdpimager_b200/io.py) against the reference's writer semantics
(reference katsdpimager/io.py:88-203): header keywords, l-axis flip, big-endian data, and the
shared-file channel cube.  astropy is not available, so the files are parsed back with the
module's own minimal reader and with raw byte checks."""
import math
import os

import numpy as np
import pytest

from katsdpimager_b200 import io, parameters as prm, polarization


def _params(pols=(1, 2, 3, 4), pixels=16):
    fixed = prm.FixedImageParameters(list(pols), np.float32)
    return prm.ImageParameters(fixed, wavelength=0.21, pixels=pixels, pixel_size=1e-4)


def test_card_format():
    assert io.format_card('SIMPLE', True) == 'SIMPLE  =                    T'.ljust(80)
    assert io.format_card('NAXIS', 4) == 'NAXIS   =                    4'.ljust(80)
    assert io.format_card('CTYPE1', 'RA---SIN') == "CTYPE1  = 'RA---SIN'".ljust(80)
    assert io.format_card('BUNIT', 'Jy') == "BUNIT   = 'Jy      '".ljust(80)
    card = io.format_card('CRVAL2', -35.0)
    assert card.startswith('CRVAL2  = ') and float(card[10:30]) == -35.0 and len(card) == 80
    assert io.format_card('END') == 'END'.ljust(80)


def test_stokes_axis():
    keys, permute = io.stokes_axis([1, 2, 3, 4])
    assert keys == {'CTYPE3': 'STOKES', 'CRPIX3': 1.0, 'CRVAL3': 1.0, 'CDELT3': 1.0}
    assert list(permute) == [0, 1, 2, 3]
    keys, permute = io.stokes_axis([polarization.STOKES_XX, polarization.STOKES_YY])
    assert keys['CRVAL3'] == -5.0 and keys['CDELT3'] == -1.0 and list(permute) == [1, 0]
    with pytest.raises(ValueError):
        io.stokes_axis([polarization.STOKES_I, polarization.STOKES_Q, polarization.STOKES_V])


def test_write_fits_image(tmp_path):
    ip = _params()
    rs = np.random.RandomState(1)
    image = rs.standard_normal((4, 16, 16)).astype(np.float32)
    beam = type('Beam', (), {'major': 3.0, 'minor': 2.0, 'theta': 0.5})()
    path = str(tmp_path / 'image-%d.fits')
    stored, cards = io.write_fits_image(image, ip, path, 7, (52.5, -35.0), beam)
    name = path % 7
    size = os.path.getsize(name)
    assert size % 2880 == 0
    header, data = io.read_fits(name)
    assert data.shape == (1, 4, 16, 16)
    np.testing.assert_array_equal(data[0], image[:, :, ::-1])          # io.py:191
    # raw bytes are big-endian (io.py:200)
    raw = open(name, 'rb').read()
    first = np.frombuffer(raw[io.header_bytes(cards).__len__():][:4], '>f4')[0]
    assert first == image[0, 0, -1]
    delt = math.degrees(math.asin(1e-4))
    assert header['SIMPLE'] is True and header['BITPIX'] == -32 and header['NAXIS'] == 4
    assert (header['NAXIS1'], header['NAXIS2'], header['NAXIS3'], header['NAXIS4']) == (16, 16, 4, 1)
    assert header['CRPIX1'] == 8.0 and header['CRPIX2'] == 9.0 and header['CRPIX4'] == 1.0
    assert header['CDELT1'] == pytest.approx(-delt) and header['CDELT2'] == pytest.approx(delt)
    assert header['CTYPE1'] == 'RA---SIN' and header['CTYPE2'] == 'DEC--SIN'
    assert header['CTYPE3'] == 'STOKES' and header['CTYPE4'] == 'FREQ'
    assert header['CRVAL1'] == 52.5 and header['CRVAL2'] == -35.0
    assert header['CRVAL4'] == pytest.approx(299792458.0 / 0.21)
    assert header['BUNIT'] == 'Jy/beam' and header['ORIGIN'] == 'katsdpimager'
    assert header['RADESYS'] == 'FK5' and header['EQUINOX'] == 2000.0
    assert header['BMAJ'] == pytest.approx(3.0 * math.degrees(1e-4))
    assert header['BPA'] == pytest.approx(math.degrees(0.5))
    assert header['DATAMIN'] == pytest.approx(float(image.min()))
    assert header['DATAMAX'] == pytest.approx(float(image.max()))


def test_cube_shared_between_workers(tmp_path):
    """Two 'workers' map the same cube and store their own channel blocks."""
    ip = _params(pols=(1, 2), pixels=8)
    name = str(tmp_path / 'cube.fits')
    io.FitsCube.create(name, 5, ip, 856e6, 1e6, (10.0, 20.0)).close()
    rs = np.random.RandomState(2)
    planes = rs.standard_normal((5, 2, 8, 8)).astype(np.float32)
    a, b = io.FitsCube(name), io.FitsCube(name)
    for c in (0, 1, 2):
        a.store(c, planes[c])
    for c in (3, 4):
        b.store(c, planes[c])
    a.close()
    b.close()
    header, data = io.read_fits(name)
    assert header['NAXIS4'] == 5 and header['CDELT4'] == 1e6 and header['CRVAL4'] == 856e6
    np.testing.assert_array_equal(data, planes[:, :, :, ::-1])
    assert os.path.getsize(name) % 2880 == 0
