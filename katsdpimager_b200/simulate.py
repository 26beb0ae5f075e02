"""Synthetic MeerKAT-shaped visibilities for tests and benchmarks.

Plays the role of the reference's ``tests/simulate.py`` (MeerKAT-64 layout,
phase centre RA 3h30m Dec -35 deg, 4 s dumps, point sources from
``tests/lsm.txt``) without RASCIL/katpoint/casacore: UVW tracks come from the
standard hour-angle rotation of the ENU baselines and visibilities from a direct
DFT with the measurement equation documented at image.py:54-62.  Output order is
baseline-major, time-minor, which is what the reference's loader produces
(loader_ms.py:466-467) and what the gridder's accumulator reuse relies on.
"""
import os

import numpy as np

LATITUDE = -(30 + 42 / 60 + 39.8 / 3600) * np.pi / 180     # MeerKAT array reference
DISH_DIAMETER = 13.5
SIDEREAL_RATE = 2 * np.pi / 86164.0905                      # rad/s

#: tests/lsm.txt of the reference: (ra hours, dec degrees, I, Q, U, V)
LSM = [
    ((3, 30, 0.0), (-35, 0, 0.0), (1.0, 0.0, 0.0, 0.0)),
    ((3, 30, 30.0), (-35, 7, 0.0), (1.5, 0.0, 0.0, 0.0)),
    ((3, 32, 0.0), (-35, 2, 0.0), (1.5, 1.0, 0.0, 0.0)),
    ((3, 31, 0.0), (-35, 15, 0.0), (1.2, 0.0, -1.2, 0.0)),
    ((3, 28, 15.0), (-34, 50, 0.0), (0.75, 0.0, 0.0, -0.5)),
    ((3, 35, 0.0), (-34, 51, 0.0), (1.0, 0.2, 0.2, 0.2)),
    ((3, 35, 45.0), (-35, 23, 0.0), (1.2, 0.0, 0.0, 0.0)),
    ((3, 30, 0.0), (-34, 40, 0.0), (1.5, 0.7, 0.3, 0.0)),
]
PHASE_CENTRE = ((3, 30, 0.0), (-35, 0, 0.0))


def _hms(h):
    return (h[0] + h[1] / 60 + h[2] / 3600) * np.pi / 12


def _dms(d):
    sign = -1 if d[0] < 0 else 1
    return sign * (abs(d[0]) + d[1] / 60 + d[2] / 3600) * np.pi / 180


def meerkat_enu():
    """64 x 3 array of antenna East/North/Up positions (metres)."""
    path = os.path.join(os.path.dirname(__file__), 'data', 'meerkat_enu.txt')
    return np.loadtxt(path, usecols=(1, 2, 3))


def baselines_enu(enu=None):
    """ENU baseline vectors for all antenna pairs i < j (autocorrelations dropped,
    as loader_ms.py:411 does)."""
    if enu is None:
        enu = meerkat_enu()
    i, j = np.triu_indices(len(enu), 1)
    return enu[i] - enu[j]


def longest_baseline(enu=None):
    return float(np.max(np.linalg.norm(baselines_enu(enu), axis=1)))


def uvw_tracks(num_dumps, dump_time=4.0, start_hour_angle=-1.0, enu=None, dec=None,
               max_baselines=None):
    """UVW in metres, shape (baselines, dumps, 3), baseline-major.

    `start_hour_angle` in hours.  Uses the usual ENU -> equatorial XYZ -> UVW
    rotation (Thompson, Moran & Swenson eq. 4.1) for a source at declination
    `dec` (default: the simulate.py phase centre).
    """
    if dec is None:
        dec = _dms(PHASE_CENTRE[1])
    b = baselines_enu(enu)
    if max_baselines is not None:
        b = b[:max_baselines]
    e, n, u = b[:, 0], b[:, 1], b[:, 2]
    sl, cl = np.sin(LATITUDE), np.cos(LATITUDE)
    x = -sl * n + cl * u
    y = e
    z = cl * n + sl * u
    ha = start_hour_angle * np.pi / 12 + SIDEREAL_RATE * dump_time * (np.arange(num_dumps) + 0.5)
    sh, ch = np.sin(ha)[None, :], np.cos(ha)[None, :]
    sd, cd = np.sin(dec), np.cos(dec)
    x, y, z = x[:, None], y[:, None], z[:, None]
    uu = sh * x + ch * y
    vv = -sd * ch * x + sd * sh * y + cd * z
    ww = cd * ch * x - cd * sh * y + sd * z
    return np.stack([uu, vv, ww], axis=-1)


def lsm_lmn_flux():
    """Direction cosines (l, m, n-1) and IQUV flux densities of the lsm.txt sources."""
    ra0, dec0 = _hms(PHASE_CENTRE[0]), _dms(PHASE_CENTRE[1])
    lmn, flux = [], []
    for ra, dec, stokes in LSM:
        ra, dec = _hms(ra), _dms(dec)
        dra = ra - ra0
        l = np.cos(dec) * np.sin(dra)
        m = np.sin(dec) * np.cos(dec0) - np.cos(dec) * np.sin(dec0) * np.cos(dra)
        n = np.sqrt(1 - l * l - m * m)
        lmn.append((l, m, n - 1))
        flux.append(stokes)
    return np.array(lmn), np.array(flux)


def random_sources(num_sources, max_lm, num_polarizations, seed=3):
    """Seeded point-source model: l, m uniform in
    +-max_lm, Stokes I log-uniform 1 mJy - 1 Jy, other Stokes a random fraction."""
    rs = np.random.RandomState(seed)
    l = rs.uniform(-max_lm, max_lm, num_sources)
    m = rs.uniform(-max_lm, max_lm, num_sources)
    n1 = np.sqrt(1 - l * l - m * m) - 1
    flux = np.zeros((num_sources, num_polarizations))
    flux[:, 0] = 10 ** rs.uniform(-3, 0, num_sources)
    for p in range(1, num_polarizations):
        flux[:, p] = flux[:, 0] * rs.uniform(-0.3, 0.3, num_sources)
    return np.stack([l, m, n1], axis=1), flux


def dft_visibilities(uvw_wavelengths, lmn, flux, chunk=65536):
    """V = sum_s flux_s / 1 * exp(-2 pi i (u l + v m + w (n-1))) per polarization
    (sign convention of image.py:54-62; the 1/n factor is folded into `flux`)."""
    uvw = uvw_wavelengths.reshape(-1, 3)
    out = np.zeros((len(uvw), flux.shape[1]), np.complex64)
    for start in range(0, len(uvw), chunk):
        phase = uvw[start:start + chunk] @ lmn.T
        phase -= np.rint(phase)
        out[start:start + chunk] = (np.exp(-2j * np.pi * phase) @ flux).astype(np.complex64)
    return out.reshape(uvw_wavelengths.shape[:-1] + (flux.shape[1],))
