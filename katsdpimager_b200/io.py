"""FITS output of image planes and channel cubes.

Restates the reference's ``write_fits_image`` (reference katsdpimager/io.py:88-203): same WCS
keywords (SIN projection about the phase centre, ``CRPIX1 = N/2``, ``CRPIX2 = N/2 + 1``,
``CDELT = -/+ arcsin(pixel_size)`` in degrees, a STOKES axis via ``_fits_polarizations``,
a degenerate FREQ axis), same data order: l axis reversed (RA increases to the left,
io.py:191), big-endian float32 (io.py:200).  The reference leans on astropy.io.fits; the
format is simple enough (80-character cards, 2880-byte blocks) to be written directly.

New here: :class:`FitsCube`, one file holding every channel of a spectral cube, memory-mapped
so that each worker process (one per GPU, :mod:`.distributed`) stores the planes of its own
channel block -- the "gather" of the channel-parallel design needs no collective and no copy
through another process.  With :meth:`FitsCube.pin` the mapped planes are page-locked and the
device writes FITS-ordered bytes (``kib_fits_plane``: flip + byte swap on the GPU) straight
into the file mapping.
"""
import datetime
import math
import os

import numpy as np

from . import polarization

BLOCK = 2880
CARD = 80

#: FITS Stokes codes (io.py:21-34)
_FITS_POLARIZATIONS = {
    polarization.STOKES_I: 1, polarization.STOKES_Q: 2,
    polarization.STOKES_U: 3, polarization.STOKES_V: 4,
}
for _name, _code in (('STOKES_RR', -1), ('STOKES_LL', -2), ('STOKES_RL', -3), ('STOKES_LR', -4),
                     ('STOKES_YY', -5), ('STOKES_XX', -6), ('STOKES_YX', -7), ('STOKES_XY', -8)):
    if hasattr(polarization, _name):
        _FITS_POLARIZATIONS[getattr(polarization, _name)] = _code


def _format_value(value):
    if isinstance(value, bool):
        return '{:>20s}'.format('T' if value else 'F')
    if isinstance(value, (int, np.integer)):
        return '{:>20d}'.format(int(value))
    if isinstance(value, (float, np.floating)):
        text = '{:.16G}'.format(float(value))
        if '.' not in text and 'E' not in text and 'N' not in text:
            text += '.0'
        return '{:>20s}'.format(text)
    text = str(value).replace("'", "''")
    return "'{:<8s}'".format(text)


def format_card(key, value=None, comment=None):
    """One 80-character header card."""
    if key in ('HISTORY', 'COMMENT'):
        card = '{:<8s}{}'.format(key, value)
    elif key == 'END':
        card = 'END'
    else:
        card = '{:<8s}= {}'.format(key, _format_value(value))
        if comment:
            card += ' / ' + comment
    if len(card) > CARD:
        raise ValueError('FITS card too long: {!r}'.format(card))
    return card.ljust(CARD)


def stokes_axis(polarizations):
    """Keywords of the STOKES axis and the permutation that orders the planes for it
    (``_fits_polarizations``, io.py:37-85); raises ValueError if the polarizations do not
    form a linear sequence in FITS enumeration."""
    codes = np.array([_FITS_POLARIZATIONS[p] for p in polarizations])
    permute = np.argsort(codes) if codes[0] >= 0 else np.argsort(-codes)
    codes = codes[permute]
    ref = int(codes[0])
    delta = int(codes[1] - codes[0]) if len(codes) > 1 else 1
    if np.any(codes != np.arange(len(codes)) * delta + ref):
        raise ValueError('Polarizations do not form a linear sequence in FITS enumeration')
    return {'CTYPE3': 'STOKES', 'CRPIX3': 1.0, 'CRVAL3': float(ref), 'CDELT3': float(delta)}, permute


def image_header(shape, image_parameters, phase_centre_deg, frequency_hz, channel_width_hz=1.0,
                 beam=None, bunit='Jy/beam', extra=None, data_range=None):
    """List of (key, value) for an image of `shape` = (channels, polarizations, height, width),
    following io.py:129-183.  `beam`, if given, has ``major``/``minor`` (pixels) and ``theta``
    (radians)."""
    channels, pols, height, width = shape
    delt = math.degrees(math.asin(float(image_parameters.pixel_size)))
    cards = [('SIMPLE', True), ('BITPIX', -32), ('NAXIS', 4), ('NAXIS1', width),
             ('NAXIS2', height), ('NAXIS3', pols), ('NAXIS4', channels), ('EXTEND', True)]
    if bunit is not None:
        cards.append(('BUNIT', bunit))
    cards += [('ORIGIN', 'katsdpimager'),
              ('HISTORY', 'Created by katsdpimager_b200 (B200 hot path of katsdpimager)'),
              ('TIMESYS', 'UTC'),
              ('DATE', datetime.datetime.now(datetime.timezone.utc).strftime('%Y-%m-%dT%H:%M:%S')),
              ('CRPIX1', width * 0.5), ('CRPIX2', height * 0.5 + 1.0), ('CRPIX4', 1.0),
              ('CDELT1', -delt), ('CDELT2', delt), ('CDELT4', float(channel_width_hz)),
              ('EQUINOX', 2000.0), ('RADESYS', 'FK5'),
              ('CUNIT1', 'deg'), ('CUNIT2', 'deg'), ('CUNIT4', 'Hz'),
              ('CTYPE1', 'RA---SIN'), ('CTYPE2', 'DEC--SIN'), ('CTYPE4', 'FREQ'),
              ('CRVAL1', float(phase_centre_deg[0])), ('CRVAL2', float(phase_centre_deg[1])),
              ('CRVAL4', float(frequency_hz))]
    if beam is not None:
        scale = math.degrees(float(image_parameters.pixel_size))
        cards += [('BMAJ', beam.major * scale), ('BMIN', beam.minor * scale),
                  ('BPA', math.degrees(beam.theta))]
    stokes, _ = stokes_axis(image_parameters.fixed.polarizations)
    cards += list(stokes.items())
    if data_range is not None and not math.isnan(data_range[0]):
        cards += [('DATAMIN', float(data_range[0])), ('DATAMAX', float(data_range[1]))]
    if extra:
        cards += list(extra.items())
    return cards


def header_bytes(cards):
    text = ''.join(format_card(k, v) for k, v in cards) + format_card('END')
    text += ' ' * (-len(text) % BLOCK)
    return text.encode('ascii')


def fits_order(image):
    """The array the reference hands to astropy (io.py:191-200): frequency axis added, l axis
    reversed, big-endian, contiguous."""
    image = image[np.newaxis, :, :, ::-1]
    return np.require(image, image.dtype.newbyteorder('>'), 'C')


def write_fits_image(image, image_parameters, filename, channel, phase_centre_deg=(0.0, 0.0),
                     beam=None, bunit='Jy/beam', extra_fits_headers=None):
    """Write one channel's image (polarizations x m x l, float32) to ``filename % channel``
    (reference io.py:88-203).  Returns (array as stored, header cards)."""
    if image.dtype != np.float32:
        raise TypeError('only float32 images are supported')
    frequency = 299792458.0 / float(image_parameters.wavelength)
    datamin = float(np.fmin.reduce(image, axis=None))
    datamax = float(np.fmax.reduce(image, axis=None))
    cards = image_header((1,) + image.shape, image_parameters, phase_centre_deg, frequency,
                         1.0, beam, bunit, extra_fits_headers, (datamin, datamax))
    stored = fits_order(image)
    try:
        path = filename % channel
    except TypeError:
        path = filename
    with open(path, 'wb') as f:
        f.write(header_bytes(cards))
        f.write(stored.tobytes())
        f.write(b'\0' * (-stored.nbytes % BLOCK))
    return stored, cards


def read_fits(filename):
    """Minimal reader for files written by this module: (dict of header values, data as a
    native-endian array of shape NAXIS4 x ... x NAXIS1).  Used by the tests."""
    header = {}
    with open(filename, 'rb') as f:
        offset = 0
        done = False
        while not done:
            block = f.read(BLOCK).decode('ascii')
            offset += BLOCK
            for i in range(0, BLOCK, CARD):
                card = block[i:i + CARD]
                key = card[:8].strip()
                if key == 'END':
                    done = True
                    break
                if card[8:10] != '= ':
                    continue
                value = card[10:].split(' / ')[0].strip()
                if value.startswith("'"):
                    header[key] = value[1:value.rindex("'")].rstrip()
                elif value in ('T', 'F'):
                    header[key] = value == 'T'
                else:
                    header[key] = float(value) if ('.' in value or 'E' in value) else int(value)
        shape = tuple(header['NAXIS{}'.format(i)] for i in range(header['NAXIS'], 0, -1))
        data = np.frombuffer(f.read(int(np.prod(shape)) * 4), '>f4').reshape(shape)
    return header, data.astype(np.float32)


#: bytes per piece of an overlapped device-to-host copy (see FitsCube.store_device)
COPY_PIECE = 8 << 20


class FitsCube:
    """A spectral cube (channels x polarizations x m x l) in one FITS file, shared between
    worker processes through the file mapping.

    ``FitsCube.create`` (one process) writes the header and sizes the file;
    ``FitsCube(filename)`` (every worker) maps it; ``cube.plane(channel)`` is the big-endian
    float32 view ``[polarizations, N, N]`` a worker fills, either from a host image with
    :meth:`store` (flip and byte swap on the host, as the reference does) or directly from
    the device with :meth:`store_device`.
    """

    def __init__(self, filename, mode='r+'):
        self.filename = filename
        with open(filename, 'rb') as f:
            self.header_size = 0
            while True:
                block = f.read(BLOCK)
                self.header_size += BLOCK
                if any(block[i:i + 8] == b'END     ' for i in range(0, BLOCK, CARD)):
                    break
        header, _ = _read_header_only(filename)
        self.shape = tuple(header['NAXIS{}'.format(i)] for i in range(4, 0, -1))
        self.data = np.memmap(filename, dtype='>f4', mode=mode, offset=self.header_size,
                              shape=self.shape)
        self._pinned = None
        self._bounce = False
        self._staging = None
        self._stored = None          # event: last overlapped device-to-host copy
        self._host = None

    @classmethod
    def create(cls, filename, num_channels, image_parameters, first_frequency_hz,
               channel_width_hz, phase_centre_deg=(0.0, 0.0), bunit='Jy/beam', extra=None):
        pols = len(image_parameters.fixed.polarizations)
        n = image_parameters.pixels
        shape = (num_channels, pols, n, n)
        head = header_bytes(image_header(shape, image_parameters, phase_centre_deg,
                                         first_frequency_hz, channel_width_hz, None, bunit,
                                         extra))
        nbytes = int(np.prod(shape)) * 4
        with open(filename, 'wb') as f:
            f.write(head)
            f.truncate(len(head) + nbytes + (-nbytes % BLOCK))
        return cls(filename)

    def plane(self, channel):
        return self.data[channel]

    def store(self, channel, image):
        """Host path: flip the l axis and convert to big-endian (io.py:191-200)."""
        self.data[channel] = image[:, :, ::-1]

    # ---- device path
    def pin(self, first_channel, last_channel):
        """Page-lock the planes of channels [first, last) for direct device copies."""
        from . import _lib
        self.unpin()
        plane_bytes = int(np.prod(self.shape[1:])) * 4
        start = self.data.ctypes.data + first_channel * plane_bytes
        nbytes = (last_channel - first_channel) * plane_bytes
        try:
            _lib.call('kib_host_register', start, nbytes)
            self._pinned = start
        except _lib.KibError:
            # the driver refuses to page-lock some file mappings (e.g. files on an overlay or
            # network file system; tmpfs works): fall back to a pinned bounce buffer
            self._pinned = None
            self._bounce = True

    def unpin(self):
        if self._pinned is not None:
            from . import _lib
            _lib.call('kib_host_unregister', self._pinned)
            self._pinned = None

    def store_device(self, channel, image, queue, copy_queue=None):
        """Enqueue on `queue`: reorder `image` (DeviceArray, polarizations x N x N float32)
        into FITS order on the device and copy it into the mapped plane of `channel`.
        The plane is valid once the queue has finished.

        With `copy_queue` (pinned cube only) the device-to-host copy is enqueued there, behind
        the reordering kernel, so that later work on `queue` overlaps it; the plane is valid
        once the returned event has completed (the next call waits for it before it reuses
        the staging buffer)."""
        from . import _lib, accel
        pols, height, width = image.shape
        if (pols, height, width) != self.shape[1:]:
            raise ValueError('image shape does not match the cube')
        nbytes = pols * height * width * 4
        if self._staging is None or self._staging.shape[0] < nbytes:
            self._staging = accel.DeviceArray(queue.context, (nbytes,), np.uint8)
        if self._stored is not None:
            queue.enqueue_wait_for_events([self._stored])      # staging still being read
            self._stored = None
        _lib.call('kib_fits_plane', self._staging.ptr, image.ptr, image.padded_shape[2],
                  image.padded_shape[1] * image.padded_shape[2], width, height, pols,
                  _lib.dtype_code(image.dtype), queue.stream)
        if self._pinned is not None:
            if copy_queue is not None and copy_queue is not queue:
                copy_queue.enqueue_wait_for_events([queue.enqueue_marker()])
                # In pieces: the copy engine serves its streams piece by piece, so the small
                # device-to-host reads of the next channel (PSF peak, noise estimate, CLEAN
                # results) wait for one piece instead of the whole plane.
                dst = self.data[channel].ctypes.data
                src = self._staging.ptr.value or 0
                for offset in range(0, nbytes, COPY_PIECE):
                    _lib.call('kib_memcpy_d2h_async', dst + offset, src + offset,
                              min(COPY_PIECE, nbytes - offset), copy_queue.stream)
                self._stored = copy_queue.enqueue_marker()
                return self._stored
            _lib.call('kib_memcpy_d2h_async', self.data[channel].ctypes.data, self._staging.ptr,
                      nbytes, queue.stream)
        else:
            if self._host is None or self._host.shape[0] < nbytes:
                self._host = accel.HostArray((nbytes,), np.uint8, context=queue.context)
            _lib.call('kib_memcpy_d2h_async', self._host.ctypes.data, self._staging.ptr,
                      nbytes, queue.stream)
            queue.finish()
            self.data[channel].view(np.uint8).reshape(-1)[:] = self._host[:nbytes]
        return queue.enqueue_marker()

    def flush(self):
        self.data.flush()

    def close(self):
        self.unpin()
        self.flush()
        del self.data


def _read_header_only(filename):
    header = {}
    with open(filename, 'rb') as f:
        while True:
            block = f.read(BLOCK).decode('ascii')
            for i in range(0, BLOCK, CARD):
                card = block[i:i + CARD]
                key = card[:8].strip()
                if key == 'END':
                    return header, None
                if card[8:10] != '= ':
                    continue
                value = card[10:].split(' / ')[0].strip()
                if value.startswith("'"):
                    header[key] = value[1:value.rindex("'")].rstrip()
                elif value in ('T', 'F'):
                    header[key] = value == 'T'
                else:
                    header[key] = float(value) if ('.' in value or 'E' in value) else int(value)
    return header, None
