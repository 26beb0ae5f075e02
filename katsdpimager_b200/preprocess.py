"""Visibility quantisation into the gridder's input records (numpy).

This is the producer side of the hot path's input contract: the reference does
it in C++ (``katsdpimager/preprocess.cpp:401-513`` quantise, ``:335-397``
compress + bucket by W slice) and stores structured records with fields
``uv, sub_uv, w_plane, weights, vis`` (``preprocess.py:42-56``,
``preprocess.cpp:39-52``).  Only the parts the imaging path needs are restated
here (no Mueller/feed-angle transform, no HDF5 store): enough to feed the
kernels, the oracle and the benchmarks with correctly formed records.
"""
import numpy as np


def make_dtype(num_polarizations):
    """Record layout of ``vis_t<P>`` minus ``w_slice`` (preprocess.cpp:39-52)."""
    P = num_polarizations
    return np.dtype(dict(
        names=['uv', 'sub_uv', 'w_plane', 'weights', 'vis'],
        formats=[('i2', (2,)), ('i2', (2,)), 'i2', ('f4', (P,)), ('c8', (P,))],
        offsets=[0, 4, 8, 12, 12 + 4 * P],
        itemsize=12 + 12 * P))


def subpixel_coord(x, oversample):
    """``subpixel_coord`` (preprocess.cpp:313-323): floor semantics for negative x."""
    xs = np.floor(x.astype(np.float32) * np.float32(oversample)).astype(np.int32)
    return (xs // oversample).astype(np.int16), (xs % oversample).astype(np.int16)


def quantise(uvw, weights, vis, image_parameters, grid_parameters):
    """Quantise one channel (preprocess.cpp:435-507).

    uvw : (N, 3) float32 metres;  weights : (N, P) float32;  vis : (N, P) complex64
    (already in the image's Stokes frame).  Visibilities are pre-multiplied by
    their weights, w < 0 is flipped with conjugation, and samples with a zero
    weight are dropped.  Returns (records, w_slice).
    """
    ip, gp = image_parameters, grid_parameters
    P = vis.shape[1]
    uvw = np.asarray(uvw, np.float32)
    weights = np.asarray(weights, np.float32)
    vis = np.asarray(vis, np.complex64)
    keep = ~np.any(weights == 0, axis=1)
    uvw, weights, vis = uvw[keep], weights[keep], vis[keep]
    flip = uvw[:, 2] < 0
    uvw = np.where(flip[:, None], -uvw, uvw)
    vis = np.where(flip[:, None], np.conj(vis), vis)
    vis = (vis * weights).astype(np.complex64)
    bad = ~(np.isfinite(vis.real) & np.isfinite(vis.imag))
    vis[bad] = 0
    weights = np.where(bad, np.float32(0), weights)

    uv_scale = np.float32(1.0) / np.float32(ip.cell_size)
    w_scale = np.float32((np.float32(gp.w_slices) - np.float32(0.5)) * np.float32(gp.w_planes)
                         / np.float32(gp.fixed.max_w))
    max_slice_plane = gp.w_slices * gp.w_planes - 1
    w = np.trunc(uvw[:, 2] * w_scale + np.float32(gp.w_planes * 0.5))
    w_slice_plane = np.minimum(w.astype(np.int64), max_slice_plane)
    out = np.zeros(len(uvw), make_dtype(P)).view(np.recarray)
    out.uv[:, 0], out.sub_uv[:, 0] = subpixel_coord(uvw[:, 0] * uv_scale, gp.fixed.oversample)
    out.uv[:, 1], out.sub_uv[:, 1] = subpixel_coord(uvw[:, 1] * uv_scale, gp.fixed.oversample)
    out.w_plane = w_slice_plane % gp.w_planes
    out.weights = weights
    out.vis = vis
    return out, (w_slice_plane // gp.w_planes).astype(np.int16)


def compress(records, w_slice):
    """Merge adjacent records with identical coordinates (preprocess.cpp:335-373)."""
    # elements whose first weight is zero are flagged (a NaN visibility squashed by quantise)
    # and skipped before runs are detected (preprocess.cpp:341-352)
    keep = records.weights[:, 0] != 0
    if not np.all(keep):
        records, w_slice = records[keep].view(np.recarray), w_slice[keep]
    if len(records) == 0:
        return records, w_slice
    key = np.empty((len(records), 6), np.int16)
    key[:, 0:2] = records.uv
    key[:, 2:4] = records.sub_uv
    key[:, 4] = records.w_plane
    key[:, 5] = w_slice
    new_run = np.ones(len(records), bool)
    new_run[1:] = np.any(key[1:] != key[:-1], axis=1)
    starts = np.flatnonzero(new_run)
    out = records[starts].copy().view(np.recarray)
    out.vis = np.add.reduceat(records.vis, starts, axis=0)
    out.weights = np.add.reduceat(records.weights, starts, axis=0)
    return out, w_slice[starts]


def bucket_by_slice(records, w_slice, num_slices):
    """Stable split into per-W-slice runs (preprocess.cpp:375-397)."""
    order = np.argsort(w_slice, kind='stable')
    sorted_records = records[order].view(np.recarray)
    counts = np.bincount(w_slice, minlength=num_slices)
    bounds = np.concatenate(([0], np.cumsum(counts)))
    return [sorted_records[bounds[s]:bounds[s + 1]] for s in range(num_slices)]


class VisibilityReaderMem:
    """In-memory equivalent of ``VisibilityReaderMem`` (preprocess.py:390-420):
    per (channel, w_slice) record arrays with ``iter_slice`` chunking."""

    def __init__(self, slices_per_channel):
        self._data = slices_per_channel

    @property
    def num_channels(self):
        return len(self._data)

    def num_w_slices(self, channel):
        return len(self._data[channel])

    def len(self, channel, w_slice):
        return len(self._data[channel][w_slice])

    def iter_slice(self, channel, w_slice, block_size=None):
        data = self._data[channel][w_slice]
        if block_size is None:
            block_size = max(1, len(data))
        for start in range(0, len(data), block_size):
            yield data[start:start + block_size]
