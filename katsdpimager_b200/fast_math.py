"""Range-reduced complex exponential (reference katsdpimager/fast_math.py:7-16)."""
import numpy as np


def expj2pi(x):
    """exp(2j*pi*x) for real x, with x reduced to [-0.5, 0.5] first."""
    x = np.asarray(x)
    turns = x - np.rint(x)
    angle = (2 * np.pi) * turns
    out = np.empty(x.shape, np.complex64 if x.dtype == np.float32 else np.complex128)
    np.cos(angle, out=out.real)
    np.sin(angle, out=out.imag)
    return out
