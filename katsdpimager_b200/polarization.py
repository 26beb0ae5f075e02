"""Stokes parameter enumeration (CASA values, reference katsdpimager/polarization.py:32-49)."""
STOKES_I = 1
STOKES_Q = 2
STOKES_U = 3
STOKES_V = 4
STOKES_IQUV = [STOKES_I, STOKES_Q, STOKES_U, STOKES_V]
STOKES_NAMES = [None, 'I', 'Q', 'U', 'V', 'RR', 'RL', 'LR', 'LL', 'XX', 'XY', 'YX', 'YY']
