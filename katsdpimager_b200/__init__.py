"""B200-native implementation of katsdpimager's imaging hot path.

W-projection gridding/degridding, the grid<->image FFT stage and the Hogbom CLEAN
minor cycle as hand-written sm_100a CUDA kernels behind a C ABI
(``libkatimager_b200.so``, see ``include/katimager_b200.h``), driven through the
same operation API as the reference (``*Template`` -> ``instantiate()`` -> slots /
``__call__``).  See DESIGN.md and INTEGRATION.md.
"""
__version__ = '0.1.0'
