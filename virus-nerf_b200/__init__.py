"""virus-nerf_b200 -- B200-native (sm_100a) implementation of the VIRUS-NeRF Instant-NGP
training / render hot path behind the reference's ``modules/*`` Python API.

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI (``include/virusnerf.h``), built into
``lib/libvirusnerf_sm100.so``; ``_lib.py`` the ctypes binding; ``modules/`` the host-side mirror
of the reference's module interface; ``engine.py`` the train-step driver used by ``bench.py``.
There is no CPU fallback: every op raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
