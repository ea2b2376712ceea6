"""laboetie_b200 -- B200 (sm_100a) implementation of laboetie's time-stepping hot path.

The product is the C-ABI shared library built from csrc/ (include/laboetie_gpu.h).
This package is the thin host side: `api` binds the C ABI with ctypes, `driver`
mirrors the reference's two phase drivers (equilibration.f90, drop_tracers.f90)
on top of it, `slab` holds the z-slab plumbing for one-process-per-GPU runs.
There is no CPU fallback: importing works anywhere, but every compute entry
point needs the CUDA library and a GPU and fails loudly otherwise.
"""
from .api import LaboetieGPU, LbgError, lib_path, load_library  # noqa: F401

__version__ = "0.1.0"
