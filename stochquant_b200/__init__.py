"""stochquant_b200 -- B200-native (sm_100a) Langevin hot path of SebTanz/StochQuant.

The product is the C-ABI CUDA library ``libsq.so`` (include/sq.h) plus the drop-in host
binary ``tauhost.o``.  This package is the thin ctypes view of that C-ABI used by the
tests and by bench.py; it mirrors the reference's host-side vocabulary (frames, loops,
runs, stable -- /root/reference/tauhost.c:479-560).  There is no CPU fallback: importing
works anywhere, but creating a context without the CUDA library or a GPU raises.
"""
from .lib import (  # noqa: F401
    SQ_KERNEL_COMPAT1D, SQ_KERNEL_LATTICE, SQ_REAL_F32, SQ_REAL_F64, SQ_MATH_ACCURATE,
    SQ_MATH_FAST, SQ_POT_HARMONIC, SQ_POT_DOUBLEWELL, SQ_POT_PHI4, SQ_FLAG_NO_OBSERVABLES,
    SQ_FLAG_FORCE_STREAMING, SQ_FLAG_GENERIC_KERNEL, SQ_FLAG_ROWBLOCK_KERNEL,
    SqError, SqParams, SqObs, SqRngEntry, SqFrameRec, SQ_FRAMES_MAX, Context, Session, rng_resolve, load, library_path, build,
    exported_symbols,
)

__all__ = ["Context", "Session", "SqError", "load", "build", "library_path", "rng_resolve"]
