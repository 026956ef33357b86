"""cuhe_b200 -- B200-native engine for the cuHE hot path.

`csrc/` holds the hand-written sm_100a kernels and the C ABI
(include/cuhe_b200.h -> libcuhe_b200.so); this package is the host-side mirror
of the reference's public interface (cuhe/CuHE.h) over that ABI.  There is no
CPU fallback: without the CUDA library every call raises CuHEError."""
from ._lib import CuHEError, LIB_PATH, SYMBOLS, cuhe_params, load_library  # noqa: F401
from .api import (  # noqa: F401
    CuCtxt, CuPolynomial, CuPtxt, GlobalParameters, cAnd, cNot, cXor, copy, copyTo, crtPrimes, ctx,
    initCuHE, initRelinearization, initRelinearizationRaw, launch_count, moveTo, mulZZX, multiGPUs,
    numGPUs, param, resetParameters, setParameters, startAllocator, stopAllocator,
)
