"""Host-side helpers a caller of the engine needs but that are not GPU work:
the cyclotomic polynomial Phi_m (what examples/DHS/DHS.cu:283-309, genPolyMod_,
builds with NTL before calling initCuHE)."""
from __future__ import annotations

from typing import List

import numpy as np


def _divide_exact(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    """num / den for integer polynomials (ascending, int64) with den[0] = +-1 and
    an exact quotient: forward substitution from the constant term."""
    dn = len(den) - 1
    qn = len(num) - 1 - dn
    work = num.astype(np.int64).copy()
    q = np.zeros(qn + 1, dtype=np.int64)
    d0 = int(den[0])
    if d0 not in (1, -1):
        raise ValueError("divisor must have constant term +-1")
    for i in range(qn + 1):
        c = int(work[i]) * d0
        q[i] = c
        if c:
            hi = min(i + dn + 1, len(work))
            work[i:hi] -= c * den[: hi - i]
    return q


def cyclotomic(m: int) -> List[int]:
    """Phi_m(x) as ascending integer coefficients, via
    Phi_{np}(x) = Phi_n(x^p) / Phi_n(x)  (p prime, p does not divide n)
    Phi_{np}(x) = Phi_n(x^p)              (p divides n)."""
    if m < 1:
        raise ValueError("m must be positive")
    phi = np.array([-1, 1], dtype=np.int64)
    rest, p = m, 2
    while rest > 1:
        if rest % p == 0:
            first = True
            while rest % p == 0:
                rest //= p
                up = np.zeros((len(phi) - 1) * p + 1, dtype=np.int64)
                up[::p] = phi
                phi = _divide_exact(up, phi) if first else up
                first = False
        p += 1
    return [int(c) for c in phi]
