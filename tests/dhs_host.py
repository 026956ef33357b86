"""Host half of the DHS/LTV scheme, written against the cuhe_b200 Python mirror,
for the integration test that mirrors examples/DHS/simple_DHS.cu:49-163
(decrypt o (cXor | cNot | cAnd + relin + modSwitch) o encrypt == plaintext op).

Test infrastructure only: it restates examples/DHS/DHS.cu:206-385 (keyGen,
encrypt, decrypt, genEk) with Python integers; NTL's ZZ_pE inverse
(DHS.cu:361-375) is replaced by an extended Euclid in Z_p[x] per CRT prime
(numpy) followed by a coefficient-wise CRT.  All ring products go through the
GPU (`mulZZX`), exactly as the reference's scheme code does."""
from __future__ import annotations

import random
from typing import List

import numpy as np


def _poly_trim(a: np.ndarray) -> np.ndarray:
    nz = np.nonzero(a)[0]
    return a[: nz[-1] + 1] if len(nz) else a[:0]


def poly_inverse_mod(f: List[int], phi: List[int], p: int) -> List[int]:
    """f^-1 in Z_p[x]/(phi), phi monic, p prime; raises if not invertible."""
    r0 = _poly_trim(np.array([c % p for c in phi], dtype=np.int64))
    r1 = _poly_trim(np.array([c % p for c in f], dtype=np.int64))
    t0 = np.zeros(1, dtype=np.int64)
    t1 = np.ones(1, dtype=np.int64)
    n = len(phi) - 1
    while len(r1) > 0:
        if len(r1) == 1:
            inv = pow(int(r1[0]), -1, p)
            out = (t1 * inv) % p
            res = np.zeros(n, dtype=np.int64)
            res[: len(out)] = out[:n]
            return [int(v) for v in res]
        # r0 = q*r1 + r2 ; t2 = t0 - q*t1   (long division, a few terms of q per step)
        lead_inv = pow(int(r1[-1]), -1, p)
        r = r0.copy()
        tq = np.zeros(max(len(r0) - len(r1) + 1, 1), dtype=np.int64)
        while len(r) >= len(r1) and len(r) > 0:
            d = len(r) - len(r1)
            c = (int(r[-1]) * lead_inv) % p
            tq[d] = c
            r[d:] = (r[d:] - c * r1) % p
            r = _poly_trim(r)
        # t2 = t0 - tq*t1
        prod = np.zeros(len(tq) + len(t1) - 1, dtype=np.int64)
        for d in np.nonzero(tq)[0]:
            prod[d: d + len(t1)] = (prod[d: d + len(t1)] + int(tq[d]) * t1) % p
        m = max(len(t0), len(prod))
        t2 = np.zeros(m, dtype=np.int64)
        t2[: len(t0)] += t0
        t2[: len(prod)] -= prod
        t2 %= p
        r0, r1, t0, t1 = r1, r, t1, t2
    raise ArithmeticError("polynomial is not invertible")


def _inverse_job(args):
    f, phi, p = args
    try:
        return poly_inverse_mod(f, phi, p)
    except ArithmeticError:
        return None


def poly_inverse_all(f: List[int], phi: List[int], primes: List[int]) -> List[List[int]]:
    """f^-1 modulo every CRT prime; the primes are independent, so large rings (Prince: n = 16384,
    25 primes, ~4 s each) are spread over host processes."""
    jobs = [(f, phi, p) for p in primes]
    if len(phi) < 10000 or len(primes) < 4:
        res = [_inverse_job(j) for j in jobs]
    else:
        import multiprocessing as mp
        import os
        from concurrent.futures import ProcessPoolExecutor
        workers = min(len(primes), max(1, (os.cpu_count() or 2) - 1))
        with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_inverse_job, jobs))
    if any(r is None for r in res):
        raise ArithmeticError("polynomial is not invertible")
    return res


def crt_combine(residues: List[List[int]], primes: List[int]) -> List[int]:
    q = 1
    for p in primes:
        q *= p
    out = [0] * len(residues[0])
    for r, p in zip(residues, primes):
        m = q // p
        e = m * pow(m % p, -1, p)
        for i, v in enumerate(r):
            out[i] += v * e
    return [v % q for v in out]


class DHS:
    """CuDHS (examples/DHS/DHS.h:46-108) over the cuhe_b200 host mirror."""

    B = 1                                                   # examples/DHS/DHS.h:44

    def __init__(self, ch, d, p, w, mn, cut, m, phi, seed=1):
        self.ch = ch
        self.rng = random.Random(seed)
        ch.resetParameters()
        ch.multiGPUs(1)
        ch.setParameters(d, p, w, mn, cut, m)
        self.par = ch.param
        self.phi = list(phi)
        self.coeffMod = ch.initCuHE(self.phi)               # DHS.cu:34-55
        self.primes = ch.crtPrimes()
        self.n = self.par.modLen
        self.keygen()

    # DHS.cu:352-357
    def sample(self):
        return [self.rng.randint(-self.B, self.B) for _ in range(self.n)]

    def reduce(self, x, lvl):
        q = self.coeffMod[lvl]
        return [c % q for c in x]

    def mul(self, a, b, lvl):
        return self.ch.mulZZX(self.reduce(a, lvl), self.reduce(b, lvl), lvl, 0)

    # DHS.cu:311-344 genPkSk, 345-370 genEk
    def keygen(self):
        p = self.par.modMsg
        L0 = self.par.numCrtPrime
        while True:
            ft = self.sample()
            f = [p * c for c in ft]
            f[0] += 1
            try:
                invs = poly_inverse_all(f, self.phi, self.primes[:L0])
                break
            except ArithmeticError:
                continue
        f_inv = crt_combine(invs, self.primes[:L0])
        one = self.mul(f, f_inv, 0)
        assert one[0] == 1 and not any(one[1:]), "f * f^-1 != 1 (through the GPU multiply)"
        g = self.sample()
        pk0 = [p * c for c in self.mul(g, f_inv, 0)]
        self.sk = [self.reduce(f, i) for i in range(self.par.depth)]
        self.pk = [self.reduce(pk0, i) for i in range(self.par.depth)]
        if self.par.logRelin > 0:
            wbase = 1 << self.par.logRelin
            tw = 1
            eks = []
            for _ in range(self.par.numEvalKey):
                s, e = self.sample(), self.sample()
                t = self.mul(self.pk[0], s, 0)
                ek = [t[i] + p * e[i] + self.sk[0][i] * tw for i in range(self.n)]
                eks.append(self.reduce(ek, 0))
                tw *= wbase
            self.ch.initRelinearization(eks)                # DHS.cu:369

    # DHS.cu:206-221
    def encrypt(self, msg, lvl):
        p = self.par.modMsg
        s, e = self.sample(), self.sample()
        t = self.mul(self.pk[lvl], s, lvl)
        return self.reduce([t[i] + p * e[i] + (msg[i] if i < len(msg) else 0) for i in range(self.n)], lvl)

    # DHS.cu:222-247
    def decrypt(self, c, lvl):
        q, p = self.coeffMod[lvl], self.par.modMsg
        t = self.mul(c, self.sk[lvl], lvl)
        out = []
        for x in t:
            if x > (q - 1) // 2:
                x -= q
            out.append(x % p)
        return out
