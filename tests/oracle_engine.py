"""The cuHE public interface (cuhe/CuHE.h:46-208) served by the CPU ORACLE instead of the GPU, so
that caller-level code (tests/dhs_host.py, tests/prince_he.py) can be run against the oracle alone.
Used to pin the oracle to the reference's known answer (the homomorphic PRINCE vector) without a GPU,
and to cross-check the GPU path ciphertext-for-ciphertext.

Test infrastructure only -- never imported by the product (tests/test_abi.py enforces that)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from oracle import oracle as orc
from oracle.oracle import Oracle


class OracleEngine:
    """Module-like object: setParameters / initCuHE / initRelinearization / mulZZX / CuCtxt /
    cAnd / cXor / cNot / copy, with the domain rules of cuhe/CuHE.cu:81-268,272-606."""

    class Error(RuntimeError):
        pass

    def __init__(self):
        self.o = None
        self.param = None
        self._ps = None
        eng = self

        class CuCtxt:
            def __init__(self):
                self.reset()

            def reset(self):
                self.level_, self.domain_, self.isProd_ = -1, 0, False
                self.z, self.c, self.x = [], None, None           # ZZX, CRT u32[L][H], NTT u64[L][N]

            def setLevel(self, lvl, dev, val):
                self.reset()
                self.level_, self.domain_, self.z = lvl, 0, list(val)

            def level(self):
                return self.level_

            def domain(self):
                return self.domain_

            def logq(self):
                return eng.param._logCoeff(self.level_)

            def isProd(self):
                return self.isProd_

            def zRep(self):
                return self.z

            def _multi(self):
                return self.logq() > eng.param.logCrtPrime

            def x2c(self):
                o = eng.o
                if self.domain_ == 2:
                    return
                if self.domain_ == 0:
                    raw = o.to_raw(self.z, self.level_)
                    self.c = o.crt(raw, self.level_) if self._multi() else np.ascontiguousarray(raw.reshape(1, o.H))
                    self.z = []
                elif self.domain_ == 3:
                    self.c = o.intt_mod(self.x) if self.isProd_ else o.intt(self.x)
                    self.isProd_, self.x = False, None
                self.domain_ = 2

            def x2n(self):
                if self.domain_ == 3:
                    return
                self.x2c()
                self.x = eng.o.ntt(self.c)
                self.c, self.domain_ = None, 3

            def _raw(self):
                o = eng.o
                self.x2c()
                return o.icrt(self.c, self.level_) if self._multi() else np.ascontiguousarray(self.c.reshape(o.H, 1))

            def x2z(self):
                if self.domain_ == 0:
                    return
                self.z = eng.o.from_raw(self._raw())
                self.c, self.domain_ = None, 0

            def relin(self):
                o = eng.o
                if o.ek is None:
                    raise OracleEngine.Error("initRelinearization has not been called")
                raw = self._raw()
                self.x = o.relin_mac(raw, self.level_)
                self.c, self.domain_, self.isProd_ = None, 3, True
                self.x2c()

            def dropToLevel(self, lvl):
                if lvl < self.level_:
                    raise OracleEngine.Error("Error: dropToLevel cannot raise the modulus!")
                if lvl == self.level_:
                    return
                self.x2c()
                self.c = np.ascontiguousarray(self.c[:eng.param._numCrtPrime(lvl)])
                self.level_ = lvl

            def modSwitch(self):
                par = eng.param
                if self.logq() < par.logCoeffMin + par.logCoeffCut:
                    raise OracleEngine.Error("Error: Cannot do modSwitch on last level!")
                self.x2c()
                self.c = eng.o.modswitch(self.c, self.level_)
                self.level_ += 1

        self.CuCtxt = CuCtxt

    # ---- set-up ------------------------------------------------------------------------------
    def resetParameters(self):
        self.o = self.param = None

    def multiGPUs(self, n):
        pass

    def setParameters(self, d, p, w, mn, cut, m):
        from oracle import pyoracle as po
        self._ps = (d, p, w, mn, cut, m)
        self.param = po.set_param(d, p, w, mn, cut, m)

    def initCuHE(self, phi: Sequence[int]) -> List[int]:
        self.o = Oracle(*self._ps, phi=list(phi))
        self.param = self.o.par
        self.o.barrett_tables()
        return list(self.o.moduli[:self.param.depth])

    def crtPrimes(self):
        return list(self.o.primes)

    def initRelinearization(self, evalkey):
        self.o.init_relin([self.o.to_raw(ek, 0) for ek in evalkey])

    # ---- operations --------------------------------------------------------------------------
    def mulZZX(self, a, b, lvl, dev=0):
        o = self.o
        if o.par._logCoeff(lvl) > o.par.logCrtPrime:
            return o.from_raw(o.icrt(o.mul_raw_to_crt(o.to_raw(a, lvl), o.to_raw(b, lvl), lvl), lvl))
        x, y = self.CuCtxt(), self.CuCtxt()
        x.setLevel(lvl, dev, a)
        y.setLevel(lvl, dev, b)
        x.x2n()
        y.x2n()
        self.cAnd(x, x, y)
        x.x2z()
        return x.zRep()

    def copy(self, dst, src):
        if dst is src:
            return
        dst.reset()
        dst.level_, dst.domain_, dst.isProd_ = src.level_, src.domain_, src.isProd_
        dst.z = list(src.z)
        dst.c = None if src.c is None else src.c.copy()
        dst.x = None if src.x is None else src.x.copy()

    def cAnd(self, out, a, b):
        if a.domain_ != 3 or b.domain_ != 3:
            raise OracleEngine.Error("Error: Multiplication of non-NTT domain!")
        if a.level_ != b.level_:
            raise OracleEngine.Error("Error: Multiplication of different levels!")
        x = orc.ntt_mul(a.x, b.x)
        lvl = a.level_
        if out is not a:
            out.reset()
        out.level_, out.domain_, out.x, out.isProd_ = lvl, 3, x, True

    def cXor(self, out, a, b):
        if a.level_ != b.level_:
            raise OracleEngine.Error("Error: Addition of different levels!")
        if a.domain_ != b.domain_ or a.domain_ not in (2, 3):
            raise OracleEngine.Error("Error: Addition of non-CRT-nor-NTT domain!")
        lvl, dom = a.level_, a.domain_
        if dom == 2:
            res, prod = self.o.crt_add(a.c, b.c), False
        else:
            res, prod = orc.ntt_add(a.x, b.x), (a.isProd_ or b.isProd_)
        if out is not a:
            out.reset()
        out.level_, out.domain_, out.isProd_ = lvl, dom, prod
        if dom == 2:
            out.c = res
        else:
            out.x = res

    def cNot(self, out, a):
        if a.domain_ != 2:
            raise OracleEngine.Error("Error: cNot of non-CRT domain!")
        res = self.o.crt_add_int(a.c, self.param.modMsg - 1)
        lvl = a.level_
        if out is not a:
            out.reset()
        out.level_, out.domain_, out.c = lvl, 2, res
