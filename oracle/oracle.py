"""
oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py header).

`Oracle(d, p, w, min, cut, m)` builds every table of the reference's init path
(cuhe/CuHE.cu:36-50 -> initNtt / initCrt / initBarrett) with Python big ints
(pyoracle) and runs the array-sized arithmetic through the C restatement
(coracle.c, loaded with ctypes).  All arrays are numpy, layouts as in the
reference: raw u32[H][W], crt u32[L][H], ntt u64[L][N].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

from . import pyoracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile coracle.c + zzx_gmp.c (building the checker is not using it)."""
    out = os.path.join(_HERE, "_build", "libcoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("coracle.c", "zzx_gmp.c", "Makefile")]
    if (not os.path.exists(out)) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("orc_add_modP", "orc_sub_modP", "orc_mul_modP",
                     "orc_mul_modP_slow"):
            f = getattr(_LIB, name)
            f.restype = C.c_uint64
            f.argtypes = [C.c_uint64, C.c_uint64]
        _LIB.orc_ls_modP.restype = C.c_uint64
        _LIB.orc_ls_modP.argtypes = [C.c_uint64, C.c_int]
        _LIB.orc_max_threads.restype = C.c_int
        _LIB.zzx_mul_mod_batch.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def roots(N: int) -> np.ndarray:
    r = np.empty(N, dtype=np.uint64)
    lib().orc_make_roots(_p(r), C.c_int(N))
    return r


_ROOTS = {}


def _roots(N: int) -> np.ndarray:
    if N not in _ROOTS:
        _ROOTS[N] = roots(N)
    return _ROOTS[N]


def ntt_ext(x: np.ndarray, N: int) -> np.ndarray:
    """x: u32[..., >=N/2] (first N/2 of each row used) -> u64[..., N]"""
    x = np.ascontiguousarray(x, dtype=np.uint32)
    lead = x.shape[:-1]
    xs = x.reshape(-1, x.shape[-1])
    out = np.empty((xs.shape[0], N), dtype=np.uint64)
    r = _roots(N)
    for i in range(xs.shape[0]):
        row = np.ascontiguousarray(xs[i, :N // 2])
        lib().orc_ntt_ext(_p(out[i]), _p(row), C.c_int(N), _p(r))
    return out.reshape(*lead, N)


def intt_modp(X: np.ndarray, N: int, primes: Sequence[int]) -> np.ndarray:
    """X: u64[L][N] -> u32[L][N] = (N^-1 INTT(X)) % p_l"""
    X = np.ascontiguousarray(X, dtype=np.uint64).reshape(-1, N)
    out = np.empty(X.shape, dtype=np.uint32)
    r = _roots(N)
    for i in range(X.shape[0]):
        lib().orc_intt_modp(_p(out[i]), _p(X[i]), C.c_int(N), _p(r),
                            C.c_uint32(int(primes[i])))
    return out


def intt_u64(X: np.ndarray, N: int) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.uint64).reshape(-1, N)
    out = np.empty(X.shape, dtype=np.uint64)
    r = _roots(N)
    for i in range(X.shape[0]):
        lib().orc_intt_u64(_p(out[i]), _p(X[i]), C.c_int(N), _p(r))
    return out


def ntt_mul(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    y = np.ascontiguousarray(y, dtype=np.uint64)
    z = np.empty_like(x)
    lib().orc_ntt_mul(_p(z), _p(x), _p(y), C.c_size_t(x.size))
    return z


def ntt_add(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    y = np.ascontiguousarray(y, dtype=np.uint64)
    z = np.empty_like(x)
    lib().orc_ntt_add(_p(z), _p(x), _p(y), C.c_size_t(x.size))
    return z


class Oracle:
    """All tables + ops of one parameter set (the reference's global state:
    cuHE::param, crtPrime[], coeffModulus[], icrtConst[][], d_u_ntt, d_m_ntt,
    d_m_crt, h_ek)."""

    def __init__(self, d, p, w, mn, cut, m, phi: Sequence[int] | None = None):
        self.par = par = po.set_param(d, p, w, mn, cut, m)
        self.primes: List[int] = po.gen_crt_primes(par)
        self.primes_np = np.array(self.primes, dtype=np.uint32)
        self.moduli = po.gen_coeff_moduli(par, self.primes)
        self.invp = po.gen_crt_inv_primes(par, self.primes)
        self.N, self.H, self.n = par.nttLen, par.crtLen, par.modLen
        self.L0 = par.numCrtPrime
        self._icrt = {}
        self.phi = list(phi) if phi is not None else po.cyclotomic(m)
        assert len(self.phi) == self.n + 1 and self.phi[-1] == 1
        self._barrett = None
        self.ek = None

    # ---- level helpers -------------------------------------------------
    def L(self, lvl):
        return self.par._numCrtPrime(lvl)

    def W(self, lvl):
        return self.par._wordsCoeff(lvl)

    def K(self, lvl):
        return self.par._numEvalKey(lvl)

    def icrt_const(self, lvl) -> po.IcrtConst:
        if lvl not in self._icrt:
            self._icrt[lvl] = po.gen_icrt(self.par, self.primes, self.moduli, lvl)
        return self._icrt[lvl]

    # ---- raw <-> python ints --------------------------------------------
    def to_raw(self, coeffs: Sequence[int], lvl: int) -> np.ndarray:
        """z2r (cuhe/CuHE.cu:317-332): u32[H][W], BytesFromZZ per coefficient."""
        W = self.W(lvl)
        raw = np.zeros((self.H, W), dtype=np.uint32)
        for i, c in enumerate(coeffs[:self.H]):
            raw[i] = po.words_from_zz(int(c), W)
        return raw

    def from_raw(self, raw: np.ndarray) -> List[int]:
        """r2z (cuhe/CuHE.cu:333-348): first modLen coefficients."""
        return [po.zz_from_words(raw[i]) for i in range(self.n)]

    # ---- CRT / ICRT -----------------------------------------------------
    def crt(self, raw: np.ndarray, lvl: int) -> np.ndarray:
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        L, W = self.L(lvl), self.W(lvl)
        assert raw.shape == (self.H, W)
        out = np.zeros((L, self.H), dtype=np.uint32)
        lib().orc_crt(_p(out), _p(raw), C.c_int(L), C.c_int(W), C.c_int(self.n),
                      C.c_int(self.H), _p(self.primes_np))
        return out

    def icrt(self, c: np.ndarray, lvl: int) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=np.uint32)
        L, W, Wp = self.L(lvl), self.W(lvl), self.W(lvl + 1)
        ic = self.icrt_const(lvl)
        out = np.zeros((self.H, W), dtype=np.uint32)
        qp = np.ascontiguousarray(ic.qp)
        lib().orc_icrt(_p(out), _p(c), C.c_int(L), C.c_int(W), C.c_int(Wp),
                       C.c_int(self.n), C.c_int(self.H), _p(self.primes_np),
                       _p(ic.q), _p(qp), _p(ic.qpinv))
        return out

    # ---- NTT domain -------------------------------------------------------
    def ntt(self, c: np.ndarray) -> np.ndarray:
        """c2n: u32[L][H] -> u64[L][N] (cuhe/Operations.cu:394-398)"""
        return ntt_ext(c, self.N)

    def intt(self, X: np.ndarray) -> np.ndarray:
        """n2c without Barrett (cuhe/Operations.cu:420-427): low crtLen of each"""
        L = X.shape[0]
        full = intt_modp(X, self.N, self.primes[:L])
        return np.ascontiguousarray(full[:, :self.H])

    def intt_hold(self, X: np.ndarray) -> np.ndarray:
        L = X.shape[0]
        return intt_modp(X, self.N, self.primes[:L])

    # ---- Barrett ------------------------------------------------------------
    def barrett_tables(self):
        """setPolyModulus (cuhe/Operations.cu:213-238): u = x^(2n-1)/Phi and
        m' = Phi - x^n, both mod q0, pushed through crt + ntt at level 0."""
        if self._barrett is None:
            n, q0 = self.n, self.moduli[0]
            u = po.barrett_u_fast(self.phi, n)
            zm = [c % q0 for c in self.phi[:n]]          # x^n coefficient dropped
            zu = [c % q0 for c in u]
            m_crt = self.crt(self.to_raw(zm, 0), 0)
            u_crt = self.crt(self.to_raw(zu, 0), 0)
            self._barrett = dict(m_crt=m_crt, m_ntt=self.ntt(m_crt),
                                 u_ntt=self.ntt(u_crt))
        return self._barrett

    def barrett(self, hold: np.ndarray) -> np.ndarray:
        """hold: u32[L][N] -> u32[L][H] (cuhe/Operations.cu:460-501)"""
        t = self.barrett_tables()
        hold = np.ascontiguousarray(hold, dtype=np.uint32)
        L = hold.shape[0]
        out = np.zeros((L, self.H), dtype=np.uint32)
        lib().orc_barrett(_p(out), _p(hold), C.c_int(L), C.c_int(self.N),
                          C.c_int(self.H), C.c_int(self.n), _p(self.primes_np),
                          _p(t["u_ntt"]), _p(t["m_ntt"]), _p(t["m_crt"]),
                          _p(_roots(self.N)))
        return out

    def intt_mod(self, X: np.ndarray) -> np.ndarray:
        """inttMod (cuhe/Operations.cu:429-434)"""
        return self.barrett(self.intt_hold(X))

    # ---- modswitch ------------------------------------------------------------
    def modswitch(self, c: np.ndarray, lvl: int) -> np.ndarray:
        """crtModSwitch (cuhe/Operations.cu:296-303): returns u32[L-1][H]"""
        c = np.ascontiguousarray(c, dtype=np.uint32)
        L = self.L(lvl)
        out = c.copy()
        lib().orc_modswitch(_p(out), _p(c), C.c_int(L), C.c_int(self.n),
                            C.c_int(self.H), C.c_int(self.par.modMsg),
                            _p(self.primes_np), _p(self.invp))
        return np.ascontiguousarray(out[:L - 1])

    # ---- CRT-domain adds -----------------------------------------------------
    def crt_add(self, a, b):
        L = a.shape[0]
        out = np.zeros_like(a)
        lib().orc_crt_add(_p(out), _p(np.ascontiguousarray(a)),
                          _p(np.ascontiguousarray(b)), C.c_int(L), C.c_int(self.n),
                          C.c_int(self.H), _p(self.primes_np))
        return out

    def crt_add_int(self, x, a: int):
        L = x.shape[0]
        out = np.ascontiguousarray(x).copy()
        lib().orc_crt_add_int(_p(out), _p(np.ascontiguousarray(x)), C.c_uint(a),
                              C.c_int(L), C.c_int(self.H), _p(self.primes_np))
        return out

    def crt_add_nx1(self, a, scalar):
        L = a.shape[0]
        out = np.zeros_like(a)
        lib().orc_crt_add_nx1(_p(out), _p(np.ascontiguousarray(a)),
                              _p(np.ascontiguousarray(scalar, dtype=np.uint32)),
                              C.c_int(L), C.c_int(self.n), C.c_int(self.H),
                              _p(self.primes_np))
        return out

    # ---- relinearization ------------------------------------------------------
    def init_relin(self, evalkeys_raw: Sequence[np.ndarray]):
        """initRelin (cuhe/Relinearization.cu:43-56): ek[l][k][N] at level 0."""
        K0, L0 = self.par.numEvalKey, self.L0
        assert len(evalkeys_raw) == K0
        ek = np.empty((L0, K0, self.N), dtype=np.uint64)
        for k in range(K0):
            ek[:, k, :] = self.ntt(self.crt(evalkeys_raw[k], 0))
        self.ek = ek

    def relin_mac(self, raw: np.ndarray, lvl: int) -> np.ndarray:
        """relinearization() (cuhe/Relinearization.cu:76-88): raw u32[H][W] ->
        u64[L][N]."""
        L, K, W = self.L(lvl), self.K(lvl), self.W(lvl)
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        ek = np.ascontiguousarray(self.ek[:L, :K, :])
        out = np.empty((L, self.N), dtype=np.uint64)
        lib().orc_relin(_p(out), _p(raw), C.c_int(L), C.c_int(K), C.c_int(W),
                        C.c_int(self.par.logRelin), C.c_int(self.N), C.c_int(self.H),
                        _p(ek), _p(_roots(self.N)))
        return out

    def digits(self, raw: np.ndarray, lvl: int, k: int) -> np.ndarray:
        W = self.W(lvl)
        out = np.empty(self.H, dtype=np.uint32)
        lib().orc_digits(_p(out), _p(np.ascontiguousarray(raw)), C.c_int(self.H),
                         C.c_int(W), C.c_int(self.par.logRelin), C.c_int(k))
        return out

    # ---- whole multiply ---------------------------------------------------------
    def mul_raw_to_crt(self, a_raw, b_raw, lvl: int) -> np.ndarray:
        t = self.barrett_tables()
        L, W = self.L(lvl), self.W(lvl)
        out = np.zeros((L, self.H), dtype=np.uint32)
        lib().orc_mul_raw_to_crt(
            _p(out), _p(np.ascontiguousarray(a_raw)), _p(np.ascontiguousarray(b_raw)),
            C.c_int(L), C.c_int(W), C.c_int(self.N), C.c_int(self.H), C.c_int(self.n),
            _p(self.primes_np), _p(t["u_ntt"]), _p(t["m_ntt"]), _p(t["m_crt"]),
            _p(_roots(self.N)))
        return out

    def mul_raw_batch(self, a_raw: np.ndarray, b_raw: np.ndarray, lvl: int) -> np.ndarray:
        """batch of products RAW -> RAW with every (polynomial, residue) pair an OpenMP task
        (the CPU arm of bench.py).  a_raw, b_raw: u32[batch][H][W]"""
        t = self.barrett_tables()
        L, W, Wp = self.L(lvl), self.W(lvl), self.W(lvl + 1)
        ic = self.icrt_const(lvl)
        a_raw = np.ascontiguousarray(a_raw, dtype=np.uint32)
        b_raw = np.ascontiguousarray(b_raw, dtype=np.uint32)
        batch = a_raw.shape[0]
        out = np.zeros_like(a_raw)
        lib().orc_mul_raw_batch(
            _p(out), _p(a_raw), _p(b_raw), C.c_int(batch), C.c_int(L), C.c_int(W), C.c_int(Wp), C.c_int(self.N),
            C.c_int(self.H), C.c_int(self.n), _p(self.primes_np), _p(t["u_ntt"]), _p(t["m_ntt"]), _p(t["m_crt"]),
            _p(_roots(self.N)), _p(ic.q), _p(np.ascontiguousarray(ic.qp)), _p(ic.qpinv))
        return out

    def inverse_series(self) -> np.ndarray:
        """rev(Phi)^-1 mod x^(m-n) over Z (Phi monic => integer coefficients; small for cyclotomic Phi,
        where it is -(x^m - 1)/Phi up to the truncation).  int64[m-n]."""
        if getattr(self, "_invser", None) is None:
            n, d = self.n, self.par.mSize - self.n
            rev = np.array(self.phi[::-1], dtype=np.int64)               # rev[0] = 1
            inv = np.zeros(d, dtype=np.int64)
            inv[0] = 1
            for k in range(1, d):
                j = min(k, n)
                inv[k] = -int(np.dot(rev[1:j + 1], inv[k - 1::-1][:j] if k - 1 >= 0 else inv[:0]))
            assert np.abs(inv).max() < 2**31, "inverse series is not small: not a cyclotomic modulus?"
            self._invser = inv
        return self._invser

    def mul_raw_batch_zzx(self, a_raw: np.ndarray, b_raw: np.ndarray, lvl: int) -> np.ndarray:
        """The reference's NTL host path (t = a*b; t %= polyMod; coeffReduce -- examples/DHS/DHS.cu:219-221)
        restated on GMP (oracle/zzx_gmp.c): Kronecker product, division by Phi_m through the inverse
        series, coefficients mod q_lvl; one product per OpenMP thread.  RAW in / RAW out like mul_raw_batch."""
        W = self.W(lvl)
        a_raw = np.ascontiguousarray(a_raw, dtype=np.uint32)
        b_raw = np.ascontiguousarray(b_raw, dtype=np.uint32)
        assert a_raw.shape[1:] == (self.H, W) and a_raw.shape == b_raw.shape
        out = np.zeros_like(a_raw)
        phi = np.array(self.phi, dtype=np.int64)
        qw = np.ascontiguousarray(po.words_from_zz(self.moduli[lvl], W), dtype=np.uint32)
        rc = lib().zzx_mul_mod_batch(_p(out), _p(a_raw), _p(b_raw), C.c_int(a_raw.shape[0]), C.c_int(self.H), C.c_int(W),
                                     C.c_int(self.n), C.c_int(self.par.mSize), _p(phi), _p(self.inverse_series()), _p(qw))
        if rc != 0:
            raise RuntimeError("libgmp.so.10 could not be loaded for the GMP CPU path")
        return out

    def mul_exact(self, a: Sequence[int], b: Sequence[int], lvl: int) -> List[int]:
        """(a*b mod Phi_m) mod q_lvl with big ints / GMP -- the NTL host path
        of examples/DHS/DHS.cu:219-221."""
        return po.mul_mod(a, b, self.phi, self.par.mSize, self.moduli[lvl])
