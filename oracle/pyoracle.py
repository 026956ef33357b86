"""
oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Big-integer CPU restatement of the cuHE hot path, written from the reference's
arithmetic definitions.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product
(cuhe_b200/) must never import, link or execute anything under oracle/.

Every function cites the reference file:line (relative to /root/reference) it
restates.  Pinning status: the reference cannot be built here (no NTL/GMP
headers, texture references rejected by nvcc 12.9), it stores no golden
vectors and seeds everything with time(NULL); what its own tests pin is
  * tests/test_ModP.cu:57-137  -- mod-P primitives == big-int arithmetic mod P
  * tests/test_ntt.cu:38-64    -- ext-NTT == O(N^2) DFT with g, w0=g^(65536/N)
  * examples/Prince/Prince.cu:96,109-144 -- the only FIXED known answer: homomorphic PRINCE of
    0^64 under k0 = 1^64, k1 = 0^64 decrypts to 9fb51935fc3df524 (+ 12 per-round states)
The first two are checked for this oracle in tests/test_oracle.py.  The third is checked in
tests/test_prince_circuit.py: the whole homomorphic PRINCE (1920 cAnd, 1152 relin, 2688 modSwitch,
24 levels, real DHS keys at (25,2,16,25,25,21845)) run on THIS oracle behind the cuHE interface
(tests/oracle_engine.py) decrypts to the reference's vector and round states -- log of the run in
tests/golden/prince_kat_oracle.log (opt-in test, CUHE_B200_SLOW=1, ~11 min of CPU).  That run goes
through every domain of the path (CRT, NTT, Barrett, ICRT, relin, modswitch), so the oracle is
PINNED end to end by a reference known answer; a wrong intermediate table (CRT primes, ICRT
constants, modswitch rounding, Barrett polynomials) destroys the decryption.  What no reference
test pins is the exact VALUE of intermediate ciphertext words (the reference stores none and seeds
keys with time(NULL)); those rest on the literal restatement here plus exact ring arithmetic
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

# cuhe/ModP.h:25,33 ; cuhe/Base.cu:64-65
P = 0xFFFFFFFF00000001
G = 15893793146607301539


# --------------------------------------------------------------------------
# small number theory helpers (NTL stand-ins: ProbPrime, NumBits, SqrRoot ...)
# --------------------------------------------------------------------------
def num_bits(x: int) -> int:
    """NTL NumBits(|x|); NumBits(0) == 0."""
    return abs(int(x)).bit_length()


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin, exact for n < 3.3e24 (stands in for
    NTL ProbPrime(n, 10), cuhe/Operations.cu:45,58,71)."""
    n = int(n)
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def next_prime(n: int) -> int:
    """NTL NextPrime(n): smallest prime >= n."""
    n = max(int(n), 2)
    while not is_prime(n):
        n += 1
    return n


def euler_totient(x: int) -> int:
    """cuhe/Parameters.cu:34-51 (literal: returns x itself for x < 3)."""
    if x < 3:
        return x
    res = x
    t = 2
    while x != 1:
        hit = False
        while math.gcd(x, t) == t:
            x //= t
            hit = True
        if hit:
            res = res * (t - 1) // t
        t = next_prime(t + 1)
    return res


def bytes_from_zz(a: int, n: int) -> bytes:
    """NTL BytesFromZZ: low-order n bytes of |a|, little endian."""
    a = abs(int(a))
    return (a & ((1 << (8 * n)) - 1)).to_bytes(n, "little")


def words_from_zz(a: int, nwords: int) -> np.ndarray:
    return np.frombuffer(bytes_from_zz(a, 4 * nwords), dtype="<u4").copy()


def zz_from_words(w: Sequence[int]) -> int:
    return int.from_bytes(np.asarray(w, dtype="<u4").tobytes(), "little")


# --------------------------------------------------------------------------
# Parameters  (cuhe/Parameters.h:34-64, cuhe/Parameters.cu:53-145)
# --------------------------------------------------------------------------
@dataclass
class Params:
    depth: int = 0
    modMsg: int = 0
    logRelin: int = 0
    logCoeffMin: int = 0
    logCoeffCut: int = 0
    mSize: int = 0
    logCoeffMax: int = 0
    modLen: int = 0
    modLen2: int = 0
    rawLen: int = 0
    crtLen: int = 0
    nttLen: int = 0
    logMsg: int = 0
    wordsMsg: int = 0
    numEvalKey: int = 0
    logCrtPrime: int = 0
    numCrtPrime: int = 0

    # cuhe/Parameters.cu:107-145
    def _numCrtPrime(self, lvl: int) -> int:
        if lvl == -1:
            return 1
        if lvl >= self.depth:
            raise ValueError(f"numCrtPrime(lvl) has lvl: {lvl}")
        return self.numCrtPrime - lvl

    def _logCoeff(self, lvl: int) -> int:
        if lvl == -1:
            return self.logMsg
        if lvl < self.depth:
            return self.logCoeffMax - lvl * self.logCoeffCut
        if lvl == self.depth:
            return self.logCoeffMin - self.logCrtPrime
        raise ValueError("lvl cannot be more than depth")

    def _wordsCoeff(self, lvl: int) -> int:
        t = (self._logCoeff(lvl) + 31) // 32
        return t if t > 1 else 1

    def _numEvalKey(self, lvl: int) -> int:
        return (self._logCoeff(lvl) + self.logRelin - 1) // self.logRelin

    def _getLevel(self, logq: int) -> int:
        if logq >= self.logCoeffMin:
            return (self.logCoeffMax - logq) // self.logCoeffCut
        return -1


def set_param(d: int, p: int, w: int, mn: int, cut: int, m: int) -> Params:
    """cuhe/Parameters.cu:53-85, line by line."""
    par = Params()
    par.depth, par.modMsg, par.logRelin = d, p, w
    par.logCoeffMin, par.logCoeffCut, par.mSize = mn, cut, m
    par.logCoeffMax = par.logCoeffMin + par.logCoeffCut * (par.depth - 1)
    par.modLen = euler_totient(par.mSize)
    par.modLen2 = 1 << num_bits(par.modLen - 1)
    if par.modLen2 < 8192:
        par.modLen2 = 8192
    par.rawLen = par.crtLen = par.modLen2
    par.nttLen = 2 * par.modLen2
    par.logMsg = num_bits(par.modMsg - 1)
    par.wordsMsg = (par.logMsg + 31) // 32
    par.numEvalKey = ((par.logCoeffMax + par.logRelin - 1) // par.logRelin
                      if par.logRelin != 0 else 0)
    par.logCrtPrime = num_bits(math.isqrt(P // par.modLen))
    par.numCrtPrime = (par.logCoeffMin + par.logCrtPrime - 1) // par.logCrtPrime
    par.logCrtPrime = 0
    while par.logCrtPrime * par.numCrtPrime < par.logCoeffMin:
        par.logCrtPrime += 1
    par.numCrtPrime += par.depth - 1
    return par


# --------------------------------------------------------------------------
# CRT precompute  (cuhe/Operations.cu:37-144)
# --------------------------------------------------------------------------
def gen_crt_primes(par: Params) -> List[int]:
    """cuhe/Operations.cu:37-80 (descending search, 'mid' prime, cut primes
    == 1 mod modMsg)."""
    pnum = par.numCrtPrime
    pr = [0] * pnum
    logmid = par.logCoeffMin - (pnum - par.depth) * par.logCrtPrime
    temp = (1 << par.logCrtPrime) - 1
    for i in range(0, pnum - par.depth):
        while not is_prime(temp):
            temp -= 1
        pr[i] = temp
        temp -= 1
    tmid = (1 << logmid) - 1 if logmid != par.logCrtPrime else temp
    while not is_prime(tmid):
        tmid -= 1
    pr[pnum - par.depth] = tmid
    if par.logCoeffCut == logmid:
        temp = tmid - 1
    elif par.logCoeffCut == par.logCrtPrime:
        temp -= 1
    else:
        temp = (1 << par.logCoeffCut) - 1
    for i in range(pnum - par.depth + 1, pnum):
        while (not is_prime(temp)) or temp % par.modMsg != 1:
            temp -= 1
        pr[i] = temp
        temp -= 1
    return pr


def gen_coeff_moduli(par: Params, primes: Sequence[int]) -> List[int]:
    """cuhe/Operations.cu:81-90: q_lvl = prod_{j < pnum-lvl} p_j."""
    out = []
    for i in range(par.depth):
        q = 1
        for j in range(par.numCrtPrime - i):
            q *= primes[j]
        out.append(q)
    return out


def gen_crt_inv_primes(par: Params, primes: Sequence[int]) -> np.ndarray:
    """cuhe/Operations.cu:91-100: invp[i*(i-1)/2+j] = (p_i mod p_j)^-1 mod p_j."""
    pnum = par.numCrtPrime
    out = np.zeros(max(pnum * (pnum - 1) // 2, 1), dtype=np.uint32)
    for i in range(1, pnum):
        for j in range(i):
            out[i * (i - 1) // 2 + j] = pow(primes[i] % primes[j], -1, primes[j])
    return out


@dataclass
class IcrtConst:
    q: np.ndarray       # u32[words_q]           M = q_lvl
    qp: np.ndarray      # u32[pnum][words_qp]    M_i = M / p_i (byte-truncated)
    qpinv: np.ndarray   # u32[pnum]              b_i = M_i^-1 mod p_i


def gen_icrt(par: Params, primes: Sequence[int], moduli: Sequence[int],
             lvl: int) -> IcrtConst:
    """cuhe/Operations.cu:107-144 incl. the truncating BytesFromZZ."""
    pnum = par._numCrtPrime(lvl)
    words_q = par._wordsCoeff(lvl)
    words_qp = par._wordsCoeff(lvl + 1)
    M = moduli[lvl]
    q = words_from_zz(M, words_q)
    qp = np.zeros((pnum, words_qp), dtype=np.uint32)
    qpinv = np.zeros(pnum, dtype=np.uint32)
    for i in range(pnum):
        Mi = M // primes[i]
        qp[i] = words_from_zz(Mi, words_qp)
        qpinv[i] = pow(Mi % primes[i], -1, primes[i])
    return IcrtConst(q, qp, qpinv)


# --------------------------------------------------------------------------
# cyclotomic polynomial + Barrett tables
# --------------------------------------------------------------------------
def _poly_divexact(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    """Exact division of integer polynomials (ascending coeffs, den monic-ish
    with den[0] == +-1) by power-series long division from the low end."""
    num = num.astype(object).copy()
    dn = len(den) - 1
    qn = len(num) - 1 - dn
    q = np.zeros(qn + 1, dtype=object)
    d0 = int(den[0])
    assert d0 in (1, -1)
    den = den.astype(object)
    for i in range(qn + 1):
        c = int(num[i]) * d0
        q[i] = c
        if c:
            hi = min(i + dn + 1, len(num))
            num[i:hi] -= c * den[:hi - i]
    return q


def cyclotomic(m: int) -> List[int]:
    """Phi_m(x), ascending integer coefficients.  Same polynomial as the
    Moebius product built by examples/DHS/DHS.cu:283-309 (genPolyMod_)."""
    phi = np.array([-1, 1], dtype=object)      # Phi_1
    n = 1
    mm = m
    p = 2
    while mm > 1:
        if mm % p == 0:
            first = True
            while mm % p == 0:
                mm //= p
                # Phi_{np}(x) = Phi_n(x^p)/Phi_n(x) (p !| n) or Phi_n(x^p) (p | n)
                up = np.zeros((len(phi) - 1) * p + 1, dtype=object)
                up[::p] = phi
                phi = _poly_divexact(up, phi) if first else up
                first = False
                n *= p
        p += 1
    return [int(c) for c in phi]


def barrett_u(phi: Sequence[int], n: int) -> List[int]:
    """u = floor(x^(2n-1) / Phi) over Z (cuhe/Operations.cu:216-219, 'zu /= zm').
    deg u = n-1.  Computed as reversed power-series inverse of the reversed Phi."""
    assert len(phi) == n + 1 and phi[n] == 1
    rev = [int(c) for c in phi[::-1]]           # rev[0] == 1
    inv = [0] * n
    inv[0] = 1
    nz = [(k, c) for k, c in enumerate(rev) if k > 0 and c != 0]
    for i in range(1, n):
        s = 0
        for k, c in nz:
            if k > i:
                break
            s += c * inv[i - k]
        inv[i] = -s
    return inv[::-1]                              # u_j = inv[n-1-j]


def barrett_u_fast(phi: Sequence[int], n: int) -> List[int]:
    """Same as barrett_u, vectorised (int64; asserts no overflow risk)."""
    rev = np.array(phi[::-1], dtype=np.int64)
    inv = np.zeros(n, dtype=np.int64)
    inv[0] = 1
    # forward substitution, column oriented: inv[i+1:] -= inv[i]*rev[1:...]
    for i in range(n):
        c = inv[i]
        if c:
            hi = min(n, i + n + 1)
            inv[i + 1:hi] -= c * rev[1:hi - i]
        if (i & 1023) == 0:
            assert np.abs(inv).max() < (1 << 40)
    return [int(c) for c in inv[::-1]]


# --------------------------------------------------------------------------
# mod-P primitives (definitions checked by tests/test_ModP.cu:57-137)
# --------------------------------------------------------------------------
def add_modP(x: int, y: int) -> int:      # cuhe/ModP.h:230-239
    return (x + y) % P


def sub_modP(x: int, y: int) -> int:      # cuhe/ModP.h:240-247
    return (x - y) % P


def mul_modP(x: int, y: int) -> int:      # cuhe/ModP.h:248-289
    return (x * y) % P


def ls_modP(x: int, l: int) -> int:       # cuhe/ModP.h:68-229
    return (x << l) % P


def root_of_unity(N: int) -> int:
    """w0 = g^(65536/N)  (cuhe/Base.cu:64-67, tests/test_ntt.cu:39-42)."""
    assert N in (16384, 32768, 65536)
    return pow(G, 65536 // N, P)


N_INV = {16384: 0xFFFBFFFF00040001,      # cuhe/Base.cu:489
         32768: 0xFFFDFFFF00020001,      # cuhe/Base.cu:656
         65536: 0xFFFEFFFF00010001}      # cuhe/Base.cu:841


# --------------------------------------------------------------------------
# NTT by definition and fast (pure python big ints; small cases only)
# --------------------------------------------------------------------------
def ntt_ext_def(x: Sequence[int], N: int, outs: Sequence[int]) -> List[int]:
    """X[i] = sum_{j<N/2} x[j] w^(ij) mod P for i in outs
    (tests/test_ntt.cu:38-64, the O(N^2) check)."""
    w = root_of_unity(N)
    res = []
    xs = [int(v) for v in x[:N // 2]]
    for i in outs:
        wi = pow(w, i, P)
        acc, t = 0, 1
        for v in xs:
            acc += v * t
            t = t * wi % P
        res.append(acc % P)
    return res


def _ntt_pow2(a: List[int], w: int) -> List[int]:
    """Natural-in / natural-out cyclic NTT of len(a) (power of two) mod P."""
    n = len(a)
    a = list(a)
    # bit reversal
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j ^= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        wl = pow(w, n // length, P)
        half = length // 2
        tw = [1] * half
        for k in range(1, half):
            tw[k] = tw[k - 1] * wl % P
        for s in range(0, n, length):
            for k in range(half):
                u = a[s + k]
                v = a[s + k + half] * tw[k] % P
                a[s + k] = (u + v) % P
                a[s + k + half] = (u - v) % P
        length <<= 1
    return a


def ntt_ext(x: Sequence[int], N: int) -> List[int]:
    """Forward zero-padded NTT: u32[N/2] -> u64[N]
    (cuhe/Base.cu:309-437 et al.; definition tests/test_ntt.cu:38-64)."""
    a = [int(v) for v in x[:N // 2]] + [0] * (N - min(len(x), N // 2))
    a = a[:N]
    return _ntt_pow2(a, root_of_unity(N))


def intt(X: Sequence[int], N: int) -> List[int]:
    """x[j] = N^-1 sum_i X[i] w^(-ij) mod P  (cuhe/Base.cu:438-490:
    index-reversed forward transform times the N^-1 constant)."""
    w = root_of_unity(N)
    a = _ntt_pow2([int(v) for v in X], pow(w, -1, P))
    ninv = N_INV[N]
    assert ninv == pow(N, -1, P)
    return [v * ninv % P for v in a]


# --------------------------------------------------------------------------
# CRT / ICRT / modswitch / Barrett / relin on python ints
# --------------------------------------------------------------------------
def crt(raw: np.ndarray, primes: Sequence[int], L: int, n: int) -> np.ndarray:
    """cuhe/Base.cu:857-879: raw u32[H][W] (little-endian words per
    coefficient) -> u32[L][H]; only idx < modLen is written."""
    H, W = raw.shape
    out = np.zeros((L, H), dtype=np.uint32)
    for i in range(n):
        c = zz_from_words(raw[i])
        for l in range(L):
            out[l, i] = c % primes[l]
    return out


def icrt(c: np.ndarray, primes: Sequence[int], ic: IcrtConst, n: int,
         W: int) -> np.ndarray:
    """cuhe/Base.cu:880-924, literal accumulate-and-conditionally-subtract,
    with the byte-truncated M_i table.  u32[L][H] -> u32[H][W]."""
    L, H = c.shape[0], c.shape[1]
    M = zz_from_words(ic.q)
    out = np.zeros((H, W), dtype=np.uint32)
    mis = [zz_from_words(ic.qp[l]) for l in range(len(ic.qpinv))]
    top = 1 << (32 * (W + 1))
    for i in range(n):
        s = 0
        for l in range(len(ic.qpinv)):
            tt = (int(c[l, i]) % primes[l]) * int(ic.qpinv[l]) % primes[l]
            s = (s + tt * mis[l]) % top
            # leq_M(): subtract once if s >= M (cuhe/Base.cu:846-856)
            if s >= M:
                s -= M
        out[i] = words_from_zz(s, W)
    return out


def _c_int32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def modswitch(c: np.ndarray, primes: Sequence[int], invp: np.ndarray,
              L: int, n: int, modmsg: int) -> np.ndarray:
    """cuhe/Base.cu:1112-1138, C integer semantics followed literally
    (signed 32-bit 'dirty'/'temp', unsigned products)."""
    out = c.copy()
    pt = primes[L - 1]
    for idx in range(n):
        dirty = _c_int32(int(c[L - 1, idx]))
        ep = dirty - modmsg * int(dirty / modmsg)      # C '%' truncation
        if ep != 0:
            # int > unsigned comparison is done in unsigned
            if (dirty & 0xFFFFFFFF) > ((pt - 1) // 2):
                dirty = _c_int32((dirty & 0xFFFFFFFF) - ((ep * pt) & 0xFFFFFFFF))
            else:
                dirty = _c_int32((dirty & 0xFFFFFFFF) + ((ep * pt) & 0xFFFFFFFF))
        for i in range(L - 1):
            temp = _c_int32(int(c[i, idx]))
            while temp < dirty:
                temp = _c_int32((temp & 0xFFFFFFFF) + primes[i])
            temp = _c_int32(temp - dirty)
            tt = (temp & 0xFFFFFFFFFFFFFFFF) if temp >= 0 else (temp + (1 << 64))
            tt = (tt * int(invp[(L - 1) * (L - 2) // 2 + i])) & 0xFFFFFFFFFFFFFFFF
            out[i, idx] = tt % primes[i]
    out[L - 1, :] = c[L - 1, :]          # reference leaves the dropped row as is
    return out


def digit(raw_row: np.ndarray, w: int, wid: int, w32: int) -> int:
    """cuhe/Base.cu:361-371: bits [w*wid, w*wid+w) of a raw coefficient."""
    lo = (w * wid) >> 5
    if lo + 1 < w32:
        s = (int(raw_row[lo + 1]) << 32) + int(raw_row[lo])
    else:
        s = int(raw_row[lo])
    s >>= (w * wid) & 0x1F
    return s & ((1 << w) - 1)


# --------------------------------------------------------------------------
# GMP-through-ctypes Kronecker multiply (libgmp.so.10 is in the image; there
# are no headers, so mpz_t is declared by hand).  Used for exact ZZX products
# at full sizes -- the role NTL's ZZX multiply plays in the reference
# (examples/DHS/DHS.cu:219-221).
# --------------------------------------------------------------------------
class _Mpz(ctypes.Structure):
    _fields_ = [("alloc", ctypes.c_int), ("size", ctypes.c_int),
                ("d", ctypes.c_void_p)]


_gmp = None


def _load_gmp():
    global _gmp
    if _gmp is None:
        name = ctypes.util.find_library("gmp") or "libgmp.so.10"
        try:
            _gmp = ctypes.CDLL(name)
        except OSError:
            _gmp = False
    return _gmp


def bigmul(a: int, b: int) -> int:
    """a*b for non-negative ints, through GMP when available."""
    g = _load_gmp()
    if not g or a.bit_length() < 200_000:
        return a * b
    def imp(v):
        z = _Mpz()
        g.__gmpz_init(ctypes.byref(z))
        raw = v.to_bytes((v.bit_length() + 7) // 8 or 1, "little")
        g.__gmpz_import(ctypes.byref(z), ctypes.c_size_t(len(raw)), -1,
                        ctypes.c_size_t(1), 0, ctypes.c_size_t(0), raw)
        return z
    za, zb = imp(a), imp(b)
    zc = _Mpz()
    g.__gmpz_init(ctypes.byref(zc))
    g.__gmpz_mul(ctypes.byref(zc), ctypes.byref(za), ctypes.byref(zb))
    g.__gmpz_sizeinbase.restype = ctypes.c_size_t
    nb = (g.__gmpz_sizeinbase(ctypes.byref(zc), 2) + 7) // 8
    buf = ctypes.create_string_buffer(nb)
    cnt = ctypes.c_size_t(0)
    g.__gmpz_export(buf, ctypes.byref(cnt), -1, ctypes.c_size_t(1), 0,
                    ctypes.c_size_t(0), ctypes.byref(zc))
    out = int.from_bytes(buf.raw[:cnt.value], "little")
    for z in (za, zb, zc):
        g.__gmpz_clear(ctypes.byref(z))
    return out


def polymul_kronecker(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Exact product of polynomials with non-negative integer coefficients."""
    if not len(a) or not len(b):
        return []
    bits = (max(int(v).bit_length() for v in a) + max(int(v).bit_length() for v in b)
            + min(len(a), len(b)).bit_length() + 1)
    nbytes = (bits + 7) // 8
    def pack(v):
        return int.from_bytes(b"".join(int(c).to_bytes(nbytes, "little") for c in v),
                              "little")
    prod = bigmul(pack(a), pack(b))
    raw = prod.to_bytes((len(a) + len(b)) * nbytes, "little")
    return [int.from_bytes(raw[i * nbytes:(i + 1) * nbytes], "little")
            for i in range(len(a) + len(b) - 1)]


def poly_mod_phi(f: Sequence[int], phi: Sequence[int]) -> List[int]:
    """f mod Phi (Phi monic, small integer coeffs) over Z, then caller reduces
    coefficients.  Uses x^m == 1 folding when Phi | x^m - 1 is supplied via
    phi_m, else plain long division from the top."""
    n = len(phi) - 1
    f = [int(c) for c in f]
    nz = [(k, int(c)) for k, c in enumerate(phi[:-1]) if c != 0]
    for i in range(len(f) - 1, n - 1, -1):
        c = f[i]
        if c:
            base = i - n
            for k, pc in nz:
                f[base + k] -= c * pc
            f[i] = 0
    return f[:n] + [0] * max(0, n - len(f))


def mul_mod(a: Sequence[int], b: Sequence[int], phi: Sequence[int], m: int,
            q: int) -> List[int]:
    """(a*b mod Phi_m) mod q, coefficients in [0,q): the NTL host path
    't = a*b; t %= polyMod_; coeffReduce' of examples/DHS/DHS.cu:219-221.
    Folds modulo x^m - 1 first (Phi_m | x^m - 1) so the long division is short."""
    n = len(phi) - 1
    prod = polymul_kronecker(a, b)
    fold = [0] * m
    for i, c in enumerate(prod):
        fold[i % m] += c
    r = poly_mod_phi([c % q for c in fold], phi)
    return [c % q for c in r[:n]]
