"""Pins the CPU oracle (oracle/) before anything trusts it.  The reference has no
stored vectors (every test seeds with time(NULL)); what its own tests assert is
 * tests/test_ModP.cu:57-137  primitives == big-int arithmetic mod P
 * tests/test_ntt.cu:38-64    ext-NTT == O(N^2) DFT with g = 15893793146607301539
so those properties are checked here on the same input distributions, plus the
oracle's internal consistency against exact big-integer ring arithmetic and the
committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from common import C2, PRINCE, ROOT, SIMPLE_DHS, SMALL_RELIN, get_oracle

from oracle import oracle as orc
from oracle import pyoracle as po

P = po.P


def test_constants():
    # cuhe/ModP.h:33, cuhe/Base.cu:65,489,656,841
    assert P == 2**64 - 2**32 + 1 and po.is_prime(P)
    assert pow(po.G, 65536, P) == 1 and pow(po.G, 32768, P) == P - 1
    assert pow(po.G, 1024, P) == 8                      # the 64th root of unity is 2^3
    assert pow(2, 96, P) == P - 1
    for N, c in po.N_INV.items():
        assert c * N % P == 1


def test_modp_primitives_vs_bigint():
    """tests/test_ModP.cu: rand_array / rand_offset operand distributions."""
    lib = orc.lib()
    rng = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**63, P - 2**32]
    vals = edge + [rng.getrandbits(64) % P for _ in range(3000)] + [rng.getrandbits(32) for _ in range(1000)]
    for _ in range(20000):
        x, y = rng.choice(vals), rng.choice(vals)
        assert lib.orc_add_modP(x, y) == (x + y) % P
        assert lib.orc_sub_modP(x, y) == (x - y) % P
        assert lib.orc_mul_modP(x, y) == (x * y) % P == lib.orc_mul_modP_slow(x, y)
    for a in range(8):
        for b in range(8):
            l = 3 * a * b                                   # the shifts the reference uses
            for x in vals[:200]:
                assert lib.orc_ls_modP(x, l) == (x << l) % P


@pytest.mark.parametrize("N", [16384, 32768, 65536])
def test_ntt_ext_equals_dft(N):
    """tests/test_ntt.cu:38-64 on rand() (31-bit) inputs."""
    rng = np.random.default_rng(N)
    x = rng.integers(0, 1 << 31, size=N // 2, dtype=np.uint32)
    X = orc.ntt_ext(x, N)
    idx = [0, 1, 2, 3, 63, 64, 1023, 1024, N // 2 - 1, N // 2, N - 2, N - 1]
    assert [int(X[i]) for i in idx] == po.ntt_ext_def(x, N, idx)
    r = orc.roots(N)
    assert int(r[1]) == pow(po.G, 65536 // N, P) and int(r[N - 1]) == pow(int(r[1]), N - 1, P)
    back = orc.intt_u64(X, N)[0]
    assert np.array_equal(back[: N // 2], x.astype(np.uint64)) and not back[N // 2:].any()


def test_ntt_c_equals_python_full():
    N = 16384
    rng = np.random.default_rng(3)
    x = rng.integers(0, 1 << 32, size=N // 2, dtype=np.uint32)
    assert [int(v) for v in orc.ntt_ext(x, N)] == po.ntt_ext(x, N)


def test_parameter_sets_known_values():
    """Values the survey derived from cuhe/Parameters.cu + cuhe/Operations.cu:37-80."""
    o = get_oracle(SIMPLE_DHS)
    assert (o.n, o.N, o.L0, o.par.numEvalKey) == (8190, 16384, 7, 141)
    assert o.primes == [2097143, 2097133, 524287, 1048573, 1048571, 1048559, 1048549]
    assert o.moduli[0].bit_length() == 141
    pp = po.set_param(*PRINCE)
    pr = po.gen_crt_primes(pp)
    assert (pp.modLen, pp.nttLen, pp.numCrtPrime, pp.numEvalKey, pp._wordsCoeff(0)) == (16384, 32768, 25, 40, 20)
    assert pr[0] == 33554393 and po.gen_coeff_moduli(pp, pr)[0].bit_length() == 625
    pc = po.set_param(*C2)
    assert (pc.modLen, pc.nttLen, pc.numCrtPrime, pc._wordsCoeff(0), pc.numEvalKey) == (27000, 65536, 24, 18, 36)
    for ps in (SIMPLE_DHS, PRINCE, C2):
        q = po.set_param(*ps)
        prs = po.gen_crt_primes(q)
        assert len(set(prs)) == len(prs) and all(po.is_prime(v) for v in prs)
        assert all(v % q.modMsg == 1 for v in prs[q.numCrtPrime - q.depth + 1:])
        assert q.modLen * (max(prs) - 1) ** 2 < P           # no wrap in the NTT convolution


def test_cyclotomic_and_barrett_u():
    for m in (8191, 21845, 32767, 105):
        phi = po.cyclotomic(m)
        n = len(phi) - 1
        assert n == po.euler_totient(m) and phi[-1] == 1 and phi == phi[::-1]
    phi = po.cyclotomic(105)                               # has a -2 coefficient
    assert min(phi) == -2
    n = len(phi) - 1
    u = po.barrett_u(phi, n)
    assert u == po.barrett_u_fast(phi, n)
    # x^(2n-1) = u*phi + rho with deg rho < n
    prod = [0] * (2 * n)
    for i, a in enumerate(u):
        for j, b in enumerate(phi):
            prod[i + j] += a * b
    rho = [(1 if k == 2 * n - 1 else 0) - prod[k] for k in range(2 * n)]
    assert all(v == 0 for v in rho[n:])


def test_icrt_constants_and_roundtrip():
    o = get_oracle(SIMPLE_DHS)
    for lvl in range(o.par.depth):
        ic = o.icrt_const(lvl)
        M = po.zz_from_words(ic.q)
        assert M == o.moduli[lvl]
        for i in range(o.L(lvl)):
            Mi = M // o.primes[i]
            assert po.zz_from_words(ic.qp[i]) == Mi          # no truncation at these sizes
            assert int(ic.qpinv[i]) * Mi % o.primes[i] == 1
    rng = random.Random(4)
    coeffs = [rng.randrange(o.moduli[1]) for _ in range(o.n)]
    raw = o.to_raw(coeffs, 1)
    c = o.crt(raw, 1)
    assert np.array_equal(c[:, :16], po.crt(raw, o.primes, o.L(1), 16)[:, :16])
    assert o.from_raw(o.icrt(c, 1)) == coeffs
    assert np.array_equal(o.icrt(c, 1)[:8], po.icrt(c, o.primes, o.icrt_const(1), 8, o.W(1))[:8])


def test_mul_pipeline_equals_exact_ring_arithmetic():
    """CRT -> NTT -> mul -> INTT -> Barrett (literal step order of
    cuhe/Operations.cu:460-501) == (a*b mod Phi_m) mod q with big ints."""
    o = get_oracle(SIMPLE_DHS)
    rng = random.Random(2)
    for lvl in (0, 2):
        q = o.moduli[lvl]
        a = [rng.randrange(q) for _ in range(o.n)]
        b = [rng.randrange(q) for _ in range(o.n)]
        ex = o.mul_exact(a, b, lvl)
        c = o.mul_raw_to_crt(o.to_raw(a, lvl), o.to_raw(b, lvl), lvl)
        for l in range(o.L(lvl)):
            assert np.array_equal(c[l, :o.n], np.array([v % o.primes[l] for v in ex], dtype=np.uint32))
        assert not c[:, o.n:].any()
        assert o.from_raw(o.icrt(c, lvl)) == ex


def test_modswitch_c_equals_literal_python_and_keeps_parity():
    o = get_oracle(SIMPLE_DHS)
    rng = random.Random(8)
    lvl = 0
    coeffs = [rng.randrange(o.moduli[lvl]) for _ in range(o.n)]
    c = o.crt(o.to_raw(coeffs, lvl), lvl)
    ms = o.modswitch(c, lvl)
    lit = po.modswitch(c, o.primes, o.invp, o.L(lvl), 128, o.par.modMsg)
    assert np.array_equal(ms[:, :128], lit[: o.L(lvl) - 1, :128])
    # the switched value is (c - delta)/p_last with delta == c (mod p_last), delta == 0 (mod 2)
    new = o.from_raw(o.icrt(np.ascontiguousarray(ms), lvl + 1))
    pl = o.primes[o.L(lvl) - 1]
    q1 = o.moduli[lvl + 1]
    for i in range(64):
        d = int(c[o.L(lvl) - 1, i])
        ep = d % 2
        if ep:
            d = d - pl if d > (pl - 1) // 2 else d + pl
        assert new[i] == ((coeffs[i] - d) // pl) % q1 and (coeffs[i] - d) % pl == 0


def test_relin_digits_recompose():
    o = get_oracle(SMALL_RELIN)
    rng = random.Random(6)
    lvl = 0
    coeffs = [rng.randrange(o.moduli[lvl]) for _ in range(o.n)]
    raw = o.to_raw(coeffs, lvl)
    w = o.par.logRelin
    acc = [0] * 32
    for k in range(o.K(lvl)):
        d = o.digits(raw, lvl, k)
        assert int(d.max()) < (1 << w)
        for i in range(32):
            acc[i] += int(d[i]) << (w * k)
            assert int(d[i]) == po.digit(raw[i], w, k, o.W(lvl))
    assert acc == coeffs[:32]


def test_golden_fixtures():
    """Hashes of oracle outputs on seeded inputs, committed with their generator
    (tests/golden/make_golden.py).  Guards the oracle itself against drift."""
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    g = json.load(open(path))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    compute_case = mg.compute_case
    for case in g["cases"]:
        if case["size"] == "full" and not os.environ.get("CUHE_GOLDEN_FULL"):
            continue
        got = compute_case(case["params"], case["seed"], exact=False)
        for k in ("crt_sha", "ntt_sha", "mul_crt_sha", "mul_raw_sha", "modswitch_sha"):
            assert got[k] == case[k], (case["name"], k)


@pytest.mark.parametrize("ps", [SIMPLE_DHS, (4, 2, 16, 50, 25, 21845)], ids=["m8191_prime", "m21845_composite"])
def test_gmp_zzx_path_equals_ntt_pipeline(ps):
    """oracle/zzx_gmp.c -- the reference's NTL host path (t = a*b; t %= polyMod; coeffReduce,
    examples/DHS/DHS.cu:219-221) restated on GMP, bench.py's CPU arm -- gives, word for word, what the
    CRT/NTT/Barrett/ICRT pipeline gives, and what exact big-integer ring arithmetic gives."""
    o = get_oracle(ps)
    W, H, n = o.W(0), o.H, o.n
    rng = np.random.default_rng(11)
    top_bits = o.moduli[0].bit_length() - 1 - 32 * (W - 1)

    def polys(batch):
        x = rng.integers(0, 1 << 32, size=(batch, H, W), dtype=np.uint32)
        x[:, :, W - 1] &= np.uint32((1 << top_bits) - 1)
        x[:, n:, :] = 0
        return x
    a, b = polys(3), polys(3)
    a[2, :n] = np.array(po.words_from_zz(o.moduli[0] - 1, W), dtype=np.uint32)      # maximal coefficients
    b[2, :n] = a[2, :n]
    got = o.mul_raw_batch_zzx(a, b, 0)
    assert np.array_equal(got, o.mul_raw_batch(a, b, 0))
    assert o.from_raw(got[0]) == o.mul_exact(o.from_raw(a[0]), o.from_raw(b[0]), 0)
    inv = o.inverse_series()
    assert len(inv) == o.par.mSize - n and int(np.abs(inv).max()) <= 2
