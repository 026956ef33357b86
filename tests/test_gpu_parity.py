"""GPU parity tests: every C-ABI entry point of the hot path against the CPU
oracle on the same seeded inputs, bit-exact (integer arithmetic; no tolerance).
They mirror the semantics of the reference's own tests:
tests/test_ModP.cu (primitives), tests/test_ntt.cu (ext-NTT == DFT) and the
domain machine exercised by examples/DHS/simple_DHS.cu."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from common import C2, MID32K, MID64K, ROOT, SIMPLE_DHS, SMALL_RELIN, get_oracle

pytestmark = pytest.mark.gpu

P = 0xFFFFFFFF00000001


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


class Eng:
    """Thin harness over the raw C ABI for one parameter set."""

    def __init__(self, lib, ps, rank=0, world=1, device=0, with_polymod=True):
        from cuhe_b200._lib import check, cuhe_params
        self.lib, self.check = lib, check
        self.torch = _torch()
        self.par = cuhe_params()
        check(lib.cuhe_set_parameters(C.byref(self.par), *ps))
        self.h = C.c_void_p()
        check(lib.cuhe_ctx_create(C.byref(self.h), C.byref(self.par), device, rank, world))
        self.orc = get_oracle(ps)
        self._keep = []
        self.dev = f"cuda:{device}"
        if with_polymod:
            phi = np.array(self.orc.phi, dtype=np.int64)
            check(lib.cuhe_ctx_set_poly_modulus_host(self.h, phi.ctypes.data_as(C.c_void_p), len(phi)))

    def close(self):
        self._keep.clear()
        self.lib.cuhe_ctx_destroy(self.h)

    def st(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def up(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        view = a.view(np.int32) if a.dtype == np.uint32 else a.view(np.int64)
        t = self.torch.from_numpy(view).to(self.dev)
        self._keep.append(t)          # p(e.up(x)) passes only the address: keep the tensor alive until close()
        if len(self._keep) > 64:
            self.torch.cuda.synchronize()
            del self._keep[:32]
        return t

    def empty(self, shape, dt):
        t = self.torch
        return t.zeros(shape, dtype=t.int32 if dt == np.uint32 else t.int64, device=self.dev)

    @staticmethod
    def dn(t, dt):
        return t.cpu().numpy().view(dt)

    def rows(self, lvl):
        return self.lib.cuhe_ctx_rows(self.h, lvl)

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.h, *args))


def p(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def eng16(lib):
    e = Eng(lib, SIMPLE_DHS)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng64(lib):
    e = Eng(lib, MID64K)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng32(lib):
    e = Eng(lib, MID32K)
    yield e
    e.close()


def rand_poly_raw(orc, lvl, seed):
    rng = random.Random(seed)
    q = orc.moduli[lvl]
    coeffs = [rng.randrange(q) for _ in range(orc.n)]
    return coeffs, orc.to_raw(coeffs, lvl)


# ---------------------------------------------------------------------------
# tests/test_ModP.cu semantics: device primitives == big-int arithmetic mod P
# ---------------------------------------------------------------------------
def _modp_inputs(n, seed):
    rng = np.random.default_rng(seed)
    edge = np.array([0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, 0x7FFFFFFFFFFFFFFF,
                     0x8000000000000000, 0xFFFFFFFE00000002, 0xFFFFFFFF], dtype=np.uint64)
    # rand_array (32/64-bit operands, tests/test_ModP.cu:38-49), reduced to canonical
    a = rng.integers(0, P, size=n, dtype=np.uint64)
    b = rng.integers(0, P, size=n, dtype=np.uint64)
    b[: n // 4] &= np.uint64(0xFFFFFFFF)
    a[n // 4: n // 2] &= np.uint64(0xFFFFFFFF)
    ea = np.repeat(edge, len(edge))
    eb = np.tile(edge, len(edge))
    return np.concatenate([ea, a]), np.concatenate([eb, b])


@pytest.mark.parametrize("op", [0, 1, 2])
def test_modp_add_sub_mul(eng16, op):
    x, y = _modp_inputs(1 << 20, 20260924 + op)
    dx, dy = eng16.up(x), eng16.up(y)
    out = eng16.empty(x.shape, np.uint64)
    eng16.call("cuhe_modp_batch", op, p(out), p(dx), p(dy), C.c_size_t(x.size), 0, eng16.st())
    got = Eng.dn(out, np.uint64)
    xo, yo = x.astype(object), y.astype(object)
    want = ((xo + yo) % P, (xo - yo) % P, (xo * yo) % P)[op]
    assert np.array_equal(got.astype(object), want)


def test_modp_shifts(eng16):
    x, _ = _modp_inputs(1 << 12, 5)
    dx = eng16.up(x)
    out = eng16.empty(x.shape, np.uint64)
    xo = x.astype(object)
    # the reference only uses l = 3*a*b (tests/test_ModP.cu:57-78); check every l in [0,192)
    for l in range(192):
        eng16.call("cuhe_modp_batch", 3, p(out), p(dx), None, C.c_size_t(x.size), l, eng16.st())
        got = Eng.dn(out, np.uint64)
        assert np.array_equal(got.astype(object), (xo << l) % P), f"shift {l}"


# ---------------------------------------------------------------------------
# tests/test_ntt.cu semantics: forward ext-NTT for N = 16k/32k/64k on rand() inputs
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N", [16384, 32768, 65536])
def test_ntt_ext_batch_vs_oracle(eng16, N):
    from oracle import oracle as orc
    from oracle import pyoracle as po
    rng = np.random.default_rng(N)
    cnt = 7
    # test_ntt.cu:115-118: x = rand() (31-bit), stride nttLen between polynomials
    x = rng.integers(0, 1 << 31, size=(cnt, N), dtype=np.uint32)
    x[3, : N // 2] = 0xFFFFFFFF          # edge: maximal words
    x[4, : N // 2] = 0
    dx = eng16.up(x)
    out = eng16.empty((cnt, N), np.uint64)
    eng16.call("cuhe_ntt_ext_batch", p(out), p(dx), N, cnt, C.c_longlong(N), eng16.st())
    got = Eng.dn(out, np.uint64)
    want = orc.ntt_ext(x, N)
    assert np.array_equal(got, want)
    # and against the O(N^2) definition itself on a few outputs (test_ntt.cu:38-64)
    idx = [0, 1, 2, 63, 64, 65, 4095, 4096, N // 2, N - 1]
    assert [int(got[5, i]) for i in idx] == po.ntt_ext_def(x[5], N, idx)
    # inverse round trip
    back = eng16.empty((cnt, N), np.uint64)
    eng16.call("cuhe_intt_batch", p(back), p(out), N, cnt, eng16.st())
    b = Eng.dn(back, np.uint64)
    assert np.array_equal(b[:, : N // 2], x[:, : N // 2].astype(np.uint64))
    assert not b[:, N // 2:].any()


def test_ntt_batch_empty_and_large(eng16):
    N = 16384
    x = eng16.up(np.zeros((1, N), dtype=np.uint32))
    out = eng16.empty((1, N), np.uint64)
    eng16.call("cuhe_ntt_ext_batch", p(out), p(x), N, 0, C.c_longlong(N), eng16.st())   # empty batch is a no-op
    assert eng16.lib.cuhe_ntt_ext_batch(eng16.h, p(out), p(x), 12345, 1, C.c_longlong(N), eng16.st()) != 0


# ---------------------------------------------------------------------------
# domain conversions and arithmetic, three ring sizes
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["eng16", "eng32", "eng64"])
def test_crt_ntt_intt_icrt(which, request):
    e = request.getfixturevalue(which)
    o = e.orc
    for lvl in (0, 1):
        L, W, H, N = o.L(lvl), o.W(lvl), o.H, o.N
        coeffs, raw = rand_poly_raw(o, lvl, 100 + lvl)
        d_raw = e.up(raw)
        d_crt = e.empty((L, H), np.uint32)
        e.call("cuhe_crt", p(d_crt), p(d_raw), lvl, e.st())
        want_crt = o.crt(raw, lvl)
        assert np.array_equal(Eng.dn(d_crt, np.uint32), want_crt)
        d_ntt = e.empty((L, N), np.uint64)
        e.call("cuhe_ntt", p(d_ntt), p(d_crt), lvl, e.st())
        want_ntt = o.ntt(want_crt)
        assert np.array_equal(Eng.dn(d_ntt, np.uint64), want_ntt)
        d_back = e.empty((L, H), np.uint32)
        e.call("cuhe_intt", p(d_back), p(d_ntt), lvl, e.st())
        assert np.array_equal(Eng.dn(d_back, np.uint32), want_crt)
        d_full = e.empty((L, N), np.uint32)
        e.call("cuhe_intt_double_deg", p(d_full), p(d_ntt), lvl, e.st())
        assert np.array_equal(Eng.dn(d_full, np.uint32), o.intt_hold(want_ntt))
        d_raw2 = e.empty((H, W), np.uint32)
        e.call("cuhe_icrt", p(d_raw2), p(d_crt), lvl, 0, H, e.st())
        got_raw = Eng.dn(d_raw2, np.uint32)
        assert np.array_equal(got_raw, o.icrt(want_crt, lvl))
        assert o.from_raw(got_raw) == coeffs


@pytest.mark.parametrize("ps", [SIMPLE_DHS, MID32K, MID64K])
def test_literal_icrt_kernel(lib, monkeypatch, ps):
    """The literal ICRT kernel (the reference's add one term / compare / subtract M loop, cuhe/Base.cu:880-924; the
    path a byte-truncated M_l takes, cuhe/Operations.cu:127-128) forced through CUHE_B200_LITERAL_ICRT=1, whole
    polynomials and coefficient slices, both levels."""
    monkeypatch.setenv("CUHE_B200_LITERAL_ICRT", "1")
    e = Eng(lib, ps)
    try:
        o = e.orc
        for lvl in (0, 1):
            L, W, H = o.L(lvl), o.W(lvl), o.H
            coeffs, raw = rand_poly_raw(o, lvl, 300 + lvl)
            want_crt = o.crt(raw, lvl)
            d_raw2 = e.empty((H, W), np.uint32)
            e.call("cuhe_icrt", p(d_raw2), p(e.up(want_crt)), lvl, 0, H, e.st())
            got = Eng.dn(d_raw2, np.uint32)
            assert np.array_equal(got, o.icrt(want_crt, lvl))
            assert o.from_raw(got) == coeffs
            # a coefficient range only
            d_part = e.empty((H, W), np.uint32)
            e.call("cuhe_icrt", p(d_part), p(e.up(want_crt)), lvl, 100, 1000, e.st())
            part = Eng.dn(d_part, np.uint32)
            assert np.array_equal(part[100:1000], got[100:1000]) and not part[:100].any() and not part[1000:].any()
    finally:
        e.close()


@pytest.mark.parametrize("which", ["eng16", "eng32", "eng64"])
def test_mul_barrett(which, request):
    """ctxt x ctxt: NTT-domain product -> inttMod (INTT + polynomial Barrett)."""
    e = request.getfixturevalue(which)
    o = e.orc
    lvl = 0
    L, H, N = o.L(lvl), o.H, o.N
    a, ra = rand_poly_raw(o, lvl, 1)
    b, rb = rand_poly_raw(o, lvl, 2)
    ca, cb = o.crt(ra, lvl), o.crt(rb, lvl)
    na, nb = o.ntt(ca), o.ntt(cb)
    want = o.mul_raw_to_crt(ra, rb, lvl)
    d_na, d_nb = e.up(na), e.up(nb)
    d_prod = e.empty((L, N), np.uint64)
    e.call("cuhe_ntt_mul", p(d_prod), p(d_na), p(d_nb), lvl, e.st())
    assert np.array_equal(Eng.dn(d_prod, np.uint64), __import__("oracle.oracle", fromlist=["x"]).ntt_mul(na, nb))
    d_out = e.empty((L, H), np.uint32)
    e.call("cuhe_intt_mod", p(d_out), p(d_prod), lvl, e.st())
    assert np.array_equal(Eng.dn(d_out, np.uint32), want)
    # fused cAnd + n2c
    d_out2 = e.empty((L, H), np.uint32)
    e.call("cuhe_ntt_mul_intt_mod", p(d_out2), p(d_na), p(d_nb), lvl, e.st())
    assert np.array_equal(Eng.dn(d_out2, np.uint32), want)
    # stand-alone barrett() on the oracle's hold buffer
    hold = o.intt_hold(Eng.dn(d_prod, np.uint64))
    d_out3 = e.empty((L, H), np.uint32)
    e.call("cuhe_barrett", p(d_out3), p(e.up(hold)), lvl, e.st())
    assert np.array_equal(Eng.dn(d_out3, np.uint32), want)
    if which == "eng16":
        # independent pin: exact big-int (a*b mod Phi) mod q (NTL host path, DHS.cu:219-221)
        ex = o.mul_exact(a, b, lvl)
        got = Eng.dn(d_out, np.uint32)
        for l in range(L):
            assert np.array_equal(got[l, :o.n], np.array([v % o.primes[l] for v in ex], dtype=np.uint32))


@pytest.mark.parametrize("mode", ["sparse", "ntt", "barrett"])
@pytest.mark.parametrize("ps", [SIMPLE_DHS, MID32K, MID64K, (3, 2, 16, 40, 20, 1155)])
def test_three_reductions_modulo_phi_agree(lib, monkeypatch, ps, mode):
    """inttMod through each of the three exact reductions modulo Phi_m -- strided differences / prefix sums over the
    binomial factors of Phi_m (default), the fold + inverse-series products of round 1 (CUHE_B200_REDUCE=ntt) and the
    reference's Barrett step order (CUHE_B200_REDUCE=barrett, cuhe/Operations.cu:460-501) -- against the oracle, at a
    prime m (8191), two three-prime m (21845, 32767) and m = 1155 = 3*5*7*11 (16 binomials, quotient longer than the
    remainder), at level 0 and level 1."""
    monkeypatch.setenv("CUHE_B200_REDUCE", mode)
    e = Eng(lib, ps)
    try:
        o = e.orc
        for lvl in (0, 1):
            L, H, N = o.L(lvl), o.H, o.N
            _, ra = rand_poly_raw(o, lvl, 61 + lvl)
            _, rb = rand_poly_raw(o, lvl, 71 + lvl)
            want = o.mul_raw_to_crt(ra, rb, lvl)
            na, nb = o.ntt(o.crt(ra, lvl)), o.ntt(o.crt(rb, lvl))
            d_out = e.empty((L, H), np.uint32)
            e.call("cuhe_ntt_mul_intt_mod", p(d_out), p(e.up(na)), p(e.up(nb)), lvl, e.st())
            assert np.array_equal(Eng.dn(d_out, np.uint32), want)
            # worst-case magnitudes: every coefficient p-1
            top = np.stack([np.full(H, pr - 1, dtype=np.uint32) for pr in o.primes[:L]])
            top[:, o.n:] = 0
            nt = o.ntt(top)
            e.call("cuhe_ntt_mul_intt_mod", p(d_out), p(e.up(nt)), p(e.up(nt)), lvl, e.st())
            assert np.array_equal(Eng.dn(d_out, np.uint32), o.intt_mod(__import__("oracle.oracle", fromlist=["x"]).ntt_mul(nt, nt)))
    finally:
        e.close()


def test_pointwise_and_crt_adds(eng16):
    e, o = eng16, eng16.orc
    from oracle import oracle as orc
    lvl = 1
    L, H, N = o.L(lvl), o.H, o.N
    rng = np.random.default_rng(9)
    x = rng.integers(0, P, size=(L, N), dtype=np.uint64)
    y = rng.integers(0, P, size=(L, N), dtype=np.uint64)
    dx, dy = e.up(x), e.up(y)
    dz = e.empty((L, N), np.uint64)
    e.call("cuhe_ntt_add", p(dz), p(dx), p(dy), lvl, e.st())
    assert np.array_equal(Eng.dn(dz, np.uint64), orc.ntt_add(x, y))
    e.call("cuhe_ntt_mul_nx1", p(dz), p(dx), p(dy), lvl, e.st())
    assert np.array_equal(Eng.dn(dz, np.uint64), orc.ntt_mul(x, np.broadcast_to(y[0], x.shape).copy()))
    e.call("cuhe_ntt_add_nx1", p(dz), p(dx), p(dy), lvl, e.st())
    assert np.array_equal(Eng.dn(dz, np.uint64), orc.ntt_add(x, np.broadcast_to(y[0], x.shape).copy()))
    pr = np.array(o.primes[:L], dtype=np.uint64)[:, None]
    a = (rng.integers(0, 1 << 40, size=(L, H), dtype=np.uint64) % pr).astype(np.uint32)
    b = (rng.integers(0, 1 << 40, size=(L, H), dtype=np.uint64) % pr).astype(np.uint32)
    a[:, o.n:] = 0
    b[:, o.n:] = 0
    da, db = e.up(a), e.up(b)
    ds = e.empty((L, H), np.uint32)
    e.call("cuhe_crt_add", p(ds), p(da), p(db), lvl, e.st())
    assert np.array_equal(Eng.dn(ds, np.uint32), o.crt_add(a, b))
    sc = (rng.integers(0, 2, size=H, dtype=np.uint32))
    e.call("cuhe_crt_add_nx1", p(ds), p(da), p(e.up(sc)), lvl, e.st())
    assert np.array_equal(Eng.dn(ds, np.uint32), o.crt_add_nx1(a, sc))
    ds2 = e.up(a)
    e.call("cuhe_crt_add_int", p(ds2), p(da), C.c_uint(1), lvl, e.st())
    assert np.array_equal(Eng.dn(ds2, np.uint32), o.crt_add_int(a, 1))


@pytest.mark.parametrize("which", ["eng16", "eng64"])
def test_modswitch(which, request):
    e = request.getfixturevalue(which)
    o = e.orc
    for lvl in range(0, o.par.depth - 1):
        L, H = o.L(lvl), o.H
        _, raw = rand_poly_raw(o, lvl, 40 + lvl)
        c = o.crt(raw, lvl)
        d = e.up(c)
        e.call("cuhe_mod_switch", p(d), p(d), p(d[L - 1]), lvl, e.st())
        got = Eng.dn(d, np.uint32)[: L - 1]
        assert np.array_equal(got, o.modswitch(c, lvl))
    assert e.lib.cuhe_mod_switch(e.h, p(d), p(d), p(d), o.par.depth - 1, e.st()) != 0   # last level refuses


@pytest.mark.parametrize("ps", [SMALL_RELIN, MID32K, MID64K])
def test_relin(lib, ps):
    e = Eng(lib, ps)
    try:
        o = e.orc
        K0, W0, H, N = o.par.numEvalKey, o.W(0), o.H, o.N
        rng = random.Random(11)
        eks = [o.to_raw([rng.randrange(o.moduli[0]) for _ in range(o.n)], 0) for _ in range(K0)]
        o.init_relin(eks)
        d_eks = e.up(np.stack(eks))
        e.call("cuhe_relin_init", p(d_eks), e.st())
        for lvl in (0, o.par.depth - 1):
            L = o.L(lvl)
            _, raw = rand_poly_raw(o, lvl, 70 + lvl)
            d_out = e.empty((L, N), np.uint64)
            e.call("cuhe_relin", p(d_out), p(e.up(raw)), lvl, e.st())
            want = o.relin_mac(raw, lvl)
            assert np.array_equal(Eng.dn(d_out, np.uint64), want)
            # followed by the reference's n2c (isProd=true): inttMod
            d_c = e.empty((L, H), np.uint32)
            e.call("cuhe_intt_mod", p(d_c), p(d_out), lvl, e.st())
            assert np.array_equal(Eng.dn(d_c, np.uint32), o.intt_mod(want))
    finally:
        e.close()


def test_relin_keys_export_import_through_the_binary_container(lib, tmp_path):
    """cuhe_relin_export_host after cuhe_relin_init equals the oracle's transformed keys (h_ek of
    cuhe/Relinearization.cu:43-56); written with save_rns, read back and imported into a FRESH context, the key switch
    gives the same words without any CRT / transform work at start-up."""
    from cuhe_b200 import utils
    e = Eng(lib, SMALL_RELIN)
    e2 = Eng(lib, SMALL_RELIN)
    try:
        o = e.orc
        K0, N, L = o.par.numEvalKey, o.N, o.L(0)
        rng = random.Random(21)
        eks = [o.to_raw([rng.randrange(o.moduli[0]) for _ in range(o.n)], 0) for _ in range(K0)]
        o.init_relin(eks)
        e.call("cuhe_relin_init", p(e.up(np.stack(eks))), e.st())
        words = lib.cuhe_relin_key_words(e.h)
        assert words == L * K0 * N
        host = np.zeros((L, K0, N), dtype=np.uint64)
        e.call("cuhe_relin_export_host", host.ctypes.data_as(C.c_void_p), C.c_size_t(words), e.st())
        assert np.array_equal(host, o.ek)
        path = str(tmp_path / "ek.rns")
        utils.save_rns(path, host, SMALL_RELIN, domain=3, level=0)
        back, meta = utils.load_rns(path)
        assert meta["params"] == tuple(SMALL_RELIN)
        back = np.ascontiguousarray(back)
        e2.call("cuhe_relin_import_host", back.ctypes.data_as(C.c_void_p), C.c_size_t(words), e2.st())
        _, raw = rand_poly_raw(o, 0, 77)
        d1, d2 = e.empty((L, N), np.uint64), e2.empty((L, N), np.uint64)
        e.call("cuhe_relin", p(d1), p(e.up(raw)), 0, e.st())
        e2.call("cuhe_relin", p(d2), p(e2.up(raw)), 0, e2.st())
        assert np.array_equal(Eng.dn(d2, np.uint64), o.relin_mac(raw, 0))
        assert np.array_equal(Eng.dn(d1, np.uint64), Eng.dn(d2, np.uint64))
    finally:
        e.close()
        e2.close()


def test_mul_raw_host_end_to_end(eng16):
    e, o = eng16, eng16.orc
    lvl = 0
    a, ra = rand_poly_raw(o, lvl, 5)
    b, rb = rand_poly_raw(o, lvl, 6)
    out = np.zeros_like(ra)
    e.call("cuhe_mul_raw_host", out.ctypes.data_as(C.c_void_p), ra.ctypes.data_as(C.c_void_p),
           rb.ctypes.data_as(C.c_void_p), lvl, e.st())
    assert o.from_raw(out) == o.mul_exact(a, b, lvl)


def test_sharded_contexts_match_unsharded(lib):
    """Residue sharding (rank r of G owns primes r, r+G, ...): two shard
    contexts on one device reproduce the single-context result row for row."""
    ps = SIMPLE_DHS
    full = Eng(lib, ps)
    shards = [Eng(lib, ps, rank=r, world=2) for r in range(2)]
    try:
        o = full.orc
        lvl = 0
        L, H, N = o.L(lvl), o.H, o.N
        _, ra = rand_poly_raw(o, lvl, 21)
        _, rb = rand_poly_raw(o, lvl, 22)
        want = o.mul_raw_to_crt(ra, rb, lvl)
        gathered = np.zeros((L, H), dtype=np.uint32)
        for r, e in enumerate(shards):
            rows = e.rows(lvl)
            assert rows == len(range(r, L, 2))
            da, db = e.up(ra), e.up(rb)
            ca, cb = e.empty((rows, H), np.uint32), e.empty((rows, H), np.uint32)
            e.call("cuhe_crt", p(ca), p(da), lvl, e.st())
            e.call("cuhe_crt", p(cb), p(db), lvl, e.st())
            na, nb = e.empty((rows, N), np.uint64), e.empty((rows, N), np.uint64)
            e.call("cuhe_ntt", p(na), p(ca), lvl, e.st())
            e.call("cuhe_ntt", p(nb), p(cb), lvl, e.st())
            out = e.empty((rows, H), np.uint32)
            e.call("cuhe_ntt_mul_intt_mod", p(out), p(na), p(nb), lvl, e.st())
            gathered[r::2] = Eng.dn(out, np.uint32)
        assert np.array_equal(gathered, want)
        # ICRT on the gathered residues, split by coefficient range across "ranks"
        W = o.W(lvl)
        d_all = shards[0].up(gathered)
        raw = shards[0].empty((H, W), np.uint32)
        shards[0].call("cuhe_icrt", p(raw), p(d_all), lvl, 0, H // 2, shards[0].st())
        shards[1].call("cuhe_icrt", p(raw), p(d_all), lvl, H // 2, H, shards[1].st())
        shards[0].torch.cuda.synchronize()
        assert np.array_equal(Eng.dn(raw, np.uint32), o.icrt(want, lvl))
        # sliced ICRT (the all-to-all form): each "rank" holds only its coefficient slice of every residue
        Hs = H // 2
        for r, e in enumerate(shards):
            sl = np.ascontiguousarray(gathered[None, :, r * Hs:(r + 1) * Hs])            # [1][L][Hs]
            raw_sl = e.empty((1, Hs, W), np.uint32)
            e.call("cuhe_icrt_slice_batch", p(raw_sl), p(e.up(sl)), lvl, r * Hs, Hs, 1, e.st())
            want_sl = o.icrt(want, lvl)[r * Hs:(r + 1) * Hs]
            assert np.array_equal(Eng.dn(raw_sl, np.uint32)[0], want_sl)
        # modswitch with the dropped row supplied by its owner
        last_owner = (L - 1) % 2
        d_last = shards[0].up(want[L - 1])
        ms = o.modswitch(want, lvl)
        for r, e in enumerate(shards):
            loc = e.up(np.ascontiguousarray(want[r::2]))
            e.call("cuhe_mod_switch", p(loc), p(loc), p(d_last), lvl, e.st())
            keep = len(range(r, L - 1, 2))
            assert np.array_equal(Eng.dn(loc, np.uint32)[:keep], ms[r::2])
        assert last_owner in (0, 1)
    finally:
        full.close()
        for e in shards:
            e.close()


def test_full_size_c2_properties(lib):
    """BASELINE configs[1] (N=65536, 24 primes): oracle parity on the whole
    multiply plus size-independent properties (linearity of the transform,
    NTT/INTT round trip, ICRT o CRT = identity)."""
    e = Eng(lib, C2)
    try:
        o = e.orc
        lvl = 0
        L, W, H, N = o.L(lvl), o.W(lvl), o.H, o.N
        a, ra = rand_poly_raw(o, lvl, 31)
        b, rb = rand_poly_raw(o, lvl, 32)
        out = np.zeros_like(ra)
        e.call("cuhe_mul_raw_host", out.ctypes.data_as(C.c_void_p), ra.ctypes.data_as(C.c_void_p),
               rb.ctypes.data_as(C.c_void_p), lvl, e.st())
        want_crt = o.mul_raw_to_crt(ra, rb, lvl)
        assert np.array_equal(out, o.icrt(want_crt, lvl))
        # commutativity through the host entry point
        out2 = np.zeros_like(ra)
        e.call("cuhe_mul_raw_host", out2.ctypes.data_as(C.c_void_p), rb.ctypes.data_as(C.c_void_p),
               ra.ctypes.data_as(C.c_void_p), lvl, e.st())
        assert np.array_equal(out, out2)
        # CRT -> ICRT identity and NTT linearity on device
        d_raw = e.up(ra)
        d_crt = e.empty((L, H), np.uint32)
        e.call("cuhe_crt", p(d_crt), p(d_raw), lvl, e.st())
        d_raw2 = e.empty((H, W), np.uint32)
        e.call("cuhe_icrt", p(d_raw2), p(d_crt), lvl, 0, H, e.st())
        assert np.array_equal(Eng.dn(d_raw2, np.uint32), ra)
        d_crtb = e.empty((L, H), np.uint32)
        e.call("cuhe_crt", p(d_crtb), p(e.up(rb)), lvl, e.st())
        na, nb, ns = (e.empty((L, N), np.uint64) for _ in range(3))
        e.call("cuhe_ntt", p(na), p(d_crt), lvl, e.st())
        e.call("cuhe_ntt", p(nb), p(d_crtb), lvl, e.st())
        d_sum = e.empty((L, H), np.uint32)
        e.call("cuhe_crt_add", p(d_sum), p(d_crt), p(d_crtb), lvl, e.st())
        e.call("cuhe_ntt", p(ns), p(d_sum), lvl, e.st())
        nab = e.empty((L, N), np.uint64)
        e.call("cuhe_ntt_add", p(nab), p(na), p(nb), lvl, e.st())
        # NTT(a+b mod p) differs from NTT(a)+NTT(b) by NTT of the p-multiples; compare after INTT % p
        c1, c2 = e.empty((L, H), np.uint32), e.empty((L, H), np.uint32)
        e.call("cuhe_intt", p(c1), p(ns), lvl, e.st())
        e.call("cuhe_intt", p(c2), p(nab), lvl, e.st())
        assert np.array_equal(Eng.dn(c1, np.uint32), Eng.dn(c2, np.uint32))
    finally:
        e.close()


def _relin_rows_oracle(o, eks, raw, rows):
    """Oracle key switch (cuhe/Relinearization.cu:76-88) for a subset of residue rows at level 0:
    sum_k NTT(digit_k) * NTT(ek_k mod p_l).  Only the listed rows' keys are transformed."""
    from oracle import oracle as orc
    N, K0 = o.N, o.par.numEvalKey
    rows = list(rows)
    acc = {l: np.zeros(N, dtype=np.uint64) for l in rows}
    for k in range(K0):
        D = orc.ntt_ext(o.digits(raw, 0, k), N)
        ek = orc.ntt_ext(o.crt(eks[k], 0)[rows], N)
        for i, l in enumerate(rows):
            acc[l] = orc.ntt_add(acc[l], orc.ntt_mul(D, ek[i]))
    return acc


def test_config3_relin_full_size(lib):
    """BASELINE configs[2]: key switch at N=65536, 44 primes, 66 evaluation keys (1.52 GB of keys
    resident in HBM).  Full-size parity on the first and last residue rows at level 0 (the oracle
    transforms only those rows' keys; every row goes through the same kernel)."""
    ps = (44, 2, 16, 24, 24, 32767)
    e = Eng(lib, ps)
    try:
        o = e.orc
        K0, W0, H, N, L0 = o.par.numEvalKey, o.W(0), o.H, o.N, o.L0
        assert (N, L0, K0) == (65536, 44, 66)
        rng = np.random.default_rng(5)
        eks = rng.integers(0, 1 << 32, size=(K0, H, W0), dtype=np.uint32)
        eks[:, o.n:, :] = 0
        eks[:, :, W0 - 1] = 0                           # keep every coefficient below q0 (33 words, 1056 bits)
        e.call("cuhe_relin_init", p(e.up(eks)), e.st())
        raw = rng.integers(0, 1 << 32, size=(H, W0), dtype=np.uint32)
        raw[o.n:] = 0
        raw[:, W0 - 1] = 0
        d_out = e.empty((L0, N), np.uint64)
        e.call("cuhe_relin", p(d_out), p(e.up(raw)), 0, e.st())
        got = Eng.dn(d_out, np.uint64)
        want = _relin_rows_oracle(o, eks, raw, (0, L0 - 1))
        for l, acc in want.items():
            assert np.array_equal(got[l], acc), f"relin row {l} (prime {o.primes[l]})"
    finally:
        e.close()


@pytest.mark.parametrize("ps", [(64, 2, 16, 24, 24, 32767), (25, 2, 16, 25, 25, 21845)],
                         ids=["config5_64primes_1536bit", "prince_32k_25primes"])
def test_large_prime_counts_mul(lib, ps):
    """BASELINE configs[4] (64 primes, ~1536-bit modulus) and the Prince parameter set (N=32768, 25
    primes): whole multiply RAW -> RAW against the oracle."""
    e = Eng(lib, ps)
    try:
        o = e.orc
        _, ra = rand_poly_raw(o, 0, 61)
        _, rb = rand_poly_raw(o, 0, 62)
        out = np.zeros_like(ra)
        e.call("cuhe_mul_raw_host", out.ctypes.data_as(C.c_void_p), ra.ctypes.data_as(C.c_void_p),
               rb.ctypes.data_as(C.c_void_p), 0, e.st())
        assert np.array_equal(out, o.icrt(o.mul_raw_to_crt(ra, rb, 0), 0))
    finally:
        e.close()


@pytest.mark.parametrize("batch", [1, 5, 16, 37, 70])
def test_mul_raw_host_batch_pipeline(eng16, batch):
    """cuhe_mul_raw_host_batch: the three-stream H2D | kernels | D2H pipeline with its ramped chunk
    schedule (one chunk below 16 products, 4-8-...-4 from 32, 8-16-...-8 from 64) returns, for every
    product of the batch, exactly what the oracle computes."""
    e, o = eng16, eng16.orc
    W, H, n = o.W(0), o.H, o.n
    rng = np.random.default_rng(100 + batch)
    top_bits = o.moduli[0].bit_length() - 1 - 32 * (W - 1)            # keep every coefficient below q0
    def polys():
        x = rng.integers(0, 1 << 32, size=(batch, H, W), dtype=np.uint32)
        x[:, :, W - 1] &= np.uint32((1 << top_bits) - 1)
        x[:, n:, :] = 0
        return x
    a, b = polys(), polys()
    torch = e.torch
    ah = torch.from_numpy(a.view(np.int32)).pin_memory()
    bh = torch.from_numpy(b.view(np.int32)).pin_memory()
    oh = torch.full((batch, H, W), -1, dtype=torch.int32).pin_memory()
    e.call("cuhe_mul_raw_host_batch", C.c_void_p(oh.data_ptr()), C.c_void_p(ah.data_ptr()), C.c_void_p(bh.data_ptr()),
           0, batch, e.st())
    torch.cuda.synchronize()
    got = oh.numpy().view(np.uint32)
    want = o.mul_raw_batch(a, b, 0)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("terms", [1, 2, 66, 141, 1000])
def test_modp_wide_accumulator(eng16, terms):
    """The key-switch accumulator (160-bit unreduced sum of 64x64-bit products, folded once with
    2^96 == -1, 2^128 == -2^32): equals big-integer arithmetic, including all-maximal operands."""
    x, y = _modp_inputs(4096, 7 + terms)
    n = x.size
    x[200:264] = P - 1
    y[200:264] = P - 1
    out = eng16.empty(x.shape, np.uint64)
    dx, dy = eng16.up(x), eng16.up(y)                       # keep both alive: p() only takes the address
    eng16.call("cuhe_modp_batch", 4, p(out), p(dx), p(dy), C.c_size_t(n), terms, eng16.st())
    got = Eng.dn(out, np.uint64)
    prod = (x.astype(object) * y.astype(object))
    ext = np.concatenate([prod, prod[:terms]])
    want = np.array([int(sum(ext[i:i + terms])) % P for i in range(n)], dtype=object)
    assert np.array_equal(got.astype(object), want)


def test_modp_canonical_residue(eng16):
    x = np.array([0, 1, P - 1, P, P + 1, 2**64 - 1, 2**64 - 2**32, 2**63, 0xFFFFFFFF, 0xFFFFFFFF00000000], dtype=np.uint64)
    x = np.concatenate([x, np.random.default_rng(3).integers(0, 2**64, size=4086, dtype=np.uint64)])
    out = eng16.empty(x.shape, np.uint64)
    dx = eng16.up(x)
    eng16.call("cuhe_modp_batch", 5, p(out), p(dx), None, C.c_size_t(x.size), 0, eng16.st())
    assert np.array_equal(Eng.dn(out, np.uint64).astype(object), x.astype(object) % P)


def test_modp_primitives_equal_the_reference_header(eng16):
    """Differential test against the REFERENCE's own device code: oracle/_ref/libref_modp.so wraps
    _add/_sub/_mul/_ls_modP of cuhe/ModP.h (compiled for sm_100a from the reference tree by oracle/Makefile,
    target `ref`).  Inputs follow tests/test_ModP.cu:38-49; shifts are the l = 3*a*b the reference supports."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_modp.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_modp.so was not built (no reference tree at build time)")
    ref = C.CDLL(path)
    ref.ref_modp_batch.restype = C.c_int
    ref.ref_modp_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    x, y = _modp_inputs(1 << 18, 77)
    dx, dy = eng16.up(x), eng16.up(y)
    ours, theirs = eng16.empty(x.shape, np.uint64), eng16.empty(x.shape, np.uint64)
    torch = eng16.torch
    for op in (0, 1, 2):
        eng16.call("cuhe_modp_batch", op, p(ours), p(dx), p(dy), C.c_size_t(x.size), 0, eng16.st())
        assert ref.ref_modp_batch(op, p(theirs), p(dx), p(dy), x.size, 0, eng16.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ours, theirs), f"op {op}"
    for l in sorted({3 * a * b for a in range(8) for b in range(8)}):
        eng16.call("cuhe_modp_batch", 3, p(ours), p(dx), None, C.c_size_t(x.size), l, eng16.st())
        assert ref.ref_modp_batch(3, p(theirs), p(dx), None, x.size, l, eng16.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ours, theirs), f"shift {l}"


def test_gpu_outputs_match_the_committed_golden_fixtures(lib):
    """The GPU path against tests/golden/golden.json directly (no oracle call on this side): SHA-256 of cRep,
    nRep, product cRep, product rRep and the modswitch output for the seeded inputs of make_golden.py, whose
    products were verified against exact big-integer ring arithmetic when the fixtures were written --
    including the full BASELINE size (24 primes, N = 65536)."""
    import hashlib
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]

    def sha(t, dt):
        return hashlib.sha256(np.ascontiguousarray(Eng.dn(t, dt)).tobytes()).hexdigest()
    for case in gold:
        e = Eng(lib, tuple(case["params"]))
        try:
            o = e.orc                                           # only for sizes and the seeded inputs
            a, b = mg.inputs(o, case["seed"])
            ra, rb = o.to_raw(a, 0), o.to_raw(b, 0)
            L, H, N, W = o.L(0), o.H, o.N, o.W(0)
            d_ra, d_rb = e.up(ra), e.up(rb)
            d_crt = e.empty((L, H), np.uint32)
            e.call("cuhe_crt", p(d_crt), p(d_ra), 0, e.st())
            assert sha(d_crt, np.uint32) == case["crt_sha"], case["name"]
            d_ntt = e.empty((L, N), np.uint64)
            e.call("cuhe_ntt", p(d_ntt), p(d_crt), 0, e.st())
            assert sha(d_ntt, np.uint64) == case["ntt_sha"], case["name"]
            d_mc = e.empty((1, L, H), np.uint32)
            e.call("cuhe_mul_crt_batch", p(d_mc), p(d_ra), p(d_rb), 0, 1, e.st())
            assert sha(d_mc, np.uint32) == case["mul_crt_sha"], case["name"]
            d_mr = e.empty((H, W), np.uint32)
            e.call("cuhe_icrt", p(d_mr), p(d_mc), 0, 0, H, e.st())
            assert sha(d_mr, np.uint32) == case["mul_raw_sha"], case["name"]
            if case["modswitch_sha"]:
                d_ms = d_mc[0].clone()
                e.call("cuhe_mod_switch", p(d_ms), p(d_ms), p(d_ms[L - 1]), 0, e.st())
                assert hashlib.sha256(np.ascontiguousarray(Eng.dn(d_ms, np.uint32)[: L - 1]).tobytes()).hexdigest() \
                    == case["modswitch_sha"], case["name"]
        finally:
            e.close()
