"""Differential tests against the REFERENCE'S OWN DEVICE KERNELS: oracle/_ref/libref_base.so is
cuhe/Base.cu (all 18 NTT/INTT kernels, crt, icrt, the five Barrett kernels, relinMulAddPerCrt,
ntt_mul/add[_nx1], crt_add*, modswitch) compiled for sm_100a from the reference tree by
oracle/Makefile (target `ref`: NTL root generation -> 128-bit C, texture references -> pointers,
nothing else changed; wrapper oracle/ref_base.cu restates the launch sequences of cuhe/Operations.cu).
Every test runs the shipped C ABI and the reference kernels on the same device buffers and compares
bit for bit; tables are loaded into the reference from the oracle's restatement of
cuhe/Operations.cu:37-144, so the tests also pin those tables to what the reference kernels expect."""
import ctypes as C
import os

import numpy as np
import pytest

from common import MID32K, MID64K, ROOT, SIMPLE_DHS, SMALL_RELIN
from test_gpu_parity import Eng, p, rand_poly_raw

pytestmark = pytest.mark.gpu
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libref_base.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_PATH):
        pytest.skip("oracle/_ref/libref_base.so was not built (no reference tree at build time)")
    r = C.CDLL(REF_PATH)
    for name in ("ref_base_preload_ntt", "ref_base_preload_primes", "ref_base_load_icrt", "ref_base_preload_barrett",
                 "ref_base_ntt_ext", "ref_base_nttw", "ref_base_intt_modcrt", "ref_base_crt", "ref_base_icrt",
                 "ref_base_barrett", "ref_base_relin_mac", "ref_base_pointwise", "ref_base_crt_add", "ref_base_modswitch"):
        getattr(r, name).restype = C.c_int
    r._ntt_loaded = set()
    return r


def _np32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _load_tables(ref, e, lvl=0):
    """what initNtt / initCrt / loadIcrtConst do (cuhe/Operations.cu:37-184), from the oracle's tables"""
    o = e.orc
    if o.N not in ref._ntt_loaded:
        assert ref.ref_base_preload_ntt(o.N) == 0
        ref._ntt_loaded.add(o.N)
    primes, invp = _np32(o.primes), _np32(o.invp)
    assert ref.ref_base_preload_primes(primes.ctypes.data_as(C.c_void_p), len(primes),
                                       invp.ctypes.data_as(C.c_void_p), len(invp)) == 0
    ic = o.icrt_const(lvl)
    M, mi, bi = _np32(ic.q), _np32(ic.qp), _np32(ic.qpinv)
    assert ref.ref_base_load_icrt(M.ctypes.data_as(C.c_void_p), M.size, mi.ctypes.data_as(C.c_void_p), mi.size,
                                  bi.ctypes.data_as(C.c_void_p), bi.size) == 0


@pytest.mark.parametrize("ps", [SIMPLE_DHS, MID32K, MID64K])
def test_forward_and_inverse_transforms_equal_the_reference_kernels(lib, ref, ps):
    """ntt_{1,2,3}_*_ext and intt_1 / ntt_2 / intt_3_*_modcrt (cuhe/Base.cu:309-842), batched through gridDim.y
    like tests/test_ntt.cu:67-100, on that test's input distribution (31-bit values, garbage in the unread half)."""
    e = Eng(lib, ps)
    try:
        _load_tables(ref, e)
        o, torch = e.orc, e.torch
        N, B = o.N, 6
        rng = np.random.default_rng(11)
        x = rng.integers(0, 1 << 31, size=(B, N), dtype=np.uint32)      # upper half: garbage, never read
        x[0, : N // 2] = 0
        x[1, : N // 2] = 0x7FFFFFFF
        dx = e.up(x)
        ours, theirs, swap = e.empty((B, N), np.uint64), e.empty((B, N), np.uint64), e.empty((B, N), np.uint64)
        e.call("cuhe_ntt_ext_batch", p(ours), p(dx), N, B, C.c_longlong(N), e.st())
        assert ref.ref_base_ntt_ext(p(theirs), p(swap), p(dx), N, B, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ours, theirs)
        # inverse + % p_0 of those spectra: cuhe_intt_double_deg works on the rows of a level, so compare row 0..
        L = o.L(0)
        spec = ours[:1].repeat(L, 1).contiguous()                       # the same spectrum for every residue row
        o32, t32 = e.empty((L, N), np.uint32), e.empty((L, N), np.uint32)
        e.call("cuhe_intt_double_deg", p(o32), p(spec), 0, e.st())
        for l in range(L):
            assert ref.ref_base_intt_modcrt(p(t32[l]), p(swap), p(spec[l]), N, 1, l, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(o32, t32)
    finally:
        e.close()


@pytest.mark.parametrize("ps", [SIMPLE_DHS, MID32K, MID64K])
def test_crt_icrt_pointwise_adds_modswitch_equal_the_reference_kernels(lib, ref, ps):
    e = Eng(lib, ps)
    try:
        _load_tables(ref, e)
        o, torch = e.orc, e.torch
        L, H, N, W, n = o.L(0), o.H, o.N, o.W(0), o.n
        _, ra = rand_poly_raw(o, 0, 31)
        _, rb = rand_poly_raw(o, 0, 32)
        d_ra, d_rb = e.up(ra), e.up(rb)
        # crt (cuhe/Base.cu:857-879)
        ca, ta = e.empty((L, H), np.uint32), e.empty((L, H), np.uint32)
        cb = e.empty((L, H), np.uint32)
        e.call("cuhe_crt", p(ca), p(d_ra), 0, e.st())
        e.call("cuhe_crt", p(cb), p(d_rb), 0, e.st())
        assert ref.ref_base_crt(p(ta), p(d_ra), L, W, n, H, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ca, ta)
        # icrt (cuhe/Base.cu:880-924)
        ro, rt = e.empty((H, W), np.uint32), e.empty((H, W), np.uint32)
        e.call("cuhe_icrt", p(ro), p(ca), 0, 0, H, e.st())
        assert ref.ref_base_icrt(p(rt), p(ca), L, W, o.W(1), n, H, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ro[:n], rt[:n])
        assert torch.equal(ro[:n].cpu(), torch.from_numpy(ra[:n].view(np.int32)))
        # CRT-domain adds (cuhe/Base.cu:1088-1109)
        so, st_ = e.empty((L, H), np.uint32), e.empty((L, H), np.uint32)
        e.call("cuhe_crt_add", p(so), p(ca), p(cb), 0, e.st())
        assert ref.ref_base_crt_add(0, p(st_), p(ca), p(cb), 0, L, n, H, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(so[:, :n], st_[:, :n])
        e.call("cuhe_crt_add_nx1", p(so), p(ca), p(cb[0]), 0, e.st())
        assert ref.ref_base_crt_add(1, p(st_), p(ca), p(cb[0]), 0, L, n, H, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(so[:, :n], st_[:, :n])
        so.copy_(ca); st_.copy_(ca)
        e.call("cuhe_crt_add_int", p(so), p(ca), 1, 0, e.st())
        assert ref.ref_base_crt_add(2, p(st_), p(ca), None, 1, L, n, H, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(so, st_)
        # NTT-domain pointwise (cuhe/Base.cu:1036-1075)
        na, nb = e.empty((L, N), np.uint64), e.empty((L, N), np.uint64)
        e.call("cuhe_ntt", p(na), p(ca), 0, e.st())
        e.call("cuhe_ntt", p(nb), p(cb), 0, e.st())
        zo, zt = e.empty((L, N), np.uint64), e.empty((L, N), np.uint64)
        for op, name, y in ((0, "cuhe_ntt_mul", nb), (1, "cuhe_ntt_add", nb), (2, "cuhe_ntt_mul_nx1", nb[0]), (3, "cuhe_ntt_add_nx1", nb[0])):
            e.call(name, p(zo), p(na), p(y), 0, e.st())
            assert ref.ref_base_pointwise(op, p(zt), p(na), p(y), L, N, e.st()) == 0
            torch.cuda.synchronize()
            assert torch.equal(zo, zt), name
        # modswitch (cuhe/Base.cu:1112-1138), in place like crtModSwitch
        if o.par.depth > 1:
            mo, mt = ca.clone(), ca.clone()
            e.call("cuhe_mod_switch", p(mo), p(mo), p(mo[L - 1]), 0, e.st())
            assert ref.ref_base_modswitch(p(mt), p(mt), L, n, H, o.par.modMsg, e.st()) == 0
            torch.cuda.synchronize()
            assert torch.equal(mo[: L - 1, :n], mt[: L - 1, :n])
    finally:
        e.close()


@pytest.mark.parametrize("ps", [SIMPLE_DHS, MID32K, MID64K])
def test_barrett_reduction_equals_the_reference_kernel_sequence(lib, ref, ps):
    """inttMod = per-residue inverse transforms + the Barrett launch sequence of cuhe/Operations.cu:460-501 on the
    reference's kernels, against cuhe_ntt_mul_intt_mod (default: fold / inverse-series reduction) and cuhe_barrett
    (the literal step order): three implementations, one canonical remainder."""
    e = Eng(lib, ps)
    try:
        _load_tables(ref, e)
        o, torch = e.orc, e.torch
        L, H, N, n = o.L(0), o.H, o.N, o.n
        t = o.barrett_tables()
        d_u, d_m, d_mc = e.up(t["u_ntt"]), e.up(t["m_ntt"]), e.up(t["m_crt"])
        assert ref.ref_base_preload_barrett(p(d_u), p(d_m), p(d_mc), L, N, H) == 0
        _, ra = rand_poly_raw(o, 0, 41)
        _, rb = rand_poly_raw(o, 0, 42)
        ca, cb = e.empty((L, H), np.uint32), e.empty((L, H), np.uint32)
        e.call("cuhe_crt", p(ca), p(e.up(ra)), 0, e.st())
        e.call("cuhe_crt", p(cb), p(e.up(rb)), 0, e.st())
        na, nb = e.empty((L, N), np.uint64), e.empty((L, N), np.uint64)
        e.call("cuhe_ntt", p(na), p(ca), 0, e.st())
        e.call("cuhe_ntt", p(nb), p(cb), 0, e.st())
        ours = e.empty((L, H), np.uint32)
        e.call("cuhe_ntt_mul_intt_mod", p(ours), p(na), p(nb), 0, e.st())
        # reference: ntt_mul, L x _intt into hold, barrett()
        prod, swap = e.empty((L, N), np.uint64), e.empty((N,), np.uint64)
        assert ref.ref_base_pointwise(0, p(prod), p(na), p(nb), L, N, e.st()) == 0
        hold, hold2 = e.empty((L, N), np.uint32), e.empty((L, N), np.uint32)
        for l in range(L):
            assert ref.ref_base_intt_modcrt(p(hold[l]), p(swap), p(prod[l]), N, 1, l, e.st()) == 0
        hold2.copy_(hold)
        lit = e.empty((L, H), np.uint32)
        e.call("cuhe_barrett", p(lit), p(hold2), 0, e.st())
        theirs, pcrt, pntt = e.empty((L, H), np.uint32), e.empty((L, N), np.uint32), e.empty((L, N), np.uint64)
        assert ref.ref_base_barrett(p(theirs), p(hold), p(pcrt), p(pntt), p(swap), L, N, H, n, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ours, theirs)
        assert torch.equal(lit, theirs)
    finally:
        e.close()


@pytest.mark.parametrize("ps", [SMALL_RELIN, MID64K])
def test_relin_digits_and_inner_product_equal_the_reference_kernels(lib, ref, ps):
    """nttw (ntt_1_*_ext_block, cuhe/Base.cu:345-385...) + relinMulAddPerCrt (cuhe/Base.cu:1024-1033) per residue,
    with the key layout of cuhe/Relinearization.cu:43-88, against cuhe_relin."""
    e = Eng(lib, ps)
    try:
        _load_tables(ref, e)
        o, torch = e.orc, e.torch
        L, H, N, W, K = o.L(0), o.H, o.N, o.W(0), o.K(0)
        rng = np.random.default_rng(5)
        q0 = o.moduli[0]
        import random
        pr = random.Random(9)
        eks = np.stack([o.to_raw([pr.randrange(q0) for _ in range(o.n)], 0) for _ in range(K)])
        e.call("cuhe_relin_init", p(e.up(eks)), e.st())
        _, raw = rand_poly_raw(o, 0, 51)
        d_raw = e.up(raw)
        ours = e.empty((L, N), np.uint64)
        e.call("cuhe_relin", p(ours), p(d_raw), 0, e.st())
        # reference: K digit transforms, then per residue ek[l] = NTT(crt(key_k) row l) and the MAC kernel
        D, swap = e.empty((K, N), np.uint64), e.empty((N,), np.uint64)
        for k in range(K):
            assert ref.ref_base_nttw(p(D[k]), p(swap), p(d_raw), N, o.par.logRelin, k, W, e.st()) == 0
        ek = e.empty((L, K, N), np.uint64)
        crt_k, ntt_k = e.empty((L, H), np.uint32), e.empty((L, N), np.uint64)
        d_eks = e.up(eks)
        for k in range(K):
            e.call("cuhe_crt", p(crt_k), p(d_eks[k]), 0, e.st())
            e.call("cuhe_ntt", p(ntt_k), p(crt_k), 0, e.st())
            ek[:, k, :] = ntt_k
        theirs = e.empty((L, N), np.uint64)
        for l in range(L):
            assert ref.ref_base_relin_mac(p(theirs[l]), p(D), p(ek[l]), K, N, e.st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(ours, theirs)
        del rng
    finally:
        e.close()
