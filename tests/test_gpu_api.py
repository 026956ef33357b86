"""GPU tests of the host-side mirror of cuhe/CuHE.h (cuhe_b200/api.py): the
reference's public call sequences (mulZZX, the CuCtxt domain machine, cAnd /
cXor / cNot, relin + modSwitch as in examples/DHS/simple_DHS.cu:49-163) against
the oracle, plus the reference's error behaviour (terminate() -> CuHEError)."""
import random

import numpy as np
import pytest

from common import SMALL_RELIN, get_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ch():
    import torch
    assert torch.cuda.is_available()
    import cuhe_b200 as ch
    ch.resetParameters()
    ch.multiGPUs(1)
    ch.setParameters(*SMALL_RELIN)
    o = get_oracle(SMALL_RELIN)
    coeff_mod = ch.initCuHE(o.phi)
    assert coeff_mod == o.moduli                      # initCuHE fills coeffMod (cuhe/Operations.cu:157-160)
    assert ch.crtPrimes() == o.primes                 # genCrtPrimes (cuhe/Operations.cu:37-80)
    yield ch
    ch.resetParameters()


def rnd(o, lvl, seed):
    rng = random.Random(seed)
    return [rng.randrange(o.moduli[lvl]) for _ in range(o.n)]


def u32(t):
    return t.cpu().numpy().view(np.uint32)


def u64(t):
    return t.cpu().numpy().view(np.uint64)


def test_param_mirror(ch):
    o = get_oracle(SMALL_RELIN)
    p = ch.param
    assert (p.modLen, p.nttLen, p.numCrtPrime, p.numEvalKey, p.logCrtPrime) == \
        (o.par.modLen, o.par.nttLen, o.par.numCrtPrime, o.par.numEvalKey, o.par.logCrtPrime)
    for lvl in range(p.depth):
        assert p._numCrtPrime(lvl) == o.L(lvl) and p._wordsCoeff(lvl) == o.W(lvl)
        assert p._getLevel(p._logCoeff(lvl)) == lvl


def test_mulzzx_equals_exact(ch):
    o = get_oracle(SMALL_RELIN)
    for lvl in (0, 1):
        a, b = rnd(o, lvl, 1 + lvl), rnd(o, lvl, 5 + lvl)
        assert ch.mulZZX(a, b, lvl, 0) == o.mul_exact(a, b, lvl)


def test_domain_machine_round_trips(ch):
    o = get_oracle(SMALL_RELIN)
    a = rnd(o, 0, 3)
    c = ch.CuCtxt()
    c.setLevel(0, 0, a)
    assert (c.domain(), c.level(), c.logq()) == (0, 0, o.par._logCoeff(0))
    c.x2r()
    assert c.domain() == 1 and np.array_equal(u32(c.rRep()), o.to_raw(a, 0))
    c.x2c()
    assert c.domain() == 2 and c.rRep() is None
    assert np.array_equal(u32(c.cRep()), o.crt(o.to_raw(a, 0), 0))
    c.x2n()
    assert c.domain() == 3 and c.cRep() is None
    assert np.array_equal(u64(c.nRep()), o.ntt(o.crt(o.to_raw(a, 0), 0)))
    assert not c.isProd()
    c.x2z()                                            # n2c (plain intt), c2r, r2z
    assert c.domain() == 0 and c.zRep() == a
    c.reset()
    c.reset()                                          # idempotent (explicit destructor calls in Prince.cu:298-318)
    assert c.domain() == -1


def test_cand_cxor_cnot(ch):
    o = get_oracle(SMALL_RELIN)
    a, b = rnd(o, 0, 7), rnd(o, 0, 8)
    ca, cb, cx = ch.CuCtxt(), ch.CuCtxt(), ch.CuCtxt()
    ca.setLevel(0, 0, a)
    cb.setLevel(0, 0, b)
    ca.x2c()
    cb.x2c()
    ra, rb = o.crt(o.to_raw(a, 0), 0), o.crt(o.to_raw(b, 0), 0)
    ch.cXor(cx, ca, cb)                                # CRT-domain add
    assert cx.domain() == 2 and np.array_equal(u32(cx.cRep()), o.crt_add(ra, rb))
    ch.cNot(cx, cx)                                    # in place, as the examples use it
    assert np.array_equal(u32(cx.cRep()), o.crt_add_int(o.crt_add(ra, rb), ch.param.modMsg - 1))
    ca.x2n()
    cb.x2n()
    cs = ch.CuCtxt()
    ch.cXor(cs, ca, cb)                                # NTT-domain add
    from oracle import oracle as orc
    assert cs.domain() == 3 and np.array_equal(u64(cs.nRep()), orc.ntt_add(o.ntt(ra), o.ntt(rb)))
    ch.cAnd(ca, ca, cb)                                # &out == &in0 supported (cuhe/CuHE.cu:114-117)
    assert ca.isProd() and ca.domain() == 3
    ca.x2c()                                           # inttMod (Barrett) because isProd
    assert not ca.isProd()
    assert np.array_equal(u32(ca.cRep()), o.mul_raw_to_crt(o.to_raw(a, 0), o.to_raw(b, 0), 0))
    # ctxt x ptxt (ntt_mul_nx1) and ctxt + ptxt (crt_add_nx1)
    m = [random.Random(9).randrange(2) for _ in range(o.n)]
    pt = ch.CuPtxt()
    pt.setLogq(ch.param.logMsg, 0, m)
    pt.x2n()
    cb2 = ch.CuCtxt()
    cb2.setLevel(0, 0, b)
    cb2.x2n()
    prod = ch.CuCtxt()
    ch.cAnd(prod, cb2, pt)
    mraw = np.zeros((o.H,), dtype=np.uint32)
    mraw[:o.n] = m
    want = orc.ntt_mul(o.ntt(rb), np.broadcast_to(orc.ntt_ext(mraw, o.N), (o.L(0), o.N)).copy())
    assert np.array_equal(u64(prod.nRep()), want)
    # plaintext domain machine: level -1 is one residue (cuhe/Parameters.cu:107-109)
    pt2 = ch.CuPtxt()
    pt2.setLogq(ch.param.logMsg, 0, m)
    pt2.x2n()
    assert pt2.nRep().shape[0] == 1
    pt2.x2z()
    assert pt2.zRep() == m


def test_relin_modswitch_chain(ch):
    """cAnd -> relin -> modSwitch as in examples/DHS/simple_DHS.cu:120-163 (with
    random evaluation keys: parity of every domain value, not decryption)."""
    o = get_oracle(SMALL_RELIN)
    rng = random.Random(12)
    eks = [[rng.randrange(o.moduli[0]) for _ in range(o.n)] for _ in range(o.par.numEvalKey)]
    ch.initRelinearization(eks)
    o.init_relin([o.to_raw(e, 0) for e in eks])
    a, b = rnd(o, 0, 13), rnd(o, 0, 14)
    ca, cb = ch.CuCtxt(), ch.CuCtxt()
    ca.setLevel(0, 0, a)
    cb.setLevel(0, 0, b)
    ca.x2n()
    cb.x2n()
    ch.cAnd(ca, ca, cb)
    ca.relin()                                         # x2r (inttMod + icrt), digit NTTs, MAC, n2c
    prod_crt = o.mul_raw_to_crt(o.to_raw(a, 0), o.to_raw(b, 0), 0)
    want = o.intt_mod(o.relin_mac(o.icrt(prod_crt, 0), 0))
    assert ca.domain() == 2 and np.array_equal(u32(ca.cRep()), want)
    ca.modSwitch()
    assert ca.level() == 1 and ca.logq() == o.par._logCoeff(1)
    assert np.array_equal(u32(ca.cRep()), o.modswitch(want, 0))
    ca.x2z()
    assert ca.zRep() == o.from_raw(o.icrt(np.ascontiguousarray(o.modswitch(want, 0)), 1))


def test_copy_and_errors(ch):
    o = get_oracle(SMALL_RELIN)
    a = rnd(o, 0, 21)
    ca, cc = ch.CuCtxt(), ch.CuCtxt()
    ca.setLevel(0, 0, a)
    ca.x2c()
    ch.copy(cc, ca)
    assert cc.domain() == 2 and np.array_equal(u32(cc.cRep()), u32(ca.cRep())) and cc.cRep() is not ca.cRep()
    ch.moveTo(cc, 0)                                   # same device: no-op
    cb = ch.CuCtxt()
    cb.setLevel(1, 0, rnd(o, 1, 22))
    cb.x2c()
    with pytest.raises(ch.CuHEError):                  # "Multiplication of non-NTT domain!"
        ch.cAnd(cc, ca, cb)
    with pytest.raises(ch.CuHEError):                  # "Addition of different levels!"
        ch.cXor(cc, ca, cb)
    ca.x2n()
    with pytest.raises(ch.CuHEError):                  # "cNot of non-CRT domain!"
        ch.cNot(ca, ca)
    last = ch.CuCtxt()
    last.setLevel(ch.param.depth - 1, 0, rnd(o, ch.param.depth - 1, 23))
    with pytest.raises(ch.CuHEError):                  # "Cannot do modSwitch on last level!"
        last.modSwitch()
    with pytest.raises(ch.CuHEError):                  # z2r from the wrong domain
        ca.z2r()


def test_operations_on_an_explicit_non_default_stream(ch):
    """Every operation takes an optional stream (cuhe/CuHE.h:149-208).  With a stream that is NOT torch's current one
    the buffers are still allocated / zero-filled on the current stream and dropped right after the call, so the
    mirror has to order the launch after the fill and block like the reference does (cudaStreamSynchronize(st) after
    every operation, cuhe/CuHE.cu:81-268).  A busy current stream makes a missing dependency visible."""
    import torch
    o = get_oracle(SMALL_RELIN)
    st = torch.cuda.Stream()
    a, b = rnd(o, 0, 17), rnd(o, 0, 18)
    want = o.mul_raw_to_crt(o.to_raw(a, 0), o.to_raw(b, 0), 0)
    ballast = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        for _ in range(20):
            ballast.add_(1)                            # keep the current stream busy ahead of the allocations
        ca, cb = ch.CuCtxt(), ch.CuCtxt()
        ca.setLevel(0, 0, a)
        cb.setLevel(0, 0, b)
        ca.x2n(st)
        cb.x2n(st)
        cs = ch.CuCtxt()
        ch.cXor(cs, ca, cb, st)
        ch.cAnd(ca, ca, cb, st)
        ca.x2c(st)
        assert np.array_equal(u32(ca.cRep()), want)
        cs.x2c(st)
        ra, rb = o.crt(o.to_raw(a, 0), 0), o.crt(o.to_raw(b, 0), 0)
        assert np.array_equal(u32(cs.cRep()), o.crt_add(ra, rb))
        ch.cNot(cs, cs, st)
        assert np.array_equal(u32(cs.cRep()), o.crt_add_int(o.crt_add(ra, rb), ch.param.modMsg - 1))
        ca.x2z(st)
        assert ca.zRep() == o.mul_exact(a, b, 0)
