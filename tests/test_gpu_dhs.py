"""End-to-end homomorphic test with REAL keys, mirroring examples/DHS/simple_DHS.cu:49-163:
decrypt o op o encrypt == plaintext op for cXor, cNot and cAnd + relin + modSwitch, on
polynomial plaintexts (coefficients mod 2).  Every ring product of key generation,
encryption and decryption goes through the GPU path (mulZZX), as in the reference's
DHS.cu; the pass criterion is the reference's own (decryption equals the plaintext
operation), which also validates the noise behaviour of relin and modSwitch."""
import random

import numpy as np
import pytest

from common import SMALL_RELIN, get_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dhs():
    import torch
    assert torch.cuda.is_available()
    import cuhe_b200 as ch
    from dhs_host import DHS
    o = get_oracle(SMALL_RELIN)
    d = DHS(ch, *SMALL_RELIN, phi=o.phi, seed=5)
    yield d
    ch.resetParameters()


def plain_mul_mod2(a, b, phi):
    n = len(phi) - 1
    prod = np.convolve(np.array(a, dtype=np.int64), np.array(b, dtype=np.int64))
    f = [int(v) for v in prod]
    from oracle import pyoracle as po
    return [c % 2 for c in po.poly_mod_phi(f, phi)[:n]]


def test_encrypt_decrypt_roundtrip(dhs):
    rng = random.Random(1)
    for lvl in range(dhs.par.depth):
        m = [rng.randrange(2) for _ in range(dhs.n)]
        assert dhs.decrypt(dhs.encrypt(m, lvl), lvl) == m


def test_homomorphic_xor_not(dhs):
    ch = dhs.ch
    rng = random.Random(2)
    m0 = [rng.randrange(2) for _ in range(dhs.n)]
    m1 = [rng.randrange(2) for _ in range(dhs.n)]
    c0, c1 = ch.CuCtxt(), ch.CuCtxt()
    c0.setLevel(0, 0, dhs.encrypt(m0, 0))
    c1.setLevel(0, 0, dhs.encrypt(m1, 0))
    c0.x2c()
    c1.x2c()
    cx = ch.CuCtxt()
    ch.cXor(cx, c0, c1)                                   # simple_DHS.cu:62-80
    ch.cNot(c1, c1)                                       # simple_DHS.cu:98
    cx.x2z()
    c1.x2z()
    assert dhs.decrypt(cx.zRep(), 0) == [(a + b) % 2 for a, b in zip(m0, m1)]
    want_not = list(m1)
    want_not[0] = (want_not[0] + 1) % 2                   # crt_add_int touches coefficient 0 (cuhe/Base.cu:1096-1100)
    assert dhs.decrypt(c1.zRep(), 0) == want_not


def test_homomorphic_and_relin_modswitch(dhs):
    """simple_DHS.cu:120-163: x2n, cAnd, relin, modSwitch, decrypt at the next level."""
    ch = dhs.ch
    rng = random.Random(3)
    o = get_oracle(SMALL_RELIN)
    m0 = [rng.randrange(2) for _ in range(dhs.n)]
    m1 = [rng.randrange(2) for _ in range(dhs.n)]
    c0, c1 = ch.CuCtxt(), ch.CuCtxt()
    c0.setLevel(0, 0, dhs.encrypt(m0, 0))
    c1.setLevel(0, 0, dhs.encrypt(m1, 0))
    c0.x2n()
    c1.x2n()
    ch.cAnd(c0, c0, c1)
    c0.relin()
    c0.modSwitch()
    assert c0.level() == 1
    c0.x2z()
    want = plain_mul_mod2(m0, m1, o.phi)
    assert dhs.decrypt(c0.zRep(), 1) == want
    # a second multiplicative level
    c2 = ch.CuCtxt()
    c2.setLevel(1, 0, dhs.encrypt(m1, 1))
    c3 = ch.CuCtxt()
    c3.setLevel(1, 0, c0.zRep())
    c2.x2n()
    c3.x2n()
    ch.cAnd(c3, c3, c2)
    c3.relin()
    c3.modSwitch()
    c3.x2z()
    assert dhs.decrypt(c3.zRep(), 2) == plain_mul_mod2(want, m1, o.phi)
