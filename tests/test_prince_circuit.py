"""The PRINCE known answer of the reference (examples/Prince/Prince.cu:96,109-144) -- CPU half.

1. the cipher restated in tests/prince_he.py is PRINCE (specification test vectors);
2. its per-round states for the reference's inputs are the twelve strings the reference prints;
3. the S-box polynomials the homomorphic schedule evaluates are the S-box (ANF == table) and are the
   ones the reference's schedule adds up (Prince.cu:245-292, 378-426);
4. the homomorphic schedule itself, run on the CPU oracle behind the cuHE interface
   (tests/oracle_engine.py) with real DHS keys, decrypts to the S-box of the plaintext bits;
5. (opt-in, ~20 min of CPU: CUHE_B200_SLOW=1) the whole homomorphic PRINCE on the oracle decrypts to
   9fb51935fc3df524 -- this is what pins the ORACLE to the reference's only fixed vector; the log of
   the committed run is tests/golden/prince_kat_oracle.log."""
import os
import random

import pytest

import prince_he as ph
from common import get_oracle

# examples/Prince/Prince.cu:109-144 -- state after each of the 12 S-box layers for m = 0^64, k0 = 1^64, k1 = 0^64
REFERENCE_ROUND_STATES = [
    "0100010001000100010001000100010001000100010001000100010001000100",
    "1100000111000101111011011001100010100001001010100010000110111011",
    "0001010111110110111001101000001101110010101111110010111100010111",
    "0000111110110100100011001100001110111010101010110110101101110000",
    "0011100101111101011100000001110101111100101110010111101100111110",
    "0110001011001101101111001000001100011000011100100010110011100011",
    "1111000000000111010110001001011111100101001011001111001001101110",
    "1110011011001010101100101000110011100000011101111010000011101110",
    "1111010001000111111011111011110001100001100000001111011100100100",
    "0010001000000000000010101101010110110101101010110110011110101111",
    "0011101110000011000111101111001010110001111011110111111101111011",
    "1010000011100110110011110111110111001010101111100101101000000111",
]
REFERENCE_FINAL = "1001111110110101000110010011010111111100001111011111010100100100"   # Prince.cu:96

SPEC_VECTORS = [  # (plaintext, k0, k1, ciphertext) -- PRINCE specification, appendix A
    ("0000000000000000", "0000000000000000", "0000000000000000", "818665aa0d02dfda"),
    ("ffffffffffffffff", "0000000000000000", "0000000000000000", "604ae6ca03c20ada"),
    ("0000000000000000", "ffffffffffffffff", "0000000000000000", "9fb51935fc3df524"),
    ("0000000000000000", "0000000000000000", "ffffffffffffffff", "78a54cbe737bb7ef"),
    ("0123456789abcdef", "0000000000000000", "fedcba9876543210", "ae25ad3ca8fa9ccf"),
]


@pytest.mark.parametrize("ops", [ph.BitOps, ph.AnfBitOps], ids=["table", "anf"])
def test_cipher_is_prince(ops):
    for pt, k0, k1, ct in SPEC_VECTORS:
        out = ph.prince_eval(ops(), ph.hex_to_bits(pt), ph.hex_to_bits(k0), ph.hex_to_bits(k1))
        assert ph.bits_to_hex(out) == ct


def test_reference_known_answer_and_round_states():
    assert ph.bits_to_hex([int(c) for c in REFERENCE_FINAL]) == ph.KAT_HEX
    states = ph.kat_round_states()
    assert sorted(states) == list(range(12))
    for r, want in enumerate(REFERENCE_ROUND_STATES):
        assert "".join(map(str, states[r])) == want, f"state after S-box layer {r}"
    out = ph.prince_eval(ph.BitOps(), [0] * 64, [1] * 64, [0] * 64)
    assert "".join(map(str, out)) == REFERENCE_FINAL


def test_sbox_polynomials_are_the_reference_schedule():
    a, b, c, d = 0, 1, 2, 3
    one = ()
    fwd = [  # Prince.cu:245-292
        {(a,), (c,), (a, b), (b, c), one, (a, b, d), (a, c, d), (b, c, d)},
        {(a,), (d,), (a, c), (a, d), (c, d), (a, b, c), (a, c, d)},
        {(a, c), (b, c), (b, d), one, (a, b, c), (b, c, d)},
        {(a,), (b,), (a, b), (a, d), (b, c), (c, d), one, (b, c, d)},
    ]
    inv = [  # Prince.cu:378-426
        {(c,), (d,), (a, b), (b, c), (b, d), (c, d), one, (a, b, c), (a, b, d), (b, c, d)},
        {(b,), (d,), (a, c), (b, c), (b, d), (c, d), (a, c, d), (b, c, d)},
        {(a, b), (a, c), (b, c), (b, d), one, (b, c, d)},
        {(a,), (a, b), (b, c), (c, d), one, (a, b, d), (a, c, d)},
    ]
    assert [set(m) for m in ph.ANF_FWD] == fwd
    assert [set(m) for m in ph.ANF_INV] == inv


def _engine_and_keys(ps, seed):
    from dhs_host import DHS
    from oracle_engine import OracleEngine
    eng = OracleEngine()
    o = get_oracle(ps)
    return eng, DHS(eng, *ps, phi=o.phi, seed=seed)


def test_homomorphic_sbox_on_the_oracle_engine():
    """One S-box and one inverse S-box (two levels each) through the reference's schedule, real keys."""
    ps = (5, 2, 16, 25, 25, 8191)
    eng, dhs = _engine_and_keys(ps, seed=3)
    ops = ph.HomOps(eng, dhs)
    rng = random.Random(9)
    bits = [rng.randrange(2) for _ in range(4)]
    cts = [dhs.encrypt([b], 0) for b in bits]
    out = ops._sbox(cts, 0, ph.ANF_FWD)
    got = [dhs.decrypt(x, 2)[0] for x in out]
    v = ph.SBOX[int("".join(map(str, bits)), 2)]
    assert got == [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
    back = ops._sbox(out, 2, ph.ANF_INV)
    assert [dhs.decrypt(x, 4)[0] for x in back] == bits
    assert ops.counts == dict(cAnd=20, relin=12, modSwitch=28, sbox=2)   # 10 / 6 / 14 per S-box


def test_device_resident_sbox_and_key_addition_on_the_oracle_engine():
    """The device-resident form: ciphertexts stay CuCtxt objects in the CRT domain, fresh
    key bits are brought to the current level with dropToLevel, linear steps are cXor / cNot."""
    ps = (5, 2, 16, 25, 25, 8191)
    eng, dhs = _engine_and_keys(ps, seed=4)
    ops = ph.DeviceHomOps(eng, dhs)
    rng = random.Random(10)
    bits = [rng.randrange(2) for _ in range(4)]
    key = [rng.randrange(2) for _ in range(4)]
    cts = [ops.upload(dhs.encrypt([b], 0)) for b in bits]
    kts = [ops.upload(dhs.encrypt([b], 0)) for b in key]           # level 0, used again two levels down
    out = ops._sbox(cts, 0, ph.ANF_FWD)
    assert all(c.level() == 2 and c.domain() == 2 for c in out)
    v = ph.SBOX[int("".join(map(str, bits)), 2)]
    sb = [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
    assert [dhs.decrypt(ops.to_zzx(c), 2)[0] for c in out] == sb
    ops.level = 2
    mixed = [ops.add_const(ops.add(c, k), 1) for c, k in zip(out, kts)]       # state + key + 1 at level 2
    assert all(k.level() == 0 for k in kts)                                    # the keys themselves are untouched
    want = [s ^ k ^ 1 for s, k in zip(sb, key)]
    assert [dhs.decrypt(ops.to_zzx(c), 2)[0] for c in mixed] == want
    back = ops._sbox(mixed, 2, ph.ANF_INV)
    w = ph.SBOX_INV[int("".join(map(str, want)), 2)]
    assert [dhs.decrypt(ops.to_zzx(c), 4)[0] for c in back] == [(w >> 3) & 1, (w >> 2) & 1, (w >> 1) & 1, w & 1]


@pytest.mark.skipif(os.environ.get("CUHE_B200_SLOW") != "1", reason="~20 min of CPU; set CUHE_B200_SLOW=1")
@pytest.mark.parametrize("resident", [False, True], ids=["host_linear_layers", "device_resident"])
def test_full_prince_on_the_oracle_engine(resident):
    import time
    eng, dhs = _engine_and_keys(ph.PRINCE_PARAMS, seed=2026)
    t0 = time.time()
    log_path = os.path.join(os.path.dirname(__file__), "golden",
                            "prince_kat_oracle_resident.log" if resident else "prince_kat_oracle.log")
    lines = []

    def log(msg):
        lines.append(f"[{time.time() - t0:7.1f} s] {msg}")
        print(lines[-1], flush=True)

    bits, ops = ph.hom_prince(eng, dhs, [0] * 64, [1] * 64, [0] * 64, check_rounds=(0, 5, 11), log=log, resident=resident)
    log("decrypted: " + "".join(map(str, bits)))
    log("expected : " + REFERENCE_FINAL)
    log(f"op counts: {ops.counts}")
    with open(log_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    assert ops.round_bits == ops.round_want
    assert "".join(map(str, bits)) == REFERENCE_FINAL
