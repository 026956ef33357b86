"""BASELINE configs[3] and the reference's only fixed known answer, on the GPU: homomorphic PRINCE
over DHS at (25, 2, 16, 25, 25, 21845) -- N = 32768, 25 primes, 40 evaluation keys, 24 levels --
through the cuHE interface (examples/Prince/Prince.cu:56-98).  1920 cAnd, 1152 relin and 2688
modSwitch calls must leave ciphertexts that decrypt to 9fb51935fc3df524.

Also the same S-box schedule cross-checked ciphertext-for-ciphertext against the CPU oracle engine
(same keys, same randomness => identical ZZX outputs, not merely identical decryptions)."""
import random
import time

import pytest

import prince_he as ph
from common import get_oracle

pytestmark = pytest.mark.gpu


def _gpu():
    import torch
    assert torch.cuda.is_available()
    import cuhe_b200 as ch
    return ch


def test_sbox_ciphertexts_equal_oracle_engine():
    from dhs_host import DHS
    from oracle_engine import OracleEngine
    ps = (5, 2, 16, 25, 25, 8191)
    o = get_oracle(ps)
    ch = _gpu()
    try:
        gpu = DHS(ch, *ps, phi=o.phi, seed=3)
        cpu = DHS(OracleEngine(), *ps, phi=o.phi, seed=3)
        assert gpu.pk[0] == cpu.pk[0] and gpu.sk[0] == cpu.sk[0]
        rng = random.Random(9)
        bits = [rng.randrange(2) for _ in range(4)]
        cg = [gpu.encrypt([b], 0) for b in bits]
        cc = [cpu.encrypt([b], 0) for b in bits]
        assert cg == cc
        og = ph.HomOps(ch, gpu)._sbox(cg, 0, ph.ANF_FWD)
        oc = ph.HomOps(cpu.ch, cpu)._sbox(cc, 0, ph.ANF_FWD)
        assert og == oc, "S-box output ciphertexts differ between the GPU path and the oracle"
        v = ph.SBOX[int("".join(map(str, bits)), 2)]
        assert [gpu.decrypt(x, 2)[0] for x in og] == [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
    finally:
        ch.resetParameters()


@pytest.fixture(scope="module")
def prince_keys():
    from dhs_host import DHS
    ch = _gpu()
    o = get_oracle(ph.PRINCE_PARAMS)
    t0 = time.time()
    dhs = DHS(ch, *ph.PRINCE_PARAMS, phi=o.phi, seed=2026)
    print(f"\nkeygen {time.time() - t0:.1f} s")
    yield ch, dhs
    ch.resetParameters()


@pytest.mark.parametrize("resident", [False, True, "batched"], ids=["host_linear_layers", "device_resident", "batched_layers"])
def test_homomorphic_prince_known_answer(prince_keys, resident):
    """host_linear_layers: the reference's flow (ZZX values on the host between S-box layers,
    Prince.cu:146-189, 460-468).  device_resident: SURVEY 8(f) N2 -- every ciphertext stays on the device
    from encryption to decryption (cXor / cNot / dropToLevel for the linear layers)."""
    ch, dhs = prince_keys
    ch.launch_count(reset=True)
    t1 = time.time()
    bits, ops = ph.hom_prince(ch, dhs, [0] * 64, [1] * 64, [0] * 64, check_rounds=(0, 11), resident=resident)
    t_eval = time.time() - t1
    import json
    import os
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "prince_kat_timings.jsonl"), "a") as f:
            f.write(json.dumps({"mode": str(resident), "seconds": t_eval, "split": ops.seconds, "launches": ch.launch_count(), "ops": ops.counts}) + "\n")
    except OSError:
        pass
    print(f"\n{ {False: 'host linear layers', True: 'device-resident', 'batched': 'device-resident, 16 S-boxes per launch set'}[resident]}: encrypt+evaluate+decrypt {t_eval:.1f} s, "
          f"split {ops.seconds}, ops {ops.counts}, kernel launches {ch.launch_count()}")
    assert ops.counts == dict(cAnd=1920, relin=1152, modSwitch=2688, sbox=192)
    assert ops.round_bits == ops.round_want
    assert ph.bits_to_hex(bits) == ph.KAT_HEX
    if resident == "batched":
        assert ch.launch_count() < 10000 and ops.seconds["evaluate"] < 6.0


def test_batched_sbox_layer_equals_the_per_ciphertext_path():
    """One S-box layer (16 S-boxes) through cuhe_b200.circuit (batched entry points) and through the per-ciphertext
    CuCtxt operations on the same encrypted bits: identical ciphertext words for every output bit."""
    from dhs_host import DHS
    ps = (5, 2, 16, 25, 25, 8191)
    o = get_oracle(ps)
    ch = _gpu()
    try:
        dhs = DHS(ch, *ps, phi=o.phi, seed=5)
        rng = random.Random(10)
        bits = [rng.randrange(2) for _ in range(64)]
        one, bat = ph.DeviceHomOps(ch, dhs), ph.BatchedHomOps(ch, dhs)
        state = [one.upload(dhs.encrypt([b], 0)) for b in bits]
        a = one.sbox_layer(list(state), False)
        b = bat.sbox_layer(list(state), False)
        assert len(a) == len(b) == 64
        for x, y in zip(a, b):
            assert x.level() == y.level() == 2 and x.domain() == y.domain() == 2
            assert bool((x.cRep() == y.cRep()).all())
        for i in range(16):
            v = ph.SBOX[int("".join(map(str, bits[4 * i:4 * i + 4])), 2)]
            got = [dhs.decrypt(bat.to_zzx(c), 2)[0] for c in b[4 * i:4 * i + 4]]
            assert got == [(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1]
    finally:
        ch.resetParameters()
