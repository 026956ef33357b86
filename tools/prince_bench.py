#!/usr/bin/env python
"""BASELINE configs[3]: homomorphic PRINCE end to end on one GPU (examples/Prince/Prince.cu:96,109-144): DHS keys at
(25, 2, 16, 25, 25, 21845) -- N = 32768, 25 CRT primes, 40 evaluation keys, 24 levels -- encryption of the 192 input
bits, evaluation with every ciphertext resident on the device and the 16 S-boxes of a layer going through one launch
set (cuhe_b200/circuit.py), decryption, and the comparison with the reference's known answer 9fb51935fc3df524.
The circuit and the DHS host side are the callers restated for the tests (tests/prince_he.py, tests/dhs_host.py).
Prints one JSON line; a child process of bench.py (key `prince_config4`)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import cuhe_b200 as ch
    import prince_he as ph
    from cuhe_b200.hostmath import cyclotomic
    from dhs_host import DHS
    assert torch.cuda.is_available()
    ch.resetParameters()
    ch.multiGPUs(1)
    t0 = time.time()
    dhs = DHS(ch, *ph.PRINCE_PARAMS, phi=cyclotomic(ph.PRINCE_PARAMS[5]), seed=2026)
    t_keys = time.time() - t0
    ch.launch_count(reset=True)
    t1 = time.time()
    bits, ops = ph.hom_prince(ch, dhs, [0] * 64, [1] * 64, [0] * 64, check_rounds=(), resident="batched")
    torch.cuda.synchronize()
    t_all = time.time() - t1
    got = ph.bits_to_hex(bits)
    print(json.dumps({
        "workload": "homomorphic PRINCE, N = 32768, 25 CRT primes, 40 evaluation keys, 24 levels (BASELINE configs[3])",
        "prince_s": t_all, "split_s": ops.seconds, "keygen_s": t_keys, "launches": ch.launch_count(), "ops": ops.counts,
        "decrypts_to": got, "known_answer": ph.KAT_HEX, "known_answer_reproduced": got == ph.KAT_HEX,
        "note": "prince_s = encrypt 192 bits (Python host) + evaluate (device resident, 16 S-boxes per launch set) + "
                "decrypt 64 bits (Python host)"}))
    ch.resetParameters()
    return 0 if got == ph.KAT_HEX else 1


if __name__ == "__main__":
    sys.exit(main())
