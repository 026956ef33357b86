"""Generates tests/golden/golden.json: SHA-256 of the oracle's outputs at every
domain boundary (cRep, nRep, product cRep, product rRep, modswitch) for seeded
inputs.  With exact=True the product is additionally verified against exact
big-integer ring arithmetic ((a*b mod Phi_m) mod q through GMP Kronecker
multiplication -- the role of NTL's ZZX multiply in the reference,
examples/DHS/DHS.cu:219-221) before its hash is recorded.

The reference itself cannot be run here (needs NTL; texture references are
rejected by nvcc 12.9) and stores no vectors, so these fixtures are oracle
outputs pinned by exact arithmetic, not reference outputs.

    python tests/golden/make_golden.py          # rewrites golden.json
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CASES = [
    dict(name="simple_dhs_16k", params=[5, 2, 1, 61, 20, 8191], seed=20260924, size="small"),
    dict(name="mid_32k", params=[4, 2, 16, 50, 25, 21845], seed=20260925, size="small"),
    dict(name="mid_64k", params=[3, 2, 16, 48, 24, 32767], seed=20260926, size="small"),
    dict(name="c2_64k_24primes", params=[24, 2, 16, 24, 24, 32767], seed=20260927, size="full"),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(o, seed, lvl=0):
    rng = random.Random(seed)
    q = o.moduli[lvl]
    a = [rng.randrange(q) for _ in range(o.n)]
    b = [rng.randrange(q) for _ in range(o.n)]
    return a, b


def compute_case(params, seed, exact=False):
    from common import get_oracle
    o = get_oracle(tuple(params))
    a, b = inputs(o, seed)
    ra, rb = o.to_raw(a, 0), o.to_raw(b, 0)
    ca = o.crt(ra, 0)
    na = o.ntt(ca)
    mc = o.mul_raw_to_crt(ra, rb, 0)
    mr = o.icrt(mc, 0)
    out = dict(crt_sha=sha(ca), ntt_sha=sha(na), mul_crt_sha=sha(mc), mul_raw_sha=sha(mr),
               modswitch_sha=sha(o.modswitch(mc, 0)) if o.par.depth > 1 else "")
    if exact:
        ex = o.mul_exact(a, b, 0)
        assert o.from_raw(mr) == ex, "oracle product differs from exact ring arithmetic"
        out["exact_verified"] = True
    return out


def main():
    cases = []
    for c in CASES:
        r = compute_case(c["params"], c["seed"], exact=True)
        cases.append({**c, **r})
        print(c["name"], "ok")
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
