#!/usr/bin/env python
"""The reference's own multi-GPU semantics on hardware (cuhe/DeviceManager.cu:50-70, cuhe/CuHE.cu:217-257): one
process, whole ciphertexts per device, tables replicated by initCuHE on every device, moveTo / copyTo between them.
Needs two GPUs.  Checked against the oracle: a ciphertext copied to device 1 keeps its words; a product computed on
device 1 from operands that were set up on device 0 equals the exact ring product; mixing devices raises."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import cuhe_b200 as ch  # noqa: E402
from common import SMALL_RELIN, get_oracle  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2, "needs two GPUs"
    o = get_oracle(SMALL_RELIN)
    ch.resetParameters()
    ch.multiGPUs(2)
    ch.setParameters(*SMALL_RELIN)
    assert ch.initCuHE(o.phi) == o.moduli
    rng = random.Random(77)
    a = [rng.randrange(o.moduli[0]) for _ in range(o.n)]
    b = [rng.randrange(o.moduli[0]) for _ in range(o.n)]
    u32 = lambda t: t.cpu().numpy().view(np.uint32)  # noqa: E731
    ca, cb, cc = ch.CuCtxt(), ch.CuCtxt(), ch.CuCtxt()
    ca.setLevel(0, 0, a)
    ca.x2c()
    ch.copyTo(cc, ca, 1)                               # CRT domain, device 0 -> 1
    assert cc.device() == 1 and cc.cRep().device.index == 1 and ca.device() == 0
    assert np.array_equal(u32(cc.cRep()), u32(ca.cRep()))
    cb.setLevel(0, 0, b)
    cb.x2n()                                           # NTT domain on device 0 ...
    ch.moveTo(cb, 1)                                   # ... moved as it is
    assert cb.device() == 1 and cb.nRep().device.index == 1
    try:
        ch.cAnd(ca, ca, cb)
        raise SystemExit("cAnd across devices did not raise")
    except ch.CuHEError:
        pass
    cc.x2n()                                           # transforms on device 1 (its own tables)
    ch.cAnd(cc, cc, cb)
    cc.x2z()
    want = o.mul_exact(a, b, 0)
    assert cc.zRep() == want, "product computed on device 1 differs from the exact ring product"
    ch.moveTo(cb, 0)                                   # and back: device 0 still multiplies
    ca.x2n()
    ch.cAnd(ca, ca, cb)
    ca.x2z()
    assert ca.zRep() == want
    assert ch.mulZZX(a, b, 0, 1) == want               # mulZZX(..., dev = 1)
    ch.resetParameters()
    print('{"multidev_check": "ok"}')


if __name__ == "__main__":
    main()
