import sys, os, random, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from test_gpu_parity import Eng, p, rand_poly_raw
from common import SMALL_RELIN
from cuhe_b200 import load_library
lib = load_library()
e = Eng(lib, SMALL_RELIN); o = e.orc
K0, N = o.par.numEvalKey, o.N
rng = random.Random(11)
eks = [o.to_raw([rng.randrange(o.moduli[0]) for _ in range(o.n)], 0) for _ in range(K0)]
o.init_relin(eks)
e.call("cuhe_relin_init", p(e.up(np.stack(eks))), e.st())
for lvl in (0, o.par.depth - 1):
    L = o.L(lvl)
    _, raw = rand_poly_raw(o, lvl, 70 + lvl)
    d_out = e.empty((L, N), np.uint64)
    e.call("cuhe_relin", p(d_out), p(e.up(raw)), lvl, e.st())
    got = Eng.dn(d_out, np.uint64); want = o.relin_mac(raw, lvl)
    bad = got != want
    print("RB", os.environ.get("CUHE_B200_RELIN_RB"), "lvl", lvl, "L", L, "K", o.K(lvl), "mismatches per row", bad.sum(axis=1), "first cols", np.nonzero(bad.any(axis=0))[0][:8])
    if bad.any():
        r, c = np.argwhere(bad)[0]
        print("  got", hex(int(got[r, c])), "want", hex(int(want[r, c])), "diff", hex((int(got[r, c]) - int(want[r, c])) % (2**64 - 2**32 + 1)))
