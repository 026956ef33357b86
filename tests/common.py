"""Shared parameter sets and oracle cache for the tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# parameter sets (d, p, w, min, cut, m)
SIMPLE_DHS = (5, 2, 1, 61, 20, 8191)        # examples/DHS/simple_DHS.cu:218 -> N=16384, L=7
SMALL_RELIN = (3, 2, 16, 40, 20, 8191)      # N=16384, small relin (K=5)
PRINCE = (25, 2, 16, 25, 25, 21845)         # examples/Prince/Prince.cu -> N=32768, L=25
MID32K = (4, 2, 16, 50, 25, 21845)          # N=32768, L=5, K=8
C2 = (24, 2, 16, 24, 24, 32767)             # BASELINE configs[1] -> N=65536, L=24
MID64K = (3, 2, 16, 48, 24, 32767)          # N=65536, L=4, K=6

_ORACLES = {}


def get_oracle(params):
    from oracle.oracle import Oracle
    params = tuple(params)
    if params not in _ORACLES:
        _ORACLES[params] = Oracle(*params)
    return _ORACLES[params]
