#!/usr/bin/env python
"""Relinearization (key switch) timing at BASELINE config 3: setParameters(44,2,16,24,24,32767)
-> N=65536, L=44 primes, K=66 evaluation keys (1.52 GB resident in HBM).
Random keys (the arithmetic does not depend on key values).  CUDA events."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuhe_b200._lib import check, cuhe_params, load_library  # noqa: E402
from cuhe_b200.hostmath import cyclotomic  # noqa: E402


def main():
    ps = (44, 2, 16, 24, 24, 32767)
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), *ps))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    phi = np.array(cyclotomic(ps[5]), dtype=np.int64)
    check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    dev = torch.device("cuda", 0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L, K, N, H = par.numCrtPrime, par.numEvalKey, par.nttLen, par.crtLen
    W = lib.cuhe_param_words_coeff(C.byref(par), 0)
    eks = torch.randint(0, 2**31 - 1, (K, H, W), dtype=torch.int32, device=dev)
    eks[:, par.modLen:, :] = 0
    check(lib.cuhe_relin_init(h, p(eks), st()))
    del eks
    raw = torch.randint(0, 2**31 - 1, (H, W), dtype=torch.int32, device=dev)
    raw[par.modLen:] = 0
    out = torch.zeros((L, N), dtype=torch.int64, device=dev)
    crt = torch.zeros((L, H), dtype=torch.int32, device=dev)

    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_relin = timeit(lambda: check(lib.cuhe_relin(h, p(out), p(raw), 0, st())))
    ms_chain = timeit(lambda: (check(lib.cuhe_relin(h, p(out), p(raw), 0, st())),
                               check(lib.cuhe_intt_mod(h, p(crt), p(out), 0, st()))))
    mac_bytes = 8 * N * (L * K + K + L)
    print(json.dumps({"config": "relin N=65536 L=44 K=66", "relin_ms": ms_relin, "relin_per_s": 1e3 / ms_relin,
                      "relin_plus_inttmod_ms": ms_chain, "keyswitch_per_s": 1e3 / ms_chain,
                      "mac_algorithmic_GB": mac_bytes / 1e9,
                      "digit_ntt_plus_mac_GBps_if_all_time_were_mac": mac_bytes / ms_relin / 1e6}))


if __name__ == "__main__":
    main()
