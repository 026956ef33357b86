#!/usr/bin/env python
"""CUDA-event timings of the HBM-bound kernels of the path (pointwise NTT-domain ops, CRT, ICRT,
modswitch, relin MAC) against their algorithmic bytes, at BASELINE config 2 sizes
(N=65536, L=24, W=18, batch 8) and config 3 for the MAC (L=44, K=66)."""
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


def timeit(fn, reps=20):
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


def main():
    lib = load_library()
    dev = torch.device("cuda", 0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {"hbm_peak_gbs": peak, "kernels": {}}

    def rec(name, nbytes, ms):
        gbs = nbytes / ms / 1e6
        out["kernels"][name] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": gbs, "frac_of_hbm_peak": gbs / peak}

    # ---- config 2 ----
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    L, N, H, n = par.numCrtPrime, par.nttLen, par.crtLen, par.modLen
    W = lib.cuhe_param_words_coeff(C.byref(par), 0)
    B = 8
    # pointwise: rotate over 6 buffer sets (> L2)
    xs = [torch.randint(0, 2**62, (L, N), dtype=torch.int64, device=dev) for _ in range(6)]
    ys = [torch.randint(0, 2**62, (L, N), dtype=torch.int64, device=dev) for _ in range(6)]
    zs = [torch.zeros((L, N), dtype=torch.int64, device=dev) for _ in range(6)]
    k = [0]

    def pw(fn):
        def run():
            i = k[0] % 6
            k[0] += 1
            check(fn(h, p(zs[i]), p(xs[i]), p(ys[i]), 0, st()))
        return run
    rec("ntt_mul (24 residues)", 24 * N * L, timeit(pw(lib.cuhe_ntt_mul)))
    rec("ntt_add (24 residues)", 24 * N * L, timeit(pw(lib.cuhe_ntt_add)))
    raw = torch.randint(0, 2**31 - 1, (4, B, H, W), dtype=torch.int32, device=dev)
    raw[:, :, n:, :] = 0
    crt = torch.zeros((4, B, L, H), dtype=torch.int32, device=dev)
    j = [0]

    def crt_run():
        i = j[0] % 4
        j[0] += 1
        check(lib.cuhe_crt(h, p(crt[i, 0]), p(raw[i, 0]), 0, st()))
    rec("crt (1 polynomial)", 4 * H * W + 4 * H * L, timeit(crt_run))
    rawo = torch.zeros((4, B, H, W), dtype=torch.int32, device=dev)
    crt.random_(0, 2**23)

    def icrt_run():
        i = j[0] % 4
        j[0] += 1
        check(lib.cuhe_icrt_batch(h, p(rawo[i]), p(crt[i]), 0, 0, H, B, st()))
    rec("icrt (batch 8)", B * (4 * H * W + 4 * H * L), timeit(icrt_run))

    def ms_run():
        i = j[0] % 4
        j[0] += 1
        check(lib.cuhe_mod_switch(h, p(crt[i, 0]), p(crt[i, 0]), p(crt[i, 0, L - 1]), 0, st()))
    rec("modswitch (1 polynomial)", 2 * 4 * H * (L - 1) + 4 * H, timeit(ms_run))
    check(lib.cuhe_ctx_destroy(h))
    del xs, ys, zs, raw, crt, rawo

    # ---- config 3: relin MAC ----
    par3 = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par3), 44, 2, 16, 24, 24, 32767))
    h3 = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h3), C.byref(par3), 0, 0, 1))
    L3, K3, W3 = par3.numCrtPrime, par3.numEvalKey, lib.cuhe_param_words_coeff(C.byref(par3), 0)
    eks = torch.randint(0, 2**31 - 1, (K3, H, W3), dtype=torch.int32, device=dev)
    eks[:, n:, :] = 0
    check(lib.cuhe_relin_init(h3, p(eks), st()))
    del eks
    raw3 = torch.randint(0, 2**31 - 1, (H, W3), dtype=torch.int32, device=dev)
    raw3[n:] = 0
    out3 = torch.zeros((L3, N), dtype=torch.int64, device=dev)
    ms = timeit(lambda: check(lib.cuhe_relin(h3, p(out3), p(raw3), 0, st())), reps=10)
    rec("relin: 66 digit NTTs + MAC (L=44,K=66)", 8 * N * (L3 * K3 + K3 + L3) + K3 * 10 * N, ms)
    check(lib.cuhe_ctx_destroy(h3))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
