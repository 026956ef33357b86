#!/usr/bin/env python
"""Micro-benchmark of the batched transforms through the C ABI (CUDA events on
the launching stream; inputs + outputs larger than L2).  Used for A/B runs of
kernel variants: CUHE_B200_LIB=<path to .so> python tools/ntt_bench.py

  --sweep   the table the reference publishes (doc/Perf_NTT.txt, from tests/test_ntt.cu:67-100): time per
            forward zero-padded transform for N in {16384, 32768, 65536} x batch in {1, 8, 64, 512}, plus the
            inverse and BASELINE configs[0] (one forward + inverse pair at N = 16384: latency)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuhe_b200._lib import LIB_PATH, check, cuhe_params, load_library  # noqa: E402


def sweep():
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    dev = torch.device("cuda", 0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    flush = torch.zeros(160 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.add_(1)                                                    # evict the previous iteration from L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    rows = []
    for N in (16384, 32768, 65536):
        H = N // 2
        for cnt in (1, 8, 64, 512):
            src = torch.randint(0, 2**31 - 1, (cnt, H), dtype=torch.int32, device=dev)
            dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
            back = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
            ms_f = timed(lambda: check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src), N, cnt, C.c_longlong(H), st())), 10)
            ms_i = timed(lambda: check(lib.cuhe_intt_batch(h, p(back), p(dst), N, cnt, st())), 10)
            rows.append({"N": N, "batch": cnt, "fwd_ms_per_transform": ms_f / cnt, "inv_ms_per_transform": ms_i / cnt,
                         "fwd_per_s": cnt / ms_f * 1e3})
    pair = next(r for r in rows if r["N"] == 16384 and r["batch"] == 1)
    print(json.dumps({"l2": "flushed between timed launches", "rows": rows,
                      "config0_fwd_plus_inv_latency_ms": pair["fwd_ms_per_transform"] + pair["inv_ms_per_transform"],
                      "reference_doc_Perf_NTT_ms_per_transform_batch512": {"16384": 0.00408, "32768": 0.00805, "65536": 0.02266}}))
    lib.cuhe_ctx_destroy(h)


def one():
    """a few batched forward 64K launch pairs (batch 512, alternating inputs > L2): the workload of bench.py's ncu child"""
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N, H, cnt = 65536, 32768, 512
    src = torch.randint(0, 2**31 - 1, (2, cnt, H), dtype=torch.int32, device=dev)
    dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
    for i in range(8):
        check(lib.cuhe_ntt_ext_batch(h, C.c_void_p(dst.data_ptr()), C.c_void_p(src[i % 2].data_ptr()), N, cnt, C.c_longlong(H), st))
    torch.cuda.synchronize()
    lib.cuhe_ctx_destroy(h)


def main():
    if "--sweep" in sys.argv:
        return sweep()
    if "--one" in sys.argv:
        return one()
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    dev = torch.device("cuda", 0)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    out = {"lib": os.path.basename(os.path.dirname(LIB_PATH)) + "/" + os.path.basename(LIB_PATH)}
    for N in (65536, 16384):
        H = N // 2
        cnt = 512 if N == 65536 else 2048
        src = torch.randint(0, 2**31 - 1, (2, cnt, H), dtype=torch.int32, device=dev)
        dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
        back = torch.zeros((cnt, N), dtype=torch.int64, device=dev)

        def timeit(fn, reps=10):
            for _ in range(3):
                fn(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_f = timeit(lambda i: check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src[i % 2]), N, cnt, C.c_longlong(H), st())))
        ms_i = timeit(lambda i: check(lib.cuhe_intt_batch(h, p(back), p(dst), N, cnt, st())))
        out[f"fwd_{N}"] = {"ntt_per_s": cnt / ms_f * 1e3, "GBps": 10 * N * cnt / ms_f / 1e6, "ms": ms_f}
        out[f"inv_{N}"] = {"ntt_per_s": cnt / ms_i * 1e3, "GBps": 16 * N * cnt / ms_i / 1e6, "ms": ms_i}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
