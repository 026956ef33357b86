#!/usr/bin/env python
"""The reference's own NTT micro-benchmark (tests/test_ntt.cu:67-154 -> doc/Perf_NTT.txt) run on THIS GPU with
the reference's own kernels (oracle/_ref/libref_base.so = cuhe/Base.cu compiled for sm_100a), beside the
shipped engine on the same protocol: 1024 zero-padded forward transforms per length, issued in bundles of
num = 1, 2, ..., 512 (reference: gridDim.y = num, three launches per bundle; ours: cuhe_ntt_ext_batch with
count = num), back to back on one stream, CUDA events around the whole loop, milliseconds per transform.
Outputs are compared bit for bit before timing.  Test infrastructure / measurement only."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuhe_b200._lib import check, cuhe_params, load_library  # noqa: E402

PUBLISHED = {  # doc/Perf_NTT.txt (hardware not stated), ms per transform
    16384: [0.0486284, 0.0242168, 0.0128587, 0.00765705, 0.00774383, 0.00576811, 0.00490982, 0.00444013, 0.00419698, 0.00407564],
    32768: [0.051598, 0.0258971, 0.0150039, 0.0130533, 0.0100444, 0.00896879, 0.00848012, 0.00812238, 0.00804524, 0.00804859],
    65536: [0.064822, 0.0403285, 0.0354673, 0.0289758, 0.0260014, 0.0243423, 0.0234518, 0.0230109, 0.0227886, 0.0226647],
}


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_base.so"))
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    cnt = 1024
    rows = []
    for N in (16384, 32768, 65536):
        assert ref.ref_base_preload_ntt(N) == 0
        src = torch.randint(0, 2**31 - 1, (cnt, N), dtype=torch.int32, device=dev)     # rand(): 31-bit values
        dst_r = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
        tmp = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
        dst_o = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
        ms = C.c_float()
        assert ref.ref_base_time_ntt(512, N, cnt, p(dst_r), p(tmp), p(src), C.byref(ms), st) == 0
        check(lib.cuhe_ntt_ext_batch(h, p(dst_o), p(src), N, cnt, C.c_longlong(N), st))
        torch.cuda.synchronize()
        assert torch.equal(dst_r, dst_o), "shipped transform differs from the reference kernels"
        for i in range(10):
            num = 1 << i
            best_r, best_o = 1e9, 1e9
            for _ in range(3):
                assert ref.ref_base_time_ntt(num, N, cnt, p(dst_r), p(tmp), p(src), C.byref(ms), st) == 0
                best_r = min(best_r, ms.value)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                d0, s0, fn, stride = dst_o.data_ptr(), src.data_ptr(), lib.cuhe_ntt_ext_batch, C.c_longlong(N)
                for j in range(cnt // num):       # raw addresses: keep interpreter work out of the timed loop
                    fn(h, C.c_void_p(d0 + j * num * N * 8), C.c_void_p(s0 + j * num * N * 4), N, num, stride, st)
                e1.record()
                torch.cuda.synchronize()
                best_o = min(best_o, e0.elapsed_time(e1) / cnt)
            rows.append({"N": N, "num": num, "reference_kernels_ms": best_r, "cuhe_b200_ms": best_o,
                         "speedup": best_r / best_o, "doc_Perf_NTT_ms": PUBLISHED[N][i]})
        del src, dst_r, tmp, dst_o
    print(json.dumps({"protocol": "tests/test_ntt.cu:67-100, 1024 transforms per length, best of 3", "gpu": torch.cuda.get_device_name(0),
                      "rows": rows}))
    lib.cuhe_ctx_destroy(h)


if __name__ == "__main__":
    main()
