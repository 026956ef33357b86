#!/usr/bin/env python
"""One small batch of forward + inverse 64K transforms and one multiply through the C ABI, checked against the
oracle: the workload for `compute-sanitizer --tool memcheck|racecheck` captures (profiles/*_sanitizer_*.log)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cuhe_b200._lib import check, cuhe_params, load_library  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cnt = 3
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), 24, 2, 16, 24, 24, 32767))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rng = np.random.default_rng(5)
    x = rng.integers(0, 1 << 31, size=(cnt, N // 2), dtype=np.uint32)
    src = torch.from_numpy(x.view(np.int32)).to(dev)
    dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
    back = torch.zeros((cnt, N), dtype=torch.int64, device=dev)
    check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src), N, cnt, C.c_longlong(N // 2), st))
    check(lib.cuhe_intt_batch(h, p(back), p(dst), N, cnt, st))
    torch.cuda.synchronize()
    got = back.cpu().numpy().view(np.uint64)
    assert np.array_equal(got[:, : N // 2], x.astype(np.uint64)) and not got[:, N // 2:].any(), "inverse(forward(x)) != x"
    from oracle import oracle as orc
    assert np.array_equal(dst.cpu().numpy().view(np.uint64), orc.ntt_ext(x, N)), "forward transform differs from the oracle"
    print("one_ntt ok: N =", N, "launches:", lib.cuhe_launch_count(0))
    lib.cuhe_ctx_destroy(h)


if __name__ == "__main__":
    main()
