"""tools/e2e_sweep.py -- A/B the host pipeline schedule of cuhe_mul_raw_host_batch (chunk size and
ramp are read from the environment once per process, so every setting runs in its own process).
  python tools/e2e_sweep.py            # sweep
  python tools/e2e_sweep.py --one 128  # one measurement in this process: products per call"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(Be, calls=8):
    import numpy as np
    import torch
    import bench
    from cuhe_b200._lib import check, cuhe_params, load_library
    from cuhe_b200.hostmath import cyclotomic
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), *bench.WORKLOAD))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), 0, 0, 1))
    phi = np.array(cyclotomic(bench.WORKLOAD[5]), dtype=np.int64)
    check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    W, H, n = lib.cuhe_param_words_coeff(C.byref(par), 0), par.crtLen, par.modLen
    qw = np.zeros(W + 1, dtype=np.uint32)
    check(lib.cuhe_ctx_coeff_modulus_host(h, 0, qw.ctypes.data_as(C.c_void_p), W + 1))
    info = dict(q0=int.from_bytes(qw.tobytes(), "little"), W=W, H=H, n=n)
    a_np, b_np = bench.gen_raw(info, Be, 1, 1)
    ah = torch.from_numpy(a_np[0].view(np.int32)).pin_memory()
    bh = torch.from_numpy(b_np[0].view(np.int32)).pin_memory()
    oh = torch.zeros((Be, H, W), dtype=torch.int32).pin_memory()
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(calls):
        check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st))
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(json.dumps({"products_per_call": Be, "chunk": os.environ.get("CUHE_B200_HOST_CHUNK", "default"),
                      "ramp": os.environ.get("CUHE_B200_HOST_RAMP", "default"), "mul_per_s": Be * calls / el,
                      "ms_per_call": 1e3 * el / calls}))
    lib.cuhe_ctx_destroy(h)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one(int(sys.argv[sys.argv.index("--one") + 1]))
    else:
        for Be in (128, 32):
            for chunk, ramp in (("8", "0"), ("8", "1"), ("16", "0"), ("16", "1"), ("32", "0"), ("32", "1"), (None, None)):
                env = dict(os.environ)
                if chunk:
                    env["CUHE_B200_HOST_CHUNK"], env["CUHE_B200_HOST_RAMP"] = chunk, ramp
                subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(Be)], env=env, check=False)
