#!/usr/bin/env python
"""Residue-sharded multiply (cuhe_mul_raw_sharded_batch: NCCL exchange inside the library) checked ON THE GPUS against the
unsharded path of the same library and, for a small case, the oracle.  Run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/sharded_check.py

Every rank multiplies its own (rank-seeded) ciphertext pairs through the sharded context and through a private unsharded
context on its own GPU; the RAW results must be identical words.  Levels 0 and 1 (at level 1 the prime count is not a
multiple of the rank count, so ranks own different numbers of residues)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cuhe_b200._lib import check, cuhe_params, load_library  # noqa: E402
from cuhe_b200.hostmath import cyclotomic  # noqa: E402


def make_ctx(lib, ps, dev, rank, world):
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), *ps))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), dev, rank, world))
    phi = np.array(cyclotomic(ps[5]), dtype=np.int64)
    check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    return h, par


def comm_init(lib, h, rank, dev):
    """rank 0 draws the NCCL id, torch.distributed carries the 128 bytes, every rank joins"""
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib.cuhe_comm_unique_id(idbuf))
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
    check(lib.cuhe_ctx_comm_init(h, idbuf))


def rand_raw(rng, par, lib, lvl, nb, q_words):
    """nb polynomials with n coefficients uniform below q_lvl, RAW layout u32[nb][H][W]"""
    W, H, n = q_words.size, par.crtLen, par.modLen
    q = int.from_bytes(q_words.tobytes(), "little")
    out = np.zeros((nb, H, W), dtype=np.uint32)
    for b in range(nb):
        vals = [int.from_bytes(rng.bytes(4 * W + 8), "little") % q for _ in range(n)]
        out[b, :n] = np.frombuffer(b"".join(v.to_bytes(4 * W, "little") for v in vals), dtype=np.uint32).reshape(n, W)
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    cases = [((3, 2, 16, 48, 24, 32767), 3), ((5, 2, 1, 61, 20, 8191), 5)]
    report = []
    for ps, nb in cases:
        hs, par = make_ctx(lib, ps, local, rank, world)
        hu, _ = make_ctx(lib, ps, local, 0, 1)
        if world > 1:
            comm_init(lib, hs, rank, dev)
        H = par.crtLen
        for lvl in (0, 1):
            W = lib.cuhe_param_words_coeff(C.byref(par), lvl)
            L = lib.cuhe_param_num_crt_prime(C.byref(par), lvl)
            qw = np.zeros(W, dtype=np.uint32)
            check(lib.cuhe_ctx_coeff_modulus_host(hu, lvl, qw.ctypes.data_as(C.c_void_p), W))
            rng = np.random.default_rng(1000 * rank + 10 * lvl + len(report))
            a = torch.from_numpy(rand_raw(rng, par, lib, lvl, nb, qw).view(np.int32)).to(dev)
            b = torch.from_numpy(rand_raw(rng, par, lib, lvl, nb, qw).view(np.int32)).to(dev)
            out_s = torch.zeros((nb, H, W), dtype=torch.int32, device=dev)
            out_u = torch.zeros((nb, H, W), dtype=torch.int32, device=dev)
            crt_u = torch.zeros((nb, L, H), dtype=torch.int32, device=dev)
            check(lib.cuhe_mul_raw_sharded_batch(hs, p(out_s), p(a), p(b), lvl, nb, st))
            check(lib.cuhe_mul_crt_batch(hu, p(crt_u), p(a), p(b), lvl, nb, st))
            check(lib.cuhe_icrt_batch(hu, p(out_u), p(crt_u), lvl, 0, H, nb, st))
            torch.cuda.synchronize()
            same = bool(torch.equal(out_s, out_u))
            report.append({"params": list(ps), "lvl": lvl, "L": L, "batch_own": nb, "equal": same})
            assert same, f"rank {rank}: sharded result differs from the unsharded path: {report[-1]}"
        lib.cuhe_ctx_destroy(hs)
        lib.cuhe_ctx_destroy(hu)
    ok = torch.tensor([1], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"sharded_check": "ok", "world": world, "cases": report}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
