#!/usr/bin/env python
"""bench.py -- homomorphic ctxt x ctxt multiply throughput on B200 (BASELINE.json
configs[1]: ring degree 2^15 -> NTT length 65536, 24 CRT primes), with the
64K-point NTT roofline, the end-to-end host-buffer number and the CPU baseline.

  python bench.py --gpus 1 --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # residues sharded over N GPUs

A "step" = one batch of B = batch x N_gpus independent ciphertext products, RAW operands
resident in HBM -> RAW product in HBM (crt x2, forward NTT x2L, fused pointwise
mul + inverse NTT, polynomial Barrett (2 more forward + 2 more inverse NTTs per
residue), ICRT).  With N > 1 the CRT-residue axis is sharded (rank r owns primes
r, r+N, ...), ICRT is split by coefficient range: an NCCL all-to-all hands rank j the
coefficient slice j of every residue, and the RAW slices are all-gathered.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = (24, 2, 16, 24, 24, 32767)      # setParameters(d,p,w,min,cut,m): N=65536, L=24, W=18
WORKLOAD_NAME = "ctxt x ctxt multiply, ring degree 2^15 (n=27000, nttLen=65536), 24 CRT primes (576-bit q)"
NTT_BYTES_64K = 655360                      # u32[32768] in + u64[65536] out


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def ntt_issue(batch, launch_ms):
    """Secondary roofline of the NTT kernels, which are integer-ALU bound: warp instructions per launch
    pair (smsp__inst_executed.sum of the committed ncu capture) / measured launch time, against the issue
    peak 148 SMs x 4 schedulers x SM clock; plus the ALU-pipe utilisation ncu reported for the capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ntt_traffic.json")))
        if t["batch"] != batch:
            return None
        rate = t["warp_instructions_per_launch_pair"] / (launch_ms * 1e-3)
        peak = 148 * 4 * 1.965e9
        return {"bound": "int32 alu pipe (16 lanes per scheduler: IADD3/LOP3/SHF issue every other cycle)",
                "warp_inst_per_s": rate, "issue_peak_warp_inst_per_s": peak, "issue_frac": rate / peak,
                "alu_pipe_pct_of_peak_ncu": t["alu_pipe_pct_of_peak"], "source": "profiles/r01_v2b_ntt_full.txt"}
    except Exception:
        return None


def ntt_traffic(batch):
    """dram__bytes_read+write of the two NTT pass kernels for one launch pair at this batch, from the
    committed ncu --set full capture (profiles/ntt_traffic.json); None if it was taken at another batch."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ntt_traffic.json")))
        return t["dram_bytes_per_launch_pair"] if t["batch"] == batch else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def gen_raw(o, batch, nbuf, seed):
    """nbuf x batch operand pairs, coefficients uniform in [0, q0): u32[nbuf][batch][H][W]"""
    rng = random.Random(seed)
    q0, W, H, n = o["q0"], o["W"], o["H"], o["n"]
    nb = W * 4

    def poly():
        body = b"".join(rng.randrange(q0).to_bytes(nb, "little") for _ in range(n))
        buf = np.zeros(H * W, dtype=np.uint32)
        buf[: n * W] = np.frombuffer(body, dtype="<u4")
        return buf.reshape(H, W)
    # a few distinct polynomials, tiled (content does not change the work)
    base = [poly() for _ in range(4)]
    a = np.stack([np.stack([base[(i + j) % 4] for j in range(batch)]) for i in range(nbuf)])
    b = np.stack([np.stack([base[(i + j + 1) % 4] for j in range(batch)]) for i in range(nbuf)])
    return a, b


# --------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference's own CPU path is NTL, absent here)
# --------------------------------------------------------------------------------
CPU_BATCH = 8      # products in flight per CPU step: 8 x 24 (polynomial, residue) tasks keep every host thread busy


def _host_threads() -> int:
    """One thread per physical core this process may run on (measured on the 64-core / 128-thread B200
    host: 12.3 mul/s with 64 threads, 3.4 with 128).  Set explicitly because torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would make the N>1 reference arm single-threaded."""
    if os.environ.get("CUHE_B200_CPU_THREADS"):
        return int(os.environ["CUHE_B200_CPU_THREADS"])
    cpus = sorted(os.sched_getaffinity(0))
    cores = set()
    try:
        for c in cpus:
            base = f"/sys/devices/system/cpu/cpu{c}/topology/"
            cores.add((open(base + "physical_package_id").read().strip(), open(base + "core_id").read().strip()))
        return max(1, len(cores))
    except OSError:
        return max(1, len(cpus))


def _cpu_setup():
    """Two CPU restatements of the same multiply, both in oracle/ (test infrastructure):
      * zzx  -- the reference's own host path, `t = a*b; t %= polyMod; coeffReduce` (examples/DHS/DHS.cu:219-221),
                on the GMP runtime NTL is built on (oracle/zzx_gmp.c): one polynomial product per thread;
      * ntt  -- the C port of the GPU pipeline (oracle/coracle.c), OpenMP over (polynomial, residue) pairs.
    The faster one (zzx, ~10x) is what `cpu_baseline` and `--impl reference` report; NTL itself is absent."""
    from oracle.oracle import Oracle, lib
    threads = _host_threads()
    lib().orc_set_threads(C.c_int(threads))
    o = Oracle(*WORKLOAD)
    o.barrett_tables()
    o.inverse_series()
    rng = random.Random(1)
    q0 = o.moduli[0]
    polys = [o.to_raw([rng.randrange(q0) for _ in range(o.n)], 0) for _ in range(3)]

    def operands(batch):
        return (np.stack([polys[i % 3] for i in range(batch)]), np.stack([polys[(i + 1) % 3] for i in range(batch)]))
    return o, operands, lib().orc_max_threads()


def _rate(fn, batch, min_seconds, max_steps):
    fn()                                            # warm-up (page in, build tables)
    done, t0 = 0, time.perf_counter()
    while True:
        fn()
        done += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or done >= max_steps:
            break
    return batch * done / el, batch * done, el


def cpu_mul_rate(min_seconds: float, max_steps: int = 40):
    """Both CPU arms on the box's host cores; returns the cpu_baseline object of the JSON line."""
    o, operands, cores = _cpu_setup()
    za, zb = operands(cores)                        # one product per thread
    v, done, el = _rate(lambda: o.mul_raw_batch_zzx(za, zb, 0), cores, min_seconds, max_steps)
    na, nb = operands(CPU_BATCH)
    v2, done2, el2 = _rate(lambda: o.mul_raw_batch(na, nb, 0), CPU_BATCH, min_seconds / 3, max_steps)
    return {"value": v, "unit": "mul/s", "cores": cores, "kind": "port",
            "sample": f"{done} multiplications of the same workload in {el:.1f} s: the reference's NTL host path "
                      "(t = a*b; t %= Phi_m; coefficients mod q) restated on GMP (Kronecker product + inverse-series "
                      "division), one product per thread; NTL itself is not installed",
            "ntt_port": {"value": v2, "unit": "mul/s",
                         "sample": f"{done2} multiplications in {el2:.1f} s: C port of the NTT pipeline, OpenMP over "
                                   f"(polynomial, residue) pairs, batches of {CPU_BATCH}"}}


def run_reference(args):
    """--impl reference: the reference's own CPU path is NTL's ZZX arithmetic (absent here and on the GPU
    box); this arm times its restatement on GMP (oracle/zzx_gmp.c) with one product per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o, operands, cores = _cpu_setup()
    a, b = operands(cores)
    for _ in range(max(1, min(args.warmup, 3))):
        o.mul_raw_batch_zzx(a, b, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.mul_raw_batch_zzx(a, b, 0)
    el = time.perf_counter() - t0
    val = cores * args.steps / el
    out = {
        "impl": "reference", "metric": "homomorphic ctxt x ctxt mul/s", "value": val, "unit": "mul/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "big integers (GMP), exact",
        "data": "synthetic", "config": {"workload": WORKLOAD_NAME, "batch": cores},
        "cpu_baseline": {"value": val, "unit": "mul/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of {cores} multiplications (one per thread): the reference's NTL host path "
                                   "(examples/DHS/DHS.cu:219-221) restated on GMP; NTL itself is not installed"},
        "e2e": {"value": val, "unit": "mul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cuhe_b200._lib import check, cuhe_params, load_library

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))   # fail fast on a mismatch
    lib = load_library()
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), *WORKLOAD))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), local, rank, world))
    # Phi_m from the host helper (the caller supplies polyMod, as DHS.cu does)
    from cuhe_b200.hostmath import cyclotomic
    phi = np.array(cyclotomic(WORKLOAD[5]), dtype=np.int64)
    check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    L, W, H, N, n = par.numCrtPrime, lib.cuhe_param_words_coeff(C.byref(par), 0), par.crtLen, par.nttLen, par.modLen
    rows = lib.cuhe_ctx_rows(h, 0)
    assert L % world == 0, "bench shards need numCrtPrime divisible by the GPU count"
    qw = np.zeros(W + 1, dtype=np.uint32)
    check(lib.cuhe_ctx_coeff_modulus_host(h, 0, qw.ctypes.data_as(C.c_void_p), W + 1))
    info = dict(q0=int.from_bytes(qw.tobytes(), "little"), W=W, H=H, n=n)
    # per-GPU work is held constant: every step multiplies batch*world ciphertext pairs, the residue
    # axis of all of them sharded over the ranks (weak scaling; the all-gather grows with the batch)
    B = args.batch * world
    NBUF = 4 if B <= 64 else 2      # rotating operand sets (each is B x 4.7 MB on the host and on the device)
    a_np, b_np = gen_raw(info, B, NBUF, 20260924)
    a_dev = torch.from_numpy(a_np.view(np.int32)).to(dev)
    b_dev = torch.from_numpy(b_np.view(np.int32)).to(dev)
    from cuhe_b200 import sharded as sh
    crt_loc = torch.zeros((B, rows, H), dtype=torch.int32, device=dev)
    raw_out = torch.zeros((B, H, W), dtype=torch.int32, device=dev)
    cb, ce = sh.coefficient_slice(H, rank, world)
    raw_slice = torch.zeros((B, ce - cb, W), dtype=torch.int32, device=dev)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    # CUHE_B200_OVERLAP=1 (multi-GPU, not measured yet, off by default): the batch is processed as two halves
    # on two streams, so that the collectives of one half run under the kernels of the other
    overlap = world > 1 and os.environ.get("CUHE_B200_OVERLAP") == "1" and B >= 2
    if overlap:
        halves = [(0, B // 2), (B // 2, B)]
        side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def step_overlapped(k):
        cur = torch.cuda.current_stream()
        outs = []
        for s_, (lo_, up_) in zip(side, halves):
            s_.wait_stream(cur)
            with torch.cuda.stream(s_):
                nb = up_ - lo_
                check(lib.cuhe_mul_crt_batch(h, p(crt_loc[lo_:up_]), p(a_dev[k][lo_:up_]), p(b_dev[k][lo_:up_]), 0, nb, st()))
                crt_slice = sh.exchange_for_icrt(crt_loc[lo_:up_], L, rank, world)
                check(lib.cuhe_icrt_slice_batch(h, p(raw_slice[lo_:up_]), p(crt_slice), 0, cb, ce - cb, nb, st()))
                outs.append(sh.all_gather_raw_slices(raw_slice[lo_:up_], world))
        for s_ in side:
            cur.wait_stream(s_)
        return outs

    def step(i):
        k = i % NBUF
        if overlap:
            return step_overlapped(k)
        check(lib.cuhe_mul_crt_batch(h, p(crt_loc), p(a_dev[k]), p(b_dev[k]), 0, B, st()))
        if world == 1:
            check(lib.cuhe_icrt_batch(h, p(raw_out), p(crt_loc), 0, 0, H, B, st()))
        else:
            # NCCL all-to-all over NVLink: rank j receives coefficient slice j of every residue
            crt_slice = sh.exchange_for_icrt(crt_loc, L, rank, world)
            check(lib.cuhe_icrt_slice_batch(h, p(raw_slice), p(crt_slice), 0, cb, ce - cb, B, st()))
            return sh.all_gather_raw_slices(raw_slice, world)            # complete RAW on every rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    if rank == 0:
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:    # nvidia-smi needs ~1 s to print its first row
            time.sleep(0.05)
    barrier()
    for i in range(args.warmup):                                   # every rank: the steps contain collectives
        step(i)
    barrier()
    first_sample = len(sampler.rows)
    lib.cuhe_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.cuhe_launch_count(0))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = B * args.steps / (ms * 1e-3)
    clocks = None
    # keep every GPU under the same load for ~0.4 s more so the 50 ms clock samples cover it; the
    # step count is derived from the all-reduced time, hence identical on every rank (collectives inside)
    n_extra = min(2000, max(4, int(400.0 / max(ms / args.steps, 1e-3))))
    for i in range(n_extra):
        step(i)
    barrier()
    if rank == 0:
        sampler.rows = sampler.rows[first_sample:]
        clocks = sampler.stop()

    # ---- roofline: the dominant kernels are the NTT passes; time one batched forward
    #      64K ext-NTT launch pair (pass 1 + pass 2) alone, inputs larger than L2 ----
    roof = None
    ntt_rate = None
    if rank == 0:
        cnt = 512
        src = torch.randint(0, 2**31 - 1, (2, cnt, H), dtype=torch.int32, device=dev)     # 2 x 67 MB
        dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)                         # 268 MB
        for i in range(3):
            check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src[i % 2]), N, cnt, C.c_longlong(H), st()))
        torch.cuda.synchronize()
        reps = 10
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(reps):
            check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src[i % 2]), N, cnt, C.c_longlong(H), st()))
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        pk, pk_src = peaks()
        ach = NTT_BYTES_64K * cnt / (kms * 1e-3) / 1e9
        ntt_rate = cnt / (kms * 1e-3)
        roof = {"kernel": "ntt_pass1_v2_kernel<EXT_U32> + ntt_pass2_v2_kernel<16,U64> (one batched forward 64K NTT)",
                "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "peak_source": pk_src + " (burst copy bandwidth)", "traffic": ntt_traffic(cnt),
                "algorithmic_bytes_per_launch": NTT_BYTES_64K * cnt, "launch_ms": kms, "batch": cnt,
                "secondary": ntt_issue(cnt, kms),
                "note": "INT32-ALU bound (SURVEY F9): see DESIGN.md for the instruction-issue roofline"}
        del src, dst

    # ---- e2e: host buffers through the C ABI (the device part of mulZZX), H2D + D2H inside ----
    e2e = None
    if world > 1:
        # data-parallel host side: rank r owns products [r*b, (r+1)*b) of every step.  Per step and rank:
        # H2D of the owned operands from pinned memory -> NVLink all-gather of the RAW operands -> the
        # residue-sharded multiply (all-to-all, sliced ICRT) -> all-to-all of RAW slices back to the
        # owners -> D2H of the owned products.  Every polynomial crosses PCIe once per step in the whole
        # job.  Copies run on side streams, double buffered, so step i+1's upload and step i-1's download
        # overlap step i's kernels and collectives.
        b = args.batch
        lo = rank * b
        ah = [torch.from_numpy(a_np[k][lo:lo + b].view(np.int32)).pin_memory() for k in range(2)]
        bh = [torch.from_numpy(b_np[k][lo:lo + b].view(np.int32)).pin_memory() for k in range(2)]
        oh = [torch.zeros((b, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
        a_loc = [torch.zeros((b, H, W), dtype=torch.int32, device=dev) for _ in range(2)]
        b_loc = [torch.zeros((b, H, W), dtype=torch.int32, device=dev) for _ in range(2)]
        o_loc = [None, None]
        s_in, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]          # operands of slot k are on the device
        ev_used = [torch.cuda.Event() for _ in range(2)]        # slot k's operands have been gathered
        ev_done = [torch.cuda.Event() for _ in range(2)]        # slot k's owned products are in o_loc[k]
        ev_read = [torch.cuda.Event() for _ in range(2)]        # slot k's products have left the device

        def upload(i):
            k = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_used[k])
                a_loc[k].copy_(ah[k], non_blocking=True)
                b_loc[k].copy_(bh[k], non_blocking=True)
                ev_in[k].record(s_in)

        def compute(i):
            k = i % 2
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in[k])
                a_all = sh.all_gather_operands(a_loc[k], world)
                b_all = sh.all_gather_operands(b_loc[k], world)
                ev_used[k].record(s_comp)
                check(lib.cuhe_mul_crt_batch(h, p(crt_loc), p(a_all), p(b_all), 0, B, st()))
                crt_slice = sh.exchange_for_icrt(crt_loc, L, rank, world)
                check(lib.cuhe_icrt_slice_batch(h, p(raw_slice), p(crt_slice), 0, cb, ce - cb, B, st()))
                if i >= 2:
                    s_comp.wait_event(ev_read[k])
                o_loc[k] = sh.raw_slices_to_owners(raw_slice, world)
                ev_done[k].record(s_comp)

        def download(i):
            k = i % 2
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                oh[k].copy_(o_loc[k], non_blocking=True)
                ev_read[k].record(s_out)

        def e2e_run(nsteps):
            upload(0)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    upload(i + 1)
                compute(i)
                download(i)
            for s_ in (s_in, s_comp, s_out):
                s_.synchronize()

        torch.cuda.synchronize()
        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * args.steps / float(tt.item()), "unit": "mul/s", "h2d_bytes_per_step": int(2 * B * H * W * 4),
               "d2h_bytes_per_step": int(B * H * W * 4),
               "api": "each rank: pinned host RAW of its products -> H2D -> NVLink all-gather -> sharded cuhe_mul_crt_batch / "
                      "all-to-all / cuhe_icrt_slice_batch / all-to-all to owners -> D2H (double-buffered side streams); "
                      "byte counts are whole-job totals"}
    if world == 1:
        # one e2e step = one call with Be = 4*B products from pinned host memory; the library pipelines
        # H2D | kernels | D2H over chunks of 8 products inside the call
        Be = 4 * B
        ah = torch.from_numpy(np.concatenate([a_np[i % NBUF] for i in range(4)]).view(np.int32)).pin_memory()
        bh = torch.from_numpy(np.concatenate([b_np[i % NBUF] for i in range(4)]).view(np.int32)).pin_memory()
        oh = torch.zeros((Be, H, W), dtype=torch.int32).pin_memory()
        for _ in range(2):
            check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st()))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st()))
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        e2e = {"value": Be * args.steps / el, "unit": "mul/s", "h2d_bytes_per_step": int(2 * Be * H * W * 4),
               "d2h_bytes_per_step": int(Be * H * W * 4), "batch": Be,
               "api": "cuhe_mul_raw_host_batch (pinned host RAW in/out, 3-stream pipeline inside the call)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_mul_rate(12.0)

    def child_json(script, *cli):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), *cli], capture_output=True, text=True, timeout=180)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            return json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-300:]}
        except Exception as ex:                                  # noqa: BLE001
            return {"error": repr(ex)[:300]}

    sweep = relin3 = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # BASELINE configs[2]: key switch at N = 65536, 44 primes, 66 keys (1.52 GB resident), tools/relin_bench.py
        relin3 = child_json("relin_bench.py")
        # the table the reference publishes (doc/Perf_NTT.txt): per-transform time for N x batch, and the
        # configs[0] latency.  Isolated in a child process so that nothing it does can disturb this line.
        sweep = child_json("ntt_bench.py", "--sweep")

    if rank == 0:
        out = {
            "metric": "homomorphic ctxt x ctxt mul/s", "value": value, "unit": "mul/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1) / u32 residues",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "batch": B, "batch_per_gpu": args.batch, "parallelism": f"residue-shard x{world}" if world > 1 else "single GPU",
                       "l2": f"{NBUF} rotating operand sets; per-step NTT intermediates {2 * B * L * N * 8 / 1e6:.0f} MB exceed the 126 MB L2"},
            "ntt_64k_per_s": ntt_rate, "ntt_sweep": sweep, "relin_config3": relin3, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(out))
    lib.cuhe_ctx_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="ciphertext pairs per step and per GPU (32: 9.0 k mul/s, 8: 7.9 k)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40          # bounded sample: each step is 8 CPU multiplications
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
