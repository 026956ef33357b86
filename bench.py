#!/usr/bin/env python
"""bench.py -- homomorphic ctxt x ctxt multiply throughput on B200 (BASELINE.json
configs[1]: ring degree 2^15 -> NTT length 65536, 24 CRT primes), with the
64K-point NTT roofline, the end-to-end host-buffer number and the CPU baseline.

  python bench.py --gpus 1 --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # residues sharded over N GPUs

A "step" = every GPU multiplies its own `batch` ciphertext pairs, RAW operands resident in HBM -> RAW products in
HBM (crt x2, forward NTT x2L, fused pointwise mul + inverse NTT, reduction modulo Phi_m, ICRT).  With N > 1 the
CRT-residue axis of all batch x N products is sharded (rank r owns primes r, r+N, ...): after CRT an NCCL
send/recv exchange hands every rank the rows of its primes, before ICRT a second one returns the product rows to the
owners (cuhe_mul_raw_sharded_batch, exchange inside the library).  One step's result is checked against an
unsharded context on the same GPU outside the timed region ("verified").
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
WORKLOAD_C5 = (64, 2, 16, 24, 24, 32767)   # BASELINE configs[4]: N=65536, L=64 (8 per GPU at 8 GPUs), W=48
WORKLOAD_C5_NAME = "batched ctxt x ctxt multiply, n=27000 (nttLen=65536), 64 CRT primes (1536-bit q), residues sharded over the GPUs"


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
def measure_ntt_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the two NTT pass kernels for one batched forward 64K launch pair
    (batch 512), measured NOW by an `ncu --metrics` child over tools/ntt_bench.py --one (caches left as the preceding
    launches leave them, clocks untouched).  Returns (bytes per launch pair, source string) or (None, why)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = os.path.join("/tmp", f"cuhe_b200_ncu_{os.getpid()}.csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--cache-control", "none",
           "-k", "regex:ntt4_pass", "-s", "8", "-c", "4", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "tools", "ntt_bench.py"), "--one"]
    try:
        subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        import csv
        tot, per = 0.0, {}
        with open(log) as f:
            rows = [r for r in csv.reader(f) if len(r) > 14 and r[0].isdigit()]
        for r in rows:
            val = float(r[14].replace(",", ""))
            unit = r[13].lower()
            val *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            per[r[0]] = per.get(r[0], 0.0) + val
            tot += val
        os.remove(log)
        if len(per) != 4:
            return None, f"ncu captured {len(per)} launches"
        return int(tot / 2), "measured in this run: ncu --metrics dram__bytes_{read,write}.sum --cache-control none over 2 launch pairs"
    except Exception as ex:                                      # noqa: BLE001
        return None, repr(ex)[:200]


def make_context(lib, check, cuhe_params, workload, local, rank, world):
    from cuhe_b200.hostmath import cyclotomic
    par = cuhe_params()
    check(lib.cuhe_set_parameters(C.byref(par), *workload))
    h = C.c_void_p()
    check(lib.cuhe_ctx_create(C.byref(h), C.byref(par), local, rank, world))
    phi = np.array(cyclotomic(workload[5]), dtype=np.int64)          # the caller supplies polyMod, as DHS.cu does
    check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    return h, par


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
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def join_comm(h):
        """rank 0 draws the NCCL id, torch.distributed carries the 128 bytes, every rank joins (cuhe_ctx_comm_init)"""
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            check(lib.cuhe_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        check(lib.cuhe_ctx_comm_init(h, (C.c_ubyte * 128)(*t.cpu().tolist())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def setup(workload, batch):
        """context (residue shard `rank` of `world`), the own operands of this rank and the step function.
        One step = every rank multiplies its own `batch` ciphertext pairs, RAW in HBM -> RAW in HBM; with world > 1
        the residue axis of all batch*world products is spread over the ranks inside cuhe_mul_raw_sharded_batch."""
        h, par = make_context(lib, check, cuhe_params, workload, local, rank, world)
        if world > 1:
            join_comm(h)
        L, W, H, N, n = par.numCrtPrime, lib.cuhe_param_words_coeff(C.byref(par), 0), par.crtLen, par.nttLen, par.modLen
        qw = np.zeros(W + 1, dtype=np.uint32)
        check(lib.cuhe_ctx_coeff_modulus_host(h, 0, qw.ctypes.data_as(C.c_void_p), W + 1))
        info = dict(q0=int.from_bytes(qw.tobytes(), "little"), W=W, H=H, n=n)
        nbuf = 4 if batch <= 64 else 2      # rotating operand sets
        a_np, b_np = gen_raw(info, batch, nbuf, 20260924 + rank)
        a_dev = torch.from_numpy(a_np.view(np.int32)).to(dev)
        b_dev = torch.from_numpy(b_np.view(np.int32)).to(dev)
        raw_out = torch.zeros((batch, H, W), dtype=torch.int32, device=dev)
        crt_loc = torch.zeros((batch, L, H), dtype=torch.int32, device=dev) if world == 1 else None

        def step(i):
            k = i % nbuf
            if world == 1:
                check(lib.cuhe_mul_crt_batch(h, p(crt_loc), p(a_dev[k]), p(b_dev[k]), 0, batch, st()))
                check(lib.cuhe_icrt_batch(h, p(raw_out), p(crt_loc), 0, 0, H, batch, st()))
            else:
                check(lib.cuhe_mul_raw_sharded_batch(h, p(raw_out), p(a_dev[k]), p(b_dev[k]), 0, batch, st()))
        return dict(h=h, par=par, L=L, W=W, H=H, N=N, n=n, nbuf=nbuf, a_np=a_np, b_np=b_np, a_dev=a_dev, b_dev=b_dev,
                    raw_out=raw_out, step=step, workload=workload, batch=batch)

    def verify(S):
        """outside the timed region: the sharded result of this rank's products against a private unsharded context
        on the same GPU, word for word; AND over the ranks"""
        if world == 1:
            return None
        hu, _ = make_context(lib, check, cuhe_params, S["workload"], local, 0, 1)
        nb = min(S["batch"], 4)
        crt_u = torch.zeros((nb, S["L"], S["H"]), dtype=torch.int32, device=dev)
        out_u = torch.zeros((nb, S["H"], S["W"]), dtype=torch.int32, device=dev)
        S["step"](0)
        check(lib.cuhe_mul_crt_batch(hu, p(crt_u), p(S["a_dev"][0]), p(S["b_dev"][0]), 0, nb, st()))
        check(lib.cuhe_icrt_batch(hu, p(out_u), p(crt_u), 0, 0, S["H"], nb, st()))
        torch.cuda.synchronize()
        ok = torch.tensor([1 if torch.equal(out_u, S["raw_out"][:nb]) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        lib.cuhe_ctx_destroy(hu)
        del crt_u, out_u
        return bool(ok.item())

    def timed(S, steps, warmup):
        for i in range(warmup):
            S["step"](i)
        barrier()
        lib.cuhe_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            S["step"](i)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, int(lib.cuhe_launch_count(0))

    S = setup(WORKLOAD, args.batch)
    h, L, W, H, N = S["h"], S["L"], S["W"], S["H"], S["N"]
    B = args.batch * world
    verified = verify(S)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        S["step"](i)
    barrier()
    if rank == 0:
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:    # nvidia-smi needs ~1 s to print its first row
            time.sleep(0.05)
    barrier()
    first_sample = len(sampler.rows)
    ms, launches = timed(S, args.steps, args.warmup)
    value = B * args.steps / (ms * 1e-3)
    clocks = None
    # keep every GPU under the same load for ~0.4 s more so the 50 ms clock samples cover it; the
    # step count is derived from the all-reduced time, hence identical on every rank (collectives inside)
    n_extra = min(2000, max(4, int(400.0 / max(ms / args.steps, 1e-3))))
    for i in range(n_extra):
        S["step"](i)
    barrier()
    if rank == 0:
        sampler.rows = sampler.rows[first_sample:]
        clocks = sampler.stop()

    # ---- roofline: the dominant kernels are the NTT passes; every rank times one batched forward 64K ext-NTT launch
    #      pair (pass 1 + pass 2) alone, at the same time, inputs larger than L2; the rate is summed over the ranks ----
    cnt = 512
    src = torch.randint(0, 2**31 - 1, (2, cnt, H), dtype=torch.int32, device=dev)     # 2 x 67 MB
    dst = torch.zeros((cnt, N), dtype=torch.int64, device=dev)                         # 268 MB
    for i in range(3):
        check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src[i % 2]), N, cnt, C.c_longlong(H), st()))
    barrier()
    reps = 10
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(reps):
        check(lib.cuhe_ntt_ext_batch(h, p(dst), p(src[i % 2]), N, cnt, C.c_longlong(H), st()))
    k1.record()
    torch.cuda.synchronize()
    kms = max_over_ranks(k0.elapsed_time(k1) / reps)
    del src, dst
    pk, pk_src = peaks()
    ach = NTT_BYTES_64K * cnt / (kms * 1e-3) / 1e9
    ntt_rate = world * cnt / (kms * 1e-3)
    roof = None
    if rank == 0:
        traffic, traffic_src = (None, "skipped (--no-cpu / multi-GPU run)")
        if world == 1 and not args.no_cpu:
            traffic, traffic_src = measure_ntt_traffic()
        if traffic is None:
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", "ntt_traffic.json")))
                traffic, traffic_src = t["dram_bytes_per_launch_pair"], "committed capture " + t["source"] + " (live measurement: " + traffic_src + ")"
            except Exception:
                pass
        roof = {"kernel": "ntt4_pass1_kernel<1024, EXT_U32> + ntt4_pass2_kernel<16, U64> (one batched forward 64K NTT, per GPU)",
                "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "peak_source": pk_src + " (burst copy bandwidth)", "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": NTT_BYTES_64K * cnt, "launch_ms": kms, "batch": cnt,
                "secondary": {"bound": "int32 alu pipe: 199.5 SASS instructions per point (ncu, profiles/r02_gen4_ntt_full.txt: 56.2 M + "
                                       "153.0 M warp instructions per 512 transforms), ~68 % of them on the alu pipe (64 lanes/clk/SM), "
                                       "alu pipe 81 % / 79 % busy in pass 1 / pass 2; measured issue ceiling with both pipes busy "
                                       "105 thread-instr/clk/SM (profiles/r02_pipe_issue_rates.txt)",
                              "thread_instr_per_point": 199.5, "alu_pipe_busy_pct_ncu": [81.4, 78.6],
                              "issue_ceiling_ntt_per_s_per_gpu": 148 * 1.965e9 * 105 / (199.5 * 65536),
                              "alu_ceiling_ntt_per_s_per_gpu": 148 * 1.965e9 * 64 / (0.68 * 199.5 * 65536)},
                "note": "ALU-pipe / issue bound, not HBM bound (SURVEY F9); DESIGN.md section 4.2"}

    # ---- e2e: host buffers, H2D + D2H inside the timed region ----
    e2e = None
    if world > 1:
        # every rank: pinned host RAW of its own products -> H2D -> cuhe_mul_raw_sharded_batch (exchange inside the
        # library) -> D2H of its own products.  Copies on side streams, double buffered, under the neighbouring steps.
        b = args.batch
        n_live = S["n"]
        cudart = C.CDLL("libcudart.so.12")
        cudart.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]

        def copy_live(dst, src, kind, stream):
            # [b][H][W] -> the first n_live rows of every polynomial, one 2-D copy (1 = H2D, 2 = D2H)
            rc = cudart.cudaMemcpy2DAsync(dst.data_ptr(), H * W * 4, src.data_ptr(), H * W * 4, n_live * W * 4, b, kind,
                                          C.c_void_p(stream.cuda_stream))
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy2DAsync failed: {rc}")
        a_np, b_np = S["a_np"], S["b_np"]
        ah = [torch.from_numpy(a_np[k].view(np.int32)).pin_memory() for k in range(2)]
        bh = [torch.from_numpy(b_np[k].view(np.int32)).pin_memory() for k in range(2)]
        oh = [torch.zeros((b, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
        a_loc = [torch.zeros((b, H, W), dtype=torch.int32, device=dev) for _ in range(2)]
        b_loc = [torch.zeros((b, H, W), dtype=torch.int32, device=dev) for _ in range(2)]
        o_loc = [torch.zeros((b, H, W), dtype=torch.int32, device=dev) for _ in range(2)]
        s_in, s_comp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]          # operands of slot k are on the device
        ev_used = [torch.cuda.Event() for _ in range(2)]        # slot k's operands have been consumed
        ev_done = [torch.cuda.Event() for _ in range(2)]        # slot k's products are in o_loc[k]
        ev_read = [torch.cuda.Event() for _ in range(2)]        # slot k's products have left the device

        def upload(i):
            k = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_used[k])
                copy_live(a_loc[k], ah[k], 1, s_in)       # rows modLen.. of a ring element are zero: not transferred
                copy_live(b_loc[k], bh[k], 1, s_in)
                ev_in[k].record(s_in)

        def compute(i):
            k = i % 2
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in[k])
                if i >= 2:
                    s_comp.wait_event(ev_read[k])
                check(lib.cuhe_mul_raw_sharded_batch(h, p(o_loc[k]), p(a_loc[k]), p(b_loc[k]), 0, b, st()))
                ev_used[k].record(s_comp)
                ev_done[k].record(s_comp)

        def download(i):
            k = i % 2
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                copy_live(oh[k], o_loc[k], 2, s_out)
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
        el = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": B * args.steps / el, "unit": "mul/s", "h2d_bytes_per_step": int(2 * B * S["n"] * W * 4),
               "d2h_bytes_per_step": int(B * S["n"] * W * 4),
               "api": "each rank: pinned host RAW of its own products -> H2D -> cuhe_mul_raw_sharded_batch (CRT, NCCL exchange of "
                      "residue rows, transforms, exchange back, ICRT inside the library) -> D2H (double-buffered side streams); "
                      "byte counts are whole-job totals"}
    if world == 1:
        # one e2e step = one call with Be = 8*B products from pinned host memory; the library pipelines
        # H2D | kernels | D2H over chunks of products inside the call
        Be = 8 * B
        a_np, b_np, NBUF = S["a_np"], S["b_np"], S["nbuf"]
        ah = torch.from_numpy(np.concatenate([a_np[i % NBUF] for i in range(8)]).view(np.int32)).pin_memory()
        bh = torch.from_numpy(np.concatenate([b_np[i % NBUF] for i in range(8)]).view(np.int32)).pin_memory()
        oh = torch.zeros((Be, H, W), dtype=torch.int32).pin_memory()
        for _ in range(2):
            check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st()))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            check(lib.cuhe_mul_raw_host_batch(h, p(oh), p(ah), p(bh), 0, Be, st()))
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        e2e = {"value": Be * args.steps / el, "unit": "mul/s", "h2d_bytes_per_step": int(2 * Be * S["n"] * W * 4),
               "d2h_bytes_per_step": int(Be * S["n"] * W * 4), "batch": Be,
               "h2d_GBps": 2 * Be * S["n"] * W * 4 * args.steps / el / 1e9,
               "pcie_ceiling_note": "tools/pcie_bw.py on the same box: pinned H2D 55.6 GB/s alone, 50.0 GB/s with D2H running",

               "api": "cuhe_mul_raw_host_batch (pinned host RAW in/out, 3-stream pipeline inside the call; only the modLen "
                      "coefficient rows of a ring element cross PCIe, the zero rows up to crtLen do not)"}
        del ah, bh, oh

    # ---- BASELINE configs[4]: batched multiply at 64 CRT primes (1536-bit q), residues over the ranks (8 per GPU at N = 8);
    #      same step definition, fewer steps.  Reported beside the headline so that every N has both workloads. ----
    config5 = None
    if not args.no_c5:
        for key in ("a_dev", "b_dev", "raw_out", "a_np", "b_np"):
            S.pop(key, None)
        lib.cuhe_ctx_destroy(h)
        torch.cuda.empty_cache()
        b5 = max(1, args.batch // 4)
        S5 = setup(WORKLOAD_C5, b5)
        ok5 = verify(S5)
        steps5 = max(3, args.steps // 2)
        ms5, _ = timed(S5, steps5, 3)
        config5 = {"workload": WORKLOAD_C5_NAME, "value": b5 * world * steps5 / (ms5 * 1e-3), "unit": "mul/s", "batch_per_gpu": b5,
                   "residues_per_gpu": S5["L"] // world if S5["L"] % world == 0 else f"{S5['L']}/{world}", "ms_per_step": ms5 / steps5,
                   "steps": steps5, "verified": ok5}
        lib.cuhe_ctx_destroy(S5["h"])
        h = None

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
    # BASELINE configs[3]: homomorphic PRINCE end to end against the reference's known answer (tools/prince_bench.py)
    prince = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_prince:
        prince = child_json("prince_bench.py")

    # ---- the reference interface itself: cuHE::mulZZX (ZZX in, ZZX out; cuhe/CuHE.cu:259-268) through libcuhe_compat.so,
    #      one call per product, host threads on their own streams (tools/mulzzx_bench.cpp, child process) ----
    mulzzx = None
    if rank == 0 and world == 1 and not args.no_cpu:
        exe = os.path.join(ROOT, "tools", "_mulzzx_bench")
        if os.path.exists(exe):
            try:
                nthreads = str(min(16, max(2, _host_threads())))
                r = subprocess.run([exe, nthreads, "24"], capture_output=True, text=True, timeout=240)
                lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                mulzzx = json.loads(lines[-1]) if lines else {"error": (r.stdout + r.stderr)[-300:]}
            except Exception as ex:                              # noqa: BLE001
                mulzzx = {"error": repr(ex)[:300]}
        else:
            mulzzx = {"error": "tools/_mulzzx_bench not built"}

    if rank == 0:
        out = {
            "metric": "homomorphic ctxt x ctxt mul/s", "value": value, "unit": "mul/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1) / u32 residues",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "batch": B, "batch_per_gpu": args.batch,
                       "parallelism": f"residue-shard x{world}: every rank owns {args.batch} products, residue rows exchanged by NCCL send/recv "
                                      "inside cuhe_mul_raw_sharded_batch" if world > 1 else "single GPU",
                       "l2": f"{S['nbuf']} rotating operand sets; per-step NTT intermediates {2 * B * L * N * 8 / world / 1e6:.0f} MB per GPU exceed the 126 MB L2"},
            "verified": verified, "ntt_64k_per_s": ntt_rate, "ntt_sweep": sweep, "relin_config3": relin3, "prince_config4": prince, "config5": config5,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_mulzzx": mulzzx, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(out))
    if h is not None:
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
    ap.add_argument("--no-prince", action="store_true", help="skip the homomorphic PRINCE child (BASELINE configs[3], ~40 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / relin / sweep / live ncu traffic legs")
    ap.add_argument("--no-c5", action="store_true", help="skip the BASELINE configs[4] leg (64 primes)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40          # bounded sample: each step is 8 CPU multiplications
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
