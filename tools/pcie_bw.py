#!/usr/bin/env python
"""Pinned-memory PCIe bandwidth of the box: H2D alone, D2H alone, both at once (two streams) -- the ceiling of every
end-to-end (`e2e`) number in bench.py.  python tools/pcie_bw.py [device]; start one process per GPU at the same
time to see what the host side sustains in aggregate (the `e2e` ceiling at N > 1)."""
import json
import sys
import torch

dev = torch.device("cuda", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = 1 << 28                                   # 1 GiB of int32
h_in = torch.empty(n, dtype=torch.int32).pin_memory()
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
d_in = torch.empty(n, dtype=torch.int32, device=dev)
d_out = torch.zeros(n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 4 / 1e9
s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
print(json.dumps({"device": dev.index, "h2d_GBps": gb / timed(h2d) * 1e3, "d2h_GBps": gb / timed(d2h) * 1e3,
                  "both_GBps_each_direction": gb / timed(both) * 1e3}))
