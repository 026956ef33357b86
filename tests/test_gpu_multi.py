"""Multi-GPU parity on hardware: the residue-sharded multiply with the NCCL exchange inside the library
(cuhe_mul_raw_sharded_batch) against the unsharded path, one process per GPU under torchrun
(tools/sharded_check.py).  Skipped on a single-GPU box; the world-size-1 form runs everywhere."""
import os
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


def _run(nproc):
    env = dict(os.environ, NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "sharded_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)


def test_sharded_multiply_world_1():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and '"sharded_check": "ok"' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_multiply_equals_unsharded_on_n_gpus(nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = _run(nproc)
    assert r.returncode == 0 and '"sharded_check": "ok"' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_reference_multi_device_semantics_moveto_copyto():
    """multiGPUs(2) in ONE process as the reference does it (whole ciphertexts per device, replicated tables,
    moveTo / copyTo, cuhe/CuHE.cu:217-257): tools/multidev_check.py against the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multidev_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and '"multidev_check": "ok"' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_cpp_host_layer_on_two_devices():
    """the same in the C++ host layer (cuhe_b200/host/cuhe_compat.cpp: moveTo / copyTo allocate from the destination
    device's pool and copy device to device): tests/cpp/compat_test.cpp run as `compat_test twodev`."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_abi import _build_cpp_compat_test
    exe = _build_cpp_compat_test()
    r = subprocess.run([exe, "twodev"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "compat ok" in r.stdout and "two devices" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
