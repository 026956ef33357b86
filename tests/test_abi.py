"""C-ABI surface checks that need no GPU: the library loads, exports every
symbol include/cuhe_b200.h declares, and its host-only entry points (parameter
derivation, cuhe/Parameters.cu:53-145) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np

import pytest

from common import C2, PRINCE, ROOT, SIMPLE_DHS, SMALL_RELIN, MID32K, MID64K


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cuhe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cuhe_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libcuhe_b200.so does not export {n}"


def test_binding_table_matches_header(lib):
    from cuhe_b200._lib import SYMBOLS
    assert sorted(SYMBOLS) == _declared_symbols()


@pytest.mark.parametrize("ps", [SIMPLE_DHS, SMALL_RELIN, PRINCE, MID32K, C2, MID64K, (44, 2, 16, 24, 24, 32767),
                                (64, 2, 16, 24, 24, 32767)])
def test_set_parameters_matches_oracle(lib, ps):
    from cuhe_b200._lib import cuhe_params
    from oracle import pyoracle as po
    cp = cuhe_params()
    assert lib.cuhe_set_parameters(C.byref(cp), *ps) == 0
    op = po.set_param(*ps)
    for name, _ in cuhe_params._fields_:
        assert getattr(cp, name) == getattr(op, name), name
    for lvl in range(op.depth):
        assert lib.cuhe_param_num_crt_prime(C.byref(cp), lvl) == op._numCrtPrime(lvl)
        assert lib.cuhe_param_log_coeff(C.byref(cp), lvl) == op._logCoeff(lvl)
        assert lib.cuhe_param_words_coeff(C.byref(cp), lvl) == op._wordsCoeff(lvl)
        if op.logRelin:
            assert lib.cuhe_param_num_eval_key(C.byref(cp), lvl) == op._numEvalKey(lvl)
        assert lib.cuhe_param_get_level(C.byref(cp), op._logCoeff(lvl)) == lvl
    assert lib.cuhe_param_get_level(C.byref(cp), 1) == -1


def test_bad_parameters_fail_loudly(lib):
    from cuhe_b200._lib import cuhe_params
    cp = cuhe_params()
    # nttLen would be 2^17: beyond the reference's preload_ntt (cuhe/Base.cu:58-62)
    assert lib.cuhe_set_parameters(C.byref(cp), 3, 2, 16, 30, 20, 65537) != 0
    assert b"nttLen" in lib.cuhe_last_error()
    assert lib.cuhe_param_num_crt_prime(None, 0) < 0


def test_python_mirror_raises_without_init(lib):
    import cuhe_b200 as ch
    ch.resetParameters()
    with pytest.raises(ch.CuHEError):
        ch.initCuHE([1, 1])
    with pytest.raises(ch.CuHEError):
        ch.ctx(0)


def test_no_oracle_in_product():
    """The shipped package must not import or reference oracle/."""
    pkg = os.path.join(ROOT, "cuhe_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "coracle" not in src, f


def _build_c_client():
    import subprocess
    out = os.path.join(ROOT, "tests", "_abi_client")
    libdir = os.path.join(ROOT, "cuhe_b200")
    subprocess.check_call(["gcc", "-O1", "-o", out, os.path.join(ROOT, "tests", "abi_client.c"),
                           "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           "-L", libdir, "-lcuhe_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                           f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return out


def test_plain_c_client_host_mode(lib):
    """The boundary is a C ABI: a C program includes the header, links the .so and
    derives the reference's parameters without C++/Python/GPU."""
    import subprocess
    exe = _build_c_client()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "params 8190 16384 7 141 5" in r.stdout and "host ok" in r.stdout


@pytest.mark.gpu
def test_plain_c_client_gpu_mode(lib):
    import subprocess
    exe = _build_c_client()
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu ok" in r.stdout


def _build_cpp_compat_test(name="compat_test"):
    import subprocess
    libdir = os.path.join(ROOT, "cuhe_b200")
    if not os.path.exists(os.path.join(libdir, "libcuhe_compat.so")):
        import __graft_entry__ as ge
        ge.build()
    out = os.path.join(ROOT, "tests", "_" + name)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(libdir, "host"),
                           "-I", "/usr/local/cuda/include", "-o", out,
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
                           "-L", libdir, "-lcuhe_compat", "-lcuhe_b200", f"-Wl,-rpath,{libdir}"])
    return out


def test_cpp_key_text_format(lib):
    """cuHE_Utils::Picklable / PicklableMap (cuhe/Utils.h:39-93), the key/polynomial text format of
    examples/DHS/DHS.cu:57-189 -- pure host code, runs without a GPU."""
    import subprocess
    exe = _build_cpp_compat_test("utils_test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "utils ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_binary_rns_container_python_and_cpp_agree(lib, tmp_path):
    """The binary RNS container (residue-domain data as the device holds it + parameter tuple + checksum): Python
    round trip, damage detection, and the C++ twin (cuHE_Utils::RnsBlob) reads a file written by Python and writes
    back the identical bytes."""
    import subprocess
    import numpy as np
    from cuhe_b200 import utils
    a = np.random.default_rng(3).integers(0, 1 << 26, size=(7, 8192), dtype=np.uint32)
    path = str(tmp_path / "crt.rns")
    utils.save_rns(path, a, (5, 2, 1, 61, 20, 8191), domain=2, level=0)
    b, meta = utils.load_rns(path)
    assert np.array_equal(a, b) and meta["params"] == (5, 2, 1, 61, 20, 8191) and meta["domain"] == 2
    k = np.random.default_rng(4).integers(0, 1 << 63, size=(2, 3, 1024), dtype=np.uint64)
    utils.save_rns(str(tmp_path / "ntt.rns"), k, (3, 2, 16, 48, 24, 32767), domain=3, level=0, shard=(1, 2))
    k2, meta2 = utils.load_rns(str(tmp_path / "ntt.rns"))
    assert np.array_equal(k, k2) and meta2["shard"] == (1, 2)
    raw = bytearray(open(path, "rb").read())
    raw[200] ^= 0x40
    open(str(tmp_path / "bad.rns"), "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        utils.load_rns(str(tmp_path / "bad.rns"))
    exe = _build_cpp_compat_test("utils_test")
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "utils ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(path + ".roundtrip", "rb").read() == open(path, "rb").read()


def test_cpp_host_layer_builds_and_exports_the_reference_interface(lib):
    """cuhe_b200/host: the C++ layer with the reference's names (namespace cuHE) compiles with plain
    g++ (no NTL, no nvcc) and a client written against that interface links."""
    import subprocess
    _build_cpp_compat_test()
    syms = subprocess.run(["nm", "-DC", os.path.join(ROOT, "cuhe_b200", "libcuhe_compat.so")],
                          capture_output=True, text=True).stdout
    for name in ("cuHE::setParameters(int, int, int, int, int, int)", "cuHE::initCuHE(", "cuHE::mulZZX(",
                 "cuHE::cAnd(cuHE::CuCtxt&, cuHE::CuCtxt&, cuHE::CuCtxt&", "cuHE::cAnd(cuHE::CuCtxt&, cuHE::CuCtxt&, cuHE::CuPtxt&",
                 "cuHE::cXor(", "cuHE::cNot(", "cuHE::copy(cuHE::CuCtxt&, cuHE::CuCtxt", "cuHE::moveTo(", "cuHE::copyTo(",
                 "cuHE::CuCtxt::relin(", "cuHE::CuCtxt::modSwitch(", "cuHE::CuPolynomial::x2n(", "cuHE::initRelinearization(",
                 "cuHE::multiGPUs(int)", "cuHE::numGPUs()", "cuHE::startAllocator()", "cuHE::stopAllocator()", "cuHE::param",
                 "cuHE_Utils::Picklable::pickle", "cuHE_Utils::PicklableMap::toString", "cuHE_Utils::PicklableMap::get("):
        assert name in syms, name


@pytest.mark.gpu
def test_cpp_client_of_the_reference_interface(lib):
    """tests/cpp/compat_test.cpp: simple_DHS-style call sequences through the C++ host layer,
    checked against exact host big-integer arithmetic."""
    import subprocess
    exe = _build_cpp_compat_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "compat ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_python_key_text_format_matches_cpp(lib):
    """cuhe_b200/utils.py and the C++ twin agree on the text (same strings as tests/cpp/utils_test.cpp)."""
    from cuhe_b200.utils import Picklable, PicklableMap
    big = 1234567890123456789012345678901234567890123456789012345678901234567890
    pk = Picklable.from_poly("pk0", [5, big, 0, 7, 0, 0])
    assert pk.pickle() == f"pk0,5,{big},0,7" and pk.getCoeffsLen() == 4
    d = Picklable("d", [24, 0, 0])
    assert d.pickle() == "d,24,0,0" and d.getPoly() == [24]
    semi = Picklable.parse("x;1;;2", ";")
    assert semi.pickle() == "x;1;2"
    m = PicklableMap([d, pk])
    assert m.toString() == d.pickle() + "\n" + pk.pickle()
    back = PicklableMap.parse(m.toString())
    assert back.get("pk0").getPoly() == [5, big, 0, 7] and back.get("d").getValues() == "24"
    with pytest.raises(KeyError):
        back.get("nope")
    assert PicklableMap.parse("a:1:2|b:3", "|", ":").toString() == "a:1:2|b:3"


def test_missing_library_fails_loudly():
    """No CPU fallback: without the CUDA library every entry point of the package raises."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import cuhe_b200 as ch\n"
            "try:\n"
            "    ch.setParameters(5, 2, 1, 61, 20, 8191)\n"
            "except Exception as e:\n"
            "    print('RAISED', type(e).__name__, e)\n"
            "else:\n"
            "    print('NO ERROR')\n") % ROOT
    env = dict(os.environ, CUHE_B200_LIB="/nonexistent/libcuhe_b200.so")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "RAISED" in r.stdout and "NO ERROR" not in r.stdout, r.stdout + r.stderr
    assert "libcuhe_b200" in r.stdout


def test_cpp_bigint_stand_in_against_python_integers(lib):
    """cuhe_b200/host/zz_lite.hpp (used when NTL is not installed): +, -, *, % (result in [0, m)), NumBits,
    byte import/export and comparisons equal Python integers on 300 random and edge operands."""
    import random
    import subprocess
    exe = _build_cpp_compat_test("zz_test")
    rng = random.Random(5)
    cases = [(0, 0, 1), (1, -1, 2), (-5, 3, 7), (2**64, 2**64 - 1, 2**32), (-(2**200), 2**199 + 1, 10**9 + 7),
             (10**9, 10**9, 10**9), (10**18 - 1, 1, 10**9)]
    for _ in range(300):
        ba, bb, bm = rng.choice([1, 31, 32, 33, 63, 64, 65, 576, 1152, 2000]), rng.randrange(1, 1200), rng.randrange(1, 700)
        a, b = rng.getrandbits(ba) * rng.choice([1, -1]), rng.getrandbits(bb) * rng.choice([1, -1])
        cases.append((a, b, rng.getrandbits(bm) + 1))
    text = "".join(f"{a} {b} {m}\n" for a, b, m in cases)
    r = subprocess.run([exe], input=text, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == len(cases)
    for (a, b, m), line in zip(cases, lines):
        want = f"{a + b} {a - b} {a * b} {a % m} {abs(a).bit_length()} {abs(a)} {1 if a < b else 0}{1 if a == b else 0}"
        assert line == want, (a, b, m)


def test_host_marshalling_equals_oracle_layout(lib):
    """z2r / r2z host halves of the Python mirror (cuhe/CuHE.cu:317-348: BytesFromZZ / ZZFromBytes per
    coefficient): RAW u32[rawLen][words], little-endian words, zero beyond the polynomial."""
    import random
    from cuhe_b200.api import _raw_np_to_zzx, _zzx_to_raw_np
    from common import get_oracle
    o = get_oracle(SMALL_RELIN)
    rng = random.Random(2)
    for lvl in (0, o.par.depth - 1):
        W = o.W(lvl)
        coeffs = [rng.randrange(o.moduli[lvl]) for _ in range(o.n)]
        coeffs[0], coeffs[1], coeffs[-1] = 0, o.moduli[lvl] - 1, 1
        raw = _zzx_to_raw_np(coeffs, W, o.H)
        assert raw.shape == (o.H, W) and raw.dtype == np.uint32
        assert np.array_equal(raw, o.to_raw(coeffs, lvl))
        assert _raw_np_to_zzx(raw, o.n) == coeffs
    short = _zzx_to_raw_np([5, 6], 2, o.H)                      # a short ZZX is zero-extended
    assert short[0, 0] == 5 and short[1, 0] == 6 and not short[2:].any()


def test_lazy_96bit_arithmetic_on_the_host():
    """cuhe_b200/csrc/l96.cuh + ntt96_core.cuh (the arithmetic inside the NTT kernels) compiled for the CPU with every
    intermediate an exact 128-bit integer and the 96-bit window enforced: residues, stated magnitudes, the 4/8/16-point
    blocks against the O(n^2) definition (tests/cpp/l96_host_test.cpp)."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_l96_host_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "l96_host_test.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 failures, 0 window overflows" in r.stdout, r.stdout + r.stderr


def test_ntt_pass_kernels_emulated_on_the_host():
    """cuhe_b200/csrc/ntt4.cuh: the phase bodies of the two NTT pass kernels are __host__ __device__; every CTA is run
    thread by thread, phase by phase on the CPU (exact 128-bit intermediates, 96-bit window enforced) for N = 16384,
    32768, 65536: forward outputs against the definition X[k] = sum x[j] w^(jk), inverse + %p round trip, the fused
    pointwise product against a convolution, the table epilogue (tests/cpp/ntt4_host_test.cpp)."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_ntt4_host_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", "/usr/local/cuda/include", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "ntt4_host_test.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 failures, 0 window overflows" in r.stdout, r.stdout + r.stderr


def test_reference_include_tree_parses():
    """compat/cuhe/{CuHE,Parameters,Utils,DeviceManager,Debug}.h: the include tree the reference's unchanged examples
    resolve `../../cuhe/CuHE.h` against (compat/Makefile builds examples/DHS and examples/Prince from it when NTL is
    installed); here: a caller written like examples/DHS/DHS.cu:34-55,218 parses against it without NTL."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "compat"), "-s", "check", f"BUILD={os.path.join('/tmp', 'cuhe_b200_examples_test')}"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "include tree ok" in r.stdout, r.stdout + r.stderr
