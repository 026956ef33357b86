"""Host-side mirror of the reference's public interface (cuhe/CuHE.h:46-208,
cuhe/Parameters.h:34-64) over the C ABI of libcuhe_b200.so.

Same names, argument meaning and state machine as the reference:
setParameters / initCuHE / initRelinearization / multiGPUs / numGPUs /
startAllocator / stopAllocator, CuCtxt / CuPtxt with the ZZX(0) / RAW(1) /
CRT(2) / NTT(3) domains and x2z / x2r / x2c / x2n, relin, modSwitch, and
cAnd / cXor / cNot / copy / moveTo / copyTo / mulZZX.

Differences forced by the host language: an NTL `ZZX` is a Python list of
ints (ascending coefficients); where the reference prints and calls
terminate()/exit() this raises CuHEError; initCuHE returns the coefficient
moduli instead of filling a caller array.  Device buffers are torch tensors
(int32 storage for u32 words, int64 for u64) -- torch is only the allocator
and stream provider; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import CuHEError, check, cuhe_params, load_library


# --------------------------------------------------------------------------
# cuHE::param  (cuhe/Parameters.h:34-64)
# --------------------------------------------------------------------------
class GlobalParameters:
    _FIELDS = [n for n, _ in cuhe_params._fields_]

    def __init__(self):
        self._c = cuhe_params()

    def __getattr__(self, name):
        if name in GlobalParameters._FIELDS:
            return getattr(self._c, name)
        raise AttributeError(name)

    def _q(self, fn, v):
        r = getattr(load_library(), fn)(C.byref(self._c), v)
        if r < 0 and fn != "cuhe_param_get_level":
            raise CuHEError(f"{fn}({v}) failed: {load_library().cuhe_last_error().decode()}")
        return r

    def _numCrtPrime(self, lvl):
        return self._q("cuhe_param_num_crt_prime", lvl)

    def _logCoeff(self, lvl):
        return self._q("cuhe_param_log_coeff", lvl)

    def _wordsCoeff(self, lvl):
        return self._q("cuhe_param_words_coeff", lvl)

    def _numEvalKey(self, lvl):
        return self._q("cuhe_param_num_eval_key", lvl)

    def _getLevel(self, logq):
        return self._q("cuhe_param_get_level", logq)


param = GlobalParameters()

_num_devices = 1
_ctx = {}            # device -> cuhe_ctx*
_polymod: Optional[List[int]] = None
_evalkeys_loaded = False


def _stream_ptr(st, dev):
    """Stream handle for a launch.  An explicit stream that is not torch's current one is first ordered after the
    current stream: the buffers of an operation are allocated and zero-filled there (rRepCreate / cRepCreate /
    nRepCreate), and the kernel must not overtake the fill."""
    cur = torch.cuda.current_stream(dev)
    if st is None or st.cuda_stream == cur.cuda_stream:
        return C.c_void_p(cur.cuda_stream)
    st.wait_stream(cur)
    return C.c_void_p(st.cuda_stream)


def _after(st, dev):
    """End of an operation on an explicit stream: block until it drains, as the reference does after every
    operation (cudaStreamSynchronize(st), cuhe/CuHE.cu:81-268), so that inputs dropped right afterwards
    (rRepFree / cRepFree / nRepFree hand the memory back to torch's allocator) are no longer in use."""
    if st is not None and st.cuda_stream != torch.cuda.current_stream(dev).cuda_stream:
        st.synchronize()


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def ctx(dev: int = 0):
    if dev not in _ctx:
        raise CuHEError("initCuHE has not been called for device %d" % dev)
    return _ctx[dev]


def launch_count(reset: bool = False) -> int:
    return int(load_library().cuhe_launch_count(1 if reset else 0))


# --------------------------------------------------------------------------
# init  (cuhe/CuHE.h:149-176, cuhe/CuHE.cu:36-78)
# --------------------------------------------------------------------------
def setParameters(d: int, p: int, w: int, min: int, cut: int, m: int) -> None:
    check(load_library().cuhe_set_parameters(C.byref(param._c), d, p, w, min, cut, m))


def resetParameters() -> None:
    global _polymod, _evalkeys_loaded
    lib = load_library()
    for c in _ctx.values():
        lib.cuhe_ctx_destroy(c)
    _ctx.clear()
    _polymod = None
    _evalkeys_loaded = False
    param._c = cuhe_params()


def multiGPUs(num: int) -> None:
    """setNumDevices (cuhe/DeviceManager.cu:50-70)."""
    global _num_devices
    if num < 1 or (torch.cuda.is_available() and num > torch.cuda.device_count()):
        raise CuHEError("multiGPUs: bad device count %d" % num)
    _num_devices = num


def numGPUs() -> int:
    return _num_devices


def initCuHE(polyMod: Sequence[int], shard_rank: int = 0, shard_world: int = 1,
             devices: Optional[Sequence[int]] = None) -> List[int]:
    """initCuHE(ZZ* coeffMod, ZZX polyMod) (cuhe/CuHE.cu:36-50): builds the
    NTT/CRT/Barrett tables on every device and returns coeffMod[0..depth)."""
    global _polymod
    if param.nttLen == 0:
        raise CuHEError("setParameters must be called before initCuHE")
    lib = load_library()
    devs = list(devices) if devices is not None else list(range(_num_devices))
    phi = np.ascontiguousarray(np.array([int(c) for c in polyMod], dtype=np.int64))
    for dev in devs:
        if dev in _ctx:
            lib.cuhe_ctx_destroy(_ctx.pop(dev))
        h = C.c_void_p()
        check(lib.cuhe_ctx_create(C.byref(h), C.byref(param._c), dev, shard_rank, shard_world))
        _ctx[dev] = h
        check(lib.cuhe_ctx_set_poly_modulus_host(h, phi.ctypes.data_as(C.c_void_p), len(phi)))
    _polymod = [int(c) for c in polyMod]
    out = []
    h = _ctx[devs[0]]
    for lvl in range(param.depth):
        nw = param._wordsCoeff(lvl) + 1
        buf = np.zeros(nw, dtype=np.uint32)
        check(lib.cuhe_ctx_coeff_modulus_host(h, lvl, buf.ctypes.data_as(C.c_void_p), nw))
        out.append(int.from_bytes(buf.tobytes(), "little"))
    return out


def crtPrimes(dev: int = 0) -> List[int]:
    buf = np.zeros(param.numCrtPrime, dtype=np.uint32)
    check(load_library().cuhe_ctx_crt_primes_host(ctx(dev), buf.ctypes.data_as(C.c_void_p)))
    return [int(v) for v in buf]


def startAllocator() -> None:
    """bootDeviceAllocator (cuhe/DeviceManager.cu:118-130).  The pool is a
    cudaMemPool that grows on demand; nothing to pre-carve."""


def stopAllocator() -> None:
    """haltDeviceAllocator (cuhe/DeviceManager.cu:131-138): release cached blocks."""
    for c in _ctx.values():
        check(load_library().cuhe_pool_trim(c))
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _zzx_to_raw_np(coeffs: Sequence[int], words: int, rawLen: int) -> np.ndarray:
    """z2r host half (cuhe/CuHE.cu:323-326): BytesFromZZ per coefficient."""
    nb = words * 4
    mask = (1 << (8 * nb)) - 1
    body = b"".join((abs(int(c)) & mask).to_bytes(nb, "little") for c in coeffs[:rawLen])
    buf = np.zeros(rawLen * words, dtype=np.uint32)
    arr = np.frombuffer(body, dtype="<u4")
    buf[:arr.size] = arr
    return buf.reshape(rawLen, words)


def _raw_np_to_zzx(raw: np.ndarray, modLen: int) -> List[int]:
    """r2z host half (cuhe/CuHE.cu:343-345): ZZFromBytes of the first modLen."""
    words = raw.shape[1]
    b = np.ascontiguousarray(raw[:modLen]).astype("<u4").tobytes()
    nb = words * 4
    return [int.from_bytes(b[i * nb:(i + 1) * nb], "little") for i in range(modLen)]


def initRelinearization(evalkey: Sequence[Sequence[int]], dev_list: Optional[Sequence[int]] = None) -> None:
    """initRelin (cuhe/Relinearization.cu:43-74): numEvalKey polynomials at
    level 0; transformed once and kept resident in HBM."""
    global _evalkeys_loaded
    K, W, H = param.numEvalKey, param._wordsCoeff(0), param.rawLen
    if len(evalkey) != K:
        raise CuHEError("initRelinearization needs param.numEvalKey polynomials")
    raw = np.stack([_zzx_to_raw_np(ek, W, H) for ek in evalkey])
    for dev in (dev_list if dev_list is not None else list(_ctx.keys())):
        t = torch.from_numpy(raw.view(np.int32)).to(f"cuda:{dev}")
        check(load_library().cuhe_relin_init(ctx(dev), _ptr(t), _stream_ptr(None, dev)))
        torch.cuda.synchronize(dev)
    _evalkeys_loaded = True


def initRelinearizationRaw(raw_dev: torch.Tensor, dev: int = 0) -> None:
    """Same, from a device tensor u32[numEvalKey][crtLen][words(0)]."""
    check(load_library().cuhe_relin_init(ctx(dev), _ptr(raw_dev), _stream_ptr(None, dev)))


# --------------------------------------------------------------------------
# CuPolynomial / CuCtxt / CuPtxt  (cuhe/CuHE.h:46-147, cuhe/CuHE.cu:272-606)
# --------------------------------------------------------------------------
class CuPolynomial:
    def __init__(self):
        self.logq_ = -1
        self.domain_ = -1
        self.device_ = -1
        self.zRep_: List[int] = []
        self.rRep_: Optional[torch.Tensor] = None
        self.cRep_: Optional[torch.Tensor] = None
        self.nRep_: Optional[torch.Tensor] = None
        self.isProd_ = False

    # the reference's explicit destructor calls are idempotent here
    def reset(self):
        self.zRep_ = []
        self.rRep_ = self.cRep_ = self.nRep_ = None
        self.isProd_ = False
        self.logq_ = self.domain_ = self.device_ = -1

    # set / get (cuhe/CuHE.cu:296-316): one name, optional value, like the overloads
    def logq(self, val=None):
        if val is None:
            return self.logq_
        self.logq_ = val

    def domain(self, val=None):
        if val is None:
            return self.domain_
        self.domain_ = val

    def device(self, val=None):
        if val is None:
            return self.device_
        self.device_ = val

    def isProd(self, val=None):
        if val is None:
            return self.isProd_
        self.isProd_ = bool(val)

    def zRep(self, val=None):
        if val is None:
            return self.zRep_
        self.zRep_ = list(val)

    def rRep(self, val=None):
        if val is None:
            return self.rRep_
        self.rRep_ = val

    def cRep(self, val=None):
        if val is None:
            return self.cRep_
        self.cRep_ = val

    def nRep(self, val=None):
        if val is None:
            return self.nRep_
        self.nRep_ = val

    # utilities (cuhe/CuHE.cu:521-522)
    def coeffWords(self):
        return (self.logq_ + 31) // 32

    def rRepSize(self):
        return param.rawLen * self.coeffWords() * 4

    def _rows(self):
        raise NotImplementedError

    def _lvl(self):
        return param._getLevel(self.logq_)

    def _dev(self):
        return torch.device("cuda", self.device_)

    # memory (cuhe/CuHE.cu:468-519): zero-filled buffers
    def rRepCreate(self, st=None):
        self.rRep_ = torch.zeros((param.rawLen, self.coeffWords()), dtype=torch.int32, device=self._dev())

    def cRepCreate(self, st=None):
        self.cRep_ = torch.zeros((self._rows(), param.crtLen), dtype=torch.int32, device=self._dev())

    def nRepCreate(self, st=None):
        self.nRep_ = torch.zeros((self._rows(), param.nttLen), dtype=torch.int64, device=self._dev())

    def rRepFree(self):
        self.rRep_ = None

    def cRepFree(self):
        self.cRep_ = None

    def nRepFree(self):
        self.nRep_ = None

    # ---- conversions (cuhe/CuHE.cu:317-457) ------------------------------
    def _need(self, dom, name):
        if self.domain_ != dom:
            raise CuHEError(f"Error: Not in domain {name}!")

    def z2r(self, st=None):
        self._need(0, "ZZX")
        raw = _zzx_to_raw_np(self.zRep_, self.coeffWords(), param.rawLen)
        host = torch.from_numpy(raw.view(np.int32)).pin_memory()
        self.rRep_ = host.to(self._dev(), non_blocking=False)
        self.zRep_ = []
        self.domain_ = 1

    def r2z(self, st=None):
        self._need(1, "RAW")
        raw = self.rRep_.cpu().numpy().view(np.uint32)
        self.zRep_ = _raw_np_to_zzx(raw, param.modLen)
        self.rRepFree()
        self.domain_ = 0

    def r2c(self, st=None):
        self._need(1, "RAW")
        if self.logq_ > param.logCrtPrime:
            self.cRepCreate(st)
            with torch.cuda.device(self.device_):
                check(load_library().cuhe_crt(ctx(self.device_), _ptr(self.cRep_), _ptr(self.rRep_), self._lvl(),
                                              _stream_ptr(st, self.device_)))
                _after(st, self.device_)
            self.rRepFree()
        else:
            self.cRep_ = self.rRep_.reshape(1, param.crtLen)
            self.rRep_ = None
        self.domain_ = 2

    def c2r(self, st=None):
        self._need(2, "CRT")
        if self.logq_ > param.logCrtPrime:
            self.rRepCreate(st)
            with torch.cuda.device(self.device_):
                check(load_library().cuhe_icrt(ctx(self.device_), _ptr(self.rRep_), _ptr(self.cRep_), self._lvl(), 0,
                                               param.crtLen, _stream_ptr(st, self.device_)))
                _after(st, self.device_)
            self.cRepFree()
        else:
            self.rRep_ = self.cRep_.reshape(param.rawLen, 1)
            self.cRep_ = None
        self.domain_ = 1

    def c2n(self, st=None):
        self._need(2, "CRT")
        self.nRepCreate(st)
        lib = load_library()
        with torch.cuda.device(self.device_):
            # a plaintext (logq <= logCrtPrime) is level -1: one residue (cuhe/Parameters.cu:107-109)
            check(lib.cuhe_ntt(ctx(self.device_), _ptr(self.nRep_), _ptr(self.cRep_), self._lvl(),
                               _stream_ptr(st, self.device_)))
            _after(st, self.device_)
        self.cRepFree()
        self.domain_ = 3

    def n2c(self, st=None):
        self._need(3, "NTT")
        self.cRepCreate(st)
        lib = load_library()
        with torch.cuda.device(self.device_):
            fn = lib.cuhe_intt_mod if self.isProd_ else lib.cuhe_intt
            check(fn(ctx(self.device_), _ptr(self.cRep_), _ptr(self.nRep_), self._lvl(), _stream_ptr(st, self.device_)))
            _after(st, self.device_)
        self.isProd_ = False
        self.nRepFree()
        self.domain_ = 2

    def x2z(self, st=None):
        if self.domain_ == 0:
            return
        if self.domain_ == 3:
            self.n2c(st)
        if self.domain_ == 2:
            self.c2r(st)
        self.r2z(st)

    def x2r(self, st=None):
        if self.domain_ == 1:
            return
        if self.domain_ == 0:
            return self.z2r(st)
        if self.domain_ == 3:
            self.n2c(st)
        self.c2r(st)

    def x2c(self, st=None):
        if self.domain_ == 2:
            return
        if self.domain_ == 3:
            return self.n2c(st)
        if self.domain_ == 0:
            self.z2r(st)
        self.r2c(st)

    def x2n(self, st=None):
        if self.domain_ == 3:
            return
        if self.domain_ == 0:
            self.z2r(st)
        if self.domain_ == 1:
            self.r2c(st)
        self.c2n(st)


class CuCtxt(CuPolynomial):
    def __init__(self):
        super().__init__()
        self.level_ = -1

    def _rows(self):
        return param._numCrtPrime(self.level_)

    def _lvl(self):
        return self.level_

    def setLevel(self, lvl, a, b=None, st=None):
        """setLevel(lvl, domain, device[, st]) or setLevel(lvl, device, ZZX)
        (cuhe/CuHE.cu:525-541)."""
        self.level_ = lvl
        self.logq_ = param._logCoeff(lvl)
        if isinstance(b, (list, tuple)):
            self.domain_, self.device_ = 0, a
            self.zRep_ = list(b)
            return
        self.domain_, self.device_ = a, b
        if self.domain_ == 0:
            self.zRep_ = []
        elif self.domain_ == 1:
            self.rRepCreate(st)
        elif self.domain_ == 2:
            self.cRepCreate(st)
        elif self.domain_ == 3:
            self.nRepCreate(st)

    def level(self):
        return self.level_

    def reset(self):
        super().reset()
        self.level_ = -1

    def cRepSize(self):
        return param._numCrtPrime(self.level_) * param.crtLen * 4

    def nRepSize(self):
        return param._numCrtPrime(self.level_) * param.nttLen * 8

    def modSwitch(self, st=None):
        """cuhe/CuHE.cu:543-554"""
        if self.logq_ < param.logCoeffMin + param.logCoeffCut:
            raise CuHEError("Error: Cannot do modSwitch on last level!")
        self.x2c()
        L = param._numCrtPrime(self.level_)
        with torch.cuda.device(self.device_):
            check(load_library().cuhe_mod_switch(ctx(self.device_), _ptr(self.cRep_), _ptr(self.cRep_),
                                                 _ptr(self.cRep_[L - 1]), self.level_, _stream_ptr(st, self.device_)))
            _after(st, self.device_)
        self.cRep_ = self.cRep_[:L - 1]
        self.logq_ -= param.logCoeffCut
        self.level_ += 1

    def dropToLevel(self, lvl, st=None):
        """EXTENSION (not in the reference): the same ciphertext modulo q_lvl, on the device.  q_lvl is the
        product of the first numCrtPrime(lvl) primes, so in the CRT domain `c mod q_lvl` is just the first
        rows -- the device-resident form of the host-side `coeffReduce(x, x, lvl); setLevel(lvl, ...)` that
        examples/Prince/Prince.cu:192-193,210-213 performs on ZZX values between S-box layers."""
        if lvl < self.level_:
            raise CuHEError("Error: dropToLevel cannot raise the modulus!")
        if lvl == self.level_:
            return
        self.x2c(st)
        self.cRep_ = self.cRep_[:param._numCrtPrime(lvl)].clone()
        self.level_ = lvl
        self.logq_ = param._logCoeff(lvl)

    def relin(self, st=None):
        """cuhe/CuHE.cu:570-581"""
        if not _evalkeys_loaded:
            raise CuHEError("initRelinearization has not been called")
        self.x2r()
        self.nRepCreate(st)
        with torch.cuda.device(self.device_):
            check(load_library().cuhe_relin(ctx(self.device_), _ptr(self.nRep_), _ptr(self.rRep_), self.level_,
                                            _stream_ptr(st, self.device_)))
            _after(st, self.device_)
        self.rRepFree()
        self.isProd_ = True
        self.domain_ = 3
        self.n2c()


class CuPtxt(CuPolynomial):
    def _rows(self):
        return 1

    def setLogq(self, logq, a, b=None, st=None):
        """setLogq(logq, domain, device[, st]) or setLogq(logq, device, ZZX)
        (cuhe/CuHE.cu:585-602)."""
        self.logq_ = logq
        if isinstance(b, (list, tuple)):
            self.domain_, self.device_ = 0, a
            self.zRep_ = list(b)
            return
        self.domain_, self.device_ = a, b
        if self.domain_ == 1:
            self.rRepCreate(st)
        elif self.domain_ == 2:
            self.cRepCreate(st)
        elif self.domain_ == 3:
            self.nRepCreate(st)

    def cRepSize(self):
        return param.crtLen * 4

    def nRepSize(self):
        return param.nttLen * 8


# --------------------------------------------------------------------------
# operations  (cuhe/CuHE.h:178-208, cuhe/CuHE.cu:81-268)
# --------------------------------------------------------------------------
def copy(dst: CuCtxt, src: CuCtxt, st=None):
    if dst is src:
        return
    dst.reset()
    dst.level_ = src.level_
    dst.logq_, dst.domain_, dst.device_ = src.logq_, src.domain_, src.device_
    dst.isProd_ = src.isProd_
    if src.domain_ == 0:
        dst.zRep_ = list(src.zRep_)
    elif src.domain_ == 1:
        dst.rRep_ = src.rRep_.clone()
    elif src.domain_ == 2:
        dst.cRep_ = src.cRep_.clone()
    elif src.domain_ == 3:
        dst.nRep_ = src.nRep_.clone()


def _same(in0, in1, what):
    if in0.device() != in1.device():
        raise CuHEError(f"Error: {what} of different devices!")


def cAnd(out: CuCtxt, in0: CuCtxt, in1, st=None):
    """ctxt x ctxt (cuhe/CuHE.cu:101-122) or ctxt x ptxt (:123-144)."""
    _same(in0, in1, "Multiplication")
    if in0.domain() != 3 or in1.domain() != 3:
        raise CuHEError("Error: Multiplication of non-NTT domain!")
    is_ptxt = isinstance(in1, CuPtxt)
    if not is_ptxt and in0.logq() != in1.logq():
        raise CuHEError("Error: Multiplication of different levels!")
    if out is not in0:
        out.reset()
        out.setLevel(in0.level(), 3, in0.device(), st)
    lib = load_library()
    fn = lib.cuhe_ntt_mul_nx1 if is_ptxt else lib.cuhe_ntt_mul
    with torch.cuda.device(out.device()):
        check(fn(ctx(out.device()), _ptr(out.nRep_), _ptr(in0.nRep_), _ptr(in1.nRep_), out.level(),
                 _stream_ptr(st, out.device())))
        _after(st, out.device())
    out.isProd(True)


def cXor(out: CuCtxt, in0: CuCtxt, in1, st=None):
    """cuhe/CuHE.cu:145-206"""
    _same(in0, in1, "Addition")
    is_ptxt = isinstance(in1, CuPtxt)
    if not is_ptxt and in0.logq() != in1.logq():
        raise CuHEError("Error: Addition of different levels!")
    lib = load_library()
    dom = in0.domain()
    if dom not in (2, 3) or in1.domain() != dom:
        raise CuHEError("Error: Addition of non-CRT-nor-NTT domain!")
    if out is not in0:
        out.reset()
        out.setLevel(in0.level(), dom, in0.device(), st)
        if dom == 3:
            out.isProd(in0.isProd() or in1.isProd())
    with torch.cuda.device(out.device()):
        s = _stream_ptr(st, out.device())
        if dom == 2:
            fn = lib.cuhe_crt_add_nx1 if is_ptxt else lib.cuhe_crt_add
            check(fn(ctx(out.device()), _ptr(out.cRep_), _ptr(in0.cRep_), _ptr(in1.cRep_), out.level(), s))
        else:
            fn = lib.cuhe_ntt_add_nx1 if is_ptxt else lib.cuhe_ntt_add
            check(fn(ctx(out.device()), _ptr(out.nRep_), _ptr(in0.nRep_), _ptr(in1.nRep_), out.level(), s))
        _after(st, out.device())


def cNot(out: CuCtxt, inp: CuCtxt, st=None):
    """cuhe/CuHE.cu:207-218"""
    if inp.domain() != 2:
        raise CuHEError("Error: cNot of non-CRT domain!")
    if out is not inp:
        out.reset()
        out.setLevel(inp.level(), inp.domain(), inp.device(), st)
        out.cRep_.copy_(inp.cRep_)
    with torch.cuda.device(out.device()):
        check(load_library().cuhe_crt_add_int(ctx(out.device()), _ptr(out.cRep_), _ptr(inp.cRep_),
                                              C.c_uint(param.modMsg - 1), out.level(), _stream_ptr(st, out.device())))
        _after(st, out.device())


def moveTo(tar: CuCtxt, dstDev: int, st=None):
    """cuhe/CuHE.cu:217-251 (cudaMemcpyPeerAsync)"""
    if dstDev == tar.device():
        return
    d = torch.device("cuda", dstDev)
    for name in ("rRep_", "cRep_", "nRep_"):
        t = getattr(tar, name)
        if t is not None:
            setattr(tar, name, t.to(d))
    tar.device(dstDev)


def copyTo(dst: CuCtxt, src: CuCtxt, dstDev: int, st=None):
    copy(dst, src, st)
    moveTo(dst, dstDev, st)


def mulZZX(in0: Sequence[int], in1: Sequence[int], lvl: int, dev: int = 0, st=None) -> List[int]:
    """mulZZX(out, in0, in1, lvl, dev, st) (cuhe/CuHE.cu:259-268); returns out."""
    cin0, cin1 = CuCtxt(), CuCtxt()
    cin0.setLevel(lvl, dev, list(in0))
    cin1.setLevel(lvl, dev, list(in1))
    cin0.x2n(st)
    cin1.x2n(st)
    cAnd(cin0, cin0, cin1, st)
    cin0.x2z(st)
    return cin0.zRep()
