"""Device-resident, batched evaluation of boolean circuits over DHS ciphertexts.

The reference evaluates a circuit one ciphertext operation at a time: every S-box of examples/Prince wraps four
host ZZX values into CuCtxt objects, runs ~300 kernel launches far below one wave each and copies four
polynomials back (examples/Prince/Prince.cu:204-322; 16 independent S-boxes per layer are spread over OpenMP
threads / GPUs, Prince.cu:191-201).  Here a `CtxtBatch` is a stack of B independent ciphertexts of one level
resident in HBM, and every operation -- cAnd, cXor, cNot, relin, modSwitch and the domain conversions between
them, with the reference's semantics (cuhe/CuHE.cu:81-268,543-581) -- is ONE launch set for all B through the
batched C-ABI entry points (cuhe_*_batch).  A PRINCE layer is then 16-wide: ~60 batched operations instead of
~5000 launches, and no polynomial crosses PCIe between encryption and decryption.

Only data layout and sequencing live here; every arithmetic step is a kernel of libcuhe_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import api
from ._lib import check, load_library


def _p(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


class CtxtBatch:
    """B ciphertexts of one level on one device.  domain 2: CRT, u32 [B][L][crtLen]; domain 3: NTT, u64 [B][L][nttLen]
    (stored as torch int32 / int64).  `is_prod` has the meaning of CuPolynomial::isProd (the next n2c reduces mod Phi_m)."""

    __slots__ = ("t", "level", "domain", "is_prod", "device")

    def __init__(self, t: torch.Tensor, level: int, domain: int, is_prod: bool = False, device: int = 0):
        self.t, self.level, self.domain, self.is_prod, self.device = t, level, domain, is_prod, device

    @property
    def batch(self) -> int:
        return self.t.shape[0]

    def clone(self) -> "CtxtBatch":
        return CtxtBatch(self.t.clone(), self.level, self.domain, self.is_prod, self.device)


class BatchOps:
    """The ciphertext operations of cuhe/CuHE.h on CtxtBatch values (unsharded context of `device`)."""

    def __init__(self, device: int = 0):
        self.device = device
        self.lib = load_library()
        self.par = api.param
        self.counts = dict(cAnd=0, relin=0, modSwitch=0, launches_start=self.lib.cuhe_launch_count(0))

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _ctx(self):
        return api.ctx(self.device)

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return torch.device("cuda", self.device)

    def _L(self, lvl):
        return self.par._numCrtPrime(lvl)

    def _empty(self, B, lvl, domain):
        if domain == 2:       # zero-filled like cRepCreate (cuhe/CuHE.cu:468-488): kernels write coefficients < modLen only
            return torch.zeros((B, self._L(lvl), self.par.crtLen), dtype=torch.int32, device=self._dev())
        return torch.empty((B, self._L(lvl), self.par.nttLen), dtype=torch.int64, device=self._dev())

    # ---- construction -----------------------------------------------------------------------------------------
    def stack(self, ctxts: Sequence["api.CuCtxt"]) -> CtxtBatch:
        """CRT-domain CuCtxt objects of one level -> one batch (device copy)"""
        lvl = ctxts[0].level()
        for c in ctxts:
            if c.domain() != 2 or c.level() != lvl:
                raise api.CuHEError("stack: ciphertexts must be in the CRT domain at one level")
        return CtxtBatch(torch.stack([c.cRep() for c in ctxts]), lvl, 2, False, self.device)

    def from_rows(self, t: torch.Tensor, level: int) -> CtxtBatch:
        return CtxtBatch(t.contiguous(), level, 2, False, self.device)

    def unstack(self, b: CtxtBatch) -> List["api.CuCtxt"]:
        self.to_crt(b)
        out = []
        for i in range(b.batch):
            c = api.CuCtxt()
            c.setLevel(b.level, 2, self.device)
            c.cRep_ = b.t[i].clone()
            out.append(c)
        return out

    # ---- domain conversions (cuhe/CuHE.cu:383-410) ---------------------------------------------------------------
    def to_ntt(self, b: CtxtBatch) -> CtxtBatch:
        if b.domain == 3:
            return b
        out = self._empty(b.batch, b.level, 3)
        check(self.lib.cuhe_ntt_batch(self._ctx(), _p(out), _p(b.t), b.level, b.batch, self._st()))
        b.t, b.domain = out, 3
        return b

    def to_crt(self, b: CtxtBatch) -> CtxtBatch:
        if b.domain == 2:
            return b
        if not b.is_prod:
            raise api.CuHEError("to_crt: only products leave the NTT domain in a circuit (n2c with isProd)")
        out = self._empty(b.batch, b.level, 2)
        check(self.lib.cuhe_intt_mod_batch(self._ctx(), _p(out), _p(b.t), None, b.level, b.batch, self._st()))
        b.t, b.domain, b.is_prod = out, 2, False
        return b

    # ---- operations (cuhe/CuHE.cu:101-216) -------------------------------------------------------------------------
    def band(self, x: CtxtBatch, y: CtxtBatch) -> CtxtBatch:
        """cAnd + n2c: both operands in the NTT domain; the product is fused into the inverse transform and comes back
        reduced modulo Phi_m in the CRT domain (what cAnd followed by x2c gives, CuHE.cu:101-122,394-410)"""
        if x.domain != 3 or y.domain != 3:
            raise api.CuHEError("Error: Multiplication of non-NTT domain!")
        if x.level != y.level or x.batch != y.batch:
            raise api.CuHEError("Error: Multiplication of different levels!")
        out = self._empty(x.batch, x.level, 2)
        check(self.lib.cuhe_intt_mod_batch(self._ctx(), _p(out), _p(x.t), _p(y.t), x.level, x.batch, self._st()))
        self.counts["cAnd"] += x.batch
        return CtxtBatch(out, x.level, 2, False, self.device)

    def bxor(self, x: CtxtBatch, y: CtxtBatch) -> CtxtBatch:
        """cXor in the CRT domain (CuHE.cu:141-176)"""
        if x.domain != 2 or y.domain != 2:
            raise api.CuHEError("Error: Addition of non-CRT domain!")
        if x.level != y.level or x.batch != y.batch:
            raise api.CuHEError("Error: Addition of different levels!")
        out = torch.zeros_like(x.t)
        check(self.lib.cuhe_crt_add_batch(self._ctx(), _p(out), _p(x.t), _p(y.t), x.level, x.batch, self._st()))
        return CtxtBatch(out, x.level, 2, False, self.device)

    def bxor_(self, x: CtxtBatch, y: CtxtBatch) -> CtxtBatch:
        check(self.lib.cuhe_crt_add_batch(self._ctx(), _p(x.t), _p(x.t), _p(y.t), x.level, x.batch, self._st()))
        return x

    def bnot_(self, x: CtxtBatch) -> CtxtBatch:
        """cNot: + (modMsg - 1) on coefficient 0 (CuHE.cu:207-218)"""
        if x.domain != 2:
            raise api.CuHEError("Error: cNot of non-CRT domain!")
        check(self.lib.cuhe_crt_add_int_batch(self._ctx(), _p(x.t), _p(x.t), C.c_uint(self.par.modMsg - 1), x.level, x.batch,
                                              self._st()))
        return x

    def relin_(self, x: CtxtBatch) -> CtxtBatch:
        """CuCtxt::relin (CuHE.cu:570-581): c2r, key switch into the NTT domain (isProd), n2c"""
        self.to_crt(x)
        H, W = self.par.rawLen, self.par._wordsCoeff(x.level)
        raw = torch.zeros((x.batch, H, W), dtype=torch.int32, device=self._dev())
        check(self.lib.cuhe_icrt_batch(self._ctx(), _p(raw), _p(x.t), x.level, 0, self.par.crtLen, x.batch, self._st()))
        nt = self._empty(x.batch, x.level, 3)
        check(self.lib.cuhe_relin_batch(self._ctx(), _p(nt), _p(raw), x.level, x.batch, self._st()))
        x.t, x.domain, x.is_prod = nt, 3, True
        self.counts["relin"] += x.batch
        return self.to_crt(x)

    def mod_switch_(self, x: CtxtBatch) -> CtxtBatch:
        """CuCtxt::modSwitch (CuHE.cu:543-554): one level down"""
        self.to_crt(x)
        out = self._empty(x.batch, x.level + 1, 2)
        check(self.lib.cuhe_mod_switch_batch(self._ctx(), _p(out), _p(x.t), x.level, x.batch, self._st()))
        x.t, x.level = out, x.level + 1
        self.counts["modSwitch"] += x.batch
        return x

    def drop_to_level(self, x: CtxtBatch, lvl: int) -> CtxtBatch:
        """a fresh (noise-free enough) ciphertext used at a deeper level: keep the first rows (the host-side equivalent
        is coeffReduce + re-upload, examples/Prince/Prince.cu:192-193)"""
        if x.domain != 2 or lvl < x.level:
            raise api.CuHEError("drop_to_level: CRT domain, deeper level only")
        if lvl == x.level:
            return x
        return CtxtBatch(x.t[:, : self._L(lvl)].contiguous(), lvl, 2, False, self.device)

    def launches(self) -> int:
        return int(self.lib.cuhe_launch_count(0)) - self.counts["launches_start"]
