"""Key / polynomial text format of the reference (namespace cuHE_Utils, cuhe/Utils.h:39-93,
cuhe/Utils.cu:75-152): one record per polynomial, ``key,c0,c1,...`` with decimal coefficients in
ascending order, records joined by newlines.  Host-side only; used to move DHS keys between
processes (examples/DHS/DHS.cu:57-189).  The C++ twin is cuhe_b200/host/cuhe_utils.hpp."""
from __future__ import annotations

from typing import Iterable, List, Sequence


def _tokens(text: str, delims: str) -> List[str]:
    """split on ANY character of `delims`, dropping empty tokens (strtok semantics, Utils.cu:77-93)"""
    out, cur = [], []
    for ch in text:
        if ch in delims:
            if cur:
                out.append("".join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return out


class Picklable:
    """Picklable(key, coeffs) keeps the given coefficient count (trailing zeros included, the
    reference's (key, ZZ*, len) constructor); Picklable.parse(text) and Picklable.from_poly(key, poly)
    normalise like a ZZX (trailing zeros dropped)."""

    def __init__(self, key: str, coeffs: Sequence[int], separator: str = ","):
        self.key = key
        self.coeffs = [int(c) for c in coeffs]
        self.separator = separator

    @classmethod
    def from_poly(cls, key: str, poly: Sequence[int], separator: str = ",") -> "Picklable":
        c = [int(v) for v in poly]
        while c and c[-1] == 0:
            c.pop()
        return cls(key, c, separator)

    @classmethod
    def parse(cls, data: str, separator: str = ",") -> "Picklable":
        t = _tokens(data, separator)
        key = t[0] if t else ""
        return cls.from_poly(key, [int(v) for v in t[1:]], separator)

    def setSeparator(self, sep: str) -> None:
        self.separator = sep

    def getSeparator(self) -> str:
        return self.separator

    def getKey(self) -> str:
        return self.key

    def getCoeffs(self) -> List[int]:
        return self.coeffs

    def getCoeffsLen(self) -> int:
        return len(self.coeffs)

    def getPoly(self) -> List[int]:
        c = list(self.coeffs)
        while c and c[-1] == 0:
            c.pop()
        return c

    def getValues(self) -> str:
        return self.separator.join(str(c) for c in self.coeffs)

    def pickle(self) -> str:
        return self.key + self.separator + self.getValues()


class PicklableMap:
    def __init__(self, items: Iterable[Picklable] = (), separator: str = "\n"):
        self.items = list(items)
        self.separator = separator

    @classmethod
    def parse(cls, data: str, separator: str = "\n", field_separator: str = ",") -> "PicklableMap":
        return cls([Picklable.parse(rec, field_separator) for rec in _tokens(data, separator)], separator)

    def setSeparator(self, sep: str) -> None:
        self.separator = sep

    def getSeparator(self) -> str:
        return self.separator

    def getPicklables(self) -> List[Picklable]:
        return self.items

    def toString(self) -> str:
        return self.separator.join(p.pickle() for p in self.items)

    def get(self, key: str) -> Picklable:
        for p in self.items:
            if p.key == key:
                return p
        raise KeyError("not found")


# ---------------------------------------------------------------------------------------------------------------
# Binary RNS container (no reference counterpart: cuhe/Utils.cu:75-152 only has the decimal text form, which costs
# ~2.4 bytes per bit and a big-integer parse per coefficient).  Carries residue-domain data exactly as the device
# holds it -- CRT domain u32[rows][crtLen] or NTT domain u64[rows][...][nttLen], e.g. the transformed evaluation keys
# of cuhe_relin_export_host / cuhe_relin_import_host -- with the parameter tuple that fixes its meaning and a
# checksum.  Little-endian, 80-byte header:
#   0  magic "CUHERNS1"       8  d, p, w, min, cut, m (6 x i32)     32  domain (2 = CRT u32, 3 = NTT u64), level (i32)
#   40 shard_rank, shard_world (i32)      48  ndim (u32), dims[3] (u32)      64  payload bytes (u64)    72  64-bit checksum of the payload (rns_checksum)
# The C++ twin is cuHE_Utils::RnsBlob (cuhe_b200/host/cuhe_utils.hpp); files are interchangeable.
# ---------------------------------------------------------------------------------------------------------------
import struct as _struct

RNS_MAGIC = b"CUHERNS1"
_RNS_HEADER = _struct.Struct("<8s6i2i2iI3IQQ")


def rns_checksum(data: bytes) -> int:
    """64-bit checksum of a payload, cheap enough for gigabyte key files in both languages: the payload is read as
    little-endian u64 words w_i (zero-padded to a multiple of 8 bytes); every block of 2^16 words is folded to
    x = XOR_i (w_i * (2i + 1) mod 2^64) (i = global word index) and absorbed by one FNV-1a style round
    h = (h ^ x) * 0x100000001B3 mod 2^64, starting from h = 0xCBF29CE484222325."""
    import numpy as np
    buf = np.frombuffer(data, dtype=np.uint8)
    pad = (-len(buf)) % 8
    if pad:
        buf = np.concatenate([buf, np.zeros(pad, dtype=np.uint8)])
    words = buf.view("<u8")
    h = 0xCBF29CE484222325
    prime = 0x100000001B3
    mask = (1 << 64) - 1
    block = 1 << 16
    for off in range(0, len(words), block):
        w = words[off:off + block]
        idx = np.arange(off, off + len(w), dtype=np.uint64)
        mixed = np.bitwise_xor.reduce(w * ((idx << np.uint64(1)) | np.uint64(1)))     # wraps mod 2^64
        h = ((h ^ int(mixed)) * prime) & mask
    return h


def save_rns(path: str, array, params: Sequence[int], domain: int, level: int = 0, shard=(0, 1)) -> None:
    import numpy as np
    a = np.ascontiguousarray(array)
    if domain == 2:
        a = a.view(np.uint32) if a.dtype.itemsize == 4 else a.astype(np.uint32)
    elif domain == 3:
        a = a.view(np.uint64) if a.dtype.itemsize == 8 else a.astype(np.uint64)
    else:
        raise ValueError("domain must be 2 (CRT, u32) or 3 (NTT, u64)")
    if not 1 <= a.ndim <= 3:
        raise ValueError("1 to 3 dimensions")
    dims = list(a.shape) + [1] * (3 - a.ndim)
    payload = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    head = _RNS_HEADER.pack(RNS_MAGIC, *[int(v) for v in params], int(domain), int(level), int(shard[0]), int(shard[1]),
                            a.ndim, *dims, len(payload), rns_checksum(payload))
    with open(path, "wb") as f:
        f.write(head)
        f.write(payload)


def load_rns(path: str):
    """-> (array, meta) with meta = dict(params, domain, level, shard); raises ValueError on a damaged file"""
    import numpy as np
    with open(path, "rb") as f:
        head = f.read(_RNS_HEADER.size)
        if len(head) != _RNS_HEADER.size:
            raise ValueError("truncated RNS file")
        magic, d, p, w, mn, cut, m, domain, level, sr, sw, ndim, d0, d1, d2, nbytes, digest = _RNS_HEADER.unpack(head)
        if magic != RNS_MAGIC:
            raise ValueError("not a CUHERNS1 file")
        payload = f.read()
    if len(payload) != nbytes:
        raise ValueError("truncated RNS payload")
    if rns_checksum(payload) != digest:
        raise ValueError("RNS payload checksum mismatch")
    dt = np.dtype("<u4") if domain == 2 else np.dtype("<u8")
    dims = [d0, d1, d2][:ndim]
    a = np.frombuffer(payload, dtype=dt).reshape(dims)
    return a, dict(params=(d, p, w, mn, cut, m), domain=domain, level=level, shard=(sr, sw))
