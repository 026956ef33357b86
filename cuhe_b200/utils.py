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
