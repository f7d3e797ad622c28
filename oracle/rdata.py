"""Minimal reader for R's XDR serialization (`save()` / `.RData`, format versions 2 and 3).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`derfinder_b200/`) may import this
module; it exists so that the reference's own golden fixtures
(`/root/reference/data/*.RData`, `/root/reference/inst/extdata/chr21/*.Rdata`) can be turned into
the small `.npz/.json` files committed under `tests/golden/` (see `tests/golden/make_golden.py`).

The reader understands exactly the SEXP types that occur in those files (no ALTREP, no
byte-code, no environments other than the global / namespace pseudo-SEXPs).  R objects are
returned as :class:`RObj` (value + attribute dict); S4 instances carry their slots as attributes,
which is how R serialises them.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from dataclasses import dataclass, field
from typing import Any

import numpy as np

NA_INTEGER = -2147483648

# SEXP type codes (R internals, Rinternals.h / serialize.c)
NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP = 0, 1, 2, 3, 4, 5, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP = 9, 10, 13, 14, 15, 16
DOTSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, RAWSXP, S4SXP = 17, 19, 20, 21, 22, 24, 25
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP = 255, 254, 253, 252
MISSINGARG_SXP, BASENAMESPACE_SXP, NAMESPACESXP, PACKAGESXP = 251, 250, 249, 248
PERSISTSXP, EMPTYENV_SXP, BASEENV_SXP = 247, 242, 241
ATTRLANGSXP, ATTRLISTSXP = 240, 239


@dataclass
class RObj:
    """An R value together with its attributes (slots for S4 objects)."""

    value: Any
    attr: dict = field(default_factory=dict)
    sexptype: int = -1

    def __getitem__(self, key):
        return self.attr[key]

    @property
    def rclass(self):
        c = self.attr.get("class")
        if c is None:
            return None
        v = c.value if isinstance(c, RObj) else c
        return v[0] if len(v) else None


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.refs: list = []

    def _int(self) -> int:
        (v,) = struct.unpack_from(">i", self.buf, self.pos)
        self.pos += 4
        return v

    def _length(self) -> int:
        n = self._int()
        if n == -1:  # long vector
            hi, lo = self._int(), self._int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def _bytes(self, n: int) -> bytes:
        b = self.buf[self.pos : self.pos + n]
        self.pos += n
        return b

    def header(self):
        magic = self._bytes(5)
        if magic not in (b"RDX2\n", b"RDX3\n"):
            raise ValueError(f"not an XDR .RData stream: {magic!r}")
        fmt = self._bytes(2)
        if fmt != b"X\n":
            raise ValueError(f"unsupported serialization format {fmt!r}")
        version = self._int()
        self._int()  # writer R version
        self._int()  # min reader version
        if version == 3:
            n = self._int()
            self._bytes(n)  # native encoding
        elif version != 2:
            raise ValueError(f"unsupported serialization version {version}")

    def item(self):
        flags = self._int()
        t = flags & 0xFF
        is_obj = bool(flags & (1 << 8))
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        del is_obj

        if t == NILVALUE_SXP or t == NILSXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, BASENAMESPACE_SXP):
            return RObj(None, {}, t)
        if t in (MISSINGARG_SXP, UNBOUNDVALUE_SXP):
            return RObj(None, {}, t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self._int()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()  # CHARSXP
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self._int()  # always 0
            n = self._int()
            info = [self.item() for _ in range(n)]
            obj = RObj(info, {}, t)
            self.refs.append(obj)
            return obj
        if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP, CLOSXP, PROMSXP, DOTSXP):
            # pairlist-like: iterate instead of recursing on the CDR
            if t == ATTRLISTSXP:
                t, has_attr = LISTSXP, True
            if t == ATTRLANGSXP:
                t, has_attr = LANGSXP, True
            items = []
            attr = {}
            first = True
            while True:
                if not first:
                    flags = self._int()
                    tt = flags & 0xFF
                    if tt == NILVALUE_SXP:
                        break
                    if tt not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                        # dotted pair ending in a non-list CDR: rewind and read it
                        self.pos -= 4
                        items.append((None, self.item()))
                        break
                    has_attr = bool(flags & (1 << 9)) or tt in (ATTRLISTSXP, ATTRLANGSXP)
                    has_tag = bool(flags & (1 << 10))
                first = False
                a = self._attributes() if has_attr else {}
                if a and not attr:
                    attr = a
                tag = self.item() if has_tag else None
                car = self.item()
                items.append((tag, car))
            return RObj(items, attr, t)
        if t == CHARSXP:
            n = self._int()
            if n == -1:
                return None  # NA_character_
            return self._bytes(n).decode("utf-8", errors="replace")
        if t == LGLSXP or t == INTSXP:
            n = self._length()
            v = np.frombuffer(self.buf, dtype=">i4", count=n, offset=self.pos).astype(np.int32)
            self.pos += 4 * n
            val: Any = v
        elif t == REALSXP:
            n = self._length()
            v = np.frombuffer(self.buf, dtype=">f8", count=n, offset=self.pos).astype(np.float64)
            self.pos += 8 * n
            val = v
        elif t == CPLXSXP:
            n = self._length()
            v = np.frombuffer(self.buf, dtype=">c16", count=n, offset=self.pos).astype(np.complex128)
            self.pos += 16 * n
            val = v
        elif t == STRSXP:
            n = self._length()
            val = [self.item() for _ in range(n)]
        elif t in (VECSXP, EXPRSXP):
            n = self._length()
            val = [self.item() for _ in range(n)]
        elif t == RAWSXP:
            n = self._length()
            val = self._bytes(n)
        elif t == S4SXP:
            val = None
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.pos}")
        attr = self._attributes() if has_attr else {}
        return RObj(val, attr, t)

    def _attributes(self) -> dict:
        pl = self.item()
        out = {}
        if pl is None:
            return out
        for tag, car in pl.value:
            out[tag] = car
        return out


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


def read_rdata(path: str) -> dict:
    """Return {name: RObj} for every object stored by `save()` in `path`."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    rd = _Reader(buf)
    rd.header()
    top = rd.item()
    out = {}
    if top is not None:
        for tag, car in top.value:
            out[tag] = car
    return out


# ---------------------------------------------------------------------------------------------
# Helpers that turn the S4 containers used by derfinder into numpy


def strings(obj) -> list:
    return list(obj.value) if isinstance(obj, RObj) else list(obj)


def rle_parts(obj: RObj):
    """(values, lengths) of an S4Vectors::Rle (factor-Rle values come back as integer codes)."""
    vals = obj.attr["values"]
    lens = obj.attr["lengths"]
    v = vals.value if isinstance(vals, RObj) else vals
    return np.asarray(v), np.asarray(lens.value, dtype=np.int64)


def rle_expand(obj: RObj) -> np.ndarray:
    v, l = rle_parts(obj)
    return np.repeat(v, l)


def is_rle(obj) -> bool:
    return isinstance(obj, RObj) and "values" in obj.attr and "lengths" in obj.attr


def rlist(obj: RObj) -> dict:
    """Named R list -> dict."""
    names = obj.attr.get("names")
    keys = strings(names) if names is not None else [str(i) for i in range(len(obj.value))]
    return dict(zip(keys, obj.value))


def dframe_columns(obj: RObj) -> dict:
    """S4Vectors::DataFrame/DFrame -> {colname: RObj}."""
    return rlist(obj.attr["listData"])


def as_vector(obj) -> np.ndarray:
    """Plain vector or Rle -> dense numpy array."""
    if is_rle(obj):
        return rle_expand(obj)
    return np.asarray(obj.value)
