"""Minimal reader for R's RDS serialization (XDR, version 2/3).

TEST INFRASTRUCTURE ONLY.  Used by `tests/golden/make_golden.py` (run in the
build container where `/root/reference` is mounted) to turn the reference's
shipped `*.rds` objects into small `.npz` golden vectors, and by the optional
tests that read `/root/reference` directly when it is present.  Nothing in the
product package (`clonalscope_b200/`) may import this module.

Format notes follow SURVEY.md Appendix C.  Only the SEXP types that occur in
the reference's fixtures are handled: NILVALUE, REFSXP, SYMSXP, CHARSXP,
LISTSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP and ALTREP (wrap_integer,
wrap_real, wrap_string, compact_intseq, compact_realseq, deferred_string).

R objects map to Python as:
  * atomic vectors  -> RVector(numpy array, attrs dict)   (strings: object array, NA -> None)
  * lists (VECSXP)  -> RList(list of values, attrs dict)   (names in attrs['names'])
  * pairlists       -> dict tag -> value (only used for attributes)
"""
from __future__ import annotations

import gzip
import bz2
import lzma
import struct

import numpy as np

NA_INTEGER = -2147483648


class RVector:
    __slots__ = ("data", "attrs")

    def __init__(self, data, attrs=None):
        self.data = data
        self.attrs = attrs or {}

    @property
    def names(self):
        n = self.attrs.get("names")
        return None if n is None else list(n.data)

    def __repr__(self):
        return f"RVector({self.data.dtype}, len={len(self.data)}, attrs={list(self.attrs)})"


class RList:
    __slots__ = ("items", "attrs")

    def __init__(self, items, attrs=None):
        self.items = items
        self.attrs = attrs or {}

    @property
    def names(self):
        n = self.attrs.get("names")
        return None if n is None else list(n.data)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.items[self.names.index(key)]
        return self.items[key]

    def __contains__(self, key):
        n = self.names
        return n is not None and key in n

    def __len__(self):
        return len(self.items)

    def __repr__(self):
        return f"RList(len={len(self.items)}, names={self.names})"


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.refs = []

    def i32(self):
        (v,) = struct.unpack_from(">i", self.buf, self.pos)
        self.pos += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def take(self, n):
        b = self.buf[self.pos:self.pos + n]
        self.pos += n
        return b

    def attrs_of(self, pl):
        return pl if isinstance(pl, dict) else {}

    def item(self, flags=None):
        if flags is None:
            flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))

        if typ == 254:  # NILVALUE
            return None
        if typ in (253, 252, 251, 250, 249, 248, 247, 242, 241):
            # global env / unbound / missing / base namespace / empty env / base env ...
            return None
        if typ == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.take(n).decode("utf-8", errors="replace")
        if typ in (2, 6):  # LISTSXP / LANGSXP -> dict of tag->value (attributes)
            out = {}
            anon = 0
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                if tag is None:
                    tag = f"__{anon}"
                    anon += 1
                out[tag] = car
                flags = self.i32()
                typ2 = flags & 0xFF
                if typ2 == 254:
                    break
                if typ2 not in (2, 6):
                    # dotted pair (e.g. ALTREP wrapper state CONS(x, meta))
                    out["__cdr"] = self.item(flags)
                    break
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if typ == 10 or typ == 13:  # LGLSXP / INTSXP
            n = self.length()
            data = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RVector(data, attrs)
        if typ == 14:  # REALSXP
            n = self.length()
            data = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RVector(data, attrs)
        if typ == 16:  # STRSXP
            n = self.length()
            data = np.empty(n, dtype=object)
            for i in range(n):
                data[i] = self.item()
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RVector(data, attrs)
        if typ == 19 or typ == 20:  # VECSXP / EXPRSXP
            n = self.length()
            items = [self.item() for _ in range(n)]
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RList(items, attrs)
        if typ == 238:  # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = next(iter(info.values())) if isinstance(info, dict) else info
            return self.altrep(cls, state, self.attrs_of(attr))
        if typ == 24:  # RAWSXP
            n = self.length()
            data = np.frombuffer(self.take(n), dtype=np.uint8).copy()
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RVector(data, attrs)
        raise ValueError(f"unsupported SEXP type {typ} at byte {self.pos}")

    def altrep(self, cls, state, attrs):
        if cls in ("wrap_integer", "wrap_real", "wrap_string", "wrap_logical"):
            # state = list(x, meta)
            inner = state.items[0] if isinstance(state, RList) else next(iter(state.values()))
            merged = dict(inner.attrs)
            merged.update(attrs)
            return RVector(inner.data, merged)
        if cls == "compact_intseq":
            n, start, step = (int(v) for v in state.data[:3])
            return RVector((start + step * np.arange(n)).astype(np.int32), attrs)
        if cls == "compact_realseq":
            n, start, step = state.data[:3]
            return RVector(start + step * np.arange(int(n), dtype=np.float64), attrs)
        if cls == "deferred_string":
            src = next(iter(state.values())) if isinstance(state, dict) else state.items[0]
            out = np.empty(len(src.data), dtype=object)
            for i, v in enumerate(src.data):
                if src.data.dtype.kind == "i":
                    out[i] = None if v == NA_INTEGER else str(int(v))
                else:
                    out[i] = format_r_number(float(v))
            return RVector(out, attrs)
        raise ValueError(f"unsupported ALTREP class {cls}")


def format_r_number(v: float) -> str:
    """as.character(double) with R's default 15 significant digits."""
    if v != v:
        return "NA"
    if v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return repr(float(f"{v:.15g}"))


def read_rds(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    if raw[:2] != b"X\n":
        raise ValueError("not an XDR RDS stream")
    r = _Reader(raw)
    r.pos = 2
    version = r.i32()
    r.i32()  # writer version
    r.i32()  # min reader version
    if version == 3:
        n = r.i32()
        r.take(n)  # native encoding
    elif version != 2:
        raise ValueError(f"unsupported serialization version {version}")
    return r.item()


def as_matrix(v: RVector):
    """R matrix (column-major with dim attr) -> numpy 2-D array + (rownames, colnames)."""
    dim = v.attrs["dim"].data
    m = np.asarray(v.data).reshape((int(dim[0]), int(dim[1])), order="F")
    rn = cn = None
    dn = v.attrs.get("dimnames")
    if dn is not None:
        rn = None if dn.items[0] is None else list(dn.items[0].data)
        cn = None if dn.items[1] is None else list(dn.items[1].data)
    return m, rn, cn
