"""File formats on either side of the path: `saveRDS` for the objects the reference's drivers write
(R/RunCovCluster.R:89,150,191,230 — deltas_allseg.rds, nonpara_clustering*.rds;
R/CreateSegtableNoWGS.R:57,102,159; R/Segmentation_bulk.R:256) and `readMM` for the 10x
MatrixMarket count matrix the tutorials start from (samples/P5931/scRNA/README.md:72).

saveRDS writes R's XDR serialization, format version 2, gzip-compressed — what `readRDS()` of any
R >= 2.3 reads.  Python objects map to R as: dict -> named list, list/tuple -> list, 1-D float /
int / bool / str arrays -> numeric / integer / logical / character vectors (NaN stays NaN, None
in a string vector -> NA_character_), 2-D arrays -> matrices (column-major + dim), NamedMatrix ->
matrix with dimnames, RDataFrame -> data.frame, (names, values) pairs as EstRegionCov returns
them -> named numeric vectors, None -> NULL.  No arithmetic happens here.

readMM parses the coordinate text on the host (C++ threads over the mapped file, in the shared
library) and builds the dgCMatrix slots on the device (cs_mm_read_*: column histogram, scan,
order check, and a key sort only if the file is not already column-major).
"""
from __future__ import annotations

import ctypes as C
import gzip
import struct

import numpy as np

from . import _lib
from .api import NamedMatrix, _check, _ptr

NA_INTEGER = -2147483648
_NIL, _SYM, _LIST, _CHAR, _LGL, _INT, _REAL, _STR, _VEC, _REF = 254, 1, 2, 9, 10, 13, 14, 16, 19, 255
_HAS_ATTR, _IS_OBJ, _HAS_TAG = 0x200, 0x100, 0x400


class RDataFrame(dict):
    """A dict of equal-length columns that saveRDS writes as a data.frame."""


class _Writer:
    def __init__(self):
        self.out = []
        self.syms = {}

    def i32(self, v):
        self.out.append(struct.pack(">i", int(v)))

    def chars(self, s):
        if s is None:
            self.i32(_CHAR)
            self.i32(-1)
            return
        b = str(s).encode("utf-8")
        ascii_only = all(c < 128 for c in b)
        self.i32(_CHAR | ((64 if ascii_only else 8) << 12))       # ASCII_MASK / UTF8_MASK in the gp field
        self.i32(len(b))
        self.out.append(b)

    def symbol(self, name):
        if name in self.syms:
            self.i32((self.syms[name] << 8) | _REF)
            return
        self.syms[name] = len(self.syms) + 1
        self.i32(_SYM)
        self.chars(name)

    def attrs(self, attrs):
        for k, v in attrs.items():
            self.i32(_LIST | _HAS_TAG)
            self.symbol(k)
            self.item(v)
        self.i32(_NIL)

    def header(self, sxp, attrs, is_obj=False):
        self.i32(sxp | (_HAS_ATTR if attrs else 0) | (_IS_OBJ if is_obj else 0))

    def vector(self, a, attrs=None, is_obj=False):
        attrs = attrs or {}
        a = np.asarray(a)
        if a.dtype.kind in "US" or a.dtype == object:
            vals = a.ravel(order="F").tolist()
            self.header(_STR, attrs, is_obj)
            self.i32(len(vals))
            for s in vals:
                self.chars(None if s is None else s)
        elif a.dtype.kind == "b":
            self.header(_LGL, attrs, is_obj)
            self.i32(a.size)
            self.out.append(a.ravel(order="F").astype(">i4").tobytes())
        elif a.dtype.kind in "iu":
            self.header(_INT, attrs, is_obj)
            self.i32(a.size)
            self.out.append(a.ravel(order="F").astype(">i4").tobytes())
        else:
            self.header(_REAL, attrs, is_obj)
            self.i32(a.size)
            self.out.append(a.ravel(order="F").astype(">f8").tobytes())
        if attrs:
            self.attrs(attrs)

    def item(self, obj):
        if obj is None:
            self.i32(_NIL)
        elif isinstance(obj, NamedMatrix):
            v = obj.values
            v = np.asarray(v.todense()) if hasattr(v, "todense") else np.asarray(v)
            dn = [None if n is None else np.asarray(n, dtype=object) for n in (obj.rownames, obj.colnames)]
            self.vector(v, {"dim": np.asarray(v.shape, dtype=np.int32), "dimnames": dn})
        elif isinstance(obj, RDataFrame):
            cols = list(obj.keys())
            n = len(next(iter(obj.values()))) if cols else 0
            attrs = {"names": np.asarray(cols, dtype=object), "class": np.asarray(["data.frame"], dtype=object),
                     "row.names": np.asarray([NA_INTEGER, -n], dtype=np.int32)}
            self.header(_VEC, attrs, is_obj=True)
            self.i32(len(cols))
            for c in cols:
                self.item(np.asarray(obj[c]) if not isinstance(obj[c], np.ndarray) else obj[c])
            self.attrs(attrs)
        elif isinstance(obj, dict):
            attrs = {"names": np.asarray([str(k) for k in obj.keys()], dtype=object)} if obj else {}
            self.header(_VEC, attrs)
            self.i32(len(obj))
            for v in obj.values():
                self.item(v)
            if attrs:
                self.attrs(attrs)
        elif isinstance(obj, tuple) and len(obj) == 2 and _is_named_vector(obj):
            names, vals = obj
            self.vector(np.asarray(vals), {"names": np.asarray(list(names), dtype=object)})
        elif isinstance(obj, (list, tuple)) and not _is_atomic_list(obj):
            self.header(_VEC, {})
            self.i32(len(obj))
            for v in obj:
                self.item(v)
        else:
            a = np.asarray(obj)
            if a.ndim >= 2:
                self.vector(a, {"dim": np.asarray(a.shape, dtype=np.int32)})
            else:
                self.vector(a.reshape(-1))


def _is_atomic_list(obj):
    return len(obj) > 0 and all(isinstance(v, (str, int, float, bool, np.generic)) or v is None for v in obj) \
        and not all(v is None for v in obj)


def _is_named_vector(pair):
    names, vals = pair
    try:
        return (not isinstance(names, str)) and len(names) == len(vals) and np.asarray(vals).ndim == 1 \
            and np.asarray(vals).dtype.kind in "fiub" and all(isinstance(n, (str, np.str_)) for n in list(names)[:8])
    except TypeError:
        return False


def serialize(obj) -> bytes:
    """R's serialize(obj, NULL, xdr = TRUE, version = 2) for the Python stand-ins listed in the module docstring."""
    w = _Writer()
    w.out.append(b"X\n")
    w.i32(2)
    w.i32(0x00030600)                                             # written "by" R 3.6.0
    w.i32(0x00020300)                                             # readable from R 2.3.0
    w.item(obj)
    return b"".join(w.out)


def saveRDS(obj, file, compress=True):
    data = serialize(obj)
    if compress:
        with gzip.open(file, "wb", compresslevel=6) as f:
            f.write(data)
    else:
        with open(file, "wb") as f:
            f.write(data)


# --------------------------------------------------------------------------- MatrixMarket
def readMM(file):
    """Matrix::readMM for `coordinate` real / integer / pattern general matrices (plain or .gz):
    returns a scipy.sparse.csc_matrix (the dgCMatrix the tutorials convert to), duplicates summed."""
    import scipy.sparse as sp

    L = _lib.lib()
    job = C.c_void_p()
    dims = (C.c_int64 * 3)()
    _check(L.cs_mm_read_begin(str(file).encode(), C.byref(job), dims))
    G, N, nnz = int(dims[0]), int(dims[1]), int(dims[2])
    colptr = np.zeros(N + 1, dtype=np.int32)
    rowidx = np.zeros(max(nnz, 1), dtype=np.int32)
    x = np.zeros(max(nnz, 1), dtype=np.float64)
    _check(L.cs_mm_read_fetch(job, _ptr(colptr, C.c_int32), _ptr(rowidx, C.c_int32), _ptr(x, C.c_double)))
    return sp.csc_matrix((x[:nnz], rowidx[:nnz], colptr), shape=(G, N))
