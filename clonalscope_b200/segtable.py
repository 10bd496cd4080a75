"""Host-side mirror of the callers of the gene -> bin aggregation: `Createobj_bulk`,
`Segmentation_bulk` and `CreateSegtableNoWGS` (reference R/Createobj_bulk.R:33-100,
R/Segmentation_bulk.R:22-218, R/CreateSegtableNoWGS.R:31-169).

Device work (libclonalscope_b200.so): the bin x cell counts (cs_bin_counts_begin), their pooling
into bin x cluster profiles where they lie on the device (cs_bin_job_rowsum_by_label) and, for all
clusters and chromosomes in ONE launch, the running mean + Viterbi path of the 4-state HMM
(cs_segment_hmm).  What stays here is what stays in the R shim: name parsing, the quantile / median
clipping of one pooled vector, the table of strings the reference returns and the union of the
clusters' breakpoints.  Plots and saveRDS side effects are not reproduced (`plot_seg` is accepted
and ignored; nothing is written unless `dir_path` is given, and then as .npz).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib
from .api import ClonalscopeError, NamedMatrix, _check, _first_column, _ptr, _bed_columns, gene_bin_map
from .intervals import r_as_character

_DELTA = (0.1, 0.2, 0.5, 0.2)                                  # R/Segmentation_bulk.R:159


# ------------------------------------------------------------------ device calls
def segment_hmm(seqs, nmean, means, sds, delta, Pi):
    """cs_segment_hmm on a list of 1-D arrays: (list of smoothed arrays, list of 0-based state arrays)."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    val = np.ascontiguousarray(np.concatenate(seqs) if seqs else np.zeros(0), dtype=np.float64)
    total = int(off[-1])
    smooth = np.zeros(max(total, 1))
    state = np.zeros(max(total, 1), dtype=np.int32)
    means = np.ascontiguousarray(means, dtype=np.float64)
    sds = np.ascontiguousarray(sds, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    Pi = np.ascontiguousarray(Pi, dtype=np.float64).reshape(16)
    _check(_lib.lib().cs_segment_hmm(len(seqs), _ptr(off, C.c_int64), _ptr(val, C.c_double), int(nmean),
                                     _ptr(means, C.c_double), _ptr(sds, C.c_double), _ptr(delta, C.c_double),
                                     _ptr(Pi, C.c_double), _ptr(smooth, C.c_double), _ptr(state, C.c_int)))
    return ([smooth[off[i]:off[i + 1]] for i in range(len(seqs))],
            [state[off[i]:off[i + 1]] for i in range(len(seqs))])


def rowsum_by_label(bin_mtx, label, K):
    """cs_rowsum_by_label on a bin x cell scipy CSC matrix: B x K pooled counts."""
    import scipy.sparse as sp

    m = sp.csc_matrix(bin_mtx)
    B, N = m.shape
    label = np.ascontiguousarray(label, dtype=np.int32)
    out = np.zeros((K, B))
    _check(_lib.lib().cs_rowsum_by_label(B, N, _ptr(np.ascontiguousarray(m.indptr, dtype=np.int32), C.c_int32),
                                         _ptr(np.ascontiguousarray(m.indices, dtype=np.int32), C.c_int32),
                                         _ptr(np.ascontiguousarray(m.data, dtype=np.float64), C.c_double),
                                         _ptr(label, C.c_int), int(K), _ptr(out, C.c_double)))
    return out.T                                               # column-major B x K


# ------------------------------------------------------------------ R helpers
def _quantile7(v, p):
    s = np.sort(v)
    h = (len(s) - 1) * p
    lo = int(math.floor(h))
    hi = min(lo + 1, len(s) - 1)
    return float(s[lo] + (h - lo) * (s[hi] - s[lo]))


def _median(v):
    s = np.sort(v)
    n = len(s)
    return float(s[n // 2]) if n % 2 else float((s[n // 2 - 1] + s[n // 2]) / 2)


def _r_mean(v):
    """mean() of R: long double sum, then one refinement pass (summary.c)."""
    v = np.asarray(v, dtype=np.longdouble)
    n = len(v)
    m = v.sum() / n
    m = m + (v - m).sum() / n
    return float(m)


def _r_var(v):
    """var() of R: two-pass, long double accumulators (cov.c), NA for fewer than two values."""
    n = len(v)
    if n < 2:
        return math.nan
    v = np.asarray(v, dtype=np.longdouble)
    m = v.sum() / n
    m = m + (v - m).sum() / n
    d = v - m
    return float((d * d).sum() / (n - 1))


def _strip_chr(name):
    name = str(name)
    return name.split("hr", 1)[1] if "chr" in name else name


def _counts_column(obj):
    """(values 1-D float64, rownames) of a one-column count matrix."""
    if isinstance(obj, NamedMatrix):
        v = obj.values
        v = np.asarray(v.todense()) if hasattr(v, "todense") else np.asarray(v)
        return np.asarray(v, dtype=np.float64).reshape(len(obj.rownames), -1).sum(axis=1), list(obj.rownames)
    if isinstance(obj, dict):
        return _counts_column(NamedMatrix(np.asarray(obj["values"]), obj["rownames"], obj.get("colnames", ["V1"])))
    if hasattr(obj, "index"):
        return np.asarray(obj.iloc[:, 0], dtype=np.float64), [str(i) for i in obj.index]
    raise ClonalscopeError(_lib.CS_ERR_ARG, "counts must carry row names formatted as chr1-1-20000")


# ------------------------------------------------------------------ Createobj_bulk
def Createobj_bulk(raw_counts=None, ref_counts=None, samplename="sample", genome_assembly="GRCh38", dir_path="./",
                   size=None, assay="WGS"):
    """The object Segmentation_bulk takes (reference R/Createobj_bulk.R:33-100): checks, numeric
    ordering of both count vectors by chromosome and start, `size` restricted to the chromosomes
    present and ordered numerically (a dict chromosome-name -> length)."""
    if size is None:
        raise ValueError("Please provide a matrix/ data.frame with two columns: col1: different chromosome; "
                         "col2: for the size (bp) of different chromosomes.")
    raw_v, raw_n = _counts_column(raw_counts)
    ref_v, ref_n = _counts_column(ref_counts)
    for names, what in ((raw_n, "raw_counts"), (ref_n, "ref_counts")):
        if len(names) > 1 and len(str(names[1]).split("-")) != 3:
            raise ValueError(f"The rownames for the {what} matrix should be formatted as: chr1-1-20000.")
        if len(names) == 0:
            raise ValueError(f"{what} matrix is not valid.")

    def reorder(v, names, label):
        chrs = [_strip_chr(str(n).split("-")[0]) for n in names]
        uniq = list(dict.fromkeys(chrs))
        if uniq != sorted(uniq, key=float):
            print(f"WGS of {label} sample not ordered, reordering numerically...")
            starts = [float(str(n).split("-")[1]) for n in names]
            order = sorted(range(len(names)), key=lambda i: (float(chrs[i]), starts[i]))
            return v[order], [names[i] for i in order]
        return v, names

    raw_v, raw_n = reorder(raw_v, raw_n, "tumor")
    ref_v, ref_n = reorder(ref_v, ref_n, "normal")
    if isinstance(size, dict):
        size_names, size_vals = list(size.keys()), list(size.values())
    else:
        size_names, size_vals = _first_column(size), list(np.asarray(size, dtype=object)[:, 1]) \
            if not hasattr(size, "iloc") else list(size.iloc[:, 1])
    size_names = [_strip_chr(n) for n in size_names] if len(size_names) > 1 and "chr" in str(size_names[1]) \
        else [str(n) for n in size_names]
    present = {_strip_chr(str(n).split("-")[0]) for n in raw_n}
    keep = [i for i, n in enumerate(size_names) if n in present]
    keep.sort(key=lambda i: float(size_names[i]))
    size_out = {size_names[i]: float(size_vals[i]) for i in keep}
    print("Object successfully created!")
    return {"raw_counts": NamedMatrix(raw_v.reshape(-1, 1), raw_n, ["V1"]),
            "ref_counts": NamedMatrix(ref_v.reshape(-1, 1), ref_n, ["V1"]), "size": size_out,
            "samplename": samplename, "dir_path": dir_path, "genome_assembly": genome_assembly,
            "seg_table": None, "seg_table_filtered": None, "assay": assay}


# ------------------------------------------------------------------ Segmentation_bulk
def _prepare_profile(raw, ref, raw_names, size, max_qt, rm_extreme):
    """:72-147 for one pooled profile: clipped, filtered coverage and the kept bins' coordinates."""
    raw_chr = [str(n).split("-")[0] for n in raw_names]
    raw_start = np.array([float(str(n).split("-")[1]) for n in raw_names])
    size_names = list(size.keys())
    chr_name = [n if "chr" in size_names[0] else "chr" + n for n in size_names]      # :64-68
    # pooled vectors in chr_name order (:92-98)
    order = [i for c in chr_name for i in range(len(raw_chr)) if raw_chr[i] == c]
    with np.errstate(divide="ignore", invalid="ignore"):
        cov2 = np.asarray(raw, dtype=np.float64)[order] / np.asarray(ref, dtype=np.float64)[order]
    if len(cov2) != len(raw_chr):
        # the reference pairs chr_sum (chr_name order, known chromosomes only) with chromnum (all rows)
        raise ClonalscopeError(_lib.CS_ERR_ARG, "raw_counts has bins on chromosomes that `size` does not name")
    cov2[np.isnan(cov2)] = 0.0
    cov2[cov2 == np.inf] = 100.0
    cov3 = np.minimum(cov2, _quantile7(cov2, max_qt))
    cov4 = cov3 / _median(cov3)
    if rm_extreme == 1:
        sel = np.nonzero(cov4 > 0.4)[0]
    elif rm_extreme == 2:
        sel = np.nonzero((cov4 > 0.4) & (cov2 != 100.0))[0]
    else:
        raise NameError("object 'sel_bin' not found")
    chromnum = np.array([float(_strip_chr(raw_chr[i])) for i in sel])                # (:137-140, row order)
    return cov3[sel], chromnum, raw_start[sel]


def _segment_profiles(profiles, raw_names, size, hmm_states, hmm_sd, hmm_p, nmean, adj, max_qt, rm_extreme, assay):
    """Segmentation_bulk's body for several pooled profiles against one bin list: one device call for
    every (profile, chromosome) sequence.  Returns [(seg_table, annot)]."""
    if len(hmm_states) != 3:
        raise ValueError("hmm_states should be an ordered vector for the three HMM numeric states.")
    if not isinstance(hmm_sd, (int, float)):
        raise ValueError("hmm_sd should be a numeric value.")
    if not isinstance(hmm_p, (int, float)):
        raise ValueError("hmn_p should be an ordered vector for the three HMM numeric states.")
    if assay == "scRNA":
        adj = -0.5
    pp = [float(hmm_states[2]), float(hmm_states[1]), 1.0, float(hmm_states[0])]    # ppa1, ppa2, ppn, ppd
    t = float(hmm_p)
    Pi = np.full((4, 4), t)
    np.fill_diagonal(Pi, 1 - 3 * t)
    chroms = [_strip_chr(n) for n in size.keys()]
    size_of = {_strip_chr(n): float(v) for n, v in size.items()}
    seqs, meta = [], []
    for p, (raw, ref) in enumerate(profiles):
        cov3, chromnum, start = _prepare_profile(raw, ref, raw_names, size, max_qt, rm_extreme)
        med = _median(cov3)
        for ii in chroms:
            idx = np.nonzero(chromnum == float(ii))[0]
            if len(idx) == 0:
                raise ClonalscopeError(_lib.CS_ERR_ARG, f"no bin left on chromosome {ii} after filtering")
            c4 = cov3[idx] / med + adj
            seqs.append(c4)
            meta.append((p, ii, start[idx] - 1))
    _, states = segment_hmm(seqs, nmean, pp, [float(hmm_sd)] * 4, _DELTA, Pi)
    ppv = np.array(pp)
    out = []
    for p in range(len(profiles)):
        rows = []
        for (q, ii, start), c4, st in zip(meta, seqs, states):
            if q != p:
                continue
            sv = ppv[st]
            cps = [x for x in (np.nonzero(sv[1:] != sv[:-1])[0] + 2).tolist() if x != 2]       # 1-based, :186-187
            seg_start = [start[0]] + [start[c - 1] for c in cps]
            seg_end = [start[c - 1] for c in cps] + [size_of[ii]]
            seg_state = [sv[1] if len(sv) > 1 else math.nan] + [sv[c - 1] for c in cps]
            edges = [1] + cps + [len(c4)]
            for s in range(len(cps) + 1):
                a, b = edges[s], edges[s + 1] - 1                                               # R's a:b
                vals = c4[a - 1:b] if b >= a else c4[b - 1:a][::-1]
                rows.append((ii, seg_start[s], seg_end[s], seg_state[s], seg_end[s] - seg_start[s],
                             _r_mean(vals), _r_var(vals)))
        tab = {"chr": [r[0] for r in rows]}
        for j, col in enumerate(("start", "end", "states", "length", "mean", "var"), start=1):
            tab[col] = [r_as_character(float(r[j])) for r in rows]
        annot = None
        if assay == "scRNA":                                                                    # :208-216
            if _median(np.array([float(v) for v in tab["states"]])) == 0.5:
                tab["states"] = [r_as_character(float(v) + 0.5) for v in tab["states"]]
                annot = "N"
            else:
                annot = "T"
        tab["chrr"] = [f"{c}:{s}" for c, s in zip(tab["chr"], tab["start"])]
        out.append((tab, annot))
    return out


def Segmentation_bulk(Obj_filtered=None, hmm_states=(0.5, 1.5, 1.8), hmm_sd=0.2, hmm_p=0.000001, nmean=100,
                      plot_seg=True, rds_path=None, adj=0, max_qt=0.99, rm_extreme=1, assay="WGS"):
    """HMM segmentation of one pooled coverage profile (reference R/Segmentation_bulk.R:22-218).
    Returns the object with `seg_table` (dict of equal-length string columns chr, start, end, states,
    length, mean, var, chrr — the reference's character matrix as a data.frame) and `annot`."""
    if Obj_filtered is None:
        raise ValueError("Please provide a valid Alleloscope object for Obj_filtered.")
    raw = Obj_filtered["raw_counts"]
    ref = Obj_filtered["ref_counts"]
    raw_v, raw_n = _counts_column(raw)
    ref_v, ref_n = _counts_column(ref)
    size = Obj_filtered["size"]
    last = list(size.keys())[-1]
    last = last if "chr" in last else "chr" + last
    if sum(str(n).split("-")[0] == last for n in raw_n) != sum(str(n).split("-")[0] == last for n in ref_n):
        raise ValueError("Tumor and control coverage files are with different numbers of bins.")
    (tab, annot), = _segment_profiles([(raw_v, ref_v)], raw_n, size, hmm_states, hmm_sd, hmm_p, nmean, adj, max_qt,
                                      rm_extreme, assay)
    out = dict(Obj_filtered)
    out["seg_table"] = tab
    out["annot"] = annot
    print("Segmentation done!")
    return out


# ------------------------------------------------------------------ CreateSegtableNoWGS
def merge_segment_tables(tables, merge=False):
    """R/CreateSegtableNoWGS.R:107-166: rbind, order by chromosome and start, drop duplicated
    chr-start-end, cut every chromosome at the union of the breakpoints; every piece carries the sorted
    distinct states of the segments it overlaps.  Returns an n x 4 array of strings."""
    rows = [(tab["chr"][i], tab["start"][i], tab["end"][i], tab["states"][i])
            for tab in tables for i in range(len(tab["chr"]))]
    rows.sort(key=lambda r: (float(r[0]), float(r[1])))
    seen, uniq = set(), []
    for r in rows:
        key = (r[0], r[1], r[2])
        if key not in seen:
            seen.add(key)
            uniq.append(r)
    out = []
    for chrom in range(1, 23):
        sub = [r for r in uniq if float(r[0]) == chrom]
        if not sub:
            continue
        s0 = np.array([float(r[1]) for r in sub])
        e0 = np.array([float(r[2]) for r in sub])
        stt = np.array([float(r[3]) for r in sub])
        bp = np.unique(np.concatenate([s0, e0]))
        pieces = []
        for i in range(len(bp) - 1):
            hit = (s0 <= bp[i + 1] - 1) & (e0 - 1 >= bp[i])                       # findOverlaps, type "any"
            pieces.append([sub[0][0], r_as_character(bp[i]), r_as_character(bp[i + 1]),
                           " ".join(r_as_character(v) for v in np.unique(stt[hit]))])
        if merge:                                                                 # :132-156
            merged, start_idx, end_idx, n = [], 0, 0, len(pieces)
            for i in range(n):
                if i + 1 < n:
                    if pieces[i][3] == pieces[i + 1][3]:
                        end_idx += 1
                    else:
                        merged.append([str(chrom), pieces[start_idx][1], pieces[end_idx][2], pieces[end_idx][3]])
                        start_idx = end_idx = i + 1
                else:
                    merged.append([str(chrom), pieces[start_idx][1], pieces[end_idx][2], pieces[end_idx][3]])
                    if end_idx < i:
                        merged.append([str(chrom), pieces[i][1], pieces[i][2], pieces[i][3]])
            pieces = merged
        out.extend(pieces)
    return np.array(out, dtype=object).reshape(-1, 4)


def _table_levels(z):
    """names(table(Zest)): the distinct non-NA labels, sorted as strings are by table()/factor()."""
    vals = [v for v in z if not (isinstance(v, float) and math.isnan(v))]
    try:
        return sorted(set(int(v) for v in vals))
    except (TypeError, ValueError):
        return sorted(set(str(v) for v in vals))


def CreateSegtableNoWGS(mtx=None, barcodes=None, features=None, Cov_obj=None, Zest=None, bin_bed=None,
                        celltype0=None, dir_path=None, bin_mtx=None, gtf=None, size=None, plot_seg=True,
                        hmm_states=(0.5, 1, 1.5, 2), max_qt=0.95, nmean=500, rm_extreme=2, adj=-0.5, rds_path=None,
                        merge=False, assay="WGS", return_obj=False):
    """Finer segments from the first-round clusters (reference R/CreateSegtableNoWGS.R:31-169): bin x
    cell counts -> bin x cluster pooled profiles -> per-cluster HMM segmentation -> union of the
    breakpoints.  Returns `seg_filtered` (n x 4 strings: chr, start, end, states) like the reference;
    with `return_obj=True` (not an argument of the reference, which writes them to Obj_all_cluster.rds
    instead) the per-cluster tables and the bin x cluster matrix come back as well:
    (seg_filtered, Obj_all, bin_by_cluster)."""
    import scipy.sparse as sp

    L = _lib.lib()
    cells = [str(v) for v in _first_column(celltype0)]
    ctype = [str(v) for v in (celltype0.iloc[:, 1] if hasattr(celltype0, "iloc") else np.asarray(celltype0, dtype=object)[:, 1])]
    if Zest is None:
        Zest = Cov_obj["result_final"]["result"]["Zest"]
    Zest = list(np.asarray(Zest).tolist())
    levels = _table_levels(Zest)
    index_of = {lev: k for k, lev in enumerate(levels)}                # one pass over the cells (NA -> no pool)
    numeric = bool(levels) and isinstance(levels[0], int)
    label = np.asarray([-1 if (isinstance(v, float) and math.isnan(v)) else index_of[int(v) if numeric else str(v)]
                        for v in Zest], dtype=np.int32)
    K = len(levels)
    normal = np.array([0 if c == "normal" else -1 for c in ctype], dtype=np.int32)        # (:77) pool 0 = the normal cells
    chrom, start, end = _bed_columns(bin_bed)
    if gtf is None and bin_mtx is None:
        print("Gene annotation file is empty, please provide the gene annotation file!")
    if bin_mtx is None:
        # align cells as in the celltype annotation (:37-38), bin on the device and pool where the result lies
        bc = [str(v) for v in _first_column(barcodes)]
        pos = {b: i for i, b in enumerate(bc)}
        try:
            cols = [pos[c] for c in cells]
        except KeyError as e:
            raise ClonalscopeError(_lib.CS_ERR_ARG, f"cell {e.args[0]} of celltype0 is not in barcodes") from None
        m = sp.csc_matrix(mtx)[:, cols]
        m.sum_duplicates()
        m.sort_indices()
        rownames, map_ptr, map_bin = gene_bin_map(bin_bed, features, gtf)
        G, N = m.shape
        B = len(rownames)
        if len(label) != N:
            raise ClonalscopeError(_lib.CS_ERR_ARG, f"Zest has {len(label)} labels for {N} cells")
        job, nnz = C.c_void_p(), C.c_int64(0)
        ip = np.ascontiguousarray(m.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(m.indices, dtype=np.int32)
        xx = np.ascontiguousarray(m.data, dtype=np.float64)
        mp = np.ascontiguousarray(map_ptr, dtype=np.int32)
        mb = np.ascontiguousarray(map_bin, dtype=np.int32)
        _check(L.cs_bin_counts_begin(G, N, B, _ptr(ip, C.c_int32), _ptr(ix, C.c_int32), _ptr(xx, C.c_double),
                                     _ptr(mp, C.c_int32), _ptr(mb, C.c_int32), C.byref(job), C.byref(nnz)))
        try:
            pooled = np.zeros((K, B))
            refsum = np.zeros((1, B))
            _check(L.cs_bin_job_rowsum_by_label(job, _ptr(label, C.c_int), K, _ptr(pooled, C.c_double)))
            _check(L.cs_bin_job_rowsum_by_label(job, _ptr(normal, C.c_int), 1, _ptr(refsum, C.c_double)))
        finally:
            L.cs_bin_counts_free(job)
        pooled, refsum = pooled.T, refsum[0]
    else:
        bm = bin_mtx.values if isinstance(bin_mtx, NamedMatrix) else bin_mtx
        B = bm.shape[0]
        if bm.shape[1] != len(label):
            raise ClonalscopeError(_lib.CS_ERR_ARG, f"Zest has {len(label)} labels for {bm.shape[1]} cells")
        pooled = rowsum_by_label(bm, label, K)
        refsum = rowsum_by_label(bm, normal, 1)[:, 0]
    print(f"Bin x Cluster matrix generated, with dimension of {B} bins x {K}clusters")
    # row names as the reference builds them from bin_bed in its given order (:76,:79)
    raw_names = [f"{c}-{r_as_character(float(s) + 1)}-{r_as_character(e) if not isinstance(e, str) else e}"
                 for c, s, e in zip(chrom, start, end)]
    if len(raw_names) != B:
        raise ClonalscopeError(_lib.CS_ERR_ARG, f"bin_bed has {len(raw_names)} rows, the bin matrix {B}")
    obj0 = Createobj_bulk(NamedMatrix(pooled[:, :1], raw_names, ["V1"]), NamedMatrix(refsum.reshape(-1, 1), raw_names, ["V1"]),
                          samplename="pool", dir_path=dir_path, size=size, assay="WGS")
    names = [f"C{lev}" for lev in levels]
    results = _segment_profiles([(pooled[:, k], refsum) for k in range(K)], raw_names, obj0["size"], hmm_states, 0.2,
                                0.000001, nmean, adj, max_qt, rm_extreme, assay)
    obj_all = {}
    for name, (tab, annot) in zip(names, results):
        obj_all[name] = {"seg_table": tab, "annot": annot}
        print(name)
    seg_filtered = merge_segment_tables([obj_all[n]["seg_table"] for n in names], merge=merge)
    if dir_path is not None:
        os.makedirs(dir_path, exist_ok=True)
        np.savez_compressed(os.path.join(dir_path, "seg_filtered_from_cluster.npz"), seg_filtered=seg_filtered.astype(str))
    if return_obj:
        return seg_filtered, obj_all, NamedMatrix(pooled, raw_names, names)
    return seg_filtered
