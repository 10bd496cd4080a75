"""Host-side mirror of the reference's R API for the cell-clustering hot path.

Same function names, argument names / order / defaults, error conditions and return-object
shapes as the reference's R closures (R is not installed in this image, so the host side
above the C ABI is Python; INTEGRATION.md shows the equivalent R `.Call` shim).  Everything
that is arithmetic runs on the device through `libclonalscope_b200.so`; what stays here is
what stays in the R shim: argument defaulting and validation, name matching, dimnames.

R objects map to Python as: named list -> dict, matrix -> 2-D numpy array (+ `rownames` /
`colnames` entries next to it), list of matrices -> list of arrays, NA_real_ -> NaN.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .intervals import gene_sites, overlap_map, r_as_character, sort_bins


class ClonalscopeError(RuntimeError):
    """A nonzero status from the C ABI (the R shim would raise it with stop())."""

    def __init__(self, status, message=None):
        if message is None:  # ClonalscopeError("text"): a host-side refusal, status CS_ERR_UNSUPPORTED
            status, message = _lib.CS_ERR_UNSUPPORTED, status
        super().__init__(f"[cs status {status}] {message}")
        self.status = status


def _check(rc):
    if rc != 0:
        raise ClonalscopeError(rc, _lib.lib().cs_last_error().decode("utf-8", "replace"))


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class NamedMatrix:
    """A matrix with dimnames (what the reference returns as a sparse `Matrix` or data.frame)."""

    def __init__(self, values, rownames, colnames):
        self.values = values
        self.rownames = list(rownames)
        self.colnames = list(colnames)

    @property
    def shape(self):
        return self.values.shape


# --------------------------------------------------------------------------- a1
def _first_column(obj):
    if hasattr(obj, "iloc"):
        return list(obj.iloc[:, 0])
    if isinstance(obj, dict):
        return list(next(iter(obj.values())))
    a = np.asarray(obj, dtype=object)
    return list(a[:, 0] if a.ndim == 2 else a)


def _bed_columns(bed):
    if hasattr(bed, "iloc"):
        return list(bed.iloc[:, 0]), list(bed.iloc[:, 1]), list(bed.iloc[:, 2])
    a = np.asarray(bed, dtype=object)
    return list(a[:, 0]), list(a[:, 1]), list(a[:, 2])


def gene_bin_map(bin_bed, genes, gtf):
    """Bin order, names and the gene->bin CSR overlap map (host integer logic of
    R/Gene_bin_cell_rna_filtered.R:28-64)."""
    chrom, start, end = _bed_columns(bin_bed)
    chrom, start, end, _ = sort_bins(chrom, start, end)
    gene_ids = _first_column(genes)
    gtf_cols = {k: (gtf[k] if k in gtf else None) for k in ("V1", "V3", "V4", "V5", "V9", "gene_id")} \
        if isinstance(gtf, dict) else {k: (gtf[k] if k in gtf.columns else None)
                                       for k in ("V1", "V3", "V4", "V5", "V9", "gene_id")}
    gtf_cols = {k: v for k, v in gtf_cols.items() if v is not None}
    seq, gs, ge = gene_sites(gene_ids, gtf_cols)
    map_ptr, map_bin = overlap_map(chrom, start, end, seq, gs, ge)
    rownames = [f"chr{c}-{r_as_character(s)}-{r_as_character(e)}" for c, s, e in zip(chrom, start, end)]
    return rownames, map_ptr, map_bin


def bin_counts_csc(G, N, B, colptr, rowidx, x, map_ptr, map_bin):
    """cs_bin_counts_begin / cs_bin_counts_fetch on dgCMatrix-style slots."""
    L = _lib.lib()
    colptr = np.ascontiguousarray(colptr, dtype=np.int32)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_bin = np.ascontiguousarray(map_bin, dtype=np.int32)
    job = C.c_void_p()
    nnz = C.c_int64(0)
    _check(L.cs_bin_counts_begin(G, N, B, _ptr(colptr, C.c_int32), _ptr(rowidx, C.c_int32),
                                 _ptr(x, C.c_double), _ptr(map_ptr, C.c_int32), _ptr(map_bin, C.c_int32),
                                 C.byref(job), C.byref(nnz)))
    ocp = np.zeros(N + 1, dtype=np.int32)
    ori = np.zeros(max(nnz.value, 1), dtype=np.int32)
    ox = np.zeros(max(nnz.value, 1), dtype=np.float64)
    _check(L.cs_bin_counts_fetch(job, _ptr(ocp, C.c_int32), _ptr(ori, C.c_int32), _ptr(ox, C.c_double)))
    return ocp, ori[:nnz.value], ox[:nnz.value]


def Gen_bin_cell_rna_filtered(bin_bed=None, barcodes=None, gene_matrix=None, genes=None, gtf=None):
    """Generate the bin-by-cell count matrix (reference R/Gene_bin_cell_rna_filtered.R:16-85).

    `gene_matrix` is a gene x cell scipy.sparse matrix (the reference's dgCMatrix);
    returns a NamedMatrix whose `.values` is a bin x cell scipy.sparse.csc_matrix with
    rownames `chr{chr}-{start}-{end}` and colnames = barcodes.
    """
    import scipy.sparse as sp

    barcodes = _first_column(barcodes)
    print("Read gene by cell matrix...")
    n_genes = len(_first_column(genes))
    print(f"Total {len(barcodes)} cells.")
    print(f"Total {n_genes} genes.")
    rownames, map_ptr, map_bin = gene_bin_map(bin_bed, genes, gtf)
    m = sp.csc_matrix(gene_matrix)
    m.sum_duplicates()
    m.sort_indices()
    G, N = m.shape
    if G != n_genes:
        raise ValueError(f"gene_matrix has {G} rows but genes has {n_genes}")
    B = len(rownames)
    print("Generate bin-by-cell matrix...", end="")
    ocp, ori, ox = bin_counts_csc(G, N, B, m.indptr, m.indices, m.data.astype(np.float64), map_ptr, map_bin)
    out = sp.csc_matrix((ox, ori, ocp), shape=(B, N))
    print("\nThe bin-by-cell matrix has beed successfully generated!")
    return NamedMatrix(out, rownames, barcodes)


# --------------------------------------------------------------------------- a7
def gene_moments(G, rowidx, x):
    """Per-gene sum and sum of squares over all cells (cs_gene_moments) on dgCMatrix slots."""
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, q = np.zeros(max(G, 1)), np.zeros(max(G, 1))
    _check(_lib.lib().cs_gene_moments(int(G), C.c_int64(len(x)), _ptr(rowidx, C.c_int32), _ptr(x, C.c_double),
                                      _ptr(s, C.c_double), _ptr(q, C.c_double)))
    return s[:G], q[:G]


def EstRegionCov(mtx=None, barcodes=None, features=None, bed=None, celltype0=None, var_pt=0.99, var_pt_ctrl=0.99,
                 ngene_filter=0, include="tumor", alpha_source="all", ctrl_region=None, seg_table_filtered=None,
                 size=None, plot_path=None, breaks=30):
    """Coverage change of every cell in every segment against the normal cells (reference
    R/EstRegionCov.R:24-209) without densifying the gene x cell matrix.  The matrix is uploaded once
    (cs_csc_upload; the cell alignment of :40 is a column gather on the device); the per-gene moments
    of the variance filter (:42-48, cs_csc_gene_moments) and, for all segments in one call, the
    per-segment column sums, library sizes, both medians and the ratios (:88-159, cs_csc_region_cov)
    run on that copy.  The interval overlaps and the quantile of the variance filter stay on the host.

    `mtx`: gene x cell scipy.sparse matrix; `bed`: 4 columns (chr, start, end, gene id);
    `celltype0`: 2 columns (barcode, "tumor"/"normal") or None; `seg_table_filtered`: mapping with
    columns chr, start, end, chrr (other columns ride along) or None with `size` (chromosome,
    length).  Returns dict(deltas_all, ngenes, seg_table_filtered, alpha, barcodes_normal) like the
    reference's list; every `deltas_all[name]` is a (cell names, values) pair.  `plot_path` is
    accepted and ignored (no plotting here)."""
    import scipy.sparse as sp

    if include not in ("tumor", "control", "all"):
        raise ValueError("include variable is not valid!")
    if alpha_source not in ("all", "control"):
        raise ValueError("alpha_source is not valid!")
    barcodes = [str(b) for b in _first_column(barcodes)]
    features = [str(f) for f in _first_column(features)]
    b = np.asarray(bed, dtype=object) if not hasattr(bed, "iloc") else bed.to_numpy(dtype=object)
    bchr = [str(c) for c in b[:, 0]]
    if "chr" not in bchr[0]:                                            # :28-30
        bchr = ["chr" + c for c in bchr]
    bstart, bend = b[:, 1].astype(np.float64), b[:, 2].astype(np.float64)
    bgene = [str(g) for g in b[:, 3]]
    if celltype0 is None:                                               # :32-34
        ct_bar, ct_type = list(barcodes), np.asarray(["normal"] * len(barcodes))
    else:
        ct = np.asarray(celltype0, dtype=object) if not hasattr(celltype0, "iloc") else celltype0.to_numpy(dtype=object)
        ct_bar, ct_type = [str(v) for v in ct[:, 0]], np.asarray([str(v) for v in ct[:, 1]])
    col_of = {}
    for i, bc in enumerate(barcodes):
        col_of.setdefault(bc, i)
    missing = [bc for bc in ct_bar if bc not in col_of]
    if missing:
        raise ClonalscopeError(_lib.CS_ERR_ARG, f"celltype0 names a barcode that is not in barcodes: {missing[0]}")
    m = mtx if sp.isspmatrix_csc(mtx) else sp.csc_matrix(mtx)
    if not m.has_canonical_format:                                      # (never touch the caller's matrix)
        m = m.copy()
        m.sum_duplicates()
    G, ncol = m.shape
    if G != len(features):
        raise ValueError(f"mtx has {G} rows but features has {len(features)}")
    col_sel = np.asarray([col_of[bc] for bc in ct_bar], dtype=np.int32)  # :40 mtx[, match(celltype0[,1], barcodes[,1])]
    N = len(col_sel)
    identity = N == ncol and bool(np.array_equal(col_sel, np.arange(ncol, dtype=np.int32)))
    handle = _CscHandle(m, None if identity else col_sel)
    try:
        return _est_region_cov_on(handle, G, N, features, bchr, bstart, bend, bgene, ct_bar, ct_type, var_pt,
                                  var_pt_ctrl, ngene_filter, include, alpha_source, ctrl_region, seg_table_filtered, size)
    finally:
        handle.free()


class _CscHandle:
    """A gene x cell matrix uploaded once (cs_csc_upload); moments and the per-segment computation run on it."""

    def __init__(self, m, col_sel=None):
        G, ncol = m.shape
        ip = np.ascontiguousarray(m.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(m.indices, dtype=np.int32)
        xx = np.ascontiguousarray(m.data, dtype=np.float64)
        self.G, self.N = G, ncol if col_sel is None else len(col_sel)
        self.h = C.c_void_p()
        sel = None if col_sel is None else np.ascontiguousarray(col_sel, dtype=np.int32)
        _check(_lib.lib().cs_csc_upload(G, ncol, _ptr(ip, C.c_int32), _ptr(ix, C.c_int32), _ptr(xx, C.c_double),
                                        0 if sel is None else len(sel), None if sel is None else _ptr(sel, C.c_int32),
                                        C.byref(self.h)))

    def free(self):
        if self.h:
            _lib.lib().cs_csc_free(self.h)
            self.h = C.c_void_p()

    def gene_moments(self):
        s, q = np.zeros(max(self.G, 1)), np.zeros(max(self.G, 1))
        _check(_lib.lib().cs_csc_gene_moments(self.h, _ptr(s, C.c_double), _ptr(q, C.c_double)))
        return s[:self.G], q[:self.G]

    def region_cov(self, S, map_ptr, map_seg, lib_flag, celltype, include):
        N = self.N
        deltas, sums = np.empty((S, N)), np.empty((S, N))
        alpha, nsel = np.zeros(N), np.zeros(S, dtype=np.int32)
        map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
        map_seg = np.ascontiguousarray(map_seg, dtype=np.int32)
        lib_flag = np.ascontiguousarray(lib_flag, dtype=np.uint8)
        celltype = np.ascontiguousarray(celltype, dtype=np.int32)
        _check(_lib.lib().cs_csc_region_cov(self.h, int(S), _ptr(map_ptr, C.c_int32), _ptr(map_seg, C.c_int32),
                                            _ptr(lib_flag, C.c_ubyte), _ptr(celltype, C.c_int), int(include),
                                            _ptr(deltas, C.c_double), _ptr(sums, C.c_double), _ptr(alpha, C.c_double),
                                            _ptr(nsel, C.c_int)))
        return deltas, sums, alpha, nsel


def _est_region_cov_on(handle, G, N, features, bchr, bstart, bend, bgene, ct_bar, ct_type, var_pt, var_pt_ctrl,
                       ngene_filter, include, alpha_source, ctrl_region, seg_table_filtered, size):
    # :42-48 per-gene variance and total, from device-side moments
    s1, s2 = handle.gene_moments()
    rna_var = np.maximum(s2 - s1 * s1 / N, 0.0) / (N - 1)
    keep = np.nonzero((rna_var < np.quantile(rna_var, var_pt)) & (rna_var != 0) & (s1 > ngene_filter))[0]
    keep_c = np.nonzero((rna_var < np.quantile(rna_var, var_pt_ctrl)) & (rna_var != 0) & (s1 > ngene_filter))[0]
    first = {}
    for i, g in enumerate(bgene):
        first.setdefault(g, i)
    j_of = np.asarray([first.get(features[k], -1) for k in keep], dtype=np.int64)
    bchr_a = np.asarray(bchr)
    sub_chr = np.where(j_of >= 0, bchr_a[np.maximum(j_of, 0)], "0") if len(bchr_a) else np.asarray(["0"] * len(j_of))  # :51-52
    sub_s = np.where(j_of >= 0, bstart[np.maximum(j_of, 0)], 0.0)
    sub_e = np.where(j_of >= 0, bend[np.maximum(j_of, 0)], 0.0)
    est_name = f"pt{r_as_character(var_pt)}_{include}_alpha{alpha_source}"  # :60
    if seg_table_filtered is None:                                      # :63-67
        message = "Estimation for each chromosome"
        print(message)
        sz = np.asarray(size, dtype=object) if not hasattr(size, "iloc") else size.to_numpy(dtype=object)
        names = [str(v) for v in sz[:22, 0]]
        ends = [float(v) for v in sz[:22, 1]]
        seg_table_filtered = dict(chr=[n.replace("chr", "") for n in names], start=[0.0] * len(names), end=ends,
                                  states=[0] * len(names), length=ends, mean=[0] * len(names), var=[0] * len(names),
                                  Var1=list(range(1, len(names) + 1)), Freq=[50000] * len(names), chrr=names)
    seg = {k: list(seg_table_filtered[k]) for k in (seg_table_filtered.keys() if isinstance(seg_table_filtered, dict)
                                                   else seg_table_filtered.columns)}
    seg_chr = [str(c) for c in seg["chr"]]
    seg_s, seg_e = np.asarray(seg["start"], dtype=np.float64), np.asarray(seg["end"], dtype=np.float64)
    seg_name = [str(c) for c in seg["chrr"]]
    uniq = list(dict.fromkeys(seg_name))                                # segments with the same chrr are one query
    S = len(uniq)
    # gene -> segment map over the genes that pass the filter (findOverlaps, type "any"); a gene that
    # two rows of one segment name overlap counts twice, as in the reference's concatenated hit list
    rows_of = {}
    for i, n in enumerate(seg_name):
        rows_of.setdefault(n, []).append(i)
    chr_codes = {c: i for i, c in enumerate(dict.fromkeys(sub_chr.tolist()))}
    sub_code = np.asarray([chr_codes[c] for c in sub_chr.tolist()], dtype=np.int64)
    valid = sub_s <= sub_e
    hit_gene, hit_seg = [], []
    ngene_of = np.zeros(S, dtype=np.int64)
    for si, name in enumerate(uniq):
        for r in rows_of[name]:
            code = chr_codes.get("chr" + seg_chr[r], -1)
            ok = np.nonzero((sub_code == code) & (sub_s <= seg_e[r]) & (sub_e >= seg_s[r]) & valid)[0]
            hit_gene.append(keep[ok])
            hit_seg.append(np.full(len(ok), si, dtype=np.int32))
            ngene_of[si] += len(ok)
    hit_gene = np.concatenate(hit_gene) if hit_gene else np.zeros(0, dtype=np.int64)
    hit_seg = np.concatenate(hit_seg) if hit_seg else np.zeros(0, dtype=np.int32)
    order = np.argsort(hit_gene, kind="stable")
    map_ptr = np.zeros(G + 1, dtype=np.int32)
    np.cumsum(np.bincount(hit_gene, minlength=G), out=map_ptr[1:])
    map_seg = hit_seg[order]
    # library sizes: the genes of the control filter (:131-133) or of the control regions
    if alpha_source == "all":
        lib_genes = keep_c
    else:
        lib_genes = keep[np.isin(sub_chr, [str(c) for c in (ctrl_region or [])])]
    lib_flag = np.zeros(max(G, 1), dtype=np.uint8)
    lib_flag[lib_genes] = 1
    is_normal, is_tumor = ct_type == "normal", ct_type == "tumor"
    code = np.where(is_normal, 0, np.where(is_tumor, 1, 2)).astype(np.int32)
    deltas, sums, alpha_all, nsel = handle.region_cov(max(S, 1), map_ptr, map_seg, lib_flag, code,
                                                      {"tumor": 0, "control": 1, "all": 2}[include])
    out_mask = {"tumor": is_tumor, "control": is_normal, "all": np.ones(N, dtype=bool)}[include]
    names_a = np.asarray(ct_bar, dtype=object)
    all_idx = np.nonzero(out_mask)[0]
    all_names = names_a[all_idx].tolist()
    deltas_all, ngenes, sel_ind = {}, {}, []
    last, dsel = -1, None
    for si, name in enumerate(uniq):                                    # :80
        if S == 0 or nsel[si] == 0:                                     # :93 no cell with a positive sum
            continue
        sel_ind.append(name)
        last = si
        if nsel[si] == N:                                               # every cell reports: one gather for all such segments
            if dsel is None:
                dsel = deltas[:, all_idx]
            deltas_all[est_name + name] = (all_names, dsel[si])
        else:
            idx = np.nonzero((sums[si] > 0) & out_mask)[0]
            deltas_all[est_name + name] = (names_a[idx].tolist(), deltas[si, idx])
        ngenes[est_name + name] = int(ngene_of[si])
        print(name)
    alphas, barcodes_normal = None, None
    if last >= 0:                                                       # the reference returns the last segment's
        sel_last = sums[last] > 0
        alphas = alpha_all[sel_last]
        barcodes_normal = names_a[sel_last & is_normal].tolist()
    keep_rows = [i for i, n in enumerate(seg_name) if n in set(sel_ind)]
    seg_out = {k: [v[i] for i in keep_rows] for k, v in seg.items()}
    return dict(deltas_all=deltas_all, ngenes=ngenes, seg_table_filtered=seg_out, alpha=alphas,
                barcodes_normal=barcodes_normal)


# --------------------------------------------------------------------------- a2
def PrepCovMatrix(deltas_all=None, ngene_filter=200, ncell_filter=None, prep_mode="intersect"):
    """Filter the output of EstRegionCov and assemble the cell x segment matrix
    (reference R/PrepCovMatrix.R:11-58; no arithmetic).

    deltas_all: dict(deltas_all=list of (names, values) pairs — one named vector per
    segment —, ngenes=array, seg_table_filtered=dict of equal-length columns incl.
    chr/start/end).  Returns dict(df=NamedMatrix N x R, ngene_filter, ncell_filter,
    seg_table_filtered).
    """
    df, ngenes = deltas_all["deltas_all"], deltas_all["ngenes"]
    if isinstance(df, dict):            # EstRegionCov's named list: one (names, values) pair per segment
        df = list(df.values())
    if isinstance(ngenes, dict):
        ngenes = list(ngenes.values())
    df = list(df)
    ngenes = np.asarray(ngenes)
    seg = {k: np.asarray(v) for k, v in deltas_all["seg_table_filtered"].items()}
    if ncell_filter is None:
        ncells = np.array([len(n) for n, _ in df], dtype=np.float64)
        ncell_filter = ncells.max() - ncells.max() * 0.1
    print(f"Selecting regions with > {r_as_character(ngene_filter)} genes \n")
    print(f"Selecting regions with > {r_as_character(ncell_filter)} cells \n")
    keep = np.array([len(n) > ncell_filter for n, _ in df], dtype=bool)
    df = [d for d, k in zip(df, keep) if k]
    ngenes = ngenes[keep]
    seg = {k: v[keep] for k, v in seg.items()}
    if prep_mode == "union":
        seen, cells = set(), []
        for names, _ in df:
            for b in names:
                if b not in seen:
                    seen.add(b)
                    cells.append(b)
    elif prep_mode == "intersect":
        cells = list(df[0][0]) if df else []
        for names, _ in df[1:]:
            s = set(names)
            cells = [b for b in cells if b in s]
    else:
        raise ValueError("Please specify the prep_mode as either union or intersect!")
    pos = {b: i for i, b in enumerate(cells)}
    mat = np.full((len(cells), len(df)), np.nan)
    for j, (names, vals) in enumerate(df):
        vals = np.asarray(vals, dtype=np.float64)
        seen = set()
        for b, v in zip(names, vals):  # match() takes the first occurrence of a name
            i = pos.get(b)
            if i is not None and b not in seen:
                mat[i, j] = v
                seen.add(b)
    sel = ngenes > ngene_filter
    mat = mat[:, sel]
    seg = {k: v[sel] for k, v in seg.items()}
    colnames = [f"chr{r_as_character(c)}-{r_as_character(s)}-{r_as_character(e)}"
                for c, s, e in zip(seg["chr"], seg["start"], seg["end"])]
    print("Matrix generated!")
    return dict(df=NamedMatrix(mat, cells, colnames), ngene_filter=ngene_filter, ncell_filter=ncell_filter,
                seg_table_filtered=seg)


# --------------------------------------------------------------------------- log-likelihood matrix
def loglik_matrix(Xir, U, sigmas):
    """LL[i,k] = sum_r dnorm(Xir[i,r], U[k,r], sigmas[r], log=TRUE), NA skipped (device)."""
    X = np.asfortranarray(np.asarray(Xir, dtype=np.float64))
    U = np.asfortranarray(np.atleast_2d(np.asarray(U, dtype=np.float64)))
    N, R = X.shape
    K = U.shape[0]
    if U.shape[1] != R:
        raise ValueError("U must have as many columns as Xir")
    sig = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmas, dtype=np.float64), (R,)))
    LL = np.zeros((N, K), dtype=np.float64, order="F")
    _check(_lib.lib().cs_loglik_matrix(N, R, K, _ptr(X, C.c_double), _ptr(U, C.c_double), _ptr(sig, C.c_double),
                                       _ptr(LL, C.c_double)))
    return LL


# --------------------------------------------------------------------------- a3
def _values(obj):
    if isinstance(obj, NamedMatrix):
        return np.asarray(obj.values, dtype=np.float64), obj.rownames, obj.colnames
    if hasattr(obj, "to_numpy"):
        return obj.to_numpy(dtype=np.float64), [str(i) for i in obj.index], [str(c) for c in obj.columns]
    a = np.asarray(obj, dtype=np.float64)
    return a, None, None


def _expand_U(kt, R, labels, rows):
    U = np.full((kt, R), np.nan)
    if len(labels):
        U[np.asarray(labels) - 1] = rows
    return U


def BayesNonparCluster(Xir=None, cna_states_WGS=None, alpha=2, beta=2, niter=200, sigmas0=None, U0=None,
                       Z0=None, clust_mode="all", seed=200, verbose=False, *, rng="philox", device=None,
                       kcap=None, _flags=0):
    """Nested-CRP Gibbs clustering on coverage (reference R/BayesNonparCluster.R:21-289).

    Same arguments and return object as the reference; the trailing keyword-only options
    are additions: rng="R" consumes base R's Mersenne-Twister stream exactly like the
    reference (bit-identical Zall for the same seed), rng="philox" (default) is the fast
    counter-based stream.
    Returns dict(results=dict(Zall, Uall, sigma_all, Likelihood), priors=dict(...), data).
    """
    X, rownames, colnames = _values(Xir)
    if X.ndim != 2:
        raise ValueError("Xir must be a matrix")
    N, R = X.shape
    cna = None if cna_states_WGS is None else np.asarray(cna_states_WGS, dtype=np.float64).ravel()
    # priors (:32-59)
    if U0 is None:
        wU0 = False
        if cna is not None and cna.size:
            U0m = np.vstack([np.ones(R), cna])  # matrix(c(rep(1,R), cna), byrow=T, ncol=R)
            if cna.size != R:
                raise ValueError("cna_states_WGS must have one value per column of Xir")
        else:
            U0m = np.ones((1, R))
            cna = U0m[0].copy()
    else:
        U0m = np.atleast_2d(np.asarray(_values(U0)[0], dtype=np.float64))
        if U0m.shape[1] != R:
            raise ValueError("Please specify a U0 matrix with ncol the same as that of Xir!")
        wU0 = True
        nn = np.nonzero(~np.isnan(U0m[:, 0]))[0]
        U0m = U0m[: nn.max() + 1]
        if cna is None or cna.size != R:
            print("Invalid/no cna_states_WGS detected. Use the one the same as that in the U0.")
            cna = U0m[nn[1]].copy() if U0m.shape[0] != 1 else U0m[0].copy()
    if sigmas0 is not None:
        s0 = np.asarray(sigmas0, dtype=np.float64).ravel()
        if s0.size == 1:
            s0 = np.repeat(s0, R)
        elif s0.size != R:
            raise ValueError("Please specify sigmas0 with length R or length 1.")
    else:
        s0 = np.full(R, 0.25)
    # Z0 (:91-112)
    if Z0 is None:
        live = np.nonzero(~np.isnan(U0m[:, 0]))[0]
        PP = np.full((N, U0m.shape[0]), np.nan)
        PP[:, live] = loglik_matrix(X, U0m[live], s0)
        Z0v = (np.nanargmax(PP, axis=1) + 1).astype(np.int32)
    else:
        Z0v = np.asarray(Z0).ravel()
        if Z0v.size != N:
            raise ValueError("Please specify a vector with length the same as the number of cells!")
        Z0v = np.where(np.isnan(Z0v.astype(np.float64)), 0, Z0v).astype(np.int32)
    # (NA rows of U0 below its last non-NA row — what AssignCluster's Uest has for dissolved clusters,
    #  R/RunCovCluster.R:131 — stay in place: they keep their label, hold no cell and are never
    #  scored or proposed, :93,:150-151,:193)
    K0 = U0m.shape[0]
    Xf = np.asfortranarray(X)
    U0f = np.asfortranarray(U0m)
    Zall = np.zeros((niter, N), dtype=np.int32, order="F")
    sigma_all = np.zeros((R, niter), dtype=np.float64, order="F")
    LL = np.zeros(niter, dtype=np.float64)
    L0 = C.c_double(0.0)
    kt_all = np.zeros(niter, dtype=np.int32)
    # (room for every sweep's occupied clusters; untouched pages of these buffers are never committed)
    ucap = int(niter) * min(N, 512 if str(rng).lower() == "r" else (int(kcap) if kcap else 256)) + 4096
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap, dtype=np.int32)
    u_rows = np.zeros((ucap, R), dtype=np.float64)
    nrej = C.c_int64(0)
    opts = _lib.GibbsOpts()
    opts.rng_mode = {"r": _lib.RNG_R, "philox": _lib.RNG_PHILOX}[str(rng).lower()]
    opts.device = -1 if device is None else int(device)
    opts.kcap = 0 if kcap is None else int(kcap)
    opts.flags = int(_flags)
    rng_state = np.zeros(625, dtype=np.int32)
    _check(_lib.lib().cs_ncrp_gibbs(
        N, R, _ptr(Xf, C.c_double), K0, _ptr(U0f, C.c_double), _ptr(Z0v, C.c_int), _ptr(s0, C.c_double),
        C.c_double(alpha), C.c_double(beta), int(niter), C.c_uint32(int(seed) & 0xFFFFFFFF), C.byref(opts),
        _ptr(Zall, C.c_int), _ptr(sigma_all, C.c_double), _ptr(LL, C.c_double), C.byref(L0),
        _ptr(kt_all, C.c_int), C.c_int64(ucap), _ptr(u_off, C.c_int64), _ptr(u_labels, C.c_int),
        _ptr(u_rows, C.c_double), C.byref(nrej), _ptr(rng_state, C.c_int32)))
    Uall = []
    for t in range(niter):
        a, b = int(u_off[t]), int(u_off[t + 1])
        Uall.append(_expand_U(int(kt_all[t]), R, u_labels[a:b], u_rows[a:b]))
        if verbose:
            print(f"Iteration:{t + 1}")
            print(f"nclusters={int(kt_all[t])}")
        elif (t + 1) % 20 == 0:
            print(f"Iteration:{t + 1}")
    single = np.unique(Z0v).size == 1
    results = dict(Zall=np.ascontiguousarray(Zall), Uall=Uall,
                   sigma_all=[sigma_all[:, t].copy() for t in range(niter)], Likelihood=LL,
                   dimnames=dict(Zall=([f"Iter{t + 1}" for t in range(niter)], rownames),
                                 U_cols=[f"R{r + 1}" for r in range(R)]))
    priors = dict(Z0=Z0v, U0=(np.ones((1, R)) if single else U0m), sigmas0=s0, alpha=alpha, beta=beta,
                  cna_states_WGS=cna, L0=L0.value, wU0=wU0, seed=seed)
    out = dict(results=results, priors=priors, data=Xir if Xir is not None else X)
    out["diagnostics"] = dict(n_reject=int(nrej.value), rng=str(rng),
                              rng_state=rng_state if opts.rng_mode == _lib.RNG_R else None)
    return out


# --------------------------------------------------------------------------- a4
def BayesNonparAlleleCluster(Xir_cov=None, Xir_allele=None, cna_states_WGS=None, alpha=1, beta=1, niter=200,
                             sigmas0_cov=None, sigmas0_allele=None, U0=None, Z0=None, maxcp=6, seed=200, *,
                             rng="philox", device=None, kcap=None):
    """Allele-aware nested-CRP Gibbs clustering (reference R/BayesNonparAlleleCluster.R:20-307).
    Parity of this path is unpinned by the reference (never called, no fixture)."""
    Xc, rownames, colnames = _values(Xir_cov)
    Xa = _values(Xir_allele)[0]
    if Xc.shape != Xa.shape:
        raise ValueError("Xir_cov and Xir_allele must have the same shape")
    N, R = Xc.shape
    grid = np.array([(0.5 * ii, 0.0 if j == 0 else j / ii) for ii in range(1, maxcp + 1) for j in range(ii + 1)])
    ngrid = grid.shape[0]  # (copies, major-allele fraction) states, :26-30
    if U0 is None:
        if cna_states_WGS is not None:
            U0m = np.vstack([np.full(R, 4), np.asarray(cna_states_WGS).ravel()]).astype(np.int32)
        else:
            U0m = np.full((1, R), 4, dtype=np.int32)
    else:
        U0m = np.atleast_2d(np.asarray(_values(U0)[0] if not isinstance(U0, np.ndarray) else U0)).astype(np.int32)
        if U0m.shape[1] != R:
            raise ValueError("Please specify a U0 matrix with ncol the same as that of Xir!")
        if U0m.shape[0] != 1:
            # the reference dereferences an undefined `result$Uest` here (:51,:62)
            raise NameError("object 'result' not found")
    if U0m.shape[1] != R:
        raise ValueError("Please specify a U0 matrix with ncol the same as that of Xir!")
    if U0m.min() < 1 or U0m.max() > ngrid:
        raise ValueError(f"U0 state ids must lie in 1..{ngrid}")

    def _sig(s, default, what):
        if s is None:
            return np.full(R, default)
        s = np.asarray(s, dtype=np.float64).ravel()
        if s.size == 1:
            return np.repeat(s, R)
        if s.size != R:
            raise ValueError(f"Please specify {what} with length R or length 1.")
        return s

    sc, sa = _sig(sigmas0_cov, 0.3, "sigmas0_cov"), _sig(sigmas0_allele, 0.15, "sigmas0_allele")
    if Z0 is None:
        # :99-105 argmax over the rows of U0 of the coverage + allele log-likelihood (device)
        PP = loglik_matrix(np.minimum(Xc, 3.0), grid[U0m - 1, 0], sc) + loglik_matrix(Xa, grid[U0m - 1, 1], sa)
        Z0v = (np.argmax(PP, axis=1) + 1).astype(np.int32)
    else:
        Z0v = np.asarray(Z0).ravel()
        if Z0v.size != N:
            raise ValueError("Please specify a vector with length the same as the number of cells!")
        Z0v = np.where(np.isnan(Z0v.astype(np.float64)), 0, Z0v).astype(np.int32)
    K0 = U0m.shape[0]
    Zall = np.zeros((niter, N), dtype=np.int32, order="F")
    sig_cov_all = np.zeros((R, niter), order="F")
    sig_all_all = np.zeros((R, niter), order="F")
    LL = np.zeros(niter)
    L0 = C.c_double(0.0)
    kt_all = np.zeros(niter, dtype=np.int32)
    ucap = int(niter) * 160 + 4096
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap, dtype=np.int32)
    u_state = np.zeros((ucap, R), dtype=np.int32)
    u_cov = np.zeros((ucap, R))
    u_allele = np.zeros((ucap, R))
    opts = _lib.GibbsOpts()
    opts.rng_mode = {"r": _lib.RNG_R, "philox": _lib.RNG_PHILOX}[str(rng).lower()]
    opts.device = -1 if device is None else int(device)
    opts.kcap = 0 if kcap is None else int(kcap)
    Xcf, Xaf, U0f = np.asfortranarray(Xc), np.asfortranarray(Xa), np.asfortranarray(U0m)
    _check(_lib.lib().cs_ncrp_gibbs_allele(
        N, R, _ptr(Xcf, C.c_double), _ptr(Xaf, C.c_double), K0, _ptr(U0f, C.c_int), _ptr(Z0v, C.c_int),
        _ptr(sc, C.c_double), _ptr(sa, C.c_double), C.c_double(alpha), C.c_double(beta), int(niter),
        C.c_uint32(int(seed) & 0xFFFFFFFF), int(maxcp), C.byref(opts), _ptr(Zall, C.c_int),
        _ptr(sig_cov_all, C.c_double), _ptr(sig_all_all, C.c_double), _ptr(LL, C.c_double), C.byref(L0),
        _ptr(kt_all, C.c_int), C.c_int64(ucap), _ptr(u_off, C.c_int64), _ptr(u_labels, C.c_int),
        _ptr(u_state, C.c_int), _ptr(u_cov, C.c_double), _ptr(u_allele, C.c_double)))
    Uall, Ucov, Uallele = [], [], []
    for t in range(niter):
        a, b = int(u_off[t]), int(u_off[t + 1])
        kt = int(kt_all[t])
        # emptied clusters keep all-zero rows here (the reference fills with 0, not NA, :203-205)
        Uall.append(np.nan_to_num(_expand_U(kt, R, u_labels[a:b], u_state[a:b].astype(np.float64))))
        Ucov.append(np.nan_to_num(_expand_U(kt, R, u_labels[a:b], u_cov[a:b])))
        Uallele.append(np.nan_to_num(_expand_U(kt, R, u_labels[a:b], u_allele[a:b])))
        print(f"Iteration:{t + 1}")
        print(f"nclusters={kt}")
    results = dict(Zall=np.ascontiguousarray(Zall), Uall=Uall, Ucov=Ucov, Uallele=Uallele,
                   sigma_cov_all=[sig_cov_all[:, t].copy() for t in range(niter)],
                   sigma_allele_all=[sig_all_all[:, t].copy() for t in range(niter)], Likelihood=LL)
    priors = dict(Z0=Z0v, U0=U0m, sigmas0_cov=sc, sigmas0_allele=sa, alpha=alpha, beta=beta,
                  cna_states_WGS=cna_states_WGS, L0=L0.value, seed=seed)
    return dict(results=results, priors=priors, data=dict(Xir_cov=Xir_cov, Xir_allele=Xir_allele))


# --------------------------------------------------------------------------- a5 / a6
def MCMCtrim(clusteringObj=None, burnin=None, thinning=1, allele=False):
    """Drop burn-in sweeps and thin (reference R/MCMCtrim.R:14-65).

    Kept for parity: Zall is thinned first and the ALREADY-thinned nrow(Zall) is then used
    to index the per-sweep lists (:33-36), so with thinning > 1 those lists come out shorter
    than Zall; a no-op for the default thinning = 1.
    """
    res = dict(clusteringObj["results"])
    nrow = res["Zall"].shape[0]
    if burnin is None:
        burnin = min(100, nrow / 2)
    if burnin >= nrow:
        raise ValueError("burnin values is too large! Please specify a number < number of iterations.")
    if burnin < 0:
        raise ValueError("can't mix positive and negative subscripts")  # what -c(1:burnin) gives in R
    # -c(1:burnin): a fractional burnin truncates like R's integer indexing; 1:0 is c(1, 0), so
    # burnin = 0 (and anything below 1) still drops the first sweep (:27-30)
    b = max(1, int(burnin))
    lists = ["Uall", "Ucov", "Uallele", "sigma_cov_all", "sigma_allele_all", "Likelihood"] if allele \
        else ["Uall", "sigma_all", "Likelihood"]
    res["Zall"] = res["Zall"][b:]
    for k in lists:
        res[k] = res[k][b:]

    def thin_idx(n):
        return [i * thinning for i in range(int(math.ceil(n / thinning)))]

    res["Zall"] = res["Zall"][thin_idx(res["Zall"].shape[0])]
    idx = thin_idx(res["Zall"].shape[0])  # reference quirk: nrow of the thinned Zall
    for k in lists:
        v = res[k]
        if isinstance(v, np.ndarray):
            res[k] = v[[i for i in idx if i < len(v)]]
        else:
            res[k] = [v[i] for i in idx if i < len(v)]
    if "dimnames" in res:
        dn = dict(res["dimnames"])
        it, cells = dn["Zall"]
        it = it[b:]
        dn["Zall"] = ([it[i] for i in thin_idx(len(it))], cells)
        res["dimnames"] = dn
    out = dict(clusteringObj)
    out["results"] = res
    out["trim"] = dict(burnin=burnin, thinning=thinning)
    print("Object was succefully trimmed and thinned!")
    return out


def majority_vote(Zall):
    """Per-cell most frequent label over the kept sweeps, ties -> smallest label
    (reference R/AssignCluster.R:37; the posterior assignment tally)."""
    Z = np.asfortranarray(np.asarray(Zall, dtype=np.int32))
    T, N = Z.shape
    out = np.zeros(N, dtype=np.int32)
    _check(_lib.lib().cs_majority_vote(T, N, _ptr(Z, C.c_int), _ptr(out, C.c_int)))
    return out


# --------------------------------------------------------------------------- AssignCluster
def _label_stats(Xd, z, K, U=None):
    """Per-cluster sums / non-NA counts (cs_cluster_stats_dev) and, given U, per-segment squared
    residuals (cs_sigma_stats_dev) of device-resident data; labels outside 1..K are skipped."""
    import torch
    L = _lib.lib()
    N, R = Xd.shape
    zd = torch.as_tensor(np.ascontiguousarray(z, dtype=np.int32), device=Xd.device)
    sum_x = torch.zeros((K, R), dtype=torch.float64, device=Xd.device)
    cnt_x, nk = torch.zeros_like(sum_x), torch.zeros(K, dtype=torch.float64, device=Xd.device)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if U is None:
        _check(L.cs_cluster_stats_dev(N, R, K, vp(Xd), vp(zd), vp(sum_x), vp(cnt_x), vp(nk), st))
        torch.cuda.synchronize()
        return sum_x.cpu().numpy(), cnt_x.cpu().numpy()
    Ud = torch.as_tensor(np.ascontiguousarray(U, dtype=np.float64), device=Xd.device)
    sq, cnt = torch.zeros(R, dtype=torch.float64, device=Xd.device), torch.zeros(R, dtype=torch.float64, device=Xd.device)
    _check(L.cs_sigma_stats_dev(N, R, K, vp(Xd), vp(zd), vp(Ud), vp(sq), vp(cnt), st))
    torch.cuda.synchronize()
    return sq.cpu().numpy()


def _genotype_grid(maxcp=6):
    """The (rho, theta) grid of R/genotype_neighbor.R:9-13 (rows in the reference's order)."""
    rows = []
    for ii in range(1, maxcp + 1):
        rho = [0.5 * ii] * (ii + 1)
        theta = [0.0] + [j / ii for j in range(1, ii + 1)]
        rows.extend(zip(rho, theta))
    return np.asarray(rows, dtype=np.float64)


def _genotype_neighbor(cov, allele, maxcp=6):
    """1-based index of the nearest grid state of every (cov, allele) pair (first minimum)."""
    mu = _genotype_grid(maxcp)
    d = np.sqrt((np.asarray(cov)[:, None] - mu[None, :, 0]) ** 2 + (np.asarray(allele)[:, None] - mu[None, :, 1]) ** 2)
    return np.argmin(d, axis=1) + 1


def _km_cutoff(stream, corrs):
    """cutoff = 'km' (R/AssignCluster.R:191-194): kmeans(annot[annot != 100], 2) on R's stream."""
    from .rrng import kmeans_hartigan_wong
    vals = [float(v) for v in corrs if v != 100]
    try:
        cen = kmeans_hartigan_wong(stream, vals, 2)
    except ValueError as e:
        raise ClonalscopeError(_lib.CS_ERR_ARG, str(e)) from None
    return float((np.longdouble(cen[0]) + np.longdouble(cen[1])) / 2)


def _assign_cluster_allele(cluster_obj, mincell, cutoff, cna_states_WGS, rm_extreme, cap_corr):
    """The allele branch (reference R/AssignCluster.R:216-414): as the coverage branch on both
    modalities, orphans scored on coverage + allele, final means with the coverage capped at 3,
    the genotype of every (cluster, region) from the nearest grid state."""
    import torch
    from .rrng import RStream, sample_int_prob_one
    res, priors, data = cluster_obj["results"], cluster_obj["priors"], cluster_obj["data"]
    Xc_v, _, colnames = _values(data["Xir_cov"])
    Xa_v, _, _ = _values(data["Xir_allele"])
    Xc = np.ascontiguousarray(Xc_v, dtype=np.float64)
    Xa = np.ascontiguousarray(Xa_v, dtype=np.float64)
    N, R = Xc.shape
    Zall = np.asarray(res["Zall"], dtype=np.int32)
    kt = int(np.asarray(res["Uall"][-1]).shape[0])
    stream = RStream(100)                                                # :231
    zmaj = majority_vote(Zall)                                           # :232
    labs, cnts = np.unique(zmaj, return_counts=True)
    ind_clusters = labs[cnts > mincell].astype(np.int64)                 # :236
    orphan = ~np.isin(zmaj, ind_clusters)
    ind_cell = np.nonzero(orphan)[0]
    if len(ind_clusters) == 0:
        raise ClonalscopeError("no cluster has more than mincell cells")
    Xcd, Xad = torch.as_tensor(Xc, device="cuda"), torch.as_tensor(Xa, device="cuda")
    z = np.where(orphan, 0, zmaj).astype(np.int32)
    zkeep = z.copy()                                                     # labels of the cells outside ind_cell

    def means(Xd, zz):                                                   # colMeans(na.rm = T), NA -> 0; counts per cluster
        sx, c = _label_stats(Xd, zz, kt)
        U = np.full((kt, R), np.nan)
        with np.errstate(divide="ignore", invalid="ignore"):
            mm = np.where(c > 0, sx / np.maximum(c, 1), 0.0)
        U[ind_clusters - 1] = mm[ind_clusters - 1]
        return U, c

    def sigma(Xd, U, n):
        # colSums((X[-ind_cell,] - U[Zest[-ind_cell],])^2): x[-integer(0), ] selects NO row in R, so
        # without orphans the reference's sums are empty (sigma 0); reproduced as it stands
        if len(ind_cell) == 0:
            sq = np.zeros(R)
        else:
            zz = np.where(orphan, 0, z).astype(np.int32)
            sq = _label_stats(Xd, zz, kt, np.where(np.isnan(U), 0.0, U))
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.sqrt(1.0 / n * sq)

    Usubc, cc = means(Xcd, z)                                            # :240-250
    Usuba, ca = means(Xad, z)
    ncov, nallele = cc[ind_clusters - 1].sum(axis=0), ca[ind_clusters - 1].sum(axis=0)      # :252-256
    sig_c, sig_a = sigma(Xcd, Usubc, ncov), sigma(Xad, Usuba, nallele)  # :258-259
    if len(ind_cell):                                                    # :265-275
        LLo = loglik_matrix(Xc[ind_cell], Usubc[ind_clusters - 1], sig_c) + \
            loglik_matrix(Xa[ind_cell], Usuba[ind_clusters - 1], sig_a)
        for row, ii in zip(LLo, ind_cell):
            P = np.full(kt, np.nan)
            P[ind_clusters - 1] = row - np.max(row)
            P[np.isnan(P)] = np.nanmin(P) - 100
            z[ii] = sample_int_prob_one(stream, np.exp(P))
    Xc2d = torch.clamp(Xcd, max=3.0)                                     # :281 pmin(x, 3) keeps NA
    Usubc, _ = means(Xc2d, z)                                            # :282-294
    Usuba, _ = means(Xad, z)
    Usub = np.full((kt, R), np.nan)
    for kk in ind_clusters:
        Usub[kk - 1] = _genotype_neighbor(Usubc[kk - 1], Usuba[kk - 1], 6)
    ncov, nallele = (~np.isnan(Xc)).sum(axis=0), (~np.isnan(Xa)).sum(axis=0)               # :302-303
    sigmas_cov, sigmas_allele = sigma(Xcd, Usubc, ncov), sigma(Xad, Usuba, nallele)         # :305-306
    pri = priors.get("cna_states_WGS") if cna_states_WGS is None else cna_states_WGS        # :311-315
    pri = np.asarray(pri, dtype=np.float64).ravel()
    if np.all(pri == 4):                                                 # :317-320
        pri = np.full(R, 8.0)
    priorsc = _genotype_grid(6)[pri.astype(np.int64) - 1, 0]             # :333
    U2 = np.maximum(np.minimum(Usubc, cap_corr), 0.5)                    # :339-340
    corrs = np.zeros(kt)
    for k in range(kt):                                                  # :342-353
        tmp = (U2[k] - 1) * (priorsc - 1)
        tmp = np.sort(tmp[~np.isnan(tmp)])
        if len(tmp) == 0:
            continue                                                     # NA -> 0 (:357)
        if rm_extreme:
            tmp = tmp[1:len(tmp) - 1] if len(tmp) > 2 else (tmp[[1, 0]] if len(tmp) == 2 else np.array([tmp[0], np.nan]))
        corrs[k] = float(np.sum(tmp.astype(np.longdouble)))
    corrs = np.where(np.isnan(corrs), 0.0, corrs)
    corrs = corrs / corrs.max()                                          # :358
    if isinstance(cutoff, str):
        if cutoff != "km":
            raise ClonalscopeError(_lib.CS_ERR_ARG, "cutoff must be a number or 'km'")
        cutoff = _km_cutoff(stream, corrs)
    annot = np.where(corrs > cutoff, "T", "N")                           # :401-402
    print("Succeed!")
    rn = [f"Cluster{k + 1}" for k in range(kt)]
    wrap = (lambda M: NamedMatrix(M, rn, colnames)) if colnames is not None else (lambda M: M)
    return dict(Zest=z.astype(np.int32), corrs=corrs[z - 1], annot=annot[z - 1], Uest=wrap(Usub), Ucest=wrap(Usubc),
                Uaest=wrap(Usuba), sigmas_cov_est=sigmas_cov, sigmas_allele_est=sigmas_allele, Zmaj=zmaj, cutoff=cutoff,
                U0=priors.get("U0"), wU0=priors.get("wU0"))


def AssignCluster(cluster_obj=None, mincell=20, cutoff=0.2, allele=False, cna_states_WGS=None, rm_extreme=False,
                  cap_corr=1.5):
    """Assign a subclone to every cell (reference R/AssignCluster.R:20-208, coverage branch):
    majority vote over the kept sweeps, clusters of at most `mincell` cells dissolved and their
    cells re-drawn among the remaining ones under R's own random stream (`set.seed(100)`, :36,63),
    final means / sigmas, malignancy index and T/N annotation per cluster.

    Device: the tally, the per-cluster sums and squared residuals, the orphans' log-likelihoods.
    Host: counting, R's stream (one uniform per orphan), the K x R index arithmetic of :106-150.
    `allele=TRUE` takes the allele branch (:216-414, _assign_cluster_allele); `cutoff='km'` is R's
    kmeans on the same stream (:191-194, rrng.kmeans_hartigan_wong)."""
    import torch
    if allele:
        return _assign_cluster_allele(cluster_obj, mincell, cutoff, cna_states_WGS, rm_extreme, cap_corr)
    if isinstance(cutoff, str) and cutoff != "km":
        raise ClonalscopeError(_lib.CS_ERR_ARG, "cutoff must be a number or 'km'")
    if rm_extreme:
        raise ClonalscopeError("rm_extreme=TRUE returns NULL indices in the reference (:139-144); not supported")
    from .rrng import RStream, sample_int_prob_one
    res, priors = cluster_obj["results"], cluster_obj["priors"]
    Xobj = cluster_obj["data"]
    Xv, _, colnames = _values(Xobj)
    X = np.ascontiguousarray(Xv, dtype=np.float64)
    N, R = X.shape
    Zall = np.asarray(res["Zall"], dtype=np.int32)
    kt = int(np.asarray(res["Uall"][-1]).shape[0])
    stream = RStream(100)                                                # :36
    zmaj = majority_vote(Zall)                                           # :37
    labs, cnts = np.unique(zmaj, return_counts=True)
    ind_clusters = labs[cnts > mincell].astype(np.int64)                 # :41
    orphan = ~np.isin(zmaj, ind_clusters)                                # :42
    ind_cell = np.nonzero(orphan)[0]
    if len(ind_clusters) == 0:
        raise ClonalscopeError("no cluster has more than mincell cells")
    Xd = torch.as_tensor(X, device="cuda")
    z = np.where(orphan, 0, zmaj).astype(np.int32)

    def means(zz):                                                       # :46-51 colMeans(na.rm = T), NA -> 0
        s, c = _label_stats(Xd, zz, kt)
        U = np.full((kt, R), np.nan)
        with np.errstate(divide="ignore", invalid="ignore"):
            m = np.where(c > 0, s / np.maximum(c, 1), 0.0)
        U[ind_clusters - 1] = m[ind_clusters - 1]
        return U

    Usub = means(z)
    Ufill = np.where(np.isnan(Usub), 0.0, Usub)
    sig = np.sqrt(1.0 / (N - len(ind_cell)) * _label_stats(Xd, z, kt, Ufill))   # :53
    if len(ind_cell):                                                    # :56-64
        LLo = loglik_matrix(X[ind_cell], Usub[ind_clusters - 1], sig)
        for row, ii in zip(LLo, ind_cell):
            P = np.full(kt, np.nan)
            P[ind_clusters - 1] = row - np.nanmax(row)
            P[np.isnan(P)] = np.nanmin(P) - 100
            z[ii] = sample_int_prob_one(stream, np.exp(P))
    Usub = means(z)                                                      # :67-72
    sigmas = np.sqrt(1.0 / N * _label_stats(Xd, z, kt, np.where(np.isnan(Usub), 0.0, Usub)))   # :77
    pri = priors.get("cna_states_WGS") if cna_states_WGS is None else cna_states_WGS    # :80-84
    pri = np.asarray(pri, dtype=np.float64).ravel()
    if np.all(pri == 1):                                                 # :85-88
        pri = np.full(R, 1.5)
    U2 = np.maximum(np.minimum(Usub, cap_corr), 0.5)                     # :106-107
    corrs = np.zeros(kt)
    for k in range(kt):                                                  # :127-152
        x = U2[k]
        if np.any(np.isnan(x)):
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            tmp = ((x - 1) * (pri - 1)) / (np.sqrt(np.sum((x - 1) ** 2)) * np.sqrt(np.sum((pri - 1) ** 2)))
        tmp = np.sort(tmp[~np.isnan(tmp)])
        corrs[k] = float(np.sum(tmp.astype(np.longdouble))) if len(tmp) else 0.0
    if corrs.max() < 0.5:                                                # :154-156
        import warnings
        warnings.warn("WGS and scRNA do not match well. annot might not be accurate!")
    if isinstance(cutoff, str):                                          # :191-194
        cutoff = _km_cutoff(stream, corrs)
    annot = np.where(corrs > cutoff, "T", "N")                           # :196-197
    print("Succeed!")
    Uest = NamedMatrix(Usub, [f"Cluster{k + 1}" for k in range(kt)], colnames) if colnames is not None else Usub
    return dict(Zest=z.astype(np.int32), corrs=corrs[z - 1], annot=annot[z - 1], Uest=Uest, sigmas_est=sigmas,
                Zmaj=zmaj, cutoff=cutoff, U0=priors.get("U0"), mincell=mincell, wU0=priors.get("wU0"))

