"""ctypes front-end of the CPU parity oracle (libcs_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
product package `clonalscope_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RNG_R = 0
RNG_PHILOX = 1
FLAG_SORT_ONCE = 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.cso_qnorm.restype = C.c_double
        _LIB.cso_qnorm.argtypes = [C.c_double]
        _LIB.cso_unif_rand.restype = C.c_double
        _LIB.cso_rnorm.restype = C.c_double
        _LIB.cso_rnorm.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _LIB.cso_total_loglik.restype = C.c_double
        _LIB.cso_bin_counts.restype = C.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("pos", C.c_int), ("ndraws", C.c_uint64)]


def mt_uniforms(seed: int, n: int) -> np.ndarray:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    return np.array([L.cso_unif_rand(C.byref(g)) for _ in range(n)])


def mt_rnorm(seed: int, means, sds) -> np.ndarray:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    return np.array([L.cso_rnorm(C.byref(g), float(m), float(s)) for m, s in zip(means, sds)])


def sample_int_prob(seed: int, prob, ndraw: int = 1) -> list:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    out = []
    n = len(prob)
    for _ in range(ndraw):
        p = np.array(prob, dtype=np.float64)
        perm = np.zeros(n, dtype=np.int32)
        out.append(L.cso_sample_int_prob(C.byref(g), n, _p(p, C.c_double), _p(perm, C.c_int)))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().cso_philox4x32(c, k, o)
    return [int(v) for v in o]


def qnorm(p):
    L = lib()
    return np.array([L.cso_qnorm(float(v)) for v in np.atleast_1d(p)])


def loglik_matrix(X, U, sigma):
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    N, R = X.shape
    K = U.shape[0]
    LL = np.empty((N, K), dtype=np.float64)
    lib().cso_loglik_matrix(N, R, K, _p(X, C.c_double), _p(U, C.c_double), _p(sigma, C.c_double),
                            _p(LL, C.c_double))
    return LL


def total_loglik(X, U, Z, sigma):
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    N, R = X.shape
    return lib().cso_total_loglik(N, R, _p(X, C.c_double), _p(U, C.c_double), U.shape[0],
                                  _p(Z, C.c_int), _p(sigma, C.c_double))


def gibbs(X, U0, Z0, sigmas0, alpha=2.0, beta=2.0, niter=200, seed=200, rng_mode=RNG_R, flags=0,
          ucap_rows=None):
    """Returns dict(Zall, sigma_all, LL, L0, kt, u_off, u_labels, u_rows, n_reject)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    U0 = np.ascontiguousarray(np.atleast_2d(U0), dtype=np.float64)
    Z0 = np.ascontiguousarray(Z0, dtype=np.int32)
    N, R = X.shape
    K0 = U0.shape[0]
    sig0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmas0, dtype=np.float64), (R,)))
    if ucap_rows is None:
        ucap_rows = niter * 256
    Zall = np.zeros((niter, N), dtype=np.int32)
    sigma_all = np.zeros((niter, R), dtype=np.float64)
    LL = np.zeros(niter, dtype=np.float64)
    L0 = np.zeros(1, dtype=np.float64)
    kt = np.zeros(niter, dtype=np.int32)
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap_rows, dtype=np.int32)
    u_rows = np.zeros((ucap_rows, R), dtype=np.float64)
    nrej = np.zeros(1, dtype=np.int64)
    rc = lib().cso_gibbs(
        N, R, _p(X, C.c_double), K0, _p(U0, C.c_double), _p(Z0, C.c_int), _p(sig0, C.c_double),
        C.c_double(alpha), C.c_double(beta), niter, C.c_uint32(int(seed)), rng_mode, flags,
        _p(Zall, C.c_int), _p(sigma_all, C.c_double), _p(LL, C.c_double), _p(L0, C.c_double),
        _p(kt, C.c_int), C.c_int64(ucap_rows), _p(u_off, C.c_int64), _p(u_labels, C.c_int),
        _p(u_rows, C.c_double), _p(nrej, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"cso_gibbs failed with code {rc}")
    n = int(u_off[-1])
    return dict(Zall=Zall, sigma_all=sigma_all, LL=LL, L0=float(L0[0]), kt=kt, u_off=u_off,
                u_labels=u_labels[:n], u_rows=u_rows[:n], n_reject=int(nrej[0]))


def expand_U(res, t):
    """Uall[[t+1]] as the reference stores it: kt x R with NaN rows for empty clusters."""
    a, b = int(res["u_off"][t]), int(res["u_off"][t + 1])
    R = res["u_rows"].shape[1]
    U = np.full((int(res["kt"][t]), R), np.nan)
    U[res["u_labels"][a:b] - 1] = res["u_rows"][a:b]
    return U


def genotype_grid(maxcp=6):
    G = maxcp * (maxcp + 3) // 2
    mu = np.zeros((G, 2))
    assert lib().cso_genotype_grid(maxcp, _p(mu, C.c_double)) == G
    return mu


def genotype_neighbor(cov, allele, maxcp=6):
    f = lib().cso_genotype_neighbor
    f.argtypes = [C.c_double, C.c_double, C.c_int]
    return f(float(cov), float(allele), maxcp)


def gibbs_allele(Xcov, Xall, U0, Z0, sig0_cov, sig0_all, alpha=1.0, beta=1.0, niter=200, seed=200, maxcp=6,
                 rng_mode=RNG_R, ucap_rows=None):
    """Returns dict(Zall, sig_cov_all, sig_allele_all, LL, L0, kt, u_off, u_labels, u_state, u_cov, u_allele)."""
    Xcov = np.ascontiguousarray(Xcov, dtype=np.float64)
    Xall = np.ascontiguousarray(Xall, dtype=np.float64)
    U0 = np.ascontiguousarray(np.atleast_2d(U0), dtype=np.int32)
    Z0 = np.ascontiguousarray(Z0, dtype=np.int32)
    N, R = Xcov.shape
    K0 = U0.shape[0]
    sc = np.ascontiguousarray(np.broadcast_to(np.asarray(sig0_cov, dtype=np.float64), (R,)))
    sa = np.ascontiguousarray(np.broadcast_to(np.asarray(sig0_all, dtype=np.float64), (R,)))
    if ucap_rows is None:
        ucap_rows = niter * 256
    Zall = np.zeros((niter, N), dtype=np.int32)
    sca, saa = np.zeros((niter, R)), np.zeros((niter, R))
    LL, L0 = np.zeros(niter), np.zeros(1)
    kt = np.zeros(niter, dtype=np.int32)
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap_rows, dtype=np.int32)
    u_state = np.zeros((ucap_rows, R), dtype=np.int32)
    u_cov, u_all = np.zeros((ucap_rows, R)), np.zeros((ucap_rows, R))
    rc = lib().cso_gibbs_allele(
        N, R, _p(Xcov, C.c_double), _p(Xall, C.c_double), K0, _p(U0, C.c_int), _p(Z0, C.c_int), _p(sc, C.c_double),
        _p(sa, C.c_double), C.c_double(alpha), C.c_double(beta), niter, C.c_uint32(int(seed)), maxcp, rng_mode,
        _p(Zall, C.c_int), _p(sca, C.c_double), _p(saa, C.c_double), _p(LL, C.c_double), _p(L0, C.c_double),
        _p(kt, C.c_int), C.c_int64(ucap_rows), _p(u_off, C.c_int64), _p(u_labels, C.c_int), _p(u_state, C.c_int),
        _p(u_cov, C.c_double), _p(u_all, C.c_double))
    if rc != 0:
        raise RuntimeError(f"cso_gibbs_allele failed with code {rc}")
    n = int(u_off[-1])
    return dict(Zall=Zall, sig_cov_all=sca, sig_allele_all=saa, LL=LL, L0=float(L0[0]), kt=kt, u_off=u_off,
                u_labels=u_labels[:n], u_state=u_state[:n], u_cov=u_cov[:n], u_allele=u_all[:n])


def majority_vote(Zall):
    Zall = np.ascontiguousarray(Zall, dtype=np.int32)
    T, N = Zall.shape
    out = np.zeros(N, dtype=np.int32)
    lib().cso_majority_vote(T, N, _p(Zall, C.c_int), _p(out, C.c_int))
    return out


def bin_counts(G, N, B, colptr, rowidx, x, map_ptr, map_bin):
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_bin = np.ascontiguousarray(map_bin, dtype=np.int32)
    ocp = np.zeros(N + 1, dtype=np.int64)
    L = lib()
    args = [G, N, B, _p(colptr, C.c_int64), _p(rowidx, C.c_int32), _p(x, C.c_double),
            _p(map_ptr, C.c_int32), _p(map_bin, C.c_int32), _p(ocp, C.c_int64)]
    nnz = L.cso_bin_counts(*args, None, None)
    ori = np.zeros(nnz, dtype=np.int32)
    ox = np.zeros(nnz, dtype=np.float64)
    L.cso_bin_counts(*args, _p(ori, C.c_int32), _p(ox, C.c_double))
    return ocp, ori, ox


def bin_counts_dense(G, N, B, colptr, rowidx, x, map_ptr, map_bin):
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_bin = np.ascontiguousarray(map_bin, dtype=np.int32)
    out = np.zeros((N, B), dtype=np.float64)  # column-major B x N == C-order N x B
    lib().cso_bin_counts_dense(G, N, B, _p(colptr, C.c_int64), _p(rowidx, C.c_int32),
                               _p(x, C.c_double), _p(map_ptr, C.c_int32), _p(map_bin, C.c_int32),
                               _p(out, C.c_double))
    return out.T  # B x N


def _r_num_str(v):
    """as.character() of a numeric scalar the way paste0() prints it (0.99 -> '0.99', 1 -> '1')."""
    f = float(v)
    return str(int(f)) if f == int(f) else repr(f)


def est_region_cov(mtx, barcodes, features, bed, celltype0=None, var_pt=0.99, var_pt_ctrl=0.99, ngene_filter=0,
                   include="tumor", alpha_source="all", ctrl_region=None, seg_table=None, size=None):
    """Restatement of EstRegionCov (reference R/EstRegionCov.R:24-209), dense and line by line;
    test infrastructure only.  `mtx`: G x N array; `bed`: list of 4 columns (chr, start, end, gene id);
    `celltype0`: (barcodes, types) or None; `seg_table`: dict(chr, start, end, chrr) or None with
    `size` = (chromosome names, lengths).  Column sums are accumulated in long double like R's
    colSums.  Returns dict(deltas_all=[(name, cell names, values)], ngenes, sel, alpha, barcodes_normal)."""
    mtx = np.asarray(mtx, dtype=np.float64)
    barcodes = [str(b) for b in barcodes]
    features = [str(f) for f in features]
    bchr = [str(c) for c in bed[0]]
    if "chr" not in bchr[0]:                                            # :28-30
        bchr = ["chr" + c for c in bchr]
    bstart = np.asarray(bed[1], dtype=np.float64)
    bend = np.asarray(bed[2], dtype=np.float64)
    bgene = [str(g) for g in bed[3]]
    if celltype0 is None:                                               # :32-34
        celltype0 = (list(barcodes), ["normal"] * len(barcodes))
    ct_bar = [str(b) for b in celltype0[0]]
    ct_type = np.asarray([str(t) for t in celltype0[1]])
    col_of = {b: i for i, b in reversed(list(enumerate(barcodes)))}     # match(): first occurrence
    mtx = mtx[:, [col_of[b] for b in ct_bar]]                           # :40
    N = mtx.shape[1]
    mean = mtx.astype(np.longdouble).sum(axis=1) / N                    # :42 var(): two passes, long double
    rna_var = np.asarray((((mtx.astype(np.longdouble) - mean[:, None]) ** 2).sum(axis=1) / (N - 1)), dtype=np.float64)
    ngenes_row = np.asarray(mtx.astype(np.longdouble).sum(axis=1), dtype=np.float64)
    q = np.quantile(rna_var, var_pt)                                    # :47 type 7
    qc = np.quantile(rna_var, var_pt_ctrl)
    keep = np.nonzero((rna_var < q) & (rna_var != 0) & (ngenes_row > ngene_filter))[0]
    keep_c = np.nonzero((rna_var < qc) & (rna_var != 0) & (ngenes_row > ngene_filter))[0]
    rna, rna_control = mtx[keep], mtx[keep_c]
    first = {}
    for i, g in enumerate(bgene):
        first.setdefault(g, i)
    sub_chr, sub_s, sub_e = [], [], []
    for k in keep:                                                      # :51-52 (unmatched genes: all-zero row)
        j = first.get(features[k])
        sub_chr.append(bchr[j] if j is not None else "0")
        sub_s.append(bstart[j] if j is not None else 0.0)
        sub_e.append(bend[j] if j is not None else 0.0)
    sub_chr = np.asarray(sub_chr)
    sub_s, sub_e = np.asarray(sub_s), np.asarray(sub_e)
    est_name = f"pt{_r_num_str(var_pt)}_{include}_alpha{alpha_source}"  # :60
    if seg_table is None:                                               # :63-67
        names = [str(s) for s in size[0]][:22]
        seg_table = dict(chr=[n.replace("chr", "") for n in names], start=[0.0] * len(names),
                         end=[float(v) for v in list(size[1])[:22]], chrr=names)
    seg_chr = [str(c) for c in seg_table["chr"]]
    seg_s = np.asarray(seg_table["start"], dtype=np.float64)
    seg_e = np.asarray(seg_table["end"], dtype=np.float64)
    seg_name = [str(c) for c in seg_table["chrr"]]
    out, ngenes, sel = [], [], []
    alphas = None
    barcodes_normal = None
    for chrr in seg_name:                                               # :80
        gene_ind = []
        for r in [i for i, n in enumerate(seg_name) if n == chrr]:      # findOverlaps, type "any"
            hit = np.nonzero((sub_chr == "chr" + seg_chr[r]) & (sub_s <= seg_e[r]) & (sub_e >= seg_s[r]) & (sub_s <= sub_e))[0]
            gene_ind.extend(hit.tolist())
        rna_sub = rna[gene_ind]
        colsum = np.asarray(rna_sub.astype(np.longdouble).sum(axis=0), dtype=np.float64)
        sel_cell = np.nonzero(colsum > 0)[0]                            # :93
        if len(sel_cell) == 0:
            continue
        sel.append(chrr)
        rna_sub = rna_sub[:, sel_cell]
        types = ct_type[sel_cell]
        normal = np.nonzero(types == "normal")[0]
        tumor = np.nonzero(types == "tumor")[0]
        barcodes_normal = [ct_bar[i] for i in sel_cell[normal]]
        control = rna_sub[:, normal]
        if include == "tumor":
            tsel = tumor
        elif include == "control":
            tsel = normal
        elif include == "all":
            tsel = np.arange(len(sel_cell))
        else:
            raise ValueError("include variable is not valid!")
        test = rna_sub[:, tsel]
        if alpha_source == "all":                                       # :131-133
            alphas = np.asarray(rna_control[:, sel_cell].astype(np.longdouble).sum(axis=0), dtype=np.float64)
        elif alpha_source == "control":
            rows = np.nonzero(np.isin(sub_chr, list(ctrl_region)))[0]
            alphas = np.asarray(rna[rows][:, sel_cell].astype(np.longdouble).sum(axis=0), dtype=np.float64)
        else:
            raise ValueError("alpha_source is not valid!")
        med = np.median(alphas)
        a_ctrl = alphas[normal] / med                                   # :141
        a_test = alphas[tsel] / med
        with np.errstate(divide="ignore", invalid="ignore"):
            tsum = np.asarray((test / a_test[None, :]).astype(np.longdouble).sum(axis=0), dtype=np.float64)
            csum = np.asarray((control / a_ctrl[None, :]).astype(np.longdouble).sum(axis=0), dtype=np.float64)
            deltas = tsum / (np.median(csum) if len(csum) else np.nan)  # :159
        out.append((est_name + chrr, [ct_bar[i] for i in sel_cell[tsel]], deltas))
        ngenes.append((est_name + chrr, len(gene_ind)))
    return dict(deltas_all=out, ngenes=ngenes, sel=sel, alpha=alphas, barcodes_normal=barcodes_normal)


def assign_cluster(Zall, kt_last, X, cna_states_WGS, mincell=20, cutoff=0.2, cap_corr=1.5):
    """Restatement of AssignCluster, coverage branch (reference R/AssignCluster.R:22-208); test
    infrastructure only.  `Zall`: T x N trimmed labels; `kt_last` = nrow(Uall[[T]]); `X`: N x R.
    Returns dict(Zest, corrs, annot, Uest, sigmas_est, Zmaj, cutoff)."""
    L = lib()
    Zall = np.asarray(Zall)
    X = np.asarray(X, dtype=np.float64)
    N, R = X.shape
    g = MT()
    L.cso_set_seed(C.byref(g), C.c_uint32(100))                          # :36
    zmaj = majority_vote(Zall)                                           # :37
    Zest = zmaj.astype(np.float64)
    labs, cnts = np.unique(zmaj, return_counts=True)
    ind_clusters = labs[cnts > mincell]                                  # :41
    ind_cell = np.nonzero(~np.isin(zmaj, ind_clusters))[0]               # :42
    Zest[ind_cell] = np.nan

    def means():
        U = np.full((kt_last, R), np.nan)
        for kk in ind_clusters:                                          # :46-51 colMeans(na.rm=T), long double
            rows = X[Zest == kk].astype(np.longdouble)
            ok = ~np.isnan(rows)
            cnt = ok.sum(axis=0)
            s = np.where(ok, rows, 0).sum(axis=0)
            with np.errstate(divide="ignore", invalid="ignore"):
                m = np.asarray(s / cnt, dtype=np.float64)
            m[cnt == 0] = 0.0
            U[kk - 1] = m
        return U

    Usub = means()
    kept = np.setdiff1d(np.arange(N), ind_cell)
    d = (X[kept] - Usub[Zest[kept].astype(int) - 1]).astype(np.longdouble) ** 2   # :53
    sig = np.sqrt(1.0 / (N - len(ind_cell)) * np.asarray(np.where(np.isnan(d), 0, d).sum(axis=0), dtype=np.float64))
    for ii in ind_cell:                                                  # :56-64
        P = np.full(kt_last, np.nan)
        for kk in ind_clusters:
            P[kk - 1] = total_loglik(X[ii:ii + 1], Usub[kk - 1:kk], np.array([1], dtype=np.int32), sig)
        P[ind_clusters - 1] -= np.nanmax(P[ind_clusters - 1])
        P[np.isnan(P)] = np.nanmin(P) - 100
        P = np.exp(P)
        perm = np.zeros(kt_last, dtype=np.int32)
        p = np.ascontiguousarray(P, dtype=np.float64)
        Zest[ii] = L.cso_sample_int_prob(C.byref(g), kt_last, _p(p, C.c_double), _p(perm, C.c_int))
    Usub = means()                                                       # :67-72
    z = Zest.astype(int)
    d = (X - Usub[z - 1]).astype(np.longdouble) ** 2
    sig = np.sqrt(1.0 / N * np.asarray(np.where(np.isnan(d), 0, d).sum(axis=0), dtype=np.float64))   # :77
    priors = np.asarray(cna_states_WGS, dtype=np.float64)
    if np.all(priors == 1):                                              # :85-88
        priors = np.full(R, 1.5)
    U2 = np.maximum(np.minimum(Usub, cap_corr), 0.5)                     # :106-107
    corrs = np.zeros(kt_last)
    for k in range(kt_last):                                             # :127-150
        x = U2[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            tmp = ((x - 1) * (priors - 1)) / (np.sqrt(np.sum((x - 1) ** 2)) * np.sqrt(np.sum((priors - 1) ** 2)))
        tmp = np.sort(tmp[~np.isnan(tmp)])                               # sort() drops NA
        s = 0.0
        for v in tmp:                                                    # sum() of the sorted terms (long double)
            s = np.longdouble(s) + v
        corrs[k] = float(s) if len(tmp) and not np.any(np.isnan(x)) else 0.0
    annot = np.where(corrs > cutoff, "T", "N")
    return dict(Zest=z, corrs=corrs[z - 1], annot=annot[z - 1], Uest=Usub, sigmas_est=sig, Zmaj=zmaj, cutoff=cutoff)


def assign_cluster_allele(Zall, kt_last, Xcov, Xall, cna_states_WGS, mincell=20, cutoff=0.2, cap_corr=1.5,
                          rm_extreme=False):
    """Restatement of AssignCluster, allele branch (reference R/AssignCluster.R:216-414); test
    infrastructure only.  PARITY UNPINNED: the reference ships no AssignCluster(allele=TRUE) output.
    R's `x[-ind_cell, ]` with an empty `ind_cell` selects no row, so without dissolved clusters the
    returned sigmas are 0 — kept as the reference behaves."""
    L = lib()
    Zall = np.asarray(Zall)
    Xcov, Xall = np.asarray(Xcov, dtype=np.float64), np.asarray(Xall, dtype=np.float64)
    N, R = Xcov.shape
    g = MT()
    L.cso_set_seed(C.byref(g), C.c_uint32(100))                          # :231
    zmaj = majority_vote(Zall)
    Zest = zmaj.astype(np.float64)
    labs, cnts = np.unique(zmaj, return_counts=True)
    ind_clusters = labs[cnts > mincell]                                  # :236
    ind_cell = np.nonzero(~np.isin(zmaj, ind_clusters))[0]
    Zest[ind_cell] = np.nan
    kept = np.setdiff1d(np.arange(N), ind_cell) if len(ind_cell) else np.zeros(0, dtype=int)   # x[-integer(0)] is empty

    def means(X):
        U = np.full((kt_last, R), np.nan)
        for kk in ind_clusters:
            rows = X[Zest == kk].astype(np.longdouble)
            ok = ~np.isnan(rows)
            cnt = ok.sum(axis=0)
            s = np.where(ok, rows, 0).sum(axis=0)
            with np.errstate(divide="ignore", invalid="ignore"):
                m = np.asarray(s / cnt, dtype=np.float64)
            m[cnt == 0] = 0.0
            U[kk - 1] = m
        return U

    def sigma(X, U, n):
        d = (X[kept] - U[Zest[kept].astype(int) - 1]).astype(np.longdouble) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.sqrt(1.0 / n * np.asarray(np.where(np.isnan(d), 0, d).sum(axis=0), dtype=np.float64))

    Uc, Ua = means(Xcov), means(Xall)                                    # :240-250
    notorph = np.ones(N, dtype=bool)
    notorph[ind_cell] = False
    ncov = (~np.isnan(Xcov) & notorph[:, None]).sum(axis=0)              # :252-256
    nall = (~np.isnan(Xall) & notorph[:, None]).sum(axis=0)
    sc, sa = sigma(Xcov, Uc, ncov), sigma(Xall, Ua, nall)                # :258-259
    one = np.array([1], dtype=np.int32)
    for ii in ind_cell:                                                  # :265-275
        P = np.full(kt_last, np.nan)
        for kk in ind_clusters:
            P[kk - 1] = total_loglik(Xcov[ii:ii + 1], Uc[kk - 1:kk], one, sc) + \
                total_loglik(Xall[ii:ii + 1], Ua[kk - 1:kk], one, sa)
        P[ind_clusters - 1] -= np.max(P[ind_clusters - 1])
        P[np.isnan(P)] = np.nanmin(P) - 100
        p = np.ascontiguousarray(np.exp(P), dtype=np.float64)
        perm = np.zeros(kt_last, dtype=np.int32)
        Zest[ii] = L.cso_sample_int_prob(C.byref(g), kt_last, _p(p, C.c_double), _p(perm, C.c_int))
    Uc, Ua = means(np.minimum(Xcov, 3.0)), means(Xall)                   # :281-294 (pmin keeps NA)
    Ug = np.full((kt_last, R), np.nan)
    for kk in ind_clusters:
        for r in range(R):
            Ug[kk - 1, r] = genotype_neighbor(Uc[kk - 1, r], Ua[kk - 1, r], 6)
    ncov, nall = (~np.isnan(Xcov)).sum(axis=0), (~np.isnan(Xall)).sum(axis=0)               # :302-303
    sc, sa = sigma(Xcov, Uc, ncov), sigma(Xall, Ua, nall)                # :305-306
    priors = np.asarray(cna_states_WGS, dtype=np.float64)
    if np.all(priors == 4):
        priors = np.full(R, 8.0)
    priorsc = genotype_grid(6)[priors.astype(int) - 1, 0]                # :333
    U2 = np.maximum(np.minimum(Uc, cap_corr), 0.5)
    corrs = np.zeros(kt_last)
    for k in range(kt_last):                                             # :342-353
        tmp = (U2[k] - 1) * (priorsc - 1)
        tmp = sorted(v for v in tmp if v == v)
        if not tmp:
            continue
        if rm_extreme:
            idx = list(range(2, len(tmp))) if len(tmp) >= 3 else list(range(2, len(tmp) - 2, -1))   # 2:(n-1), 1-based
            tmp = [tmp[i - 1] if 1 <= i <= len(tmp) else np.nan for i in idx if i != 0]
        s = np.longdouble(0)
        for v in tmp:
            s += v
        corrs[k] = 0.0 if s != s else float(s)
    corrs = corrs / corrs.max()                                          # :358
    z = Zest.astype(int)
    annot = np.where(corrs > cutoff, "T", "N")
    return dict(Zest=z, corrs=corrs[z - 1], annot=annot[z - 1], Uest=Ug, Ucest=Uc, Uaest=Ua, sigmas_cov_est=sc,
                sigmas_allele_est=sa, Zmaj=zmaj, cutoff=cutoff)
