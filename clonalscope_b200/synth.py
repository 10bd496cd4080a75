"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d) (numpy PCG64, fixed seeds).

Used by bench.py and the parity tests; generation only, no arithmetic of the hot path.
"""
from __future__ import annotations

import numpy as np

# chromosome lengths (GRCh38 autosomes, bp) — for the synthetic geometry
CHR_SIZES = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
             138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
             83257441, 80373285, 58617616, 64444167, 46709983, 50818468]


def counts_csc(N, G=30000, seed=1001, mean_genes=2000, vmax=32767):
    """Sparse gene x cell count matrix in CSC slots (colptr int32, rowidx int32, x float64).

    Per-gene weight LogNormal(0, 1.5); per-cell number of expressed genes
    clip(round(LogNormal(ln mean_genes, 0.4)), 200, 8000) capped at G; genes drawn without
    replacement proportional to the weights (Gumbel top-k), ascending within a column;
    values Geometric(0.6) capped at vmax.
    """
    rng = np.random.default_rng(seed)
    w = rng.lognormal(0.0, 1.5, size=G)
    logw = np.log(w)
    m = np.clip(np.rint(rng.lognormal(np.log(mean_genes), 0.4, size=N)), min(200, G), min(8000, G)).astype(np.int64)
    colptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(m, out=colptr[1:])
    nnz = int(colptr[-1])
    rowidx = np.empty(nnz, dtype=np.int32)
    chunk = max(1, min(N, (1 << 26) // G))
    for c0 in range(0, N, chunk):
        c1 = min(N, c0 + chunk)
        keys = logw[None, :] + rng.gumbel(size=(c1 - c0, G))
        mmax = int(m[c0:c1].max())
        part = np.argpartition(-keys, mmax - 1, axis=1)[:, :mmax] if mmax < G else np.tile(np.arange(G), (c1 - c0, 1))
        pk = np.take_along_axis(keys, part, axis=1)
        order = np.argsort(-pk, axis=1)
        top = np.take_along_axis(part, order, axis=1)
        for i in range(c1 - c0):
            sel = np.sort(top[i, : m[c0 + i]])
            rowidx[colptr[c0 + i]:colptr[c0 + i + 1]] = sel
    x = np.minimum(rng.geometric(0.6, size=nnz), vmax).astype(np.float64)
    return colptr.astype(np.int32 if nnz < 2**31 else np.int64), rowidx, x


def geometry(G=30000, bin_size=200000, seed=1002):
    """Synthetic gene spans on 22 autosomes and 200-kb tiles.  Returns (bins dict, genes dict)."""
    rng = np.random.default_rng(seed)
    sizes = np.array(CHR_SIZES, dtype=np.int64)
    chrom = rng.choice(22, size=G, p=sizes / sizes.sum())
    length = np.maximum(200, rng.lognormal(np.log(25000.0), 1.2, size=G)).astype(np.int64)
    start = (rng.random(G) * np.maximum(1, sizes[chrom] - length)).astype(np.int64) + 1
    end = np.minimum(start + length, sizes[chrom])
    order = np.lexsort((start, chrom))  # GTF-like order
    chrom, start, end = chrom[order], start[order], end[order]
    bchr, bstart, bend = [], [], []
    for c in range(22):
        s = np.arange(0, sizes[c], bin_size, dtype=np.int64)
        bchr.append(np.full(len(s), c + 1))
        bstart.append(s)
        bend.append(np.minimum(s + bin_size, sizes[c]))
    bins = dict(chr=np.concatenate(bchr), start=np.concatenate(bstart), end=np.concatenate(bend))
    genes = dict(chr=chrom + 1, start=start, end=end)
    return bins, genes


def geometry_map(bins, genes):
    """gene -> bin CSR overlap map for tiles (bin interval [start+1, end], gene [start, end])."""
    from .intervals import overlap_map

    bchr = [str(int(c)) for c in bins["chr"]]
    gseq = ["chr" + str(int(c)) for c in genes["chr"]]
    return overlap_map(bchr, list(bins["start"]), list(bins["end"]), gseq, np.asarray(genes["start"]),
                       np.asarray(genes["end"]))


def sampler_input(N, R=300, seed=1003, K_true=12, na_frac=0.0):
    """Planted-cluster coverage matrix and the priors the bench uses."""
    rng = np.random.default_rng(seed)
    w = rng.dirichlet(np.ones(K_true))
    z = rng.choice(K_true, size=N, p=w)
    prof = np.ones((K_true, R))
    for k in range(K_true):
        f = rng.uniform(0.1, 0.3)
        msk = rng.random(R) < f
        prof[k, msk] = rng.choice([0.5, 1.5, 2.0], size=int(msk.sum()))
    X = np.clip(prof[z] + 0.2 * rng.standard_normal((N, R)), 0.05, 3.0)
    if na_frac > 0:
        X[rng.random((N, R)) < na_frac] = np.nan
    big = int(np.bincount(z, minlength=K_true).argmax())
    row = np.minimum(prof[big], 1.5)
    U0 = np.vstack([np.ones(R), row])
    return dict(X=X, z_true=z, U0=U0, sigmas0=np.full(R, 0.25), alpha=2.0, beta=2.0, seed=200)
