"""Host-side integer interval logic of the gene -> bin aggregation.

Mirrors the non-arithmetic part of `Gen_bin_cell_rna_filtered`
(reference R/Gene_bin_cell_rna_filtered.R:28-64): bin sorting, gene coordinate lookup
and the (bin, gene) overlap table that `findOverlaps(query, subject)` returns
(GenomicRanges default type="any": closed 1-based integer intervals, equal seqnames,
strand ignored).  The table is handed to the device as a CSR map gene -> bins.
"""
from __future__ import annotations

import math

import numpy as np


def r_as_character(v) -> str:
    """as.character() of one R scalar: integers print plainly, doubles with up to 15
    significant digits in fixed or scientific notation, whichever is narrower (fixed on
    ties) — e.g. 6e+05, 1200000, 0.5 — which is what paste0() applies to numeric columns."""
    if isinstance(v, (str, np.str_)):
        return str(v)
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    f = float(v)
    if math.isnan(f):
        return "NA"
    if math.isinf(f):
        return "Inf" if f > 0 else "-Inf"
    if f == 0:
        return "0"
    for digits in range(1, 16):
        s = f"{f:.{digits - 1}e}"
        if float(s) == f:
            break
    else:
        digits = 15
        s = f"{f:.14e}"
    mant, exp = s.split("e")
    e = int(exp)
    mant = mant.rstrip("0").rstrip(".") if "." in mant else mant
    sci = f"{mant}e{'+' if e >= 0 else '-'}{abs(e):02d}"
    nsig = len(mant.replace("-", "").replace(".", ""))
    decimals = max(0, nsig - 1 - e)
    fixed = f"{f:.{decimals}f}"
    return fixed if len(fixed) <= len(sci) else sci


def _as_numeric(col):
    """as.numeric() of a character/numeric column; non-numeric strings become NaN."""
    out = np.empty(len(col), dtype=np.float64)
    for i, v in enumerate(col):
        try:
            out[i] = float(v)
        except (TypeError, ValueError):
            out[i] = np.nan
    return out


def sort_bins(chrom, start, end):
    """:30-37 — strip a 'chr' prefix (decided from the first row only) and order by
    (as.numeric(chr), as.numeric(start)); non-numeric chromosomes sort last (NA)."""
    chrom = [str(c) for c in chrom]
    if len(chrom) and "chr" in chrom[0]:
        chrom = [c.replace("chr", "") for c in chrom]
    cnum = _as_numeric(chrom)
    snum = _as_numeric(start)
    ckey = np.where(np.isnan(cnum), np.inf, cnum)
    skey = np.where(np.isnan(snum), np.inf, snum)
    order = np.lexsort((skey, ckey))  # stable, like R's radix order
    chrom = [chrom[i] for i in order]
    start = [start[i] for i in order]
    end = [end[i] for i in order]
    return chrom, start, end, order


def gene_sites(gene_ids, gtf):
    """:43-60 — per gene (seqname, start, end) from the GTF-like table; unknown genes get
    the reference's "0-0-0" placeholder (seqname 'chr0', [0, 0])."""
    cols = {k: np.asarray(gtf[k]) for k in ("V1", "V4", "V5")}
    if "gene_id" in gtf and gtf["gene_id"] is not None:
        ensg = [str(x) for x in gtf["gene_id"]]
        keep = np.arange(len(ensg))
    else:
        keep = np.nonzero(np.asarray(gtf["V3"]) == "gene")[0]
        ensg = []
        for s in np.asarray(gtf["V9"])[keep]:
            first = str(s).split(";")[0]
            parts = first.split("gene_id ")
            ensg.append(parts[1] if len(parts) > 1 else None)
    lookup = {}
    for j, e in zip(keep, ensg):  # chr[genes[,1]] takes the first match by name
        if e is not None and e not in lookup:
            lookup[e] = j
    n = len(gene_ids)
    seq = ["0"] * n
    gs = np.zeros(n, dtype=np.int64)
    ge = np.zeros(n, dtype=np.int64)
    for i, g in enumerate(gene_ids):
        j = lookup.get(str(g))
        if j is not None:
            seq[i] = str(cols["V1"][j])
            gs[i] = int(cols["V4"][j])
            ge[i] = int(cols["V5"][j])
    # :56-60 the 'chr' prefix is decided from the first gene only
    site0 = f"{seq[0]}-{gs[0]}-{ge[0]}" if n else ""
    if "chr" not in site0:
        seq = ["chr" + s for s in seq]
    return seq, gs, ge


def overlap_map(bin_chr, bin_start, bin_end, gene_seq, gene_start, gene_end):
    """findOverlaps(query=bins [start+1, end], subject=genes [start, end]) as a CSR map
    gene -> ascending bin indices.  `bin_chr` are the stripped names (seqname 'chr'+name)."""
    B, G = len(bin_chr), len(gene_seq)
    bs1 = np.asarray(_as_numeric(bin_start), dtype=np.float64) + 1
    be = np.asarray(_as_numeric(bin_end), dtype=np.float64)
    by_seq = {}
    for i, c in enumerate(bin_chr):
        by_seq.setdefault("chr" + c, []).append(i)
    counts = np.zeros(G, dtype=np.int64)
    hits = [None] * G
    genes_by_seq = {}
    for g, s in enumerate(gene_seq):
        genes_by_seq.setdefault(s, []).append(g)
    for s, glist in genes_by_seq.items():
        if s not in by_seq:
            continue
        idx = np.asarray(by_seq[s], dtype=np.int64)
        o = np.argsort(bs1[idx], kind="stable")
        idx = idx[o]
        s1, e1 = bs1[idx], be[idx]
        glist = np.asarray(glist, dtype=np.int64)
        gs, ge = gene_start[glist].astype(np.float64), gene_end[glist].astype(np.float64)
        valid = gs <= ge  # IRanges would reject start > end + 1; treat as empty
        if np.all(np.diff(e1) >= 0):
            # sorted tiles: bins with start+1 <= ge form a prefix, those with end >= gs a suffix
            hi = np.searchsorted(s1, ge, side="right")
            lo = np.searchsorted(e1, gs, side="left")
            for k, g in enumerate(glist):
                if valid[k] and hi[k] > lo[k]:
                    hits[g] = np.sort(idx[lo[k]:hi[k]])
        else:
            for k, g in enumerate(glist):
                if not valid[k]:
                    continue
                m = (s1 <= ge[k]) & (e1 >= gs[k])
                if m.any():
                    hits[g] = np.sort(idx[m])
    for g in range(G):
        if hits[g] is not None:
            counts[g] = len(hits[g])
    map_ptr = np.zeros(G + 1, dtype=np.int32)
    np.cumsum(counts, out=map_ptr[1:])
    map_bin = np.zeros(int(map_ptr[-1]), dtype=np.int32)
    for g in range(G):
        if hits[g] is not None:
            map_bin[map_ptr[g]:map_ptr[g + 1]] = hits[g]
    return map_ptr, map_bin
