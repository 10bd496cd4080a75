"""CPU restatement of the reference's segmentation path — TEST INFRASTRUCTURE ONLY.

Follows, statement by statement and with plain Python loops,
  * R/CreateSegtableNoWGS.R:62-69,77  bin x cluster pooling (`bin_by_cluster`, `ref_counts`)
  * R/Segmentation_bulk.R:72-203      pooled coverage -> running mean -> 4-state HMM -> seg table
  * R/CreateSegtableNoWGS.R:107-166   union of the clusters' breakpoints (`seg_filtered`)
Nothing in the product package may import this module (only tests/, smoke(), bench.py's CPU leg).

Third-party algorithms that are not in the reference tree (restated from their published
behaviour, versions as on CRAN when the fixtures were written):
  * caTools::runmean(x, k, alg="C", endrule="mean", align="center") (caTools 1.18): k <- min(k, n);
    window of k values with k %/% 2 of them to the right of the centre, shrinking at both ends;
    the C code carries a running sum with an error term.
  * HiddenMarkov::Viterbi.dthmm (HiddenMarkov 1.8): nu[1,] = log(delta) + log f(x1);
    nu[i,k] = max_j(nu[i-1,j] + log Pi[j,k]) + log f_k(x_i); y[n] = which.max(nu[n,]);
    y[i] = which.max(log Pi[, y[i+1]] + nu[i,]) (first maximum).
  * stats::quantile type 7, stats::median, stats::var.

PARITY: the union step is PINNED by the reference's shipped
samples/P5847/scRNA/noWGS_output/{Obj_all_cluster,seg_filtered_from_cluster}.rds
(tests/golden/segtable_P5847.npz).  Segmentation_bulk itself is UNPINNED by reference data (the
bin x cell counts behind the shipped seg tables are in .MISSING_LARGE_BLOBS).
"""
from __future__ import annotations

import math

import numpy as np

LN_SQRT_2PI = 0.918938533204672741780329736406


def r_chr(x):
    """as.character() of an R double: 15 significant digits, fixed unless scientific is shorter."""
    x = float(x)
    if x != x:
        return "NA"
    if x == math.inf:
        return "Inf"
    if x == -math.inf:
        return "-Inf"
    if x == 0:
        return "0"
    for nd in range(1, 16):
        s = f"{x:.{nd - 1}e}"
        if float(s) == x:
            break
    else:
        nd = 15
        s = f"{x:.14e}"
    mant, ex = s.split("e")
    ex = int(ex)
    digits = mant.replace(".", "").replace("-", "").rstrip("0") or "0"
    neg = x < 0
    sci_m = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
    sci = f"{sci_m}e{'+' if ex >= 0 else '-'}{abs(ex):02d}"
    if ex >= 0:
        fixed = digits + "0" * (ex + 1 - len(digits)) if len(digits) <= ex + 1 else digits[:ex + 1] + "." + digits[ex + 1:]
    else:
        fixed = "0." + "0" * (-ex - 1) + digits
    out = fixed if len(fixed) <= len(sci) else sci  # R: fixed unless wider than scientific (scipen = 0)
    return ("-" if neg else "") + out


def pool_by_label(B, colptr, rowidx, x, label, K):
    """bin_by_cluster (R/CreateSegtableNoWGS.R:62-69): rowSums over the cells of each label."""
    out = np.zeros((B, K))
    for c in range(len(colptr) - 1):
        k = label[c]
        if k < 0 or k >= K:
            continue
        for e in range(colptr[c], colptr[c + 1]):
            out[rowidx[e], k] += x[e]
    return out


def runmean(x, k):
    """caTools::runmean, alg "C", endrule "mean", centred (see the module docstring)."""
    x = [float(v) for v in x]
    n = len(x)
    if k <= 1:
        return list(x)
    if k > n:
        k = n
    k2 = k // 2
    out = [0.0] * n
    state = {"sum": 0.0, "err": 0.0, "num": 0}

    def add(y, d):
        if math.isfinite(y):
            t = state["sum"] + y
            state["err"] += (state["sum"] - t) + y if abs(state["sum"]) >= abs(y) else (y - t) + state["sum"]
            state["sum"] = t
            state["num"] += d

    o = 0
    for i in range(k2):                      # step 1
        add(x[i], 1)
    for i in range(k2, k):                   # step 2: left edge, window grows
        add(x[i], 1)
        out[o] = (state["sum"] + state["err"]) / state["num"] if state["num"] else math.nan
        o += 1
    for i in range(k, n):                    # step 3: full windows
        add(x[i], 1)
        add(-x[i - k], -1)
        out[o] = (state["sum"] + state["err"]) / state["num"] if state["num"] else math.nan
        o += 1
    for i in range(k2):                      # step 4: right edge, window shrinks
        add(-x[n - k + i], -1)
        out[o] = (state["sum"] + state["err"]) / state["num"] if state["num"] else math.nan
        o += 1
    return out


def viterbi(x, means, sds, delta, Pi):
    """HiddenMarkov::Viterbi for a Gaussian dthmm; returns 1-based states."""
    n, m = len(x), len(means)
    logPi = [[math.log(Pi[j][k]) for k in range(m)] for j in range(m)]

    def dens(v, a):
        z = (v - means[a]) / sds[a]
        return -(LN_SQRT_2PI + 0.5 * z * z + math.log(sds[a]))

    nu = [[0.0] * m for _ in range(n)]
    for a in range(m):
        nu[0][a] = math.log(delta[a]) + dens(x[0], a)
    for i in range(1, n):
        for k in range(m):
            nu[i][k] = max(nu[i - 1][j] + logPi[j][k] for j in range(m)) + dens(x[i], k)
    if any(v == -math.inf for v in nu[n - 1]):
        raise FloatingPointError("Problems With Underflow")
    y = [0] * n
    y[n - 1] = max(range(m), key=lambda a: (nu[n - 1][a], -a))
    for i in range(n - 2, -1, -1):
        y[i] = max(range(m), key=lambda a: (logPi[a][y[i + 1]] + nu[i][a], -a))
    return [v + 1 for v in y]


def quantile7(v, p):
    s = sorted(v)
    h = (len(s) - 1) * p
    lo = math.floor(h)
    hi = min(lo + 1, len(s) - 1)
    return s[lo] + (h - lo) * (s[hi] - s[lo])


def median(v):
    s = sorted(v)
    n = len(s)
    return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])


def r_mean(vals):
    """mean() of R (summary.c): long double sum / n, then one refinement pass over the residuals."""
    ld = np.longdouble
    n = len(vals)
    s = ld(0)
    for v in vals:
        s += ld(v)
    s /= n
    t = ld(0)
    for v in vals:
        t += ld(v) - s
    return float(s + t / n)


def r_var(vals):
    """var() of R (cov.c, one column): mean as above, then sum of squared residuals / (n - 1)."""
    ld = np.longdouble
    n = len(vals)
    if n < 2:
        return math.nan
    s = ld(0)
    for v in vals:
        s += ld(v)
    s /= n
    t = ld(0)
    for v in vals:
        t += ld(v) - s
    m = s + t / n
    q = ld(0)
    for v in vals:
        q += (ld(v) - m) * (ld(v) - m)
    return float(q / (n - 1))


def r_colon(a, b):
    """a:b of R (descending when b < a), 1-based inclusive."""
    return list(range(a, b + 1)) if b >= a else list(range(a, b - 1, -1))


def segmentation_bulk(bin_chr, bin_start, bin_end, raw, ref, size, hmm_states=(0.5, 1.5, 1.8), hmm_sd=0.2,
                      hmm_p=0.000001, nmean=100, adj=0.0, max_qt=0.99, rm_extreme=1, assay="WGS"):
    """R/Segmentation_bulk.R:72-218 for one pooled profile.  `bin_chr`: chromosome numbers (ints) in
    sorted order; `bin_start`/`bin_end`: the numbers parsed from the row names; `size`: dict
    chromosome number -> length.  Returns (seg_table: dict of string columns, annot)."""
    if len(hmm_states) != 3:
        raise ValueError("hmm_states should be an ordered vector for the three HMM numeric states.")
    if assay == "scRNA":
        adj = -0.5
    n = len(raw)
    cov2 = []
    for i in range(n):                                         # :101-103
        a, b = float(raw[i]), float(ref[i])
        if b == 0.0:
            v = math.nan if a == 0.0 else math.inf
        else:
            v = a / b
        if v != v:
            v = 0.0
        if v == math.inf:
            v = 100.0
        cov2.append(v)
    q = quantile7(cov2, max_qt)
    cov3 = [min(v, q) for v in cov2]                           # :114
    med = median(cov3)
    cov4 = [v / med for v in cov3]
    if rm_extreme == 1:                                        # :118-130
        sel = [i for i in range(n) if cov4[i] > 0.4]
    elif rm_extreme == 2:
        sel = [i for i in range(n) if cov4[i] > 0.4 and cov2[i] != 100.0]
    else:
        raise NameError("object 'sel_bin' not found")
    cov3 = [cov3[i] for i in sel]
    chromnum = [bin_chr[i] for i in sel]
    raw_start = [bin_start[i] for i in sel]
    raw_end = [bin_end[i] for i in sel]
    med = median(cov3)                                         # (:151 recomputes it over the kept bins)
    chroms = sorted(size)
    ppa1, ppa2, ppn, ppd = hmm_states[2], hmm_states[1], 1.0, hmm_states[0]
    t = hmm_p
    Pi = [[1 - 3 * t if j == k else t for k in range(4)] for j in range(4)]
    delta = [0.1, 0.2, 0.5, 0.2]
    rows = []
    for ii in chroms:                                          # :150
        idx = [i for i in range(len(sel)) if chromnum[i] == ii]
        c4 = [cov3[i] / med + adj for i in idx]
        c5 = runmean(c4, nmean)
        res = viterbi(c5, [ppa1, ppa2, ppn, ppd], [hmm_sd] * 4, delta, Pi)
        states = [[ppa1, ppa2, ppn, ppd][r - 1] for r in res]
        start = [raw_start[i] - 1 for i in idx]
        cps = [xx for xx in range(2, len(states) + 1) if states[xx - 1] != states[xx - 2]]   # 1-based :186
        cps = [c for c in cps if c != 2]
        nseg = len(cps) + 1
        seg_start = [start[0]] + [start[c - 1] for c in cps]
        seg_end = [start[c - 1] for c in cps] + [size[ii]]
        seg_state = [states[1]] + [states[c - 1] for c in cps]
        edges = [1] + cps + [len(c4)]
        for s in range(nseg):
            sel_i = r_colon(edges[s], edges[s + 1] - 1)
            vals = [c4[j - 1] for j in sel_i]
            mean, var = r_mean(vals), r_var(vals)
            rows.append((str(ii), seg_start[s], seg_end[s], seg_state[s], seg_end[s] - seg_start[s], mean, var))
    tab = dict(chr=[r[0] for r in rows], start=[r_chr(r[1]) for r in rows], end=[r_chr(r[2]) for r in rows],
               states=[r_chr(r[3]) for r in rows], length=[r_chr(r[4]) for r in rows],
               mean=[r_chr(r[5]) for r in rows], var=[r_chr(r[6]) for r in rows])
    annot = None
    if assay == "scRNA":                                       # :208-216
        if median([float(v) for v in tab["states"]]) == 0.5:
            tab["states"] = [r_chr(float(v) + 0.5) for v in tab["states"]]
            annot = "N"
        else:
            annot = "T"
    tab["chrr"] = [f"{c}:{s}" for c, s in zip(tab["chr"], tab["start"])]
    return tab, annot


def merge_segments(tables, merge=False):
    """R/CreateSegtableNoWGS.R:107-166: rbind the clusters' seg tables, order, drop duplicates, cut every
    chromosome at the union of the breakpoints and label each piece with the sorted distinct states of the
    segments it overlaps.  `tables`: list of dicts of string columns.  Returns a list of 4-string rows."""
    allrows = []
    for tab in tables:
        for i in range(len(tab["chr"])):
            allrows.append((tab["chr"][i], tab["start"][i], tab["end"][i], tab["states"][i]))
    allrows.sort(key=lambda r: (float(r[0]), float(r[1])))     # order(): stable
    seen, uniq = set(), []
    for r in allrows:
        key = f"{r[0]}-{r[1]}-{r[2]}"
        if key not in seen:
            seen.add(key)
            uniq.append(r)
    out = []
    for chrom in range(1, 23):
        sub = [r for r in uniq if float(r[0]) == chrom]        # seg_all$chr == chr compares as numbers/strings alike
        if not sub:
            continue
        bp = sorted({float(v) for r in sub for v in (r[1], r[2])})
        pieces = []
        for i in range(len(bp) - 1):
            qs, qe = bp[i], bp[i + 1] - 1
            st = {float(r[3]) for r in sub if float(r[1]) <= qe and float(r[2]) - 1 >= qs}
            pieces.append([sub[0][0], r_chr(bp[i]), r_chr(bp[i + 1]), " ".join(r_chr(v) for v in sorted(st))])
        if merge:                                              # :132-156
            merged, start_idx, end_idx = [], 0, 0
            npc = len(pieces)
            for i in range(npc):
                if i + 1 < npc:
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
    return out
