"""Base R's default random stream on the host, for the few sequential draws the reference makes
outside its samplers (`set.seed(100)` + one `sample.int(n, 1, prob=)` per orphan cell in
R/AssignCluster.R:36,63): Mersenne-Twister with R's seeding (RNG.c `Randomize` / `MT_genrand` /
`fixup`) and `sample.int`'s unequal-probability walk over the probabilities sorted by R's
`revsort` (random.c `FixupProb` / `ProbSampleNoReplace`, sort.c `revsort`).  Integer and
comparison logic only; the device samplers have their own copy of the stream (csrc/cs_rstream.cuh).
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


class RStream:
    def __init__(self, seed: int):
        s = int(seed) & _M32
        for _ in range(50):                       # initial scrambling
            s = (69069 * s + 1) & _M32
        s = (69069 * s + 1) & _M32                # the slot that holds the position, overwritten
        mt = []
        for _ in range(624):
            s = (69069 * s + 1) & _M32
            mt.append(s)
        self.mt = mt
        self.pos = 624                            # refill on the first draw

    def _refill(self):
        mt = self.mt
        for kk in range(624):
            y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % 624] & 0x7FFFFFFF)
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def unif_rand(self) -> float:
        if self.pos >= 624:
            self._refill()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        v = y * 2.3283064365386963e-10
        if v <= 0.0:                              # fixup(): strictly inside (0, 1)
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v


def revsort(a, ib):
    """Sort a[] into descending order by heapsort, carrying ib[] along (R's revsort; not stable)."""
    n = len(a)
    if n <= 1:
        return
    l, ir = (n >> 1) + 1, n
    while True:
        if l > 1:
            l -= 1
            ra, ii = a[l - 1], ib[l - 1]
        else:
            ra, ii = a[ir - 1], ib[ir - 1]
            a[ir - 1], ib[ir - 1] = a[0], ib[0]
            ir -= 1
            if ir == 1:
                a[0], ib[0] = ra, ii
                return
        i, j = l, l << 1
        while j <= ir:
            if j < ir and a[j - 1] > a[j]:
                j += 1
            if ra > a[j - 1]:
                a[i - 1], ib[i - 1] = a[j - 1], ib[j - 1]
                i = j
                j += j
            else:
                j = ir + 1
        a[i - 1], ib[i - 1] = ra, ii


def sample_int_prob_one(stream: RStream, prob) -> int:
    """sample.int(length(prob), 1, replace = FALSE, prob = prob): 1-based index."""
    p = [float(v) for v in prob]
    total = 0.0
    for v in p:
        if v > 0:
            total += v
    p = [v / total for v in p]
    perm = list(range(1, len(p) + 1))
    revsort(p, perm)
    rT = stream.unif_rand()
    mass = 0.0
    j = 0
    for j in range(len(p) - 1):
        mass += p[j]
        if rT <= mass:
            break
    else:
        j = len(p) - 1
    return perm[j]


def unif_index(stream: RStream, dn: int) -> int:
    """R_unif_index(dn) with sample.kind = "Rejection" (R >= 3.6, RNG.c): whole 16-bit chunks of
    unif_rand() assembled into ceil(log2(dn)) bits, rejected while >= dn."""
    if dn <= 0:
        return 0
    bits = int(np.ceil(np.log2(dn)))
    while True:
        v = 0
        n = 0
        while n <= bits:
            v = 65536 * v + int(np.floor(stream.unif_rand() * 65536))
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        if v < dn:
            return v


def sample_int_noreplace(stream: RStream, n: int, k: int):
    """sample.int(n, k) (no prob, no replacement; random.c do_sample): 1-based indices."""
    x = list(range(n))
    out = []
    for _ in range(k):
        j = unif_index(stream, n)
        out.append(x[j] + 1)
        n -= 1
        x[j] = x[n]
    return out


def kmeans_hartigan_wong(stream: RStream, x, k=2, iter_max=10):
    """stats::kmeans(x, k) for a numeric vector with R's defaults (one random start drawn with
    sample.int on the given stream, Hartigan & Wong's AS 136 as in R's kmns.c).  Returns the k centres
    in cluster order.  Only the few per-cluster malignancy indices of AssignCluster go through this
    (R/AssignCluster.R:191-194), so it stays on the host like the rest of R's stream."""
    a = [float(v) for v in x]
    m = len(a)
    if k <= 1 or k >= m:
        raise ValueError("number of cluster centres must lie between 1 and nrow(x)")
    cen = [a[i - 1] for i in sample_int_noreplace(stream, m, k)]
    if len(set(cen)) < k:                                       # duplicated centres: draw among the distinct values
        cn = list(dict.fromkeys(a))
        if len(cn) < k:
            raise ValueError("more cluster centers than distinct data points.")
        cen = [cn[i - 1] for i in sample_int_noreplace(stream, len(cn), k)]
    big = 1e30
    ic1, ic2 = [0] * m, [0] * m
    for i in range(m):                                          # the two closest centres of every point
        d = [(a[i] - c) ** 2 for c in cen]
        i1, i2 = (0, 1) if d[0] <= d[1] else (1, 0)
        for l in range(2, k):
            if d[l] < d[i2]:
                if d[l] < d[i1]:
                    i1, i2 = l, i1
                else:
                    i2 = l
        ic1[i], ic2[i] = i1, i2
    nc = [0] * k
    cen = [0.0] * k
    for i in range(m):
        cen[ic1[i]] += a[i]
        nc[ic1[i]] += 1
    if min(nc) == 0:
        raise ValueError("empty cluster: try a better set of initial centers")
    cen = [c / n for c, n in zip(cen, nc)]
    an2 = [n / (n + 1.0) for n in nc]
    an1 = [big if n <= 1 else n / (n - 1.0) for n in nc]
    itran, ncp, live = [1] * k, [-1] * k, [0] * k
    d = [0.0] * m
    indx = 0

    def move(i, l1, l2):
        al1, al2 = float(nc[l1]), float(nc[l2])
        alw, alt = al1 - 1.0, al2 + 1.0
        cen[l1] = (cen[l1] * al1 - a[i]) / alw
        cen[l2] = (cen[l2] * al2 + a[i]) / alt
        nc[l1] -= 1
        nc[l2] += 1
        an2[l1] = alw / al1
        an1[l1] = big if alw <= 1.0 else alw / (alw - 1.0)
        an1[l2] = alt / al2
        an2[l2] = alt / (alt + 1.0)
        ic1[i], ic2[i] = l2, l1

    for _ in range(iter_max):
        # ---- optimal-transfer stage
        for l in range(k):
            if itran[l] == 1:
                live[l] = m
        done = False
        for i in range(m):
            indx += 1
            l1 = ic1[i]
            if nc[l1] != 1:
                if ncp[l1] != 0:
                    d[i] = (a[i] - cen[l1]) ** 2 * an1[l1]
                l2 = ll = ic2[i]
                r2 = (a[i] - cen[l2]) ** 2 * an2[l2]
                for l in range(k):
                    if ((i + 1 >= live[l1] and i + 1 >= live[l]) or l == l1 or l == ll):
                        continue
                    dc = (a[i] - cen[l]) ** 2
                    if dc >= r2 / an2[l]:
                        continue
                    r2 = dc * an2[l]
                    l2 = l
                if r2 >= d[i]:
                    ic2[i] = l2
                else:
                    indx = 0
                    live[l1] = live[l2] = m + i + 1
                    ncp[l1] = ncp[l2] = i + 1
                    move(i, l1, l2)
            if indx == m:
                done = True
                break
        if done:
            break
        for l in range(k):
            itran[l] = 0
            live[l] -= m
        # ---- quick-transfer stage
        icoun = istep = 0
        limit = 50 * m
        stop = False
        while not stop:
            for i in range(m):
                icoun += 1
                istep += 1
                if istep >= limit:
                    stop = True
                    break
                l1, l2 = ic1[i], ic2[i]
                if nc[l1] != 1:
                    if istep <= ncp[l1]:
                        d[i] = (a[i] - cen[l1]) ** 2 * an1[l1]
                    if istep < ncp[l1] or istep < ncp[l2]:
                        if (a[i] - cen[l2]) ** 2 < d[i] / an2[l2]:
                            icoun = indx = 0
                            itran[l1] = itran[l2] = 1
                            ncp[l1] = ncp[l2] = istep + m
                            move(i, l1, l2)
                if icoun == m:
                    stop = True
                    break
        if k == 2:
            break
        ncp = [0] * k
    cen = [0.0] * k
    for i in range(m):
        cen[ic1[i]] += a[i]
    return [c / n for c, n in zip(cen, nc)]
