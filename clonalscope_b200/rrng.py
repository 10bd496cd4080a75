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
