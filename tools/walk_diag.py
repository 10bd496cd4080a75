#!/usr/bin/env python
"""Per-sweep diagnostics of the Philox sampler's walk on the bench workload: time of the sweep
(CUDA events on the chain's stream), events (cells that move), steps of the fixed-point
iteration, clusters opened, cells evaluated.
python tools/walk_diag.py [nsweeps] [seed] [N] [R] [flags]   (flags 6 = the round-1 ring walk)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clonalscope_b200 as cs  # noqa: E402
from clonalscope_b200 import _lib, synth  # noqa: E402

nsw = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 200
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
R = int(sys.argv[4]) if len(sys.argv) > 4 else 300
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 2
L = cs.lib()
s = synth.sampler_input(N, R=R)
X = np.asfortranarray(s["X"])
U0 = np.asfortranarray(s["U0"])
sig0 = np.ascontiguousarray(s["sigmas0"])
Z0 = (cs.loglik_matrix(X, s["U0"], sig0).argmax(axis=1) + 1).astype(np.int32)
dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
chain = C.c_void_p()
assert L.cs_chain_create(N, R, 0, nsw, _lib.RNG_PHILOX, C.byref(chain)) == 0
assert L.cs_chain_init(chain, dp(X, C.c_double), 2, dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                       C.c_double(2.0), C.c_double(2.0), C.c_uint32(seed), flags) == 0
prev = np.zeros(6, dtype=np.int64)
tot = 0.0
for t in range(nsw):
    assert L.cs_chain_run(chain, 1, int(t == nsw - 1), 1) == 0, L.cs_last_error()
    ms = [C.c_double() for _ in range(5)]
    L.cs_chain_timing(chain, *[C.byref(m) for m in ms])
    v = [C.c_int64() for _ in range(7)]
    L.cs_chain_stats(chain, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]))
    L.cs_chain_walk_stats(chain, C.byref(v[3]), C.byref(v[4]), C.byref(v[5]), C.byref(v[6]))
    cur = np.array([v[1].value, v[3].value, v[4].value, v[5].value, v[6].value, v[0].value])
    d = cur - prev
    prev = cur
    tot += ms[0].value
    print("sweep %3d: run %7.2f ms  prepare %5.2f index %5.2f walk %7.3f eos %5.2f | events %6d steps %5d opened %3d evaluated %8d  us/step %6.2f"
          % (t, ms[0].value, ms[1].value, ms[2].value, ms[3].value, ms[4].value, d[0], d[1], d[3], d[5],
             1e3 * ms[3].value / max(d[1], 1)), flush=True)
    if os.environ.get("CS_FIX_PROF"):
        pr = np.zeros(1024, dtype=np.uint64)
        L.cs_chain_fix_prof(chain, dp(pr, C.c_uint64))
        pr = pr.reshape(128, 8).astype(np.int64)
        for k in range(min(int(d[1]), 64)):
            n = max(int(pr[k, 6]), 1)
            print("      step %2d: lo %6d  eval max %7.1f us  mean(busy CTAs) %7.1f us (%3d CTAs)  barrier1 %6.1f  scan %6.1f  barrier2 %6.1f | per tile: setup %5.1f pull %5.1f draw %5.1f"
                  % (k, pr[k, 5], pr[k, 0] / 1e3, pr[k, 1] / 1e3 / n, pr[k, 6], pr[k, 2] / 1e3, pr[k, 3] / 1e3, pr[k, 4] / 1e3,
                     pr[64 + k, 0] / 1e3 / n, pr[64 + k, 1] / 1e3 / n, pr[64 + k, 2] / 1e3 / n))
print("total %.1f ms" % tot)
