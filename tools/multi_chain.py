"""Independent chains run concurrently on ONE GPU (a chain's sequential walk occupies one cluster of
SMs; other chains' all-SM kernels fill the rest).  Not a benchmark: bench.py reports it.

  python tools/multi_chain.py [N] [R] [niter] [chains...]
"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clonalscope_b200 as cs
from clonalscope_b200 import _lib, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 300
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 200
counts = [int(a) for a in sys.argv[4:]] or [1, 2, 4]
L = cs.lib()
s = synth.sampler_input(N, R=R)
X = np.asfortranarray(s["X"]); U0 = np.asfortranarray(s["U0"]); sig0 = s["sigmas0"]
Z0 = (cs.loglik_matrix(X, s["U0"], sig0).argmax(axis=1) + 1).astype(np.int32)
dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))
import torch
for nch in counts:
    chains = []
    for k in range(nch):
        ch = C.c_void_p()
        assert L.cs_chain_create(N, R, 0, niter, _lib.RNG_PHILOX, C.byref(ch)) == 0, L.cs_last_error()
        assert L.cs_chain_init(ch, dp(X, C.c_double), U0.shape[0], dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                               C.c_double(2.0), C.c_double(2.0), C.c_uint32(200 + k), 0) == 0, L.cs_last_error()
        chains.append(ch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(niter):  # sweeps are enqueued round-robin so that no chain's launches wait behind another's
        for ch in chains:
            assert L.cs_chain_run(ch, 1, int(t == niter - 1), 0) == 0, L.cs_last_error()
    for ch in chains:
        assert L.cs_chain_sync(ch) == 0, L.cs_last_error()
    dt = time.perf_counter() - t0
    print(f"chains {nch}: {dt * 1e3:8.1f} ms wall  {nch * N * niter / dt / 1e6:8.2f} M cell-updates/s")
    for ch in chains:
        L.cs_chain_destroy(ch)
