"""Per-sweep diagnostics of the Philox chain on the bench workload (not a benchmark)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clonalscope_b200 as cs
from clonalscope_b200 import _lib, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 300
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 40
L = cs.lib()
s = synth.sampler_input(N, R=R)
X = np.asfortranarray(s["X"]); U0 = np.asfortranarray(s["U0"]); sig0 = s["sigmas0"]
Z0 = (cs.loglik_matrix(X, s["U0"], sig0).argmax(axis=1) + 1).astype(np.int32)
dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))
ch = C.c_void_p()
assert L.cs_chain_create(N, R, 0, niter, _lib.RNG_PHILOX, C.byref(ch)) == 0
assert L.cs_chain_init(ch, dp(X, C.c_double), 2, dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                       C.c_double(2.0), C.c_double(2.0), C.c_uint32(200), 2) == 0
pe = ps = 0
for t in range(niter):
    assert L.cs_chain_run(ch, 1, int(t == niter - 1), 1) == 0, L.cs_last_error()
    ms = [C.c_double() for _ in range(5)]
    L.cs_chain_timing(ch, *[C.byref(m) for m in ms])
    sc, ev, rj = C.c_int64(), C.c_int64(), C.c_int64()
    L.cs_chain_stats(ch, C.byref(sc), C.byref(ev), C.byref(rj))
    print(f"sweep {t:3d}: run {ms[0].value:8.2f} ms  prepare {ms[1].value:6.2f}  index {ms[2].value:6.2f}  scan {ms[3].value:8.2f}  eos {ms[4].value:6.2f}"
          f" | events {ev.value - pe:6d}  scanned {sc.value - ps:8d}")
    pe, ps = ev.value, sc.value
kt = np.zeros(niter, dtype=np.int32)
Zall = np.zeros((niter, N), dtype=np.int32, order="F")
uoff = np.zeros(niter + 1, dtype=np.int64)
L.cs_chain_fetch(ch, dp(Zall, C.c_int), None, None, None, dp(kt, C.c_int), C.c_int64(0), None, None, None)
print("kt:", [int(k) for k in kt])
