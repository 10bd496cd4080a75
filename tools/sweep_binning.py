"""Sweep of the one-pass binning kernel's compile-time options (CS_BIN_OP_OPT bitmask, threads per CTA)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clonalscope_b200 as cs
from clonalscope_b200 import synth
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
dev = torch.device("cuda", 0)
L = cs.lib()
bins, genes = synth.geometry(G=30000)
mp, mb = synth.geometry_map(bins, genes)
B = len(bins["chr"])
colptr, rowidx, x = bench.synth_counts_gpu(torch, dev, N, 30000)
dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))
vp = lambda t: C.c_void_p(t.data_ptr())
plan = C.c_void_p()
assert L.cs_bin_plan_create(30000, B, dp(mp, C.c_int32), dp(mb, C.c_int32), C.byref(plan)) == 0
ocp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
nnz = C.c_int64()
assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), None, None, 0, C.byref(nnz), None) == 0
ori = torch.empty(nnz.value, dtype=torch.int32, device=dev)
ox = torch.empty(nnz.value, dtype=torch.float64, device=dev)
ref = None
nin = int(colptr[-1])
for threads in (512,):
    for opt in range(32):
        os.environ["CS_BIN_OP_SMAP_THREADS"] = str(threads)
        os.environ["CS_BIN_OP_OPT"] = str(opt)
        best = 1e9
        for _ in range(3):
            assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), vp(ori), vp(ox), nnz.value, C.byref(nnz), None) == 0
            t = [C.c_double() for _ in range(3)]
            L.cs_bin_plan_timing(plan, *[C.byref(v) for v in t])
            best = min(best, t[2].value)
        sig = (int(ocp.sum().item()), int(ori.sum().item()), float(ox.sum().item()))
        if ref is None:
            ref = sig
        print("threads %d opt %2d  %.3f ms  %.0f GB/s  %s" % (threads, opt, best, 12 * (nin + nnz.value) / best / 1e6,
                                                              "ok" if sig == ref else "MISMATCH"), flush=True)
