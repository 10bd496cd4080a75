"""Small driver for ncu captures of the binning kernels (not a benchmark)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import clonalscope_b200 as cs
from clonalscope_b200 import synth
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
if os.environ.get("CS_BIN_COMPARE"):
    # the one-pass kernel against the count / scan / fill kernels on the same buffers
    os.environ["CS_BIN_ONEPASS"] = "0"
    ocp0, ori0, ox0 = torch.zeros_like(ocp), torch.empty_like(ori), torch.empty_like(ox)
    assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp0), vp(ori0), vp(ox0), nnz.value, C.byref(nnz), None) == 0
    os.environ["CS_BIN_ONEPASS"] = "1"
    assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), vp(ori), vp(ox), nnz.value, C.byref(nnz), None) == 0
    print("same colptr", torch.equal(ocp, ocp0), "same bins", torch.equal(ori, ori0), "same sums", torch.equal(ox, ox0), flush=True)
for _ in range(reps):
    assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), vp(ori), vp(ox), nnz.value, C.byref(nnz), None) == 0
    t = [C.c_double() for _ in range(3)]
    L.cs_bin_plan_timing(plan, *[C.byref(v) for v in t])
    nin = int(colptr[-1])
    print("count %.3f ms  scan %.3f ms  fill %.3f ms | fill %.0f GB/s  total %.0f GB/s" % (
        t[0].value, t[1].value, t[2].value, 12 * (nin + nnz.value) / t[2].value / 1e6,
        (12 * (nin + nnz.value)) / (t[0].value + t[1].value + t[2].value) / 1e6))
