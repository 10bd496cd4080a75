import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import clonalscope_b200 as cs
from clonalscope_b200 import synth
from clonalscope_b200.api import bin_counts_csc
import bench
dev = torch.device("cuda", 0)
bins, genes = synth.geometry(G=30000)
mp, mb = synth.geometry_map(bins, genes)
B = len(bins["chr"])
colptr, rowidx, x = bench.synth_counts_gpu(torch, dev, 20000, 30000)
cp_h = colptr.cpu().numpy().astype(np.int32); ri_h = rowidx.cpu().numpy(); x_h = x.cpu().numpy()
print("cores", os.cpu_count())
for thr in (4, 8, 12, 16):
    os.environ["CS_STAGE_THREADS"] = str(thr)
    bin_counts_csc(30000, 20000, B, cp_h, ri_h, x_h, mp, mb)
    ts = []
    for _ in range(4):
        t = time.perf_counter(); o = bin_counts_csc(30000, 20000, B, cp_h, ri_h, x_h, mp, mb); ts.append(time.perf_counter() - t)
    by = 12 * int(cp_h[-1]) + 12 * len(o[1])
    print(thr, "threads:", [round(v * 1e3, 1) for v in ts], "ms  median %.1f GB/s" % (by / np.median(ts) / 1e9), flush=True)
