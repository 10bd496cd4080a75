"""cProfile of EstRegionCov through the host API on the bench's 20k-cell matrix (host-side costs)."""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse as sp
import clonalscope_b200 as cs
from clonalscope_b200 import synth
import bench
dev = torch.device("cuda", 0)
G, N = 30000, 20000
bins, genes = synth.geometry(G=G)
colptr, rowidx, x = bench.synth_counts_gpu(torch, dev, N, G)
mtx = sp.csc_matrix((x.cpu().numpy(), rowidx.cpu().numpy(), colptr.cpu().numpy().astype(np.int32)), shape=(G, N))
feats = [f"g{i}" for i in range(G)]; bars = [f"c{i}" for i in range(N)]
bed = np.empty((G, 4), dtype=object)
bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3] = [str(int(c)) for c in genes["chr"]], genes["start"], genes["end"], feats
rng = np.random.default_rng(77)
ct = np.empty((N, 2), dtype=object); ct[:, 0], ct[:, 1] = bars, rng.choice(["tumor", "normal"], size=N, p=[0.7, 0.3])
seg = dict(chr=[], start=[], end=[], chrr=[])
for c, size_c in enumerate(synth.CHR_SIZES):
    nseg = 14 if c < 14 else 13
    edges = np.linspace(0, size_c, nseg + 1)
    for lo_, hi_ in zip(edges[:-1], edges[1:]):
        seg["chr"].append(str(c + 1)), seg["start"].append(float(lo_)), seg["end"].append(float(hi_)); seg["chrr"].append(f"chr{c + 1}_{int(lo_)}_{int(hi_)}")
with contextlib.redirect_stdout(io.StringIO()):
    cs.EstRegionCov(mtx, bars, feats, bed, ct, seg_table_filtered=seg)
    pr = cProfile.Profile(); t = time.perf_counter(); pr.enable()
    cs.EstRegionCov(mtx, bars, feats, bed, ct, seg_table_filtered=seg)
    pr.disable(); dt = time.perf_counter() - t
print("total %.1f ms" % (dt * 1e3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
