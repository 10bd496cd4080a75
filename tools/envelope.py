#!/usr/bin/env python
"""Distributional envelope of the Philox stream on the reference's fixtures:
for seeds 200..200+n-1, 200 sweeps, burn-in 100, majority vote -> ARI against the reference's own
posterior clustering (the shipped Zall) and the number of clusters above mincell; plus the chain's
throughput in cell-updates/s.  Writes gpurun_out/envelope.json.  Run on the GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clonalscope_b200 as cs  # noqa: E402
from sklearn.metrics import adjusted_rand_score  # noqa: E402


def main():
    nseed = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    out = {}
    for name, mincell in (("P5931", 13), ("BC_ductal2", 40), ("P5847_allcells", 50)):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"sampler_{name}.npz"))
        zr = cs.majority_vote(g["Zall"][100:].astype(np.int32))
        rows = []
        for seed in range(200, 200 + nseed):
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                o = cs.BayesNonparCluster(g["X"], cna_states_WGS=g["cna_states_WGS"], alpha=float(g["alpha"]),
                                          beta=float(g["beta"]), niter=200, sigmas0=g["sigmas0"], U0=g["U0"],
                                          Z0=g["Z0"], seed=seed, rng="philox")
                dt = time.perf_counter() - t0
                zm = cs.majority_vote(cs.MCMCtrim(o)["results"]["Zall"])
            lab, cnt = np.unique(zm, return_counts=True)
            rows.append(dict(seed=seed, ari=float(adjusted_rand_score(zm, zr)), nclusters=int((cnt > mincell).sum()),
                             kt_final=int(o["results"]["Uall"][-1].shape[0]), seconds=dt,
                             updates_per_s=g["X"].shape[0] * 200 / dt))
            print(name, rows[-1], flush=True)
        out[name] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "envelope.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
