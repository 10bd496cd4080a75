"""R-compatible sampler on the reference's fixtures: label parity with the shipped chain and updates/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clonalscope_b200 as cs
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for name in ("P5931", "BC_ductal2", "P5847_allcells"):
    g = np.load(os.path.join("tests", "golden", f"sampler_{name}.npz"))
    kw = dict(cna_states_WGS=g["cna_states_WGS"], alpha=float(g["alpha"]), beta=float(g["beta"]), niter=niter,
              sigmas0=g["sigmas0"], U0=g["U0"], Z0=g["Z0"], seed=int(g["seed"]), rng="r")
    cs.BayesNonparCluster(g["X"], **{**kw, "niter": 2})
    t = time.perf_counter()
    out = cs.BayesNonparCluster(g["X"], **kw)
    dt = time.perf_counter() - t
    same = bool(np.array_equal(out["results"]["Zall"], g["Zall"][:niter]))
    print(f"{name}: {g['X'].shape} {niter} sweeps {dt*1e3:.0f} ms  {g['X'].shape[0]*niter/dt:.0f} updates/s  same_as_reference={same}", flush=True)
