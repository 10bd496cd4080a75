import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import clonalscope_b200 as cs
from clonalscope_b200 import _lib, synth
N,R,niter=100000,300,200
L=cs.lib()
s=synth.sampler_input(N,R=R)
X=np.asfortranarray(s["X"]); U0=np.asfortranarray(s["U0"]); sig0=np.ascontiguousarray(s["sigmas0"])
Z0=(cs.loglik_matrix(X,s["U0"],sig0).argmax(axis=1)+1).astype(np.int32)
for k in range(3):
    t=time.perf_counter()
    g=cs.BayesNonparCluster(s["X"],U0=s["U0"],sigmas0=s["sigmas0"],niter=niter,seed=200,rng="philox",Z0=Z0)
    print("api call %.1f ms"%((time.perf_counter()-t)*1e3))
