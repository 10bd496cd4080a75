"""ctypes front-end of the CPU parity oracle (libcs_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
product package `clonalscope_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RNG_R = 0
RNG_PHILOX = 1
FLAG_SORT_ONCE = 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.cso_qnorm.restype = C.c_double
        _LIB.cso_qnorm.argtypes = [C.c_double]
        _LIB.cso_unif_rand.restype = C.c_double
        _LIB.cso_rnorm.restype = C.c_double
        _LIB.cso_rnorm.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _LIB.cso_total_loglik.restype = C.c_double
        _LIB.cso_bin_counts.restype = C.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("pos", C.c_int), ("ndraws", C.c_uint64)]


def mt_uniforms(seed: int, n: int) -> np.ndarray:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    return np.array([L.cso_unif_rand(C.byref(g)) for _ in range(n)])


def mt_rnorm(seed: int, means, sds) -> np.ndarray:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    return np.array([L.cso_rnorm(C.byref(g), float(m), float(s)) for m, s in zip(means, sds)])


def sample_int_prob(seed: int, prob, ndraw: int = 1) -> list:
    g = MT()
    L = lib()
    L.cso_set_seed(C.byref(g), C.c_uint32(seed))
    out = []
    n = len(prob)
    for _ in range(ndraw):
        p = np.array(prob, dtype=np.float64)
        perm = np.zeros(n, dtype=np.int32)
        out.append(L.cso_sample_int_prob(C.byref(g), n, _p(p, C.c_double), _p(perm, C.c_int)))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().cso_philox4x32(c, k, o)
    return [int(v) for v in o]


def qnorm(p):
    L = lib()
    return np.array([L.cso_qnorm(float(v)) for v in np.atleast_1d(p)])


def loglik_matrix(X, U, sigma):
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    N, R = X.shape
    K = U.shape[0]
    LL = np.empty((N, K), dtype=np.float64)
    lib().cso_loglik_matrix(N, R, K, _p(X, C.c_double), _p(U, C.c_double), _p(sigma, C.c_double),
                            _p(LL, C.c_double))
    return LL


def total_loglik(X, U, Z, sigma):
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    N, R = X.shape
    return lib().cso_total_loglik(N, R, _p(X, C.c_double), _p(U, C.c_double), U.shape[0],
                                  _p(Z, C.c_int), _p(sigma, C.c_double))


def gibbs(X, U0, Z0, sigmas0, alpha=2.0, beta=2.0, niter=200, seed=200, rng_mode=RNG_R, flags=0,
          ucap_rows=None):
    """Returns dict(Zall, sigma_all, LL, L0, kt, u_off, u_labels, u_rows, n_reject)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    U0 = np.ascontiguousarray(np.atleast_2d(U0), dtype=np.float64)
    Z0 = np.ascontiguousarray(Z0, dtype=np.int32)
    N, R = X.shape
    K0 = U0.shape[0]
    sig0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmas0, dtype=np.float64), (R,)))
    if ucap_rows is None:
        ucap_rows = niter * 256
    Zall = np.zeros((niter, N), dtype=np.int32)
    sigma_all = np.zeros((niter, R), dtype=np.float64)
    LL = np.zeros(niter, dtype=np.float64)
    L0 = np.zeros(1, dtype=np.float64)
    kt = np.zeros(niter, dtype=np.int32)
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap_rows, dtype=np.int32)
    u_rows = np.zeros((ucap_rows, R), dtype=np.float64)
    nrej = np.zeros(1, dtype=np.int64)
    rc = lib().cso_gibbs(
        N, R, _p(X, C.c_double), K0, _p(U0, C.c_double), _p(Z0, C.c_int), _p(sig0, C.c_double),
        C.c_double(alpha), C.c_double(beta), niter, C.c_uint32(int(seed)), rng_mode, flags,
        _p(Zall, C.c_int), _p(sigma_all, C.c_double), _p(LL, C.c_double), _p(L0, C.c_double),
        _p(kt, C.c_int), C.c_int64(ucap_rows), _p(u_off, C.c_int64), _p(u_labels, C.c_int),
        _p(u_rows, C.c_double), _p(nrej, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"cso_gibbs failed with code {rc}")
    n = int(u_off[-1])
    return dict(Zall=Zall, sigma_all=sigma_all, LL=LL, L0=float(L0[0]), kt=kt, u_off=u_off,
                u_labels=u_labels[:n], u_rows=u_rows[:n], n_reject=int(nrej[0]))


def expand_U(res, t):
    """Uall[[t+1]] as the reference stores it: kt x R with NaN rows for empty clusters."""
    a, b = int(res["u_off"][t]), int(res["u_off"][t + 1])
    R = res["u_rows"].shape[1]
    U = np.full((int(res["kt"][t]), R), np.nan)
    U[res["u_labels"][a:b] - 1] = res["u_rows"][a:b]
    return U


def genotype_grid(maxcp=6):
    G = maxcp * (maxcp + 3) // 2
    mu = np.zeros((G, 2))
    assert lib().cso_genotype_grid(maxcp, _p(mu, C.c_double)) == G
    return mu


def genotype_neighbor(cov, allele, maxcp=6):
    f = lib().cso_genotype_neighbor
    f.argtypes = [C.c_double, C.c_double, C.c_int]
    return f(float(cov), float(allele), maxcp)


def gibbs_allele(Xcov, Xall, U0, Z0, sig0_cov, sig0_all, alpha=1.0, beta=1.0, niter=200, seed=200, maxcp=6,
                 rng_mode=RNG_R, ucap_rows=None):
    """Returns dict(Zall, sig_cov_all, sig_allele_all, LL, L0, kt, u_off, u_labels, u_state, u_cov, u_allele)."""
    Xcov = np.ascontiguousarray(Xcov, dtype=np.float64)
    Xall = np.ascontiguousarray(Xall, dtype=np.float64)
    U0 = np.ascontiguousarray(np.atleast_2d(U0), dtype=np.int32)
    Z0 = np.ascontiguousarray(Z0, dtype=np.int32)
    N, R = Xcov.shape
    K0 = U0.shape[0]
    sc = np.ascontiguousarray(np.broadcast_to(np.asarray(sig0_cov, dtype=np.float64), (R,)))
    sa = np.ascontiguousarray(np.broadcast_to(np.asarray(sig0_all, dtype=np.float64), (R,)))
    if ucap_rows is None:
        ucap_rows = niter * 256
    Zall = np.zeros((niter, N), dtype=np.int32)
    sca, saa = np.zeros((niter, R)), np.zeros((niter, R))
    LL, L0 = np.zeros(niter), np.zeros(1)
    kt = np.zeros(niter, dtype=np.int32)
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap_rows, dtype=np.int32)
    u_state = np.zeros((ucap_rows, R), dtype=np.int32)
    u_cov, u_all = np.zeros((ucap_rows, R)), np.zeros((ucap_rows, R))
    rc = lib().cso_gibbs_allele(
        N, R, _p(Xcov, C.c_double), _p(Xall, C.c_double), K0, _p(U0, C.c_int), _p(Z0, C.c_int), _p(sc, C.c_double),
        _p(sa, C.c_double), C.c_double(alpha), C.c_double(beta), niter, C.c_uint32(int(seed)), maxcp, rng_mode,
        _p(Zall, C.c_int), _p(sca, C.c_double), _p(saa, C.c_double), _p(LL, C.c_double), _p(L0, C.c_double),
        _p(kt, C.c_int), C.c_int64(ucap_rows), _p(u_off, C.c_int64), _p(u_labels, C.c_int), _p(u_state, C.c_int),
        _p(u_cov, C.c_double), _p(u_all, C.c_double))
    if rc != 0:
        raise RuntimeError(f"cso_gibbs_allele failed with code {rc}")
    n = int(u_off[-1])
    return dict(Zall=Zall, sig_cov_all=sca, sig_allele_all=saa, LL=LL, L0=float(L0[0]), kt=kt, u_off=u_off,
                u_labels=u_labels[:n], u_state=u_state[:n], u_cov=u_cov[:n], u_allele=u_all[:n])


def majority_vote(Zall):
    Zall = np.ascontiguousarray(Zall, dtype=np.int32)
    T, N = Zall.shape
    out = np.zeros(N, dtype=np.int32)
    lib().cso_majority_vote(T, N, _p(Zall, C.c_int), _p(out, C.c_int))
    return out


def bin_counts(G, N, B, colptr, rowidx, x, map_ptr, map_bin):
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_bin = np.ascontiguousarray(map_bin, dtype=np.int32)
    ocp = np.zeros(N + 1, dtype=np.int64)
    L = lib()
    args = [G, N, B, _p(colptr, C.c_int64), _p(rowidx, C.c_int32), _p(x, C.c_double),
            _p(map_ptr, C.c_int32), _p(map_bin, C.c_int32), _p(ocp, C.c_int64)]
    nnz = L.cso_bin_counts(*args, None, None)
    ori = np.zeros(nnz, dtype=np.int32)
    ox = np.zeros(nnz, dtype=np.float64)
    L.cso_bin_counts(*args, _p(ori, C.c_int32), _p(ox, C.c_double))
    return ocp, ori, ox


def bin_counts_dense(G, N, B, colptr, rowidx, x, map_ptr, map_bin):
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_bin = np.ascontiguousarray(map_bin, dtype=np.int32)
    out = np.zeros((N, B), dtype=np.float64)  # column-major B x N == C-order N x B
    lib().cso_bin_counts_dense(G, N, B, _p(colptr, C.c_int64), _p(rowidx, C.c_int32),
                               _p(x, C.c_double), _p(map_ptr, C.c_int32), _p(map_bin, C.c_int32),
                               _p(out, C.c_double))
    return out.T  # B x N
