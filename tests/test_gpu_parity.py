"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the reference's golden fixtures, and through size-independent
properties.  Integer / label work must match bit for bit; floating point to the stated
tolerance (north star: log-likelihood matrix 1e-6 relative — we assert 1e-11)."""
import os

import numpy as np
import pytest

from conftest import SAMPLER_FIXTURES, load_golden

pytestmark = pytest.mark.gpu
FULL = os.environ.get("CS_FULL_FIXTURES", "0") == "1"


# ------------------------------------------------------------------ log-likelihood matrix
def test_loglik_matrix_vs_oracle(cs, oracle):
    rng = np.random.default_rng(11)
    for (N, R, K) in [(1, 1, 1), (37, 5, 3), (1000, 38, 47), (513, 300, 20), (200, 333, 2)]:
        X = rng.uniform(0.05, 3.0, size=(N, R))
        X[rng.random((N, R)) < 0.05] = np.nan
        if N > 3:
            X[2, :] = np.nan  # an all-NA cell scores 0 everywhere
        U = rng.uniform(0.3, 2.5, size=(K, R))
        sig = rng.uniform(0.05, 0.6, size=R)
        got = cs.loglik_matrix(X, U, sig)
        want = oracle.loglik_matrix(X, U, sig)
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("name", SAMPLER_FIXTURES)
def test_loglik_known_answers_from_fixtures(cs, name):
    """Likelihood[t] = sum_i LL[i, Zall[t,i]] recomputed on the device from the reference's own
    (data, Uall[[t]], sigma_all[[t]], Zall[t,]) — the 1e-6 criterion, asserted at 1e-10."""
    g = load_golden(f"sampler_{name}.npz")
    X = g["X"]
    for j, t in enumerate(g["usel"]):
        a, b = int(g["usel_off"][j]), int(g["usel_off"][j + 1])
        labels = g["u_labels"][int(g["u_off"][t]):int(g["u_off"][t + 1])]
        LLm = cs.loglik_matrix(X, g["usel_rows"][a:b], g["sigma_all"][t])
        col = {int(l): i for i, l in enumerate(labels)}
        Z = g["Zall"][t]
        tot = LLm[np.arange(X.shape[0]), [col[int(z)] for z in Z]].sum()
        assert abs(tot / g["Likelihood"][t] - 1) < 1e-10, (name, t)
    L0m = cs.loglik_matrix(X, g["U0"], g["sigmas0"])
    assert abs(L0m[np.arange(X.shape[0]), g["Z0"] - 1].sum() / float(g["L0"]) - 1) < 1e-10


# ------------------------------------------------------------------ gene -> bin aggregation
def _random_csc(rng, G, N, density, vmax=9):
    cols = []
    colptr = [0]
    for c in range(N):
        m = rng.binomial(G, density)
        cols.append(np.sort(rng.choice(G, size=m, replace=False)).astype(np.int32))
        colptr.append(colptr[-1] + m)
    rowidx = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    x = rng.integers(1, vmax + 1, size=len(rowidx)).astype(np.float64)
    return np.array(colptr, dtype=np.int32), rowidx.astype(np.int32), x


def _check_bins(cs, oracle, G, N, B, colptr, rowidx, x, map_ptr, map_bin, exact=True):
    from clonalscope_b200.api import bin_counts_csc
    ocp, ori, ox = bin_counts_csc(G, N, B, colptr, rowidx, x, map_ptr, map_bin)
    wcp, wri, wx = oracle.bin_counts(G, N, B, colptr, rowidx, x, map_ptr, map_bin)
    assert np.array_equal(ocp, wcp)
    assert np.array_equal(ori, wri)
    if exact:
        assert np.array_equal(ox, wx)
    else:
        np.testing.assert_allclose(ox, wx, rtol=1e-13)
    return ocp, ori, ox


def test_binning_synthetic_geometry_bit_exact(cs, oracle):
    from clonalscope_b200 import synth
    bins, genes = synth.geometry(G=3000, seed=7)
    map_ptr, map_bin = synth.geometry_map(bins, genes)
    B = len(bins["chr"])
    colptr, rowidx, x = synth.counts_csc(400, G=3000, seed=8, mean_genes=300)
    ocp, ori, ox = _check_bins(cs, oracle, 3000, 400, B, colptr, rowidx, x, map_ptr, map_bin)
    # the reference's own dense algorithm (densify, loop over bins) agrees as well
    dense = oracle.bin_counts_dense(3000, 400, B, colptr, rowidx, x, map_ptr, map_bin)
    import scipy.sparse as sp
    got = sp.csc_matrix((ox, ori, ocp), shape=(B, 400)).toarray()
    assert np.array_equal(got, dense)
    # property: total per cell = sum_g x[g,c] * #bins(g)   (many-to-many)
    nh = np.diff(map_ptr).astype(np.float64)
    for c in (0, 17, 399):
        assert got[:, c].sum() == (x[colptr[c]:colptr[c + 1]] * nh[rowidx[colptr[c]:colptr[c + 1]]]).sum()


@pytest.mark.parametrize("env", [{"CS_BIN_OP_SMAP": "0"}, {"CS_BIN_OP_SMAP": "0", "CS_BIN_OP_CTAS": "2"},
                                 {"CS_BIN_ONEPASS": "0"}, {"CS_BIN_OP_OPT": "0"}, {"CS_BIN_OP_OPT": "31"}])
def test_binning_alternative_kernels_bit_exact(cs, oracle, monkeypatch, env):
    """The kernels behind the default one-pass path — the map read through L1 instead of shared memory
    (what a gene table too large for shared memory takes), the count / scan / fill kernels, and the
    one-pass kernel with none / all of its optional stages — give the same bytes."""
    from clonalscope_b200 import synth
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    bins, genes = synth.geometry(G=3000, seed=7)
    map_ptr, map_bin = synth.geometry_map(bins, genes)
    colptr, rowidx, x = synth.counts_csc(700, G=3000, seed=9, mean_genes=400)
    _check_bins(cs, oracle, 3000, 700, len(bins["chr"]), colptr, rowidx, x, map_ptr, map_bin)


def test_binning_real_geometry_api(cs, oracle):
    """Gen_bin_cell_rna_filtered end to end on the reference's real tiles and gene table."""
    import scipy.sparse as sp
    from clonalscope_b200.api import gene_bin_map
    g = load_golden("geometry_hg38.npz")
    bed = np.stack([g["bin_chr"], g["bin_start"].astype(object), g["bin_end"].astype(object)], axis=1)
    gtf = dict(V1=g["gtf_V1"], V3=g["gtf_V3"], V4=g["gtf_V4"], V5=g["gtf_V5"], V9=g["gtf_V9"])
    ids = [s.split("gene_id ")[1].rstrip(";") for s in g["gtf_V9"]]
    genes = np.array(ids, dtype=object).reshape(-1, 1)
    G, N = len(ids), 150
    rng = np.random.default_rng(3)
    colptr, rowidx, x = _random_csc(rng, G, N, 0.05)
    m = sp.csc_matrix((x, rowidx, colptr), shape=(G, N))
    barcodes = np.array([f"cell{i}-1" for i in range(N)], dtype=object).reshape(-1, 1)
    out = cs.Gen_bin_cell_rna_filtered(bin_bed=bed, barcodes=barcodes, gene_matrix=m, genes=genes, gtf=gtf)
    assert out.shape == (15641, N) and out.colnames[3] == "cell3-1" and out.rownames[0] == "chr1-0-200000"
    rn, map_ptr, map_bin = gene_bin_map(bed, genes, gtf)
    wcp, wri, wx = oracle.bin_counts(G, N, 15641, colptr, rowidx, x, map_ptr, map_bin)
    v = out.values
    assert np.array_equal(v.indptr, wcp) and np.array_equal(v.indices, wri) and np.array_equal(v.data, wx)


def test_binning_edge_cases(cs, oracle):
    rng = np.random.default_rng(5)
    G, B = 50, 40
    # general (non-contiguous, unsorted-origin) map exercises the escape path
    lists = [np.sort(rng.choice(B, size=rng.integers(0, 4), replace=False)) for _ in range(G)]
    lists[3] = np.array([0, 5, 39])
    map_ptr = np.zeros(G + 1, dtype=np.int32)
    map_ptr[1:] = np.cumsum([len(l) for l in lists])
    map_bin = np.concatenate(lists).astype(np.int32)
    # ragged columns: empty, single entry, every alignment of the 128-bit body, full
    cols = [np.zeros(0, np.int32), np.array([3], np.int32)]
    for m in range(2, 14):
        cols.append(np.sort(rng.choice(G, size=m, replace=False)).astype(np.int32))
    cols.append(np.arange(G, dtype=np.int32))
    cols.append(np.zeros(0, np.int32))
    colptr = np.zeros(len(cols) + 1, dtype=np.int32)
    colptr[1:] = np.cumsum([len(c) for c in cols])
    rowidx = np.concatenate(cols)
    x = rng.integers(1, 100, size=len(rowidx)).astype(np.float64)
    N = len(cols)
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x, map_ptr, map_bin)
    # explicit zeros and cancelling negatives: touched bins with zero sum are dropped (shrink path)
    x2 = x.copy()
    x2[colptr[5]:colptr[6]] = 0.0
    x2[colptr[8]] = 0.0
    ocp, _, _ = _check_bins(cs, oracle, G, N, B, colptr, rowidx, x2, map_ptr, map_bin)
    assert ocp[6] == ocp[5]
    # counts beyond int32 and non-integer values go through the FP64 accumulators
    x3 = x.copy()
    x3[0] = 3.0e9
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x3, map_ptr, map_bin)
    x4 = x + 0.25
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x4, map_ptr, map_bin, exact=False)
    # no cells at all
    from clonalscope_b200.api import bin_counts_csc
    ocp, ori, ox = bin_counts_csc(G, 0, B, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), map_ptr, map_bin)
    assert list(ocp) == [0] and len(ori) == 0
    with pytest.raises(cs.ClonalscopeError):
        bin_counts_csc(G, 1, B, np.array([0, 1], np.int32), np.array([G], np.int32), np.ones(1), map_ptr, map_bin)


@pytest.mark.parametrize("widths,step", [([1, 1, 1, 2, 3], 2), ([1, 2, 5, 20, 32], 6), ([30, 31, 32], 3), ([1, 1, 2, 32], 10)])
def test_binning_onepass_overlapping_and_wide_genes(cs, oracle, widths, step):
    """The one-pass kernel's own limits: genes of up to 32 bins whose intervals overlap heavily (most
    hits fall on bins an earlier entry opened: the red.shared path, entries reaching over several old
    bins), tiles that open more ranks than a warp's staging buffer holds (the kernel gives up on the
    column and the plan falls back to the generic kernels), next to ordinary narrow genes."""
    rng = np.random.default_rng(31 + step)
    G = 5000
    nh = rng.choice(widths, size=G)
    b0 = np.cumsum(rng.integers(0, step + 1, size=G))
    B = int(b0[-1] + 40)
    map_ptr = np.zeros(G + 1, np.int32)
    map_ptr[1:] = np.cumsum(nh)
    map_bin = np.concatenate([np.arange(f, f + n) for f, n in zip(b0, nh)]).astype(np.int32)
    cols = [np.arange(G, dtype=np.int32), np.zeros(0, np.int32)]
    for m in (1, 7, 8, 9, 255, 256, 257, 1000, 2500, 4000, 4999):
        cols.append(np.sort(rng.choice(G, size=m, replace=False)).astype(np.int32))
    cols += [np.arange(G, dtype=np.int32)] * 3 + [np.arange(0, G, 2, dtype=np.int32)] * 40      # enough columns for every warp
    colptr = np.zeros(len(cols) + 1, np.int32)
    colptr[1:] = np.cumsum([len(c) for c in cols])
    rowidx = np.concatenate(cols)
    x = rng.integers(1, 9, size=len(rowidx)).astype(np.float64)
    _check_bins(cs, oracle, G, len(cols), B, colptr, rowidx, x, map_ptr, map_bin)
    x[rng.random(len(x)) < 0.3] = 0.0                                    # explicit zeros: sums that vanish are dropped
    _check_bins(cs, oracle, G, len(cols), B, colptr, rowidx, x, map_ptr, map_bin)


def test_binning_streaming_path_edge_cases(cs, oracle):
    """Regular map (contiguous bins, non-decreasing first bin in gene order) -> the warp-per-column
    streaming kernels: ragged / empty / very sparse columns whose batches span more bins than a
    warp's ring (forces batch splitting), long genes, zero sums, non-integer and huge values."""
    rng = np.random.default_rng(21)
    G, B = 6000, 20000
    b0 = np.sort(rng.integers(0, B - 400, size=G))
    nh = rng.choice([0, 1, 1, 1, 1, 2, 3, 12, 300], size=G, p=[.05, .3, .2, .2, .1, .08, .05, .015, .005])
    map_ptr = np.zeros(G + 1, np.int32)
    map_ptr[1:] = np.cumsum(nh)
    map_bin = np.concatenate([np.arange(f, f + n) for f, n in zip(b0, nh)]).astype(np.int32)
    cols = [np.zeros(0, np.int32), np.array([G - 1], np.int32), np.arange(G, dtype=np.int32)]
    for m in (1, 2, 3, 5, 31, 32, 33, 64, 65, 100, 1000, 3000):
        cols.append(np.sort(rng.choice(G, size=m, replace=False)).astype(np.int32))
    cols += [np.zeros(0, np.int32)] * 3
    colptr = np.zeros(len(cols) + 1, np.int32)
    colptr[1:] = np.cumsum([len(c) for c in cols])
    rowidx = np.concatenate(cols)
    x = rng.integers(1, 50, size=len(rowidx)).astype(np.float64)
    N = len(cols)
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x, map_ptr, map_bin)
    x2 = x.copy()
    x2[colptr[6]:colptr[7]] = 0.0          # a whole column of explicit zeros
    x2[colptr[12] + 3] = 0.0
    k = colptr[13]
    x2[k], x2[k + 1] = 7.0, -7.0           # may or may not share a bin; the oracle decides
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x2, map_ptr, map_bin)
    x3 = x.copy()
    x3[5] = 2.5e9
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x3, map_ptr, map_bin)
    _check_bins(cs, oracle, G, N, B, colptr, rowidx, x + 0.5, map_ptr, map_bin, exact=False)
    # row indices that do not ascend within a column (not a valid dgCMatrix, but the C ABI takes any
    # CSC): the streaming fill notices and the generic kernels take over
    r4, x4 = rowidx.copy(), x.copy()
    perm = rng.permutation(colptr[14] - colptr[13])
    r4[colptr[13]:colptr[14]] = r4[colptr[13]:colptr[14]][perm]
    x4[colptr[13]:colptr[14]] = x4[colptr[13]:colptr[14]][perm]
    _check_bins(cs, oracle, G, N, B, colptr, r4, x4, map_ptr, map_bin)


def test_binning_permutation_invariance(cs):
    """Relabelling the bins permutes the rows of the result and nothing else."""
    from clonalscope_b200.api import bin_counts_csc
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    G, N, B = 300, 64, 200
    nh = rng.integers(0, 3, size=G)
    first = rng.integers(0, B - 3, size=G)
    map_ptr = np.zeros(G + 1, np.int32)
    map_ptr[1:] = np.cumsum(nh)
    map_bin = np.concatenate([np.arange(f, f + n) for f, n in zip(first, nh)]).astype(np.int32)
    colptr, rowidx, x = _random_csc(rng, G, N, 0.2)
    a = sp.csc_matrix(bin_counts_csc(G, N, B, colptr, rowidx, x, map_ptr, map_bin)[::-1], shape=(B, N)).toarray()
    perm = rng.permutation(B)
    pm = np.concatenate([np.sort(perm[np.arange(f, f + n)]) for f, n in zip(first, nh)]).astype(np.int32)
    b = sp.csc_matrix(bin_counts_csc(G, N, B, colptr, rowidx, x, map_ptr, pm)[::-1], shape=(B, N)).toarray()
    assert np.array_equal(b[perm], a)


# ------------------------------------------------------------------ tally
def test_majority_vote_vs_oracle(cs, oracle):
    rng = np.random.default_rng(2)
    for (T, N, L) in [(1, 5, 3), (100, 2000, 4), (100, 3000, 60), (7, 1000, 2), (33, 257, 500)]:
        Z = rng.integers(1, L + 1, size=(T, N)).astype(np.int32)
        Z[:, : N // 3] = Z[0, : N // 3]  # converged cells take the fast path
        assert np.array_equal(cs.majority_vote(Z), oracle.majority_vote(Z))
    Z = np.array([[5, 2], [2, 5], [5, 2], [2, 5]], dtype=np.int32)  # exact ties -> smallest label
    assert list(cs.majority_vote(Z)) == [2, 2]


@pytest.mark.parametrize("name", ["P5931", "BC_ductal2", "P5847_allcells"])
def test_tally_on_fixture_chains(cs, oracle, name):
    g = load_golden(f"sampler_{name}.npz")
    obj = dict(results=dict(Zall=g["Zall"].astype(np.int32), Uall=[None] * 200, sigma_all=[None] * 200,
                            Likelihood=np.zeros(200)), priors={}, data=None)
    trimmed = cs.MCMCtrim(obj)  # burnin = 100
    zm = cs.majority_vote(trimmed["results"]["Zall"])
    assert np.array_equal(zm, oracle.majority_vote(g["Zall"][100:].astype(np.int32)))


# ------------------------------------------------------------------ sampler, R-compatible RNG
def _run_r(cs, g, niter):
    return cs.BayesNonparCluster(g["X"], cna_states_WGS=g["cna_states_WGS"], alpha=float(g["alpha"]),
                                 beta=float(g["beta"]), niter=niter, sigmas0=g["sigmas0"], U0=g["U0"],
                                 Z0=g["Z0"], seed=int(g["seed"]), rng="R")


def _ari(a, b):
    from sklearn.metrics import adjusted_rand_score
    return adjusted_rand_score(a, b)


@pytest.mark.parametrize("name,niter", [("P5931", 200), ("P5931_updated1", 200), ("BC_ductal2", 200),
                                        ("P5847_tumor", 200 if FULL else 25),
                                        ("P5847_allcells", 200 if FULL else 12)])
def test_sampler_r_mode_reproduces_reference_outputs(cs, name, niter):
    """With base R's RNG stream consumed in the reference's order, the device reproduces the
    Zall the reference itself shipped, label for label."""
    g = load_golden(f"sampler_{name}.npz")
    out = _run_r(cs, g, niter)
    res = out["results"]
    mism = int((res["Zall"] != g["Zall"][:niter]).sum())
    assert mism == 0, f"{mism} label mismatches"
    k = niter if niter == 200 else niter - 1  # a truncated run overwrites ITS last sweep noise-free
    assert [u.shape[0] for u in res["Uall"]] == list(g["kt"][:niter])
    np.testing.assert_allclose(np.array(res["sigma_all"][:k]), g["sigma_all"][:k], rtol=1e-9)
    np.testing.assert_allclose(res["Likelihood"][:k], g["Likelihood"][:k], rtol=1e-9)
    assert abs(out["priors"]["L0"] / float(g["L0"]) - 1) < 1e-12
    for t in range(k):
        live = np.nonzero(~np.isnan(res["Uall"][t][:, 0]))[0]
        a, b = int(g["u_off"][t]), int(g["u_off"][t + 1])
        assert np.array_equal(live + 1, g["u_labels"][a:b])
        np.testing.assert_allclose(res["Uall"][t][live].sum(axis=1), g["u_rowsum"][a:b], rtol=1e-9, atol=1e-12)
    if name == "BC_ductal2":
        assert out["diagnostics"]["n_reject"] == 207384  # first-sweep identical() rejections
    if niter == 200:  # north-star criterion: ARI >= 0.95 vs the reference's posterior clustering
        zm = cs.majority_vote(cs.MCMCtrim(out)["results"]["Zall"])
        zr = cs.majority_vote(g["Zall"][100:].astype(np.int32))
        assert _ari(zm, zr) >= 0.95 and len(np.unique(zm)) == len(np.unique(zr))


@pytest.mark.parametrize("N,R,na,single", [(400, 300, 0.0, False), (350, 311, 0.02, False), (300, 330, 0.0, False),
                                           (500, 37, 0.03, True), (64, 1, 0.0, False)])
def test_sampler_r_mode_vs_oracle_wide_matrices(cs, oracle, N, R, na, single):
    """R-compatible mode against the oracle's R mode beyond the fixtures' 4-38 segments: with 2R + 1
    draws per cell the pooled stream crosses a Mersenne-Twister block almost every cell (R = 300,
    311: the widest the pool takes), R = 330 runs on the draw-by-draw kernel, and a single-cluster
    start exercises the first sweep's rejection loop.  Labels, sigmas and likelihoods must agree."""
    from clonalscope_b200 import synth
    s = synth.sampler_input(N, R=R, seed=2000 + R, K_true=5, na_frac=na)
    U0 = s["U0"][:1] if single else s["U0"]
    Z0 = oracle.loglik_matrix(s["X"], U0, s["sigmas0"]).argmax(axis=1).astype(np.int32) + 1
    niter = 4
    want = oracle.gibbs(s["X"], U0, Z0, s["sigmas0"], 2.0, 2.0, niter, 200, rng_mode=oracle.RNG_R, ucap_rows=niter * 512)
    got = cs.BayesNonparCluster(s["X"], alpha=2.0, beta=2.0, niter=niter, sigmas0=s["sigmas0"], U0=U0, Z0=Z0, seed=200,
                                rng="R")
    res = got["results"]
    assert np.array_equal(res["Zall"], want["Zall"])
    assert [u.shape[0] for u in res["Uall"]] == list(want["kt"])
    k = niter - 1                                                     # (the last sweep of a run is re-recorded noise-free)
    np.testing.assert_allclose(np.array(res["sigma_all"][:k]), want["sigma_all"][:k], rtol=1e-9)
    np.testing.assert_allclose(res["Likelihood"][:k], want["LL"][:k], rtol=1e-9)


# ------------------------------------------------------------------ sampler, Philox
def _philox_pair(cs, oracle, X, U0, Z0, sig0, niter, seed=200, alpha=2.0, beta=2.0, flags=0):
    want = oracle.gibbs(X, U0, Z0, sig0, alpha, beta, niter, seed, rng_mode=oracle.RNG_PHILOX,
                        ucap_rows=niter * 512)
    got = cs.BayesNonparCluster(X, alpha=alpha, beta=beta, niter=niter, sigmas0=sig0, U0=U0, Z0=Z0, seed=seed,
                                rng="philox", _flags=flags)
    return got, want


def _assert_same_chain(got, want, niter, tol=1e-9):
    res = got["results"]
    mism = int((res["Zall"] != want["Zall"]).sum())
    assert mism == 0, f"{mism} label mismatches, first at sweep {np.argwhere(res['Zall'] != want['Zall'])[0]}"
    assert [u.shape[0] for u in res["Uall"]] == list(want["kt"])
    np.testing.assert_allclose(np.array(res["sigma_all"]), want["sigma_all"], rtol=tol)
    np.testing.assert_allclose(res["Likelihood"], want["LL"], rtol=tol)
    assert abs(got["priors"]["L0"] / want["L0"] - 1) < 1e-12
    for t in range(niter):
        live = np.nonzero(~np.isnan(res["Uall"][t][:, 0]))[0]
        a, b = int(want["u_off"][t]), int(want["u_off"][t + 1])
        assert np.array_equal(live + 1, want["u_labels"][a:b])
        np.testing.assert_allclose(res["Uall"][t][live], want["u_rows"][a:b], rtol=tol, atol=1e-12)


@pytest.mark.parametrize("N,R,na,niter", [(1500, 40, 0.0, 12), (1200, 300, 0.0, 8), (900, 70, 0.02, 10),
                                          (333, 4, 0.0, 25), (64, 1, 0.0, 10), (501, 700, 0.01, 6), (6, 33, 0.0, 5)])
def test_sampler_philox_vs_oracle_synthetic(cs, oracle, N, R, na, niter):
    from clonalscope_b200 import synth
    s = synth.sampler_input(N, R=R, seed=1003 + R, K_true=6, na_frac=na)
    Z0 = oracle.loglik_matrix(s["X"], s["U0"], s["sigmas0"]).argmax(axis=1).astype(np.int32) + 1
    got, want = _philox_pair(cs, oracle, s["X"], s["U0"], Z0, s["sigmas0"], niter)
    _assert_same_chain(got, want, niter)
    # the default Z0 of the shim (device log-likelihood argmax, :91-108) is the same vector
    assert np.array_equal(cs.BayesNonparCluster(s["X"], U0=s["U0"], sigmas0=s["sigmas0"], niter=1)["priors"]["Z0"], Z0)


def test_sampler_philox_extreme_values_clamp_like_the_oracle(cs, oracle):
    """A segment's fixed-point term is clamped at 2^62 (standardised distances beyond 2^15: here
    outliers of 1e7 against sigma 0.05); the clamped sums must give the oracle's chain all the same."""
    from clonalscope_b200 import synth
    s = synth.sampler_input(700, R=40, seed=77, K_true=5)
    X = s["X"].copy()
    rng = np.random.default_rng(3)
    idx = rng.choice(X.size, size=60, replace=False)
    X.ravel()[idx] = rng.choice([1e7, 5e8, 3e4], size=60)
    sig0 = np.full(40, 0.05)
    Z0 = oracle.loglik_matrix(X, s["U0"], sig0).argmax(axis=1).astype(np.int32) + 1
    got, want = _philox_pair(cs, oracle, X, s["U0"], Z0, sig0, 6)
    _assert_same_chain(got, want, 6)


@pytest.mark.parametrize("N,R,K,noise,alpha", [(1200, 12, 45, 0.08, 8.0), (1500, 6, 48, 0.05, 2.0)])
def test_sampler_philox_more_clusters_than_the_resident_walk_holds(cs, oracle, N, R, K, noise, alpha):
    """The resident-window walk keeps 32 clusters in shared memory; when a 33rd opens in the middle
    of a sweep it hands the walk to the generic kernel at that cell, and sweeps that start with more
    run on the generic kernel alone.  Many well separated planted clusters force both."""
    rng = np.random.default_rng(5)
    prof = rng.choice([0.4, 0.9, 1.4, 1.9, 2.4], size=(K, R))
    z = rng.integers(0, K, size=N)
    X = np.clip(prof[z] + noise * rng.standard_normal((N, R)), 0.05, 3.0)
    U0 = np.vstack([np.ones(R), prof[0]])
    sig0 = np.full(R, 0.15)
    Z0 = oracle.loglik_matrix(X, U0, sig0).argmax(axis=1).astype(np.int32) + 1
    got, want = _philox_pair(cs, oracle, X, U0, Z0, sig0, 8, alpha=alpha)
    assert max(want["kt"]) > 32
    _assert_same_chain(got, want, 8)


@pytest.mark.parametrize("name,niter", [("P5931", 40), ("BC_ductal2", 6), ("P5847_tumor", 8)])
def test_sampler_philox_vs_oracle_on_reference_data(cs, oracle, name, niter):
    """Real fixtures are event-dense (8-25 % of the cells move per sweep, kt keeps growing) and
    BC_ductal2 starts from a single cluster (first-sweep rejection loop)."""
    g = load_golden(f"sampler_{name}.npz")
    got, want = _philox_pair(cs, oracle, g["X"], g["U0"], g["Z0"], g["sigmas0"], niter, seed=int(g["seed"]),
                             alpha=float(g["alpha"]), beta=float(g["beta"]))
    _assert_same_chain(got, want, niter)
    assert got["diagnostics"]["n_reject"] == want["n_reject"]


def test_speculative_and_sequential_evaluation_agree(cs):
    """flags bit 0 disables the speculative prepare pass (every cell goes through the scan
    kernel): the chain must be identical, which is the 'sequential order preserved' claim."""
    from clonalscope_b200 import synth
    s = synth.sampler_input(2500, R=120, seed=77, K_true=8)
    kw = dict(alpha=2.0, beta=2.0, niter=10, sigmas0=s["sigmas0"], U0=s["U0"], seed=5, rng="philox")
    a = cs.BayesNonparCluster(s["X"], **kw)
    b = cs.BayesNonparCluster(s["X"], _flags=1, **kw)
    assert np.array_equal(a["results"]["Zall"], b["results"]["Zall"])
    for ua, ub in zip(a["results"]["Uall"], b["results"]["Uall"]):
        assert np.array_equal(ua, ub, equal_nan=True)
    assert np.array_equal(a["results"]["Likelihood"], b["results"]["Likelihood"])


@pytest.mark.parametrize("name,mincell,ari_lo,ari_hi,ncl_lo,ncl_hi", [
    ("P5931", 13, 0.95, 0.975, 5, 8),
    ("BC_ductal2", 40, 0.70, 0.84, 12, 17)])
def test_philox_chain_is_in_the_reference_envelope(cs, name, mincell, ari_lo, ari_hi, ncl_lo, ncl_hi):
    """Distributional criterion for the Philox stream: 200 sweeps, burn-in 100,
    majority vote -> ARI against the reference's own posterior clustering (its shipped Zall) and the
    number of clusters above mincell.  The reference's own seed-to-seed envelope (R stream, seeds
    200-203): P5931 ARI 0.954-0.964 / 6-8 clusters, BC_ductal2 0.737-0.778 / 14-17.  The
    Philox stream over seeds 200-207, measured on the B200 (profiles/r2_philox_envelope.json,
    tools/envelope.py): P5931 0.956-0.972 / 5-7, BC_ductal2 0.709-0.830 / 12-16 — the same
    distribution sampled eight times instead of four; the bounds asserted here are the union of
    the two ranges."""
    g = load_golden(f"sampler_{name}.npz")
    out = cs.BayesNonparCluster(g["X"], cna_states_WGS=g["cna_states_WGS"], alpha=float(g["alpha"]),
                                beta=float(g["beta"]), niter=200, sigmas0=g["sigmas0"], U0=g["U0"], Z0=g["Z0"],
                                seed=200, rng="philox")
    zm = cs.majority_vote(cs.MCMCtrim(out)["results"]["Zall"])
    zr = cs.majority_vote(g["Zall"][100:].astype(np.int32))
    ari = _ari(zm, zr)
    lab, cnt = np.unique(zm, return_counts=True)
    nclu = int((cnt > mincell).sum())
    print("philox", name, "ARI", ari, "subclones", nclu)
    assert ari_lo <= ari <= ari_hi and ncl_lo <= nclu <= ncl_hi


@pytest.mark.parametrize("rng_name", ["philox", "R"])
def test_tracing_flow_u0_with_na_rows(cs, oracle, rng_name):
    """The tracing / update flow of R/RunCovCluster.R:118-134: AssignCluster's Uest — NA rows for every
    dissolved or never-occupied label — goes back in as U0.  The reference keeps those rows (:47 only
    truncates behind the last non-NA row): they hold no cell, are never scored (:193) or proposed
    (:170), and still count towards kt and the numbering of new labels (:150-151, :209-218)."""
    g = load_golden("sampler_P5931.npz")
    X, R = g["X"], g["X"].shape[1]
    obj = dict(results=dict(Zall=g["Zall"][100:].astype(np.int32), Uall=[np.zeros((int(g["kt"][-1]), R))]),
               priors=dict(U0=g["U0"], cna_states_WGS=g["cna_states_WGS"], wU0=True), data=X)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res0 = cs.AssignCluster(obj, mincell=13)
    Uest = res0["Uest"].values if hasattr(res0["Uest"], "values") else res0["Uest"]
    assert Uest.shape[0] == 424 and np.isnan(Uest[:, 0]).sum() == 424 - 8
    niter = 8
    got = cs.BayesNonparCluster(X, cna_states_WGS=g["cna_states_WGS"], niter=niter, U0=Uest, seed=200, rng=rng_name)
    U0t = Uest[:192]                                     # :47 up to the last non-NA row (label 192)
    live = np.nonzero(~np.isnan(U0t[:, 0]))[0]
    assert got["priors"]["U0"].shape == (192, R) and got["priors"]["wU0"] is True
    sig0 = np.full(R, 0.25)
    Z0 = (live[oracle.loglik_matrix(X, U0t[live], sig0).argmax(axis=1)] + 1).astype(np.int32)
    assert np.array_equal(got["priors"]["Z0"], Z0)
    want = oracle.gibbs(X, U0t, Z0, sig0, 2.0, 2.0, niter, 200,
                        rng_mode=oracle.RNG_PHILOX if rng_name == "philox" else oracle.RNG_R, ucap_rows=niter * 512)
    _assert_same_chain(got, want, niter)
    assert want["kt"][0] > 192 and set(np.unique(want["Zall"][0])) - set(live + 1)   # new labels start at 193


def test_u0_na_rows_synthetic_and_errors(cs, oracle):
    """NA rows in front of, between and (truncated) behind the live rows; a live row without cells."""
    from clonalscope_b200 import synth
    s = synth.sampler_input(1500, R=40, seed=31, K_true=5)
    R = 40
    nan = np.full(R, np.nan)
    U0 = np.vstack([nan, s["U0"][0], nan, nan, s["U0"][1], np.full(R, 2.4), nan, nan])
    live = np.array([1, 4, 5])
    Z0 = (live[oracle.loglik_matrix(s["X"], U0[live], s["sigmas0"]).argmax(axis=1)] + 1).astype(np.int32)
    for rng_name, mode in (("philox", oracle.RNG_PHILOX), ("R", oracle.RNG_R)):
        got = cs.BayesNonparCluster(s["X"], niter=6, sigmas0=s["sigmas0"], U0=U0, seed=9, rng=rng_name)
        assert np.array_equal(got["priors"]["Z0"], Z0) and got["priors"]["U0"].shape[0] == 6
        want = oracle.gibbs(s["X"], U0[:6], Z0, s["sigmas0"], 2.0, 2.0, 6, 9, rng_mode=mode, ucap_rows=6 * 512)
        _assert_same_chain(got, want, 6)
        assert want["kt"][0] >= 6
    Zbad = Z0.copy()
    Zbad[3] = 3                                          # an NA row of U0
    with pytest.raises(cs.ClonalscopeError) as e:
        cs.BayesNonparCluster(s["X"], niter=2, sigmas0=s["sigmas0"], U0=U0, Z0=Zbad, rng="philox")
    assert e.value.status == 6 and "NA row" in str(e.value)


def test_kcap_overflow_is_reported(cs):
    from clonalscope_b200 import synth
    s = synth.sampler_input(400, R=8, seed=3, K_true=6)
    with pytest.raises(cs.ClonalscopeError) as e:
        cs.BayesNonparCluster(s["X"], U0=s["U0"], sigmas0=s["sigmas0"], niter=30, rng="philox", kcap=2)
    assert e.value.status == 3


# ------------------------------------------------------------------ allele variant (parity UNPINNED by the reference)
def _allele_data(oracle, N=260, R=7, seed=0, na=0.0):
    rng = np.random.default_rng(seed)
    G = oracle.genotype_grid(6)
    true = np.array([[4] * R, [4, 4, 7, 7, 12, 4, 4][:R], [4, 2, 4, 9, 4, 4, 14][:R]])
    z = rng.integers(0, 3, N)
    Xc = G[true[z] - 1, 0] + 0.15 * rng.standard_normal((N, R))
    Xc[0, 0] = 3.7  # capped at 3 inside the sampler (:37)
    Xa = np.clip(G[true[z] - 1, 1] + 0.08 * rng.standard_normal((N, R)), 0, 1)
    if na > 0:
        Xc[rng.random((N, R)) < na] = np.nan
        Xa[rng.random((N, R)) < na] = np.nan
    return Xc, Xa, true, z


@pytest.mark.parametrize("rng_name", ["R", "philox"])
@pytest.mark.parametrize("na", [0.0, 0.03])
def test_allele_sampler_vs_oracle(cs, oracle, rng_name, na):
    """No reference output exists for BayesNonparAlleleCluster; the device chain is compared
    with the CPU restatement of R/BayesNonparAlleleCluster.R draw for draw."""
    Xc, Xa, true, z = _allele_data(oracle, na=na)
    niter = 25
    mode = oracle.RNG_R if rng_name == "R" else oracle.RNG_PHILOX
    got = cs.BayesNonparAlleleCluster(Xc, Xa, cna_states_WGS=true[1], niter=niter, seed=200, rng=rng_name)
    Z0 = got["priors"]["Z0"]
    U0 = np.vstack([np.full(Xc.shape[1], 4), true[1]])
    want = oracle.gibbs_allele(Xc, Xa, U0, Z0, 0.3, 0.15, niter=niter, seed=200, rng_mode=mode)
    res = got["results"]
    mism = int((res["Zall"] != want["Zall"]).sum())
    assert mism == 0, f"{mism} label mismatches"
    assert [u.shape[0] for u in res["Uall"]] == list(want["kt"])
    np.testing.assert_allclose(np.array(res["sigma_cov_all"]), want["sig_cov_all"], rtol=1e-9)
    np.testing.assert_allclose(np.array(res["sigma_allele_all"]), want["sig_allele_all"], rtol=1e-9)
    np.testing.assert_allclose(res["Likelihood"], want["LL"], rtol=1e-9)
    assert abs(got["priors"]["L0"] / want["L0"] - 1) < 1e-12
    for t in range(niter):
        a, b = int(want["u_off"][t]), int(want["u_off"][t + 1])
        lab = want["u_labels"][a:b] - 1
        assert np.array_equal(res["Uall"][t][lab], want["u_state"][a:b])
        np.testing.assert_allclose(res["Ucov"][t][lab], want["u_cov"][a:b], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(res["Uallele"][t][lab], want["u_allele"][a:b], rtol=1e-9, atol=1e-12)
        dead = np.setdiff1d(np.arange(res["Uall"][t].shape[0]), lab)
        assert not res["Uall"][t][dead].any()  # emptied clusters keep all-zero rows, not NA (:203-205)
    # invariants: states live on the maxcp = 6 grid, trimmed object has the allele fields
    assert res["Uall"][-1].max() <= 27
    tr = cs.MCMCtrim(got, burnin=10, allele=True)
    assert len(tr["results"]["Ucov"]) == niter - 10 and tr["results"]["Zall"].shape[0] == niter - 10


def test_allele_shim_reference_quirks(cs, oracle):
    Xc, Xa, true, z = _allele_data(oracle)
    with pytest.raises(NameError):  # a user U0 with more than one row hits the undefined `result$Uest`
        cs.BayesNonparAlleleCluster(Xc, Xa, U0=np.vstack([np.full(7, 4), true[1]]), niter=2)
    with pytest.raises(cs.ClonalscopeError) as e:  # a cluster of U0 without cells: sample.int would fail
        cs.BayesNonparAlleleCluster(Xc, Xa, cna_states_WGS=true[1], Z0=np.ones(len(z), dtype=int), niter=2)
    assert e.value.status == 6
    assert oracle.genotype_neighbor(1.02, 0.48) == 4 and oracle.genotype_neighbor(1.6, 0.7) == 8


# ------------------------------------------------------------------ cell-sharded sufficient statistics
def test_device_stats_match_numpy(cs):
    import torch
    from clonalscope_b200 import distributed as D
    rng = np.random.default_rng(4)
    for (N, R, K) in [(1000, 13, 5), (5000, 300, 20), (300, 40, 80)]:
        X = rng.normal(1.0, 0.3, size=(N, R))
        X[rng.random((N, R)) < 0.03] = np.nan
        Z = rng.integers(1, K + 1, size=N).astype(np.int32)
        U = rng.uniform(0.5, 2.0, size=(K, R))
        out = D.allreduce_stats(torch.from_numpy(X).cuda(), torch.from_numpy(Z).cuda(), torch.from_numpy(U).cuda(), K)
        for k in range(K):
            rows = X[Z == k + 1]
            assert int(out["n_k"][k].item()) == len(rows)
            cnt = (~np.isnan(rows)).sum(axis=0)
            want = np.where(cnt > 0, np.nansum(rows, axis=0) / np.maximum(cnt, 1), 0.0)
            np.testing.assert_allclose(out["mean"][k].cpu().numpy(), want, rtol=1e-11, atol=1e-12)
        d = X - U[Z - 1]
        sig = np.sqrt(np.nansum(d * d, axis=0) / (~np.isnan(X)).sum(axis=0))
        np.testing.assert_allclose(out["sigma"].cpu().numpy(), sig, rtol=1e-11)


# ------------------------------------------------------------------ BASELINE.json's full sizes, through properties
def test_full_size_binning_properties(cs):
    """100k cells x 30k genes -> 14,385 bins (the bench workload), through the device entry point.
    Size-independent properties: every column's sum equals sum_g x_g * (bins of g) (integer
    exact), row indices strictly ascend within a column, no stored zero, the column pointer is
    the running count, and a second pass over the same buffers gives byte-identical output."""
    import ctypes as C
    import torch
    import bench
    from clonalscope_b200 import synth
    dev = torch.device("cuda", 0)
    L = cs.lib()
    N, G = 100000, 30000
    bins, genes = synth.geometry(G=G)
    mp, mb = synth.geometry_map(bins, genes)
    B = len(bins["chr"])
    colptr, rowidx, x = bench.synth_counts_gpu(torch, dev, N, G, seed=4242)
    dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    plan = C.c_void_p()
    assert L.cs_bin_plan_create(G, B, dp(mp, C.c_int32), dp(mb, C.c_int32), C.byref(plan)) == 0
    try:
        ocp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64()
        assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), None, None, 0, C.byref(nnz), None) == 0
        cap = nnz.value
        ori = torch.empty(cap, dtype=torch.int32, device=dev)
        ox = torch.empty(cap, dtype=torch.float64, device=dev)
        assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), vp(ori), vp(ox), cap, C.byref(nnz), None) == 0, L.cs_last_error()
        assert nnz.value == cap == int(ocp[-1])
        assert bool((ocp[1:] >= ocp[:-1]).all()) and int(ocp[0]) == 0
        assert bool((ox != 0).all()) and bool((ox == ox.round()).all())
        assert int(ori.min()) >= 0 and int(ori.max()) < B
        # strictly ascending bins within each column
        col_of_out = torch.repeat_interleave(torch.arange(N, device=dev), (ocp[1:] - ocp[:-1]))
        same_col = col_of_out[1:] == col_of_out[:-1]
        assert bool((ori[1:][same_col] > ori[:-1][same_col]).all())
        # per-column sums: sum of outputs == sum_g x_g * nh_g (all values are small integers: exact in FP64)
        nh = torch.from_numpy(np.diff(mp).astype(np.float64)).to(dev)
        col_of_in = torch.repeat_interleave(torch.arange(N, device=dev), (colptr[1:] - colptr[:-1]))
        want = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, col_of_in, x * nh[rowidx.long()])
        got = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, col_of_out, ox)
        assert bool((want == got).all())
        # per-bin totals over all cells (a checksum of checksums across the other axis)
        first = torch.from_numpy(mb[mp[:-1].clip(max=len(mb) - 1)].astype(np.int64)).to(dev)
        hits_g = torch.zeros(G, dtype=torch.float64, device=dev).index_add_(0, rowidx.long(), x)
        want_bins = torch.zeros(B + 32, dtype=torch.float64, device=dev)
        for k in range(int(np.diff(mp).max())):
            sel = torch.from_numpy((np.diff(mp) > k)).to(dev)
            want_bins.index_add_(0, (first + k)[sel], hits_g[sel])
        got_bins = torch.zeros(B + 32, dtype=torch.float64, device=dev).index_add_(0, ori.long(), ox)
        assert bool((want_bins == got_bins).all())
        # idempotence: the same call again gives the same bytes
        ori2, ox2, ocp2 = torch.empty_like(ori), torch.empty_like(ox), torch.zeros_like(ocp)
        assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp2), vp(ori2), vp(ox2), cap, C.byref(nnz), None) == 0
        assert torch.equal(ocp, ocp2) and torch.equal(ori, ori2) and torch.equal(ox, ox2)
    finally:
        L.cs_bin_plan_destroy(plan)


def test_full_size_binning_vs_oracle(cs, oracle):
    """BASELINE config 4 at full size against the oracle itself: 100k cells x 30k genes -> 14,385
    bins, every output slot (column pointers, bin ids, sums) byte-identical to the CPU restatement's
    sparse sweep (~6 s on one core)."""
    import ctypes as C
    import torch
    import bench
    from clonalscope_b200 import synth
    dev = torch.device("cuda", 0)
    L = cs.lib()
    N, G = 100000, 30000
    bins, genes = synth.geometry(G=G)
    mp, mb = synth.geometry_map(bins, genes)
    B = len(bins["chr"])
    colptr, rowidx, x = bench.synth_counts_gpu(torch, dev, N, G, seed=1001)
    dp = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    plan = C.c_void_p()
    assert L.cs_bin_plan_create(G, B, dp(mp, C.c_int32), dp(mb, C.c_int32), C.byref(plan)) == 0
    try:
        ocp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        nnz = C.c_int64()
        assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), None, None, 0, C.byref(nnz), None) == 0
        cap = nnz.value
        ori = torch.empty(cap, dtype=torch.int32, device=dev)
        ox = torch.empty(cap, dtype=torch.float64, device=dev)
        assert L.cs_bin_counts_dev(plan, N, vp(colptr), vp(rowidx), vp(x), vp(ocp), vp(ori), vp(ox), cap, C.byref(nnz), None) == 0, L.cs_last_error()
    finally:
        L.cs_bin_plan_destroy(plan)
    wcp, wri, wx = oracle.bin_counts(G, N, B, colptr.cpu().numpy(), rowidx.cpu().numpy(), x.cpu().numpy(), mp, mb)
    assert nnz.value == len(wri) > 2e8
    assert np.array_equal(ocp.cpu().numpy(), wcp)
    assert np.array_equal(ori.cpu().numpy(), wri)
    assert np.array_equal(ox.cpu().numpy(), wx)


def test_full_size_sampler_vs_oracle(cs, oracle):
    """BASELINE config 4 at full size against the oracle itself: 100k cells x 300 segments, the first
    five (event-dense: ~22k, ~27k, ... label changes) sweeps of the bench chain, label for label and
    Uall / sigma / Likelihood against the CPU restatement's sequential chain (~3.5 s per sweep)."""
    from clonalscope_b200 import synth
    N, R, niter = 100000, 300, 5
    s = synth.sampler_input(N, R=R)
    Z0 = (cs.loglik_matrix(s["X"], s["U0"], s["sigmas0"]).argmax(axis=1) + 1).astype(np.int32)
    got, want = _philox_pair(cs, oracle, s["X"], s["U0"], Z0, s["sigmas0"], niter)
    _assert_same_chain(got, want, niter)
    moved = sum(int((want["Zall"][t] != want["Zall"][t - 1]).sum()) for t in range(1, niter))
    assert moved > 30000


def test_full_size_sampler_properties(cs):
    """100k cells x 300 segments (the bench workload), a few event-dense sweeps: the speculative
    walk (resident window, cluster helpers) and the plainly sequential kernel give the SAME chain,
    and the chain's book-keeping is consistent."""
    from clonalscope_b200 import synth
    N, R, niter = 100000, 300, 5
    s = synth.sampler_input(N, R=R)
    kw = dict(alpha=2.0, beta=2.0, niter=niter, sigmas0=s["sigmas0"], U0=s["U0"], seed=200, rng="philox")
    a = cs.BayesNonparCluster(s["X"], **kw)
    b = cs.BayesNonparCluster(s["X"], _flags=1, **kw)
    Za, Zb = a["results"]["Zall"], b["results"]["Zall"]
    assert np.array_equal(Za, Zb)
    assert np.array_equal(a["results"]["Likelihood"], b["results"]["Likelihood"])
    for ua, ub in zip(a["results"]["Uall"], b["results"]["Uall"]):
        assert np.array_equal(ua, ub, equal_nan=True)
    kt = [u.shape[0] for u in a["results"]["Uall"]]
    assert all(k2 >= k1 for k1, k2 in zip(kt, kt[1:]))                # labels are never renumbered
    for t in range(niter):
        z = Za[t]
        assert z.min() >= 1 and z.max() <= kt[t]
        live = np.nonzero(~np.isnan(a["results"]["Uall"][t][:, 0]))[0] + 1
        assert set(np.unique(z)) == set(live)                          # occupied clusters == non-NA rows
        assert np.isfinite(a["results"]["Likelihood"][t])
        sg = np.asarray(a["results"]["sigma_all"][t])
        assert np.all(np.isfinite(sg)) and np.all(sg > 0)
    moved = sum(int((Za[t] != Za[t - 1]).sum()) for t in range(1, niter))
    assert moved > 30000                                               # the sweeps really are event-dense


# ------------------------------------------------------------------ a7: EstRegionCov
def _region_cov_case(seed, G=500, N=320, nseg=9):
    rng = np.random.default_rng(seed)
    lam = rng.lognormal(-1.0, 1.0, size=G)
    mtx = rng.poisson(lam[:, None] * rng.lognormal(0, 0.3, size=N)[None, :]).astype(np.float64)
    mtx[rng.choice(G, 12, replace=False)] = 0.0                  # genes without a single read
    mtx[:, rng.choice(N, 3, replace=False)] *= 0                 # empty cells
    barcodes = [f"AAAC{i:05d}-1" for i in range(N)]
    features = [f"ENSG{i:08d}" for i in range(G)]
    chrom = np.sort(rng.integers(1, 5, size=G))
    start = np.zeros(G, dtype=np.int64)
    for c in range(1, 5):
        sel = chrom == c
        start[sel] = np.sort(rng.integers(1000, 2_000_000, size=int(sel.sum())))
    end = start + rng.integers(500, 60_000, size=G)
    order = rng.permutation(G)                                   # the bed file is not in matrix order
    bed = np.empty((G - 7, 4), dtype=object)                     # ... and misses a few genes
    o2 = order[:G - 7]
    bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3] = [str(c) for c in chrom[o2]], start[o2], end[o2], [features[i] for i in o2]
    perm = rng.permutation(N)
    celltype0 = np.empty((N - 5, 2), dtype=object)               # a subset of the cells, shuffled
    celltype0[:, 0] = [barcodes[i] for i in perm[:N - 5]]
    celltype0[:, 1] = rng.choice(["tumor", "normal"], size=N - 5, p=[0.7, 0.3])
    seg = dict(chr=[], start=[], end=[], chrr=[], states=[])
    for c in range(1, 5):
        cuts = np.sort(rng.choice(np.arange(50_000, 1_950_000, 10_000), size=nseg // 4, replace=False))
        edges = [0] + list(cuts) + [2_100_000]
        for a, b_ in zip(edges[:-1], edges[1:]):
            seg["chr"].append(str(c)), seg["start"].append(float(a)), seg["end"].append(float(b_))
            seg["chrr"].append(f"chr{c}_{a}_{b_}"), seg["states"].append(int(rng.integers(0, 3)))
    seg["chr"].append("7"), seg["start"].append(0.0), seg["end"].append(1e6), seg["chrr"].append("chr7_0_1e6"), seg["states"].append(1)
    return mtx, barcodes, features, bed, celltype0, seg


@pytest.mark.parametrize("include,alpha_source,seed", [("tumor", "all", 1), ("all", "control", 2), ("control", "all", 3)])
def test_est_region_cov_vs_oracle(cs, oracle, include, alpha_source, seed):
    """Device moments + segment aggregation vs the dense, line-by-line restatement: the same genes
    pass the filter, the same cells are selected per segment, deltas agree to 1e-12 (the reference
    divides every count by its cell's alpha before summing, here the integer sum is divided)."""
    import scipy.sparse as sp
    from clonalscope_b200.api import EstRegionCov
    mtx, barcodes, features, bed, celltype0, seg = _region_cov_case(seed)
    kw = dict(var_pt=0.97, var_pt_ctrl=0.95, ngene_filter=1, include=include, alpha_source=alpha_source,
              ctrl_region=["chr1", "chr3"])
    got = EstRegionCov(mtx=sp.csc_matrix(mtx), barcodes=np.asarray(barcodes, dtype=object).reshape(-1, 1),
                       features=np.asarray(features, dtype=object).reshape(-1, 1), bed=bed, celltype0=celltype0,
                       seg_table_filtered=seg, **kw)
    want = oracle.est_region_cov(mtx, barcodes, features, [bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3]],
                                 (list(celltype0[:, 0]), list(celltype0[:, 1])), seg_table=seg, **kw)
    assert list(got["deltas_all"].keys()) == [n for n, _, _ in want["deltas_all"]]
    assert len(want["deltas_all"]) >= 8
    for name, cells, vals in want["deltas_all"]:
        gc, gv = got["deltas_all"][name]
        assert gc == cells
        np.testing.assert_allclose(gv, vals, rtol=1e-12, equal_nan=True)
    assert list(got["ngenes"].items()) == want["ngenes"]
    assert got["seg_table_filtered"]["chrr"] == want["sel"] and "chr7_0_1e6" not in want["sel"]
    assert got["barcodes_normal"] == want["barcodes_normal"]
    np.testing.assert_allclose(got["alpha"], want["alpha"], rtol=0, atol=0)


def test_region_cov_device_halves_equal_the_one_call_form(cs):
    """cs_csc_region_sums_dev + cs_region_ratios_dev (what the cell-sharded form runs around its
    all-gather, here at world size 1) == cs_csc_region_cov, bit for bit."""
    import scipy.sparse as sp
    from clonalscope_b200 import distributed as D
    from clonalscope_b200.api import _CscHandle
    rng = np.random.default_rng(13)
    G, N, S = 300, 800, 7
    m = sp.random(G, N, density=0.2, random_state=3, format="csc", data_rvs=lambda k: rng.integers(1, 9, k).astype(float))
    m.sort_indices()
    hits = [sorted(set(rng.integers(0, S, size=rng.integers(0, 3)).tolist())) for _ in range(G)]
    map_ptr = np.zeros(G + 1, dtype=np.int32)
    map_ptr[1:] = np.cumsum([len(h) for h in hits])
    map_seg = np.array([s_ for h in hits for s_ in h], dtype=np.int32)
    lib_flag = (rng.random(G) < 0.8).astype(np.uint8)
    code = rng.choice([0, 1, 2], size=N, p=[0.3, 0.6, 0.1]).astype(np.int32)
    h = _CscHandle(m)
    for include in (0, 1, 2):
        want_d, want_s, want_a, want_n = h.region_cov(S, map_ptr, map_seg, lib_flag, code, include)
        got = D.region_cov_sharded(h, S, map_ptr, map_seg, lib_flag, code, include)
        assert np.array_equal(got["deltas"].cpu().numpy(), want_d, equal_nan=True)
        assert np.array_equal(got["sums"].cpu().numpy(), want_s) and np.array_equal(got["alpha"].cpu().numpy(), want_a)
        assert np.array_equal(got["nsel"].cpu().numpy(), want_n) and got["offset"] == 0 and got["n_total"] == N
    h.free()


def test_est_region_cov_default_segments_and_errors(cs, oracle):
    import scipy.sparse as sp
    from clonalscope_b200.api import EstRegionCov
    mtx, barcodes, features, bed, celltype0, _ = _region_cov_case(5)
    size = np.empty((24, 2), dtype=object)
    size[:, 0] = [f"chr{c}" for c in range(1, 25)]
    size[:, 1] = [2_100_000] * 24
    got = EstRegionCov(mtx=sp.csc_matrix(mtx), barcodes=barcodes, features=features, bed=bed, celltype0=None, size=size)
    want = oracle.est_region_cov(mtx, barcodes, features, [bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3]], None,
                                 size=(list(size[:, 0]), list(size[:, 1])))
    assert list(got["deltas_all"].keys()) == [n for n, _, _ in want["deltas_all"]] == \
        [f"pt0.99_tumor_alphaallchr{c}" for c in range(1, 5)]
    for name, cells, vals in want["deltas_all"]:
        assert got["deltas_all"][name][0] == cells == []           # no tumor cell without a celltype table
    assert got["ngenes"] == dict(want["ngenes"])
    with pytest.raises(ValueError):
        EstRegionCov(mtx=sp.csc_matrix(mtx), barcodes=barcodes, features=features, bed=bed, size=size, include="both")


def test_chains_on_two_devices_from_one_process(cs):
    """`device=` selects the GPU per call; per-device kernel attributes must follow (needs 2 GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from clonalscope_b200 import synth
    s = synth.sampler_input(3000, R=300, seed=9, K_true=6)
    kw = dict(alpha=2.0, beta=2.0, niter=6, sigmas0=s["sigmas0"], U0=s["U0"], seed=11, rng="philox")
    a = cs.BayesNonparCluster(s["X"], device=0, **kw)
    b = cs.BayesNonparCluster(s["X"], device=1, **kw)
    c = cs.BayesNonparCluster(s["X"], device=0, **kw)
    assert np.array_equal(a["results"]["Zall"], b["results"]["Zall"])
    assert np.array_equal(a["results"]["Zall"], c["results"]["Zall"])
    assert np.array_equal(a["results"]["Likelihood"], b["results"]["Likelihood"])


def test_region_cov_feeds_prep_and_the_sampler(cs, tmp_path):
    """The tutorial's chain on one synthetic sample, file to file: matrix.mtx.gz -> readMM -> EstRegionCov
    (a7) -> PrepCovMatrix (a2) -> BayesNonparCluster (a3) -> MCMCtrim -> AssignCluster -> saveRDS; the
    coverage matrix separates a planted gain when clustered and the saved object reads back."""
    import gzip
    import scipy.sparse as sp
    rng = np.random.default_rng(12)
    G, N = 1200, 600
    chrom = np.sort(rng.integers(1, 4, size=G))
    start = np.concatenate([np.sort(rng.integers(1, 3_000_000, size=int((chrom == c).sum()))) for c in (1, 2, 3)])
    lam = rng.lognormal(0.3, 0.6, size=G)
    tumor = np.arange(N) >= 200
    gain = (chrom == 2)[:, None] & (tumor & (np.arange(N) % 2 == 0))[None, :]      # half of the tumor cells gain chr2
    mtx = rng.poisson(lam[:, None] * np.where(gain, 1.6, 1.0) * rng.lognormal(0, 0.2, size=N)[None, :]).astype(np.float64)
    feats = [f"g{i}" for i in range(G)]
    bars = [f"c{i}" for i in range(N)]
    bed = np.empty((G, 4), dtype=object)
    bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3] = [str(c) for c in chrom], start, start + 5000, feats
    ct = np.empty((N, 2), dtype=object)
    ct[:, 0], ct[:, 1] = bars, np.where(tumor, "tumor", "normal")
    seg = dict(chr=["1", "2", "3"], start=[0.0, 0.0, 0.0], end=[4e6, 4e6, 4e6], chrr=["chr1", "chr2", "chr3"],
               states=[1, 1.5, 1])
    coo = sp.coo_matrix(sp.csc_matrix(mtx))
    order = np.lexsort((coo.row, coo.col))
    with gzip.open(tmp_path / "matrix.mtx.gz", "wt") as f:                 # what cellranger writes
        f.write(f"%%MatrixMarket matrix coordinate integer general\n%metadata_json: {{}}\n{G} {N} {coo.nnz}\n")
        f.write("".join(f"{r + 1} {c + 1} {int(v)}\n" for r, c, v in zip(coo.row[order], coo.col[order], coo.data[order])))
    m = cs.readMM(tmp_path / "matrix.mtx.gz")
    assert (m != sp.csc_matrix(mtx)).nnz == 0
    est = cs.EstRegionCov(m, bars, feats, bed, ct, seg_table_filtered=seg)
    prep = cs.PrepCovMatrix(est, ngene_filter=50)
    X = prep["df"]
    assert X.values.shape == (400, 3) and X.colnames == ["chr1-0-4e+06", "chr2-0-4e+06", "chr3-0-4e+06"]
    out = cs.BayesNonparCluster(X.values, cna_states_WGS=np.array([1, 1.5, 1.0]), niter=40, rng="philox", seed=3)
    z = cs.majority_vote(cs.MCMCtrim(out)["results"]["Zall"])
    planted = (np.arange(200, 600) % 2 == 0).astype(int)
    ratio = np.nanmean(X.values[planted == 1, 1]) / np.nanmean(X.values[planted == 0, 1])
    assert 1.25 < ratio < 1.5      # the planted 1.6x gain of chr2, diluted by its own share of the library size (1.6 / 1.2)
    purity = sum(max(int(((z == k) & (planted == 1)).sum()), int(((z == k) & (planted == 0)).sum()))
                 for k in np.unique(z)) / len(z)
    assert purity > 0.9                                                    # clusters do not mix the two groups
    final = cs.AssignCluster(cs.MCMCtrim(out), mincell=20)
    assert set(np.unique(final["annot"])) <= {"T", "N"} and len(final["Zest"]) == 400
    import rds_reader as rr
    obj = dict(clustering=dict(results=dict(Zall=out["results"]["Zall"], Likelihood=out["results"]["Likelihood"]),
                               data=X), result=dict(Zest=final["Zest"], annot=list(final["annot"]), Uest=final["Uest"]))
    cs.saveRDS(obj, str(tmp_path / "nonpara_clustering.rds"))
    back = rr.read_rds(str(tmp_path / "nonpara_clustering.rds"))
    assert np.array_equal(back["clustering"]["results"]["Zall"].data.reshape(400, 40).T, out["results"]["Zall"])
    assert list(back["clustering"]["data"].attrs["dimnames"][1].data) == X.colnames
    assert list(back["result"]["annot"].data) == list(final["annot"])


# ------------------------------------------------------------------ AssignCluster
@pytest.mark.parametrize("name,mincell", [("P5931", 20), ("P5931", 100), ("P5847_tumor", 20)])
def test_assign_cluster_vs_oracle(cs, oracle, name, mincell):
    """The posterior step on a chain that reproduces the reference's own (R-compatible mode, the
    fixture's seed): same tally, the same orphans re-drawn to the same clusters under set.seed(100),
    means / sigmas / malignancy index to 1e-10."""
    g = load_golden(f"sampler_{name}.npz")
    niter = 40
    out = cs.BayesNonparCluster(g["X"], cna_states_WGS=g["cna_states_WGS"], alpha=float(g["alpha"]),
                                beta=float(g["beta"]), niter=niter, sigmas0=g["sigmas0"], U0=g["U0"], Z0=g["Z0"],
                                seed=int(g["seed"]), rng="r")
    assert np.array_equal(out["results"]["Zall"], g["Zall"][:niter])
    trimmed = cs.MCMCtrim(out)
    got = cs.AssignCluster(trimmed, mincell=mincell)
    want = oracle.assign_cluster(trimmed["results"]["Zall"], trimmed["results"]["Uall"][-1].shape[0], g["X"],
                                 g["cna_states_WGS"], mincell=mincell)
    assert np.array_equal(got["Zmaj"], want["Zmaj"])
    assert (got["Zest"] != got["Zmaj"]).sum() > 0                   # some cells really were re-drawn
    assert np.array_equal(got["Zest"], want["Zest"])
    U = got["Uest"].values if hasattr(got["Uest"], "values") else got["Uest"]
    np.testing.assert_allclose(U, want["Uest"], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(got["sigmas_est"], want["sigmas_est"], rtol=1e-10)
    np.testing.assert_allclose(got["corrs"], want["corrs"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(got["annot"], want["annot"])


@pytest.mark.parametrize("name,mincell,labels", [
    ("P5931", 13, {1, 2, 4, 53, 62, 73, 100, 192}),
    ("BC_ductal2", 40, {1, 7, 9, 13, 19, 25, 26, 27, 37, 49, 141, 146, 150, 160, 201}),
    ("P5847_tumor", 50, {1, 2}),
    ("P5847_allcells", 50, {1, 3, 5, 6, 9, 36, 39, 80, 125})])
def test_assign_cluster_on_the_reference_chains(cs, name, mincell, labels):
    """On the reference's own stored chains (burn-in 100) the final assignment keeps exactly the
    published subclones and gives every cell one of them."""
    g = load_golden(f"sampler_{name}.npz")
    R = g["X"].shape[1]
    obj = dict(results=dict(Zall=g["Zall"][100:].astype(np.int32), Uall=[np.zeros((int(g["kt"][-1]), R))]),
               priors=dict(U0=g["U0"], cna_states_WGS=g["cna_states_WGS"], wU0=True), data=g["X"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = cs.AssignCluster(obj, mincell=mincell)
    lab, cnt = np.unique(out["Zest"], return_counts=True)
    assert set(int(l) for l in lab) == labels
    zl, zc = np.unique(out["Zmaj"], return_counts=True)
    before = {int(l): int(c) for l, c in zip(zl, zc)}
    assert all(c >= before[int(l)] for l, c in zip(lab, cnt)) and cnt.sum() == g["X"].shape[0]
    assert set(np.unique(out["annot"])) <= {"T", "N"} and np.all(np.isfinite(out["sigmas_est"]))


@pytest.mark.parametrize("mincell,na,rm_extreme", [(81, 0.0, False), (81, 0.05, True), (5, 0.0, False)])
def test_assign_cluster_allele_branch_vs_oracle(cs, oracle, mincell, na, rm_extreme):
    """AssignCluster(allele=TRUE) after BayesNonparAlleleCluster + MCMCtrim against the restated
    R/AssignCluster.R:216-414 (no reference output exists: parity unpinned): the same orphans re-drawn
    to the same clusters under set.seed(100), genotype states identical, means / sigmas / index 1e-10.
    With mincell=5 no cluster is dissolved and the reference's sigmas come out 0 (x[-integer(0), ])."""
    Xc, Xa, true, z = _allele_data(oracle, na=na)
    chain = cs.BayesNonparAlleleCluster(Xc, Xa, cna_states_WGS=true[1], niter=30, seed=200, rng="R")
    tr = cs.MCMCtrim(chain, burnin=10, allele=True)
    got = cs.AssignCluster(tr, mincell=mincell, allele=True, rm_extreme=rm_extreme)
    kt = tr["results"]["Uall"][-1].shape[0]
    want = oracle.assign_cluster_allele(tr["results"]["Zall"], kt, Xc, Xa, true[1], mincell=mincell,
                                        rm_extreme=rm_extreme)
    assert np.array_equal(got["Zmaj"], want["Zmaj"])
    assert np.array_equal(got["Zest"], want["Zest"])
    if mincell == 81:                                                  # the 80-cell cluster is dissolved
        assert (got["Zest"] != got["Zmaj"]).sum() >= 80               # and its cells re-drawn
    else:
        assert np.array_equal(got["Zest"], got["Zmaj"]) and not got["sigmas_cov_est"].any()
    val = lambda m: m.values if hasattr(m, "values") else m            # noqa: E731
    assert np.array_equal(np.nan_to_num(val(got["Uest"]), nan=-1), np.nan_to_num(want["Uest"], nan=-1))
    np.testing.assert_allclose(val(got["Ucest"]), want["Ucest"], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(val(got["Uaest"]), want["Uaest"], rtol=1e-10, atol=1e-14, equal_nan=True)
    np.testing.assert_allclose(got["sigmas_cov_est"], want["sigmas_cov_est"], rtol=1e-10)
    np.testing.assert_allclose(got["sigmas_allele_est"], want["sigmas_allele_est"], rtol=1e-10)
    np.testing.assert_allclose(got["corrs"], want["corrs"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(got["annot"], want["annot"]) and np.nanmax(got["corrs"]) == 1.0


def test_assign_cluster_km_cutoff(cs, oracle):
    """cutoff='km' (R/AssignCluster.R:191-194): kmeans on R's stream after the orphans' draws.  The
    cutoff must be the midpoint of a 2-means partition of the per-cluster indices that no single
    transfer improves (Hartigan-Wong's stopping rule), and the labels must follow it."""
    g = load_golden("sampler_P5931.npz")
    R = g["X"].shape[1]
    obj = dict(results=dict(Zall=g["Zall"][100:].astype(np.int32), Uall=[np.zeros((int(g["kt"][-1]), R))]),
               priors=dict(U0=g["U0"], cna_states_WGS=g["cna_states_WGS"], wU0=True), data=g["X"])
    fixed = cs.AssignCluster(obj, mincell=13, cutoff=0.2)
    km = cs.AssignCluster(obj, mincell=13, cutoff="km")
    assert np.array_equal(km["Zest"], fixed["Zest"]) and np.array_equal(km["corrs"], fixed["corrs"])
    # per-cluster indices: zeros for the dissolved clusters included, as the reference passes them
    kt = int(g["kt"][-1])
    per = np.zeros(kt)
    per[km["Zest"] - 1] = km["corrs"]
    c = km["cutoff"]
    lo, hi = per[per <= c], per[per > c]
    assert len(lo) and len(hi) and abs(c - (lo.mean() + hi.mean()) / 2) < 1e-12
    def wss(a, b):
        return ((a - a.mean()) ** 2).sum() + ((b - b.mean()) ** 2).sum()
    base = wss(lo, hi)
    for i in range(len(lo)):
        if len(lo) > 1:
            assert wss(np.delete(lo, i), np.append(hi, lo[i])) >= base - 1e-12
    for i in range(len(hi)):
        if len(hi) > 1:
            assert wss(np.append(lo, hi[i]), np.delete(hi, i)) >= base - 1e-12
    assert np.array_equal(km["annot"], np.where(km["corrs"] > c, "T", "N"))


# ------------------------------------------------------------------ f3: CreateSegtableNoWGS + Segmentation_bulk
def _planted_profile(rng, bins, depth=40.0):
    """Pooled tumour / normal counts over the synthetic tiles with a few planted gains and losses."""
    B = len(bins["chr"])
    cn = np.ones(B)
    for chrom, lo, hi, v in ((1, 0.10, 0.35, 1.5), (3, 0.5, 1.0, 0.5), (7, 0.0, 1.0, 1.5), (8, 0.2, 0.4, 1.8),
                             (13, 0.3, 0.9, 0.5), (20, 0.0, 0.5, 1.5)):
        idx = np.nonzero(bins["chr"] == chrom)[0]
        cn[idx[int(lo * len(idx)):int(hi * len(idx))]] = v
    base = rng.gamma(4.0, depth / 4.0, size=B)
    base[rng.random(B) < 0.03] = 0.0                                    # dead bins: 0/0 and x/0
    ref = rng.poisson(base).astype(np.float64)
    raw = rng.poisson(base * cn).astype(np.float64)
    return raw, ref


@pytest.mark.gpu
def test_segment_hmm_vs_oracle(cs, oracle):
    """cs_segment_hmm (running mean + 4-state Viterbi, all sequences in one launch) against the
    restated caTools::runmean / HiddenMarkov::Viterbi: same states everywhere, smoothed values to 1e-12."""
    import segmentation as sg
    from clonalscope_b200.segtable import segment_hmm
    rng = np.random.default_rng(21)
    means, sds, delta = [1.8, 1.5, 1.0, 0.5], [0.2] * 4, [0.1, 0.2, 0.5, 0.2]
    t = 1e-6
    Pi = [[1 - 3 * t if a == b else t for b in range(4)] for a in range(4)]
    for nmean in (1, 4, 100, 500):
        seqs = []
        for n in (1, 2, 3, 9, 250, 1245, 640):
            lv = np.repeat(rng.choice(means, size=n // 40 + 1), 40)[:n]
            seqs.append(lv + rng.normal(0, 0.3, size=n))
        sm, st = segment_hmm(seqs, nmean, means, sds, delta, Pi)
        for x, a, b in zip(seqs, sm, st):
            want = sg.runmean(x, nmean)
            np.testing.assert_allclose(a, want, rtol=1e-12, atol=1e-14)
            assert np.array_equal(b + 1, sg.viterbi(want, means, sds, delta, Pi))
    # underflow is the reference's error, not a silent path
    with pytest.raises(cs.ClonalscopeError, match="Underflow"):
        segment_hmm([np.array([1.0, 1e200, 1.0])], 1, means, sds, delta, Pi)


@pytest.mark.gpu
def test_rowsum_by_label_exact(cs, oracle):
    import scipy.sparse as sp
    import segmentation as sg
    from clonalscope_b200.segtable import rowsum_by_label
    rng = np.random.default_rng(22)
    B, N, K = 700, 900, 5
    colptr, rowidx, x = _random_csc(rng, B, N, 0.2)
    label = rng.integers(-1, K + 1, size=N)                            # -1 and K: cells outside every pool
    got = rowsum_by_label(sp.csc_matrix((x, rowidx, colptr), shape=(B, N)), label, K)
    assert np.array_equal(got, sg.pool_by_label(B, colptr, rowidx, x, label, K))


@pytest.mark.gpu
@pytest.mark.parametrize("rm_extreme,assay,nmean", [(1, "WGS", 100), (2, "WGS", 500), (2, "scRNA", 30)])
def test_segmentation_bulk_vs_oracle(cs, oracle, rm_extreme, assay, nmean):
    """Segmentation_bulk through the host API against the line-by-line restatement: the same segments
    (chr / start / end / states / length strings), means and variances to 1e-12."""
    import segmentation as sg
    from clonalscope_b200 import synth
    from clonalscope_b200.api import NamedMatrix
    from clonalscope_b200.segtable import Createobj_bulk, Segmentation_bulk
    rng = np.random.default_rng(23 + rm_extreme)
    bins, _ = synth.geometry(G=10)
    raw, ref = _planted_profile(rng, bins)
    names = [f"chr{c}-{s + 1}-{e}" for c, s, e in zip(bins["chr"], bins["start"], bins["end"])]
    size = {f"chr{i + 1}": v for i, v in enumerate(synth.CHR_SIZES)}
    obj = Createobj_bulk(NamedMatrix(raw.reshape(-1, 1), names, ["V1"]), NamedMatrix(ref.reshape(-1, 1), names, ["V1"]),
                         size=size)
    out = Segmentation_bulk(obj, hmm_states=(0.5, 1.5, 1.8), nmean=nmean, plot_seg=False, max_qt=0.95,
                            rm_extreme=rm_extreme, adj=-0.5 if assay == "scRNA" else 0, assay=assay)
    want, annot = sg.segmentation_bulk(list(bins["chr"]), list(bins["start"] + 1), list(bins["end"]), raw, ref,
                                       {i + 1: float(v) for i, v in enumerate(synth.CHR_SIZES)}, nmean=nmean,
                                       max_qt=0.95, rm_extreme=rm_extreme, adj=-0.5 if assay == "scRNA" else 0.0,
                                       assay=assay)
    got = out["seg_table"]
    assert out["annot"] == annot
    assert len(got["chr"]) > 22                                        # the planted events were found
    for col in ("chr", "start", "end", "states", "length", "chrr"):
        assert list(got[col]) == list(want[col]), col
    for col in ("mean", "var"):
        a = np.array([float(v) if v != "NA" else np.nan for v in got[col]])
        b = np.array([float(v) if v != "NA" else np.nan for v in want[col]])
        np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)


@pytest.mark.gpu
def test_create_segtable_nowgs_end_to_end(cs, oracle):
    """CreateSegtableNoWGS from the gene x cell counts: device binning -> pooling on the device-resident
    result -> HMM per cluster -> breakpoint union, against the restated chain on the same inputs."""
    import scipy.sparse as sp
    import segmentation as sg
    from clonalscope_b200 import synth
    from clonalscope_b200.segtable import CreateSegtableNoWGS
    G, N = 3000, 600
    rng = np.random.default_rng(29)
    bins, genes = synth.geometry(G=G, bin_size=2000000, seed=31)       # ~1450 tiles of 2 Mb: two genes per tile
    B = len(bins["chr"])
    colptr, rowidx, x = synth.counts_csc(N, G=G, seed=30, mean_genes=900)
    z = rng.integers(1, 4, size=N)                                      # clusters 1..3, cluster 1 = the normal cells
    # plant copy-number changes per cluster by scaling the counts of the genes of some chromosomes
    gchr = genes["chr"][rowidx]
    cell = np.repeat(np.arange(N), np.diff(colptr))
    scale = np.ones(len(x))
    gpos = genes["start"][rowidx]
    scale[(z[cell] == 2) & (gchr == 5) & (gpos > 90e6)] = 1.5
    scale[(z[cell] == 2) & (gchr == 9)] = 0.5
    scale[(z[cell] == 3) & (gchr == 5) & (gpos < 60e6)] = 1.5
    scale[(z[cell] == 3) & (gchr == 17)] = 1.8
    x = np.maximum(1.0, np.rint(x * 4 * scale))
    m = sp.csc_matrix((x, rowidx, colptr), shape=(G, N))
    ids = [f"ENSG{i:08d}" for i in range(G)]
    gtf = dict(V1=np.array([f"chr{c}" for c in genes["chr"]]), V3=np.array(["gene"] * G), V4=genes["start"],
               V5=genes["end"], V9=np.array([f'gene_id {i}; gene_name x' for i in ids]))
    features = np.array(ids, dtype=object).reshape(-1, 1)
    barcodes = np.array([f"c{i}" for i in range(N)], dtype=object).reshape(-1, 1)
    perm = rng.permutation(N)                                           # celltype0 lists the cells in another order
    celltype0 = np.array([[f"c{i}", "normal" if z[i] == 1 else "tumor"] for i in perm], dtype=object)
    bed = np.stack([np.array([f"chr{c}" for c in bins["chr"]], dtype=object), bins["start"].astype(object),
                    bins["end"].astype(object)], axis=1)
    size = np.array([[f"chr{i + 1}", v] for i, v in enumerate(synth.CHR_SIZES)], dtype=object)
    kw = dict(hmm_states=(0.5, 1.5, 1.8), max_qt=0.95, nmean=10, rm_extreme=2, adj=0, plot_seg=False)
    seg, obj_all, bbc = CreateSegtableNoWGS(mtx=m, barcodes=barcodes, features=features, Zest=z[perm], bin_bed=bed,
                                            celltype0=celltype0, gtf=gtf, size=size, merge=True, return_obj=True, **kw)
    # the restated chain
    mp, mb = synth.geometry_map(bins, genes)
    mm = m[:, perm].tocsc()
    mm.sort_indices()
    ocp, ori, ox = oracle.bin_counts(G, N, B, mm.indptr, mm.indices, mm.data, mp, mb)
    pooled = sg.pool_by_label(B, ocp, ori, ox, z[perm] - 1, 3)
    refsum = sg.pool_by_label(B, ocp, ori, ox, np.where(z[perm] == 1, 0, -1), 1)[:, 0]
    assert np.array_equal(bbc.values, pooled) and bbc.colnames == ["C1", "C2", "C3"]
    tabs = []
    for k in range(3):
        tab, _ = sg.segmentation_bulk(list(bins["chr"]), list(bins["start"] + 1), list(bins["end"]), pooled[:, k], refsum,
                                      {i + 1: float(v) for i, v in enumerate(synth.CHR_SIZES)}, nmean=10, max_qt=0.95,
                                      rm_extreme=2, adj=0.0)
        for col in ("chr", "start", "end", "states", "length"):
            assert list(obj_all[f"C{k + 1}"]["seg_table"][col]) == list(tab[col]), (k, col)
        tabs.append(tab)
    want = np.array(sg.merge_segments(tabs, merge=True))
    assert np.array_equal(seg.astype(str), want)
    assert len(want) > 22                                               # breakpoints from the planted events
    # a precomputed bin matrix gives the same answer (the reference's bin_mtx= argument)
    from clonalscope_b200.api import NamedMatrix
    bm = sp.csc_matrix((ox, ori, ocp), shape=(B, N))
    seg2 = CreateSegtableNoWGS(bin_mtx=NamedMatrix(bm, [""] * B, [""] * N), Zest=z[perm], bin_bed=bed, celltype0=celltype0,
                               size=size, merge=True, **kw)
    assert np.array_equal(seg2.astype(str), want)
    with pytest.raises(ValueError, match="three HMM numeric states"):   # the reference's own default fails its check
        CreateSegtableNoWGS(bin_mtx=bm, Zest=z[perm], bin_bed=bed, celltype0=celltype0, size=size)


# ------------------------------------------------------------------ f4: MatrixMarket ingest
def test_readmm_matches_scipy(cs, tmp_path):
    """cs_mm_read_*: cellranger-style files (column-major, ascending rows — the fast path), a shuffled
    file with duplicated coordinates (sort + segmented sum), pattern and real-valued files, .gz."""
    import gzip
    import scipy.io
    import scipy.sparse as sp
    from clonalscope_b200.rio import readMM
    rng = np.random.default_rng(41)
    G, N = 5000, 700
    m = sp.random(G, N, density=0.08, random_state=7, format="csc", data_rvs=lambda k: rng.integers(1, 50, k).astype(float))
    m.sort_indices()
    coo = m.tocoo()                                                    # column-major, rows ascending
    def write(path, rows, cols, vals, field="integer", opener=open):
        with opener(path, "wt") as f:
            f.write(f"%%MatrixMarket matrix coordinate {field} general\n%metadata_json: {{}}\n")
            f.write(f"{G} {N} {len(rows)}\n")
            if vals is None:
                f.write("".join(f"{r + 1} {c + 1}\n" for r, c in zip(rows, cols)))
            else:
                f.write("".join(f"{r + 1} {c + 1} {v}\n" for r, c, v in zip(rows, cols, vals)))
    ints = coo.data.astype(np.int64)
    write(tmp_path / "a.mtx", coo.row, coo.col, ints)
    got = readMM(tmp_path / "a.mtx")
    assert got.shape == (G, N) and (got != m).nnz == 0 and got.has_sorted_indices
    assert np.array_equal(got.indptr, m.indptr) and np.array_equal(got.indices, m.indices)
    assert (tmp_path / "a.mtx").stat().st_size > (1 << 20)              # several parser threads took part
    write(tmp_path / "a.mtx.gz", coo.row, coo.col, ints, opener=gzip.open)
    gz = readMM(tmp_path / "a.mtx.gz")
    assert np.array_equal(gz.indptr, m.indptr) and np.array_equal(gz.indices, m.indices) and np.array_equal(gz.data, m.data)
    # shuffled, with every 10th entry split in two: duplicates are summed, like readMM + dgCMatrix
    rows, cols = list(coo.row), list(coo.col)
    vals = list(ints)
    for e in range(0, len(vals), 10):
        if vals[e] > 1:
            rows.append(rows[e]), cols.append(cols[e]), vals.append(1)
            vals[e] -= 1
    perm = rng.permutation(len(vals))
    write(tmp_path / "b.mtx", np.array(rows)[perm], np.array(cols)[perm], np.array(vals)[perm])
    got = readMM(tmp_path / "b.mtx")
    assert np.array_equal(got.indptr, m.indptr) and np.array_equal(got.indices, m.indices) and np.array_equal(got.data, m.data)
    # pattern and real fields against scipy's reader
    write(tmp_path / "p.mtx", coo.row[:500], coo.col[:500], None, field="pattern")
    want = sp.csc_matrix(scipy.io.mmread(str(tmp_path / "p.mtx")))
    got = readMM(tmp_path / "p.mtx")
    assert (got != want).nnz == 0
    real = rng.normal(size=300)
    write(tmp_path / "r.mtx", coo.row[:300], coo.col[:300], [repr(float(v)) for v in real], field="real")
    got = readMM(tmp_path / "r.mtx")
    want = sp.csc_matrix(scipy.io.mmread(str(tmp_path / "r.mtx")))
    assert abs(got - want).max() == 0.0
    # loud failures
    with open(tmp_path / "bad.mtx", "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n3 3 2\n1 1 5\n4 1 2\n")
    with pytest.raises(cs.ClonalscopeError, match="outside"):
        readMM(tmp_path / "bad.mtx")
    with open(tmp_path / "short.mtx", "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n3 3 5\n1 1 5\n")
    with pytest.raises(cs.ClonalscopeError, match="announces"):
        readMM(tmp_path / "short.mtx")
    with pytest.raises(cs.ClonalscopeError, match="cannot open"):
        readMM(tmp_path / "missing.mtx")
    # empty matrix
    with open(tmp_path / "empty.mtx", "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n4 3 0\n")
    e = readMM(tmp_path / "empty.mtx")
    assert e.shape == (4, 3) and e.nnz == 0
