"""CPU tests of the host-side logic and the C-ABI surface (no kernels are launched)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_builds_loads_and_exports_every_declared_symbol(cs):
    from clonalscope_b200 import _lib
    L = cs.lib()
    hdr = open(os.path.join(ROOT, "include", "clonalscope_b200.h")).read()
    declared = set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.cs_version() >= 100
    assert isinstance(L.cs_last_error(), bytes)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "clonalscope_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in ("import oracle", "from oracle", "libcs_oracle", "cs_oracle.h", "cso_", "oracle/"):
                    assert pat not in src, (pat, os.path.join(dirpath, f))


def test_r_number_formatting():
    from clonalscope_b200.intervals import r_as_character as f
    assert f(600000.0) == "6e+05" and f(200000.0) == "2e+05" and f(100000.0) == "1e+05"
    assert f(1200000.0) == "1200000" and f(123456.0) == "123456"
    assert f(600000) == "600000" and f("17") == "17"
    assert f(0.5) == "0.5" and f(1e-5) == "1e-05" and f(0.0001) == "1e-04" and f(0.1234) == "0.1234"
    assert f(1e15) == "1e+15" and f(248956422.0) == "248956422"


def test_sort_bins_matches_r_order():
    from clonalscope_b200.intervals import sort_bins
    chrom = ["chr2", "chr10", "chrX", "chr1", "chr1", "chrY", "chr2"]
    start = [200, 0, 5, 400, 200, 1, 0]
    end = [400, 200, 9, 600, 400, 2, 200]
    c, s, e, order = sort_bins(chrom, start, end)
    assert c == ["1", "1", "2", "2", "10", "Y", "X"]  # NA chromosomes last, ordered by start
    assert s == [200, 400, 0, 200, 0, 1, 5]


def test_overlap_map_any_overlap_semantics():
    from clonalscope_b200.intervals import overlap_map
    bins_chr = ["1", "1", "1", "2"]
    bs, be = [0, 100, 200, 0], [100, 200, 300, 100]
    gene_seq = ["chr1", "chr1", "chr1", "chr2", "chr0", "chr1"]
    gs = np.array([50, 100, 101, 1, 0, 301])
    ge = np.array([60, 100, 250, 100, 0, 400])
    ptr, mb = overlap_map(bins_chr, bs, be, gene_seq, gs, ge)
    hits = [list(mb[ptr[g]:ptr[g + 1]]) for g in range(6)]
    # bin i covers [start+1, end]: gene [100,100] lies in bin 0 only; [101,250] spans bins 1 and 2
    assert hits == [[0], [0], [1, 2], [3], [], []]


def test_overlap_map_general_path_equals_tiles_path():
    from clonalscope_b200.intervals import overlap_map
    rng = np.random.default_rng(5)
    bs = np.sort(rng.integers(0, 10000, size=60))
    be = bs + rng.integers(1, 900, size=60)  # overlapping windows, ends not monotone
    gs = rng.integers(0, 10000, size=200)
    ge = gs + rng.integers(0, 1500, size=200)
    ptr, mb = overlap_map(["1"] * 60, list(bs), list(be), ["chr1"] * 200, gs, ge)
    for g in range(200):
        want = [b for b in range(60) if bs[b] + 1 <= ge[g] and be[b] >= gs[g]]
        assert list(mb[ptr[g]:ptr[g + 1]]) == want


def test_real_geometry_map():
    """14,385 autosomal + X/Y/... tiles x 36,601 genes of the reference's own data files."""
    from clonalscope_b200.api import gene_bin_map
    g = load_golden("geometry_hg38.npz")
    bed = np.stack([g["bin_chr"], g["bin_start"].astype(object), g["bin_end"].astype(object)], axis=1)
    gtf = dict(V1=g["gtf_V1"], V3=g["gtf_V3"], V4=g["gtf_V4"], V5=g["gtf_V5"], V9=g["gtf_V9"])
    ids = [s.split("gene_id ")[1].rstrip(";") for s in g["gtf_V9"]]
    genes = np.array(ids, dtype=object).reshape(-1, 1)
    rownames, ptr, mb = gene_bin_map(bed, genes, gtf)
    assert len(rownames) == 15641 and rownames[0] == "chr1-0-200000" and rownames[1] == "chr1-200000-400000"
    nh = np.diff(ptr)
    assert nh.min() >= 0 and nh.max() >= 10 and 1.0 < nh[nh > 0].mean() < 1.3
    # every hit is a true overlap and bins are ascending per gene
    from clonalscope_b200.intervals import sort_bins
    c, s, e, _ = sort_bins(list(g["bin_chr"]), list(g["bin_start"]), list(g["bin_end"]))
    for gi in np.random.default_rng(0).integers(0, len(ids), 300):
        hb = mb[ptr[gi]:ptr[gi + 1]]
        assert np.all(np.diff(hb) > 0)
        for b in hb:
            assert "chr" + c[b] == g["gtf_V1"][gi] and s[b] + 1 <= g["gtf_V5"][gi] and e[b] >= g["gtf_V4"][gi]


@pytest.mark.parametrize("name,cap,sampler", [("P5931", 2.0, "P5931"), ("BC_ductal2", 3.0, "BC_ductal2")])
def test_prepcovmatrix_reproduces_fixture_data(name, cap, sampler):
    """deltas_allseg.rds -> PrepCovMatrix(ngene_filter=150, 'intersect') -> pmin(., est_cap) -> the
    cells x segments the reference clustered == fixture `data`, difference 0.0."""
    from clonalscope_b200 import PrepCovMatrix
    d = load_golden(f"deltas_{name}.npz")
    bar = d["barcodes"]
    nseg = len(d["seg_names"])
    deltas = [(list(bar[d[f"idx_{s}"]]), d[f"val_{s}"]) for s in range(nseg)]
    seg = dict(chr=d["seg_chr"], start=d["seg_start"], end=d["seg_end"])
    out = PrepCovMatrix(dict(deltas_all=deltas, ngenes=d["ngenes"], seg_table_filtered=seg), ngene_filter=150,
                        prep_mode="intersect")
    df = out["df"]
    g = load_golden(f"sampler_{sampler}.npz")
    rows = {b: i for i, b in enumerate(df.rownames)}
    cols = {c: j for j, c in enumerate(df.colnames)}
    ri = [rows[b] for b in g["cell_names"]]
    ci = [cols[c] for c in g["seg_names"]]
    sub = np.minimum(df.values[np.ix_(ri, ci)], cap)
    assert sub.shape == g["X"].shape
    assert np.array_equal(sub, g["X"])


def test_prepcovmatrix_union_and_errors():
    from clonalscope_b200 import PrepCovMatrix
    deltas = [(["a", "b", "c"], np.array([1.0, 2.0, 3.0])), (["c", "a", "d"], np.array([4.0, 5.0, 6.0]))]
    seg = dict(chr=np.array(["1", "2"]), start=np.array([0.0, 600000.0]), end=np.array([100.0, 1200000.0]))
    obj = dict(deltas_all=deltas, ngenes=np.array([300, 300]), seg_table_filtered=seg)
    u = PrepCovMatrix(obj, ngene_filter=200, ncell_filter=0, prep_mode="union")["df"]
    assert u.rownames == ["a", "b", "c", "d"] and u.colnames == ["chr1-0-100", "chr2-6e+05-1200000"]
    assert np.isnan(u.values[1, 1]) and np.isnan(u.values[3, 0]) and u.values[2, 1] == 4.0
    i = PrepCovMatrix(obj, ngene_filter=200, ncell_filter=0, prep_mode="intersect")["df"]
    assert i.rownames == ["a", "c"] and np.array_equal(i.values, [[1.0, 5.0], [3.0, 4.0]])
    with pytest.raises(ValueError):
        PrepCovMatrix(obj, prep_mode="both")


def test_mcmctrim_slicing_and_quirk():
    from clonalscope_b200 import MCMCtrim
    niter, N = 200, 7
    Z = np.arange(niter * N).reshape(niter, N)
    obj = dict(results=dict(Zall=Z, Uall=list(range(niter)), sigma_all=list(range(niter)),
                            Likelihood=np.arange(niter, dtype=float)), priors={}, data=None)
    t = MCMCtrim(obj)
    assert t["trim"] == dict(burnin=100, thinning=1)
    assert np.array_equal(t["results"]["Zall"], Z[100:]) and t["results"]["Uall"] == list(range(100, 200))
    t3 = MCMCtrim(obj, burnin=50, thinning=3)
    assert np.array_equal(t3["results"]["Zall"], Z[50::3])  # 50 rows
    # reference quirk (R/MCMCtrim.R:33-36): the lists are indexed with the thinned nrow -> 17 entries
    assert t3["results"]["Uall"] == [50 + 3 * i for i in range(17)]
    assert len(t3["results"]["Likelihood"]) == 17
    with pytest.raises(ValueError):
        MCMCtrim(obj, burnin=200)
    short = dict(results=dict(Zall=Z[:30], Uall=list(range(30)), sigma_all=list(range(30)),
                              Likelihood=np.arange(30.0)), priors={}, data=None)
    assert MCMCtrim(short)["trim"]["burnin"] == 15
    # -c(1:0) is -c(1, 0): burnin = 0 still drops the first sweep (R/MCMCtrim.R:27-30)
    t0 = MCMCtrim(obj, burnin=0)
    assert np.array_equal(t0["results"]["Zall"], Z[1:]) and t0["results"]["Uall"] == list(range(1, 200))
    assert len(t0["results"]["Likelihood"]) == 199
    assert np.array_equal(MCMCtrim(obj, burnin=2.7)["results"]["Zall"], Z[2:])


def test_documented_refusals_raise_clonalscope_error(cs):
    """The host-side refusals raise ClonalscopeError with their message (none needs a device: they
    fire before any library call)."""
    from clonalscope_b200 import _lib
    e = cs.ClonalscopeError("only a message")
    assert e.status == _lib.CS_ERR_UNSUPPORTED and "only a message" in str(e)
    assert cs.ClonalscopeError(3, "m").status == 3
    obj = dict(results=dict(Zall=np.ones((4, 5), dtype=np.int32), Uall=[np.ones((1, 2))]), priors={},
               data=np.ones((5, 2)))
    with pytest.raises(cs.ClonalscopeError, match="rm_extreme"):
        cs.AssignCluster(obj, rm_extreme=True)
    import scipy.sparse as sp
    with pytest.raises(cs.ClonalscopeError, match="not in barcodes") as ei:
        cs.EstRegionCov(sp.csc_matrix(np.ones((2, 2))), ["a", "b"], ["g1", "g2"],
                        np.array([["1", 0, 10, "g1"], ["1", 20, 30, "g2"]], dtype=object),
                        np.array([["a", "tumor"], ["zz", "normal"]], dtype=object))
    assert ei.value.status == _lib.CS_ERR_ARG


def test_bayesnonparcluster_argument_errors(cs):
    X = np.ones((5, 3))
    with pytest.raises(ValueError):
        cs.BayesNonparCluster(X, U0=np.ones((2, 4)))
    with pytest.raises(ValueError):
        cs.BayesNonparCluster(X, cna_states_WGS=[1, 1, 1], sigmas0=[0.1, 0.2])
    with pytest.raises(ValueError):
        cs.BayesNonparCluster(X, cna_states_WGS=[1, 1, 1], Z0=[1, 2])


def test_no_gpu_means_loud_failure(cs):
    """Without a CUDA device the product path must fail, never fall back to the CPU."""
    L = cs.lib()
    if L.cs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cs.ClonalscopeError):
        cs.loglik_matrix(np.ones((4, 2)), np.ones((1, 2)), 0.25)
    with pytest.raises(cs.ClonalscopeError):
        cs.majority_vote(np.ones((3, 4), dtype=np.int32))


def test_host_r_stream_matches_the_oracle(oracle):
    """clonalscope_b200.rrng (the product's host copy of R's stream, used by AssignCluster) against the
    oracle's C restatement: uniforms after set.seed, and sample.int(n, 1, prob=) draws incl. ties."""
    import ctypes as C
    from clonalscope_b200.rrng import RStream, sample_int_prob_one
    for seed in (100, 42, 200):
        r = RStream(seed)
        assert np.array_equal(np.array([r.unif_rand() for _ in range(1300)]), oracle.mt_uniforms(seed, 1300))
    L = oracle.lib()
    gm = oracle.MT()
    L.cso_set_seed(C.byref(gm), C.c_uint32(100))
    r = RStream(100)
    rng = np.random.default_rng(0)
    for t in range(300):
        n = int(rng.integers(1, 40))
        P = np.exp(rng.normal(0, 3, size=n))
        if t % 5 == 0:
            P[rng.integers(0, n, size=3)] = P.min()
        p, perm = P.copy(), np.zeros(n, dtype=np.int32)
        a = L.cso_sample_int_prob(C.byref(gm), n, p.ctypes.data_as(C.POINTER(C.c_double)),
                                  perm.ctypes.data_as(C.POINTER(C.c_int)))
        assert a == sample_int_prob_one(r, P)


def test_r_glue_compiles_against_the_c_abi_and_shims_match_it(tmp_path):
    """integration/src/cs_glue.c is compiled (not linked) against a declarations-only copy of R's C
    API, so any drift against include/clonalscope_b200.h is a compile error; every .Call in the R
    shims must name a registered routine with the registered number of arguments."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    glue = os.path.join(root, "integration", "src", "cs_glue.c")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-c", glue, "-I", os.path.join(root, "integration", "rstub"),
                           "-I", os.path.join(root, "include"), "-o", str(tmp_path / "cs_glue.o")])
    src = open(glue).read()
    registered = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(cs_R_\w+)", \(DL_FUNC\)&\1, (\d+)\}', src)}
    assert len(registered) >= 8
    shims = os.path.join(root, "integration", "R")
    files = sorted(f for f in os.listdir(shims) if f.endswith(".R"))
    assert {"Gene_bin_cell_rna_filtered.R", "EstRegionCov.R", "BayesNonparCluster.R", "BayesNonparAlleleCluster.R",
            "AssignCluster.R"} <= set(files)
    used = set()
    for f in files:
        text = open(os.path.join(shims, f)).read()
        for m in re.finditer(r'\.Call\("(\w+)"', text):
            name = m.group(1)
            # count top-level commas up to the matching parenthesis
            i, depth, nargs = m.end(), 1, 0
            while depth:
                ch = text[i]
                depth += ch in "([{"
                depth -= ch in ")]}"
                nargs += depth == 1 and ch == ","
                i += 1
            nargs -= 1  # PACKAGE=
            assert name in registered, f"{f}: .Call to unregistered {name}"
            assert nargs == registered[name], f"{f}: {name} called with {nargs} arguments, registered with {registered[name]}"
            used.add(name)
    assert used == set(registered), f"registered but never called from a shim: {set(registered) - used}"


def test_merge_segment_tables_reproduces_shipped_seg_filtered():
    """The product's host-side breakpoint union (segtable.merge_segment_tables, no device work) on the
    reference's shipped per-cluster tables == its shipped seg_filtered_from_cluster.rds."""
    from clonalscope_b200.segtable import merge_segment_tables
    d = load_golden("segtable_P5847.npz")
    tabs = [{c: [str(v) for v in d[f"{n}_{c}"]] for c in ("chr", "start", "end", "states")} for n in d["clusters"]]
    out = merge_segment_tables(tabs, merge=True)
    assert np.array_equal(out.astype(str), d["seg_filtered"])


def test_segmentation_argument_errors():
    from clonalscope_b200.segtable import Segmentation_bulk, Createobj_bulk
    with pytest.raises(ValueError, match="valid Alleloscope object"):
        Segmentation_bulk(None)
    with pytest.raises(ValueError, match="two columns"):
        Createobj_bulk(raw_counts=None, ref_counts=None, size=None)
    from clonalscope_b200.api import NamedMatrix
    names = [f"chr1-{1 + 200000 * i}-{200000 * (i + 1)}" for i in range(5)]
    obj = Createobj_bulk(NamedMatrix(np.ones((5, 1)), names, ["V1"]), NamedMatrix(np.ones((5, 1)), names, ["V1"]),
                         size={"chr1": 1000000, "chr2": 5})
    assert obj["size"] == {"1": 1000000.0}
    with pytest.raises(ValueError, match="three HMM numeric states"):
        Segmentation_bulk(obj, hmm_states=(0.5, 1, 1.5, 2))


def test_saverds_known_bytes_and_round_trip(tmp_path):
    """saveRDS: the byte stream of R's serialize(c(1, 2), NULL, version = 2) is a known answer, and the
    objects of the path (named list, matrices with dimnames, named numeric vectors, data.frame, NA
    strings) read back through the RDS reader that is pinned by the reference's own .rds files."""
    import rds_reader as rr
    from clonalscope_b200.api import NamedMatrix
    from clonalscope_b200.rio import RDataFrame, saveRDS, serialize
    assert serialize(np.array([1.0, 2.0])) == bytes.fromhex(
        "580a" "00000002" "00030600" "00020300" "0000000e" "00000002" "3ff0000000000000" "4000000000000000")
    assert serialize(None)[-4:] == bytes.fromhex("000000fe")
    clustering = dict(
        data=NamedMatrix(np.array([[1.0, np.nan], [0.5, 2.0], [1.5, 1.0]]), ["c1", "c2", "c3"], ["chr1-0-5", "chr2-0-9"]),
        results=dict(Zall=np.array([[1, 2, 2], [1, 1, 2]], dtype=np.int32), Likelihood=np.array([-3.5, -2.25]),
                     Uall=[np.ones((2, 2)), np.full((3, 2), 0.5)]),
        priors=dict(alpha=2.0, seed=np.int32(200), U0=None, names=["a", None, "é"]),
        deltas_all={"pt0.99_tumor_alphaallchr1": (["c1", "c3"], np.array([1.1, 0.9]))},
        seg=RDataFrame(chr=["1", "2"], start=np.array([0.0, 6e5]), Var1=np.array([1, 2], dtype=np.int32)))
    fn = str(tmp_path / "nonpara_clustering.rds")
    saveRDS(clustering, fn)
    r = rr.read_rds(fn)
    assert r.names == ["data", "results", "priors", "deltas_all", "seg"]
    d = r["data"]
    assert list(d.attrs["dim"].data) == [3, 2] and list(d.attrs["dimnames"][1].data) == ["chr1-0-5", "chr2-0-9"]
    np.testing.assert_array_equal(d.data.reshape(2, 3).T, clustering["data"].values)
    assert np.array_equal(r["results"]["Zall"].data.reshape(3, 2).T, clustering["results"]["Zall"])
    assert r["results"]["Zall"].data.dtype.kind == "i" and len(r["results"]["Uall"]) == 2
    assert list(r["results"]["Uall"][1].attrs["dim"].data) == [3, 2]
    assert r["priors"]["U0"] is None and list(r["priors"]["names"].data) == ["a", None, "é"]
    assert float(r["priors"]["alpha"].data[0]) == 2.0 and int(r["priors"]["seed"].data[0]) == 200
    nv = r["deltas_all"]["pt0.99_tumor_alphaallchr1"]
    assert nv.names == ["c1", "c3"] and list(nv.data) == [1.1, 0.9]
    assert list(r["seg"].attrs["class"].data) == ["data.frame"] and list(r["seg"].attrs["row.names"].data) == [-2147483648, -2]
    assert list(r["seg"]["chr"].data) == ["1", "2"] and list(r["seg"]["start"].data) == [0.0, 6e5]
    saveRDS(np.arange(3.0), str(tmp_path / "plain.rds"), compress=False)
    assert list(rr.read_rds(str(tmp_path / "plain.rds")).data) == [0.0, 1.0, 2.0]


def test_createobj_bulk_orders_chromosomes_and_sizes():
    """Createobj_bulk (R/Createobj_bulk.R:57-90): count vectors whose chromosomes are not in numeric
    order are reordered by (chromosome, start); `size` is cut to the chromosomes present and ordered."""
    from clonalscope_b200.api import NamedMatrix
    from clonalscope_b200.segtable import Createobj_bulk
    names = ["chr2-1-100", "chr2-101-200", "chr10-1-100", "chr1-101-200", "chr1-1-100"]
    raw = NamedMatrix(np.array([[1.0], [2.0], [3.0], [4.0], [5.0]]), names, ["V1"])
    ref = NamedMatrix(np.ones((5, 1)), names, ["V1"])
    size = np.array([["chr10", 500], ["chr1", 300], ["chr2", 400], ["chrX", 900]], dtype=object)
    obj = Createobj_bulk(raw, ref, size=size)
    assert obj["raw_counts"].rownames == ["chr1-1-100", "chr1-101-200", "chr2-1-100", "chr2-101-200", "chr10-1-100"]
    assert list(obj["raw_counts"].values[:, 0]) == [5.0, 4.0, 1.0, 2.0, 3.0]
    assert list(obj["size"].items()) == [("1", 300.0), ("2", 400.0), ("10", 500.0)]
    with pytest.raises(ValueError, match="chr1-1-20000"):
        Createobj_bulk(NamedMatrix(np.ones((2, 1)), ["a", "b"], ["V1"]), ref, size=size)


def test_r_stream_sample_int_and_kmeans_known_answers():
    """sample.int without prob (R >= 3.6 rejection sampling) against base R's documented outputs, and
    Hartigan-Wong on a separable vector: set.seed(42); sample.int(10, 3) is 1 5 10; set.seed(1);
    sample.int(5, 2) is 1 4; set.seed(123); sample.int(10, 3) is 3 10 2."""
    from clonalscope_b200.rrng import RStream, kmeans_hartigan_wong, sample_int_noreplace
    assert sample_int_noreplace(RStream(42), 10, 3) == [1, 5, 10]
    assert sample_int_noreplace(RStream(1), 5, 2) == [1, 4]
    assert sample_int_noreplace(RStream(123), 10, 3) == [3, 10, 2]
    for seed in range(5):
        cen = sorted(kmeans_hartigan_wong(RStream(seed), [0.1, 0.15, 0.9, 1.0, 0.05, 0.8, 0.95]))
        assert abs(cen[0] - 0.1) < 1e-12 and abs(cen[1] - 0.9125) < 1e-12
    with pytest.raises(ValueError, match="between 1 and nrow"):
        kmeans_hartigan_wong(RStream(1), [0.3, 0.4], 2)
    with pytest.raises(ValueError, match="distinct data points"):
        kmeans_hartigan_wong(RStream(1), [0.3, 0.3, 0.3], 2)
