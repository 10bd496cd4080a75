"""Generate tests/golden/*.npz from the reference's shipped RDS objects.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so every parity test reads these small
extracts instead.  Nothing here is computed: the arrays are copied verbatim out
of the reference's own outputs (SURVEY.md §4 lists what each file pins).

Per sampler fixture (nonpara_clustering*.rds = full return value of
BayesNonparCluster incl. its inputs):
    X, Z0, U0, sigmas0, alpha, beta, seed      inputs (`data`, `priors`)
    Zall (int16), sigma_all, Likelihood, L0    outputs, all 200 sweeps
    kt                                         nrow(Uall[[t]])
    u_off, u_labels, u_rowsum                  non-NA rows of every Uall[[t]]: labels + row sums
    usel, usel_off, usel_rows                  full non-NA rows for a few selected sweeps
    cell_names, seg_names                      dimnames(data)
Per coverage fixture (deltas_allseg.rds = EstRegionCov output = PrepCovMatrix input):
    barcodes, seg_names, idx_<s>, val_<s>, ngenes, seg_chr/start/end
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from rds_reader import read_rds, as_matrix  # noqa: E402

REF = "/root/reference"
SAMPLER = {
    "P5931": "samples/P5931/scRNA/output/nonpara_clustering.rds",
    "P5931_updated1": "samples/P5931/scRNA/output/nonpara_clustering_updated1.rds",
    "BC_ductal2": "samples/BC_ductal2/ST/output/nonpara_clustering.rds",
    "P5847_tumor": "samples/P5847/scRNA/noWGS_output/nonpara_clustering.rds",
    "P5847_allcells": "samples/P5847/scRNA/noWGS_output/second_round_allcells/nonpara_clustering.rds",
}
DELTAS = {
    "P5931": "samples/P5931/scRNA/output/deltas_allseg.rds",
    "BC_ductal2": "samples/BC_ductal2/ST/output/deltas_allseg.rds",
}
USEL = [0, 1, 2, 3, 4, 9, 49, 99, 149, 198, 199]


def sampler_fixture(name, path):
    o = read_rds(os.path.join(REF, path))
    X, cells, segs = as_matrix(o["data"])
    pr, rs = o["priors"], o["results"]
    U0, _, _ = as_matrix(pr["U0"])
    Zall, _, _ = as_matrix(rs["Zall"])
    niter = Zall.shape[0]
    assert Zall.max() < 32767
    kt, u_off, u_labels, u_rowsum = [], [0], [], []
    usel_off, usel_rows = [0], []
    for t in range(niter):
        U, _, _ = as_matrix(rs["Uall"][t])
        live = np.where(~np.isnan(U[:, 0]))[0]
        kt.append(U.shape[0])
        u_labels.append(live + 1)
        u_rowsum.append(U[live].sum(axis=1))
        u_off.append(u_off[-1] + len(live))
        if t in USEL:
            usel_rows.append(U[live])
            usel_off.append(usel_off[-1] + len(live))
    out = dict(
        X=X, Z0=pr["Z0"].data.astype(np.int32), U0=U0, sigmas0=pr["sigmas0"].data,
        alpha=pr["alpha"].data[0], beta=pr["beta"].data[0], seed=int(pr["seed"].data[0]),
        L0=pr["L0"].data[0], Zall=Zall.astype(np.int16),
        sigma_all=np.array([s.data for s in rs["sigma_all"].items]),
        Likelihood=rs["Likelihood"].data, kt=np.array(kt, dtype=np.int32),
        u_off=np.array(u_off, dtype=np.int64), u_labels=np.concatenate(u_labels).astype(np.int32),
        u_rowsum=np.concatenate(u_rowsum), usel=np.array(USEL, dtype=np.int32),
        usel_off=np.array(usel_off, dtype=np.int64), usel_rows=np.concatenate(usel_rows, axis=0),
        cell_names=np.array(cells), seg_names=np.array(segs),
        cna_states_WGS=pr["cna_states_WGS"].data,
    )
    fn = os.path.join(HERE, f"sampler_{name}.npz")
    np.savez_compressed(fn, **out)
    print(name, X.shape, "kt_final", kt[-1], os.path.getsize(fn) // 1024, "KiB")


def deltas_fixture(name, path):
    d = read_rds(os.path.join(REF, path))
    deltas = d["deltas_all"]
    universe = {}
    out = {}
    for s, v in enumerate(deltas.items):
        idx = np.array([universe.setdefault(b, len(universe)) for b in v.names], dtype=np.int32)
        out[f"idx_{s}"] = idx
        out[f"val_{s}"] = v.data
    st = d["seg_table_filtered"]
    out["barcodes"] = np.array(list(universe.keys()))
    out["seg_names"] = np.array(deltas.names)
    out["ngenes"] = d["ngenes"].data.astype(np.int32)
    for col in ("chr", "start", "end"):
        out[f"seg_{col}"] = np.array([str(x) for x in st[col].data]) if st[col].data.dtype == object \
            else st[col].data
    out["seg_cols_are_character"] = np.array(st["start"].data.dtype == object)
    fn = os.path.join(HERE, f"deltas_{name}.npz")
    np.savez_compressed(fn, **out)
    print(name, len(deltas), "segments", len(universe), "barcodes", os.path.getsize(fn) // 1024, "KiB")


def geometry_fixture():
    """Real geometry for the binning tests: the 200-kb tiles and the GTF-like gene table the
    reference ships (data-raw/hg38_200kb.windows.bed, data-raw/genes_filtered.rds)."""
    chrom, start, end = [], [], []
    with open(os.path.join(REF, "data-raw/hg38_200kb.windows.bed")) as fh:
        for line in fh:
            p = line.split()
            if len(p) >= 3:
                chrom.append(p[0]); start.append(int(p[1])); end.append(int(p[2]))
    g = read_rds(os.path.join(REF, "data-raw/genes_filtered.rds"))
    fn = os.path.join(HERE, "geometry_hg38.npz")
    np.savez_compressed(
        fn, bin_chr=np.array(chrom), bin_start=np.array(start, dtype=np.int64),
        bin_end=np.array(end, dtype=np.int64),
        gtf_V1=np.array([str(x) for x in g["V1"].data]), gtf_V3=np.array([str(x) for x in g["V3"].data]),
        gtf_V4=g["V4"].data.astype(np.int64), gtf_V5=g["V5"].data.astype(np.int64),
        gtf_V9=np.array([str(x).split(";")[0] + ";" for x in g["V9"].data]))
    print("geometry", len(chrom), "bins", len(g["V1"].data), "genes", os.path.getsize(fn) // 1024, "KiB")


def segtable_fixture():
    """CreateSegtableNoWGS on P5847: the per-cluster seg tables (Obj_all_cluster.rds, the input of the
    breakpoint union) and the merged table the reference saved (seg_filtered_from_cluster.rds)."""
    base = "samples/P5847/scRNA/noWGS_output"
    obj = read_rds(os.path.join(REF, base, "Obj_all_cluster.rds"))
    out = {"clusters": np.array(obj.names)}
    for name in obj.names:
        st = obj[name]["seg_table"]
        for col in ("chr", "start", "end", "states", "length", "mean", "var", "chrr"):
            out[f"{name}_{col}"] = np.array([str(v) for v in st[col].data])
        # the table Segmentation_bulk saved on its own for the same cluster must be the same object
        alone = read_rds(os.path.join(REF, base, "rds", f"seg_table_all_{name}_qt95_rm2_nmean500.rds"))
        for col in ("chr", "start", "end", "states"):
            assert [str(v) for v in alone[col].data] == list(out[f"{name}_{col}"]), (name, col)
    merged = read_rds(os.path.join(REF, base, "seg_filtered_from_cluster.rds"))
    nr, nc = (int(v) for v in merged.attrs["dim"].data)
    out["seg_filtered"] = np.array([str(v) for v in merged.data]).reshape(nc, nr).T  # column-major
    fn = os.path.join(HERE, "segtable_P5847.npz")
    np.savez_compressed(fn, **out)
    print("segtable", list(obj.names), out["seg_filtered"].shape, os.path.getsize(fn) // 1024, "KiB")


if __name__ == "__main__":
    geometry_fixture()
    segtable_fixture()
    for k, v in SAMPLER.items():
        sampler_fixture(k, v)
    for k, v in DELTAS.items():
        deltas_fixture(k, v)
