"""world_size-2 test of the cell-sharded statistics on two GPUs over NCCL: per-shard
cs_cluster_stats_dev / cs_sigma_stats_dev, ONE all-reduce of the sufficient statistics, the same
closed forms on every rank — against numpy on the whole matrix.  Skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from clonalscope_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(7)
    N, R, K = 20001, 300, 20
    X = rng.normal(1.0, 0.3, size=(N, R))
    X[rng.random((N, R)) < 0.03] = np.nan
    Z = rng.integers(1, K + 1, size=N).astype(np.int32)
    U = rng.uniform(0.5, 2.0, size=(K, R))
    lo, hi = D.shard_range(N, world, rank)
    dev = torch.device("cuda", rank)
    out = D.allreduce_stats(torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(Z[lo:hi]).to(dev),
                            torch.from_numpy(U).to(dev), K)
    ok = True
    for k in range(K):
        rows = X[Z == k + 1]
        cnt = (~np.isnan(rows)).sum(axis=0)
        want = np.where(cnt > 0, np.nansum(rows, axis=0) / np.maximum(cnt, 1), 0.0)
        ok = ok and int(out["n_k"][k].item()) == len(rows)
        ok = ok and np.allclose(out["mean"][k].cpu().numpy(), want, rtol=1e-11, atol=1e-12)
    d = X - U[Z - 1]
    sig = np.sqrt(np.nansum(d * d, axis=0) / (~np.isnan(X)).sum(axis=0))
    ok = ok and np.allclose(out["sigma"].cpu().numpy(), sig, rtol=1e-11)
    gathered = [torch.zeros_like(out["sigma"]) for _ in range(world)]
    dist.all_gather(gathered, out["sigma"])
    ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)     # every rank holds the same reduced result
    # ---- EstRegionCov sharded by cells: local sums + library sizes, one all-gather, medians over all
    # cells on every rank == the one-GPU computation on the whole matrix (bit for bit)
    import scipy.sparse as sp
    from clonalscope_b200.api import _CscHandle
    rng = np.random.default_rng(11)
    G, Nc, S = 400, 1501, 9
    m = sp.random(G, Nc, density=0.15, random_state=5, format="csc", data_rvs=lambda k: rng.integers(1, 9, k).astype(float))
    m.sort_indices()
    map_ptr = np.zeros(G + 1, dtype=np.int32)
    hits = [sorted(set(rng.integers(0, S, size=rng.integers(0, 3)).tolist())) for _ in range(G)]
    map_ptr[1:] = np.cumsum([len(h) for h in hits])
    map_seg = np.array([s_ for h in hits for s_ in h], dtype=np.int32)
    lib_flag = (rng.random(G) < 0.8).astype(np.uint8)
    code = rng.choice([0, 1, 2], size=Nc, p=[0.3, 0.6, 0.1]).astype(np.int32)
    whole = _CscHandle(m)
    want_d, want_s, want_a, want_n = whole.region_cov(S, map_ptr, map_seg, lib_flag, code, 0)
    whole.free()
    clo, chi = D.shard_range(Nc, world, rank)
    local = _CscHandle(m[:, clo:chi].tocsc())
    got = D.region_cov_sharded(local, S, map_ptr, map_seg, lib_flag, code[clo:chi], 0, device=dev)
    local.free()
    ok = ok and got["offset"] == clo and got["n_total"] == Nc
    ok = ok and np.array_equal(got["nsel"].cpu().numpy(), want_n)
    ok = ok and np.array_equal(got["alpha"].cpu().numpy(), want_a[clo:chi])
    ok = ok and np.array_equal(got["sums"].cpu().numpy(), want_s[:, clo:chi])
    ok = ok and np.array_equal(got["deltas"].cpu().numpy(), want_d[:, clo:chi], equal_nan=True)
    with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
        fh.write("1" if ok else "0")
    dist.destroy_process_group()


def test_allreduce_stats_on_two_gpus_over_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
