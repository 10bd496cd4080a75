"""world_size-2 tests of the multi-GPU host logic on CPU (gloo).  The per-shard compute is
injected from the oracle here — the product's stats_fn launches CUDA kernels and is covered
by the gpu-marked tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _numpy_stats(X, Z, U, K):
    """Reference per-shard statistics (plain numpy), same flat layout as device_stats."""
    Xn, Zn, Un = X.numpy(), Z.numpy(), U.numpy()
    n, R = Xn.shape
    sum_x, cnt_x, nk = np.zeros((K, R)), np.zeros((K, R)), np.zeros(K)
    for k in range(K):
        rows = Xn[Zn == k + 1]
        nk[k] = len(rows)
        sum_x[k] = np.nansum(rows, axis=0)
        cnt_x[k] = (~np.isnan(rows)).sum(axis=0)
    d = Xn - Un[Zn - 1]
    sq = np.nansum(d * d, axis=0)
    cnt = (~np.isnan(Xn)).sum(axis=0).astype(float)
    return torch.from_numpy(np.concatenate([sum_x.ravel(), cnt_x.ravel(), nk, sq, cnt]))


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    from clonalscope_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)  # same data on every rank, each takes its shard
    N, R, K = 1001, 13, 5
    X = rng.normal(1.0, 0.3, size=(N, R))
    X[rng.random((N, R)) < 0.03] = np.nan
    Z = rng.integers(1, K + 1, size=N).astype(np.int32)
    U = rng.uniform(0.5, 2.0, size=(K, R))
    lo, hi = D.shard_range(N, world, rank)
    out = D.allreduce_stats(torch.from_numpy(X[lo:hi]), torch.from_numpy(Z[lo:hi]), torch.from_numpy(U), K,
                            stats_fn=_numpy_stats)
    full = D.allreduce_stats.__wrapped__ if hasattr(D.allreduce_stats, "__wrapped__") else None
    ref = _numpy_stats(torch.from_numpy(X), torch.from_numpy(Z), torch.from_numpy(U), K).numpy()
    mean_ref = ref[:K * R].reshape(K, R) / np.maximum(ref[K * R:2 * K * R].reshape(K, R), 1)
    sig_ref = np.sqrt(ref[2 * K * R + K:2 * K * R + K + R] / ref[2 * K * R + K + R:])
    ok = (np.allclose(out["mean"].numpy(), mean_ref, rtol=1e-12) and np.allclose(out["sigma"].numpy(), sig_ref, rtol=1e-12)
          and np.array_equal(out["n_k"].numpy(), ref[2 * K * R:2 * K * R + K]))
    # every rank holds the same reduced result
    gathered = [torch.zeros_like(out["sigma"]) for _ in range(world)]
    dist.all_gather(gathered, out["sigma"])
    ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
    # binning shards: concatenating the ranks' column ranges restores the matrix
    colptr = np.concatenate([[0], np.cumsum(rng.integers(0, 9, size=57))]).astype(np.int64)
    rowidx = rng.integers(0, 40, size=int(colptr[-1])).astype(np.int32)
    xs = rng.integers(1, 5, size=int(colptr[-1])).astype(np.float64)
    cp, ri, xx, (clo, chi) = D.shard_csc(colptr, rowidx, xs, world, rank)
    ok = ok and cp[0] == 0 and len(cp) == chi - clo + 1 and np.array_equal(xx, xs[colptr[clo]:colptr[chi]])
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(ri)], dtype=torch.int64))
    ok = ok and sum(int(s) for s in sizes) == len(rowidx)
    ok = ok and D.chain_seed(200, rank) == 200 + rank
    with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
        fh.write("1" if ok else "0")
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    import sys
    sys.path.insert(0, ROOT)
    from clonalscope_b200.distributed import shard_range
    for n in (0, 1, 7, 100000, 1000003):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_allreduce_and_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
