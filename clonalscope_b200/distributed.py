"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed
for the plumbing, a collective only where the path has a real exchange step.

  gene->bin aggregation   cells (CSC columns) sharded in contiguous ranges, no collective;
                          every output column depends on one input column
  cluster / sigma stats   cells sharded, U and sigma replicated; ONE all-reduce of the
                          per-cluster sums, non-NA counts, sizes and per-segment squared
                          residuals (K*R*2 + K + 2R doubles, ~100 KB: latency-bound on NVLink)
  Gibbs chains            the sweep is sequential by contract -> replicas only: one
                          independent chain per GPU (seed + rank), no communication; labels
                          of different chains are not comparable and are never pooled
  tally                   cells sharded, no collective
  EstRegionCov            cells sharded for the per-segment sums and library sizes; the two medians
                          per segment (R/EstRegionCov.R:141,153) need every cell: ONE all-gather of
                          the shards' (S + 1) x n doubles, then the medians and ratios on every rank
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of n items for `rank` (the first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_csc(colptr, rowidx, x, world: int, rank: int):
    """The CSC slots of this rank's column range, colptr rebased to 0."""
    n = len(colptr) - 1
    lo, hi = shard_range(n, world, rank)
    e0, e1 = int(colptr[lo]), int(colptr[hi])
    return (np.asarray(colptr[lo:hi + 1]) - e0), rowidx[e0:e1], x[e0:e1], (lo, hi)


def chain_seed(seed: int, rank: int) -> int:
    """Independent chains: chain g runs with seed + g."""
    return (int(seed) + int(rank)) & 0xFFFFFFFF


def device_stats(X, Z, U, K):
    """Per-shard sufficient statistics on the current CUDA device (cs_cluster_stats_dev /
    cs_sigma_stats_dev).  X: torch float64 [n, R] cell-major on the device, Z: int32 labels 1..K,
    U: [K, R].  Returns one flat float64 tensor (sum_x | cnt_x | n_k | sq | cnt)."""
    import torch
    from . import _lib

    L = _lib.lib()
    n, R = X.shape
    buf = torch.zeros(2 * K * R + K + 2 * R, dtype=torch.float64, device=X.device)
    p = buf.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rc = L.cs_cluster_stats_dev(n, R, K, vp(X), vp(Z), C.c_void_p(p), C.c_void_p(p + 8 * K * R),
                                C.c_void_p(p + 16 * K * R), st)
    if rc == 0:
        rc = L.cs_sigma_stats_dev(n, R, K, vp(X), vp(Z), vp(U), C.c_void_p(p + 8 * (2 * K * R + K)),
                                  C.c_void_p(p + 8 * (2 * K * R + K + R)), st)
    if rc != 0:
        raise RuntimeError(L.cs_last_error().decode())
    return buf


def allreduce_stats(X, Z, U, K, stats_fn=device_stats, group=None):
    """Cluster means, sizes and per-segment sigmas of cell-sharded data
    (R/BayesNonparCluster.R:221-234 at fixed labels): local statistics, one all-reduce, then
    the same closed forms on every rank."""
    import torch
    import torch.distributed as dist

    R = X.shape[1]
    buf = stats_fn(X, Z, U, K)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    sum_x = buf[:K * R].reshape(K, R)
    cnt_x = buf[K * R:2 * K * R].reshape(K, R)
    n_k = buf[2 * K * R:2 * K * R + K]
    sq = buf[2 * K * R + K:2 * K * R + K + R]
    cnt = buf[2 * K * R + K + R:]
    mean = torch.where(cnt_x > 0, sum_x / torch.clamp(cnt_x, min=1), torch.zeros_like(sum_x))
    sigma = torch.sqrt((1.0 / cnt) * sq)
    return dict(mean=mean, n_k=n_k, sigma=sigma, sum_x=sum_x, cnt_x=cnt_x)


def region_cov_sharded(handle, S, map_ptr, map_seg, lib_flag, celltype, include, device=None, group=None):
    """EstRegionCov's per-segment computation (R/EstRegionCov.R:88-159) on cell-sharded data.

    `handle`: this rank's `api._CscHandle` (its cells, uploaded once); `celltype`: int32 codes of this
    rank's cells (0 normal, 1 tumor, other neither); map / flags as for cs_csc_region_cov.  Local
    sums and library sizes (cs_csc_region_sums_dev) -> one all-gather over the ranks (cells in rank
    order) -> medians and ratios over ALL cells on every rank (cs_region_ratios_dev).  Returns torch
    tensors on the device: deltas and sums of THIS rank's cells (S x n_local), alpha (n_local), nsel
    (S, counted over all cells), and the cell offset of the shard."""
    import torch
    import torch.distributed as dist
    from . import _lib

    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    n = handle.N
    map_ptr = np.ascontiguousarray(map_ptr, dtype=np.int32)
    map_seg = np.ascontiguousarray(map_seg, dtype=np.int32)
    lib_flag = np.ascontiguousarray(lib_flag, dtype=np.uint8)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    hp = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sums = torch.empty((S, n), dtype=torch.float64, device=dev)
    alpha = torch.empty(n, dtype=torch.float64, device=dev)
    rc = L.cs_csc_region_sums_dev(handle.h, int(S), hp(map_ptr, C.c_int32), hp(map_seg, C.c_int32),
                                  hp(lib_flag, C.c_ubyte), vp(sums), vp(alpha), st)
    if rc != 0:
        raise RuntimeError(L.cs_last_error().decode())
    ct = torch.as_tensor(np.ascontiguousarray(celltype, dtype=np.int32), device=dev)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    offset = 0
    if world > 1:
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
        counts = [int(c.item()) for c in counts]
        nmax, offset = max(counts), sum(counts[:rank])
        # cell-major rows of (S sums | alpha | celltype), padded to the largest shard
        pack = torch.zeros((nmax, S + 2), dtype=torch.float64, device=dev)
        pack[:n, :S] = sums.t()
        pack[:n, S] = alpha
        pack[:n, S + 1] = ct.to(torch.float64)
        gathered = torch.empty((world * nmax, S + 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, pack, group=group)
        keep = torch.cat([torch.arange(r * nmax, r * nmax + counts[r], device=dev) for r in range(world)])
        full = gathered[keep]
        sums_all = full[:, :S].t().contiguous()
        alpha_all = full[:, S].contiguous()
        ct_all = full[:, S + 1].to(torch.int32).contiguous()
    else:
        sums_all, alpha_all, ct_all = sums, alpha, ct
    N = sums_all.shape[1]
    deltas_all = torch.empty((S, N), dtype=torch.float64, device=dev)
    nsel = torch.zeros(S, dtype=torch.int32, device=dev)
    rc = L.cs_region_ratios_dev(int(S), int(N), vp(sums_all), vp(alpha_all), vp(ct_all), int(include), vp(deltas_all),
                                vp(nsel), st)
    if rc != 0:
        raise RuntimeError(L.cs_last_error().decode())
    torch.cuda.synchronize()
    return dict(deltas=deltas_all[:, offset:offset + n].contiguous(), sums=sums, alpha=alpha, nsel=nsel, offset=offset,
                n_total=N)
