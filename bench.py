#!/usr/bin/env python
"""bench.py — headline benchmark of the cell-clustering hot path (BASELINE.json).

Workload (BASELINE.json configs[3]): synthetic 100k cells x 30k genes -> 15k 200-kb bins
-> 300 segments, 1 GPU.  A "step" is one complete BayesNonparCluster chain (niter = 200
sweeps over 100,000 cells x 300 segments, Philox mode, sequential within-sweep order kept);
the metric is Gibbs cell-updates/s = N * niter / time.

  value      device-resident: inputs already in HBM, CUDA events on the chain's stream
  e2e        through the C ABI with HOST buffers (cs_ncrp_gibbs): H2D of X, the chain,
             D2H of Zall / Uall / sigma / Likelihood inside the timed region
  roofline   dominant kernel of the step (prepare_kernel: every cell against every live
             cluster + proposal + speculative draw); algorithmic bytes = N*S*8 per launch;
             fix_walk_kernel (the sweep's fixed-point iteration) nested under it
  ranks      per-rank ms_per_step / events (independent chains: seed + rank)
  seeds      the same chain under seeds 200..207 (min / median / max): how much the
             throughput depends on the trajectory
  fixtures   Philox-mode and R-compatible chains on the reference's real data sets
  kernels    log-likelihood matrix (K = 2, 20), posterior tally (T = 100), allele sampler
  binning    gene->bin aggregation GB/s against the measured HBM peak (+ end to end)
  config5    BASELINE.json configs[4]: --spots cells sharded over the ranks: per-rank
             binning + log-likelihood matrix + sufficient statistics, the NCCL all-reduce of
             the statistics (timed on its own), sharded tally — strong scaling
  cpu_baseline  the CPU oracle (validated restatement of the reference's R algorithm) on a
             bounded sample of the same workload, 1 core

`--impl reference` times the reference's own algorithm on the host cores (oracle port in
R-compatible mode, one independent chain per core) on the same workload/metric.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nCRP Gibbs cell-updates/sec at 100k cells x 300 segs; bin-count GB/s vs HBM"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# captures of sweep 10 of this workload (profiles/r2_prepare_full.txt, profiles/r2_fix_walk_full.txt)
PREPARE_TRAFFIC_BYTES = 243742208 + 36662016   # 1.17x the 240 MB of X: the Q / E / J write-back
WALK_TRAFFIC_BYTES = 606408960 + 60434176      # sweep 10 is event-dense; a stationary sweep touches far less
KERNELS_PER_SWEEP = 10  # prepare, fix_walk, 8 end-of-sweep kernels


def workload_config(N, R, niter, NB, G, B):
    return dict(workload=f"synthetic {N} cells x {R} segments (K_true=12, alpha=beta=2, sigmas0=0.25), "
                         f"niter={niter}; one independent chain per GPU",
                step="one full chain of niter sweeps (reference arm: a bounded sample of it, see cpu_baseline.sample)",
                l2="per-sweep working set (X 240 MB + Q + J) exceeds the 126 MB L2",
                binning=f"{NB} cells x {G} genes -> {B} bins per GPU")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------ CPU arms
def _cpu_chain(args):
    """one R-compatible chain of `niter` sweeps on the CPU oracle (a worker process)."""
    N, R, niter, seed = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from clonalscope_b200 import synth
    s = synth.sampler_input(N, R=R)
    Z0 = O.loglik_matrix(s["X"], s["U0"], s["sigmas0"]).argmax(axis=1).astype(np.int32) + 1
    t = time.perf_counter()
    O.gibbs(s["X"], s["U0"], Z0, s["sigmas0"], 2.0, 2.0, niter, seed, rng_mode=O.RNG_R, ucap_rows=niter * 512)
    return time.perf_counter() - t


def cpu_gibbs_rate(N, R, niter, procs):
    """cell-updates/s of the CPU port: `procs` independent chains, one per core."""
    import multiprocessing as mp
    if procs == 1:
        dt = _cpu_chain((N, R, niter, 200))
        return N * niter / dt, dt
    with mp.get_context("spawn").Pool(procs) as pool:
        t = time.perf_counter()
        pool.map(_cpu_chain, [(N, R, niter, 200 + i) for i in range(procs)])
        dt = time.perf_counter() - t
    return procs * N * niter / dt, dt


def run_reference(a):
    """--impl reference: the reference's own (single-threaded R) algorithm, restated in C and
    validated bit-for-bit against its shipped outputs, one independent chain per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    N, R, sweeps = a.cells, a.segs, 1
    for _ in range(a.warmup):
        cpu_gibbs_rate(max(1000, N // 50), R, 1, cores)
    tot_t, tot_u = 0.0, 0.0
    for _ in range(a.steps):
        rate, dt = cpu_gibbs_rate(N, R, sweeps, cores)
        tot_t += dt
        tot_u += rate * dt
    value = tot_u / tot_t
    sample = (f"{cores} independent chains x {sweeps} sweep(s) of the {N} x {R} workload per step "
              f"(R-compatible RNG, every sample.int heap-sorted as in base R; the cost of a cell-update does not "
              f"depend on the sweep, so one sweep per chain stands for the niter={a.niter} of the workload)")
    from clonalscope_b200 import synth
    bins, genes = synth.geometry(G=a.genes)
    B = len(bins["chr"])
    line = dict(metric=METRIC, value=value, unit="cell-updates/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1e3 * tot_t / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=workload_config(N, R, a.niter, a.bin_cells, a.genes, B),
                cpu_baseline=dict(value=value, unit="cell-updates/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="cell-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


# ------------------------------------------------------------------------ GPU arm
def synth_counts_gpu(torch, dev, N, G, seed=1001, mean_genes=2000):
    """gene x cell counts on the device (CSC slots), SURVEY §8d shape: per-gene weights
    LogNormal(0,1.5), per-cell number of expressed genes ~ LogNormal(ln 2000, 0.4) clipped to
    [200, 8000], Poisson-sampled inclusion proportional to the weights, values Geometric(0.6)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    w = torch.exp(1.5 * torch.randn(G, generator=g, device=dev, dtype=torch.float32))
    m = torch.exp(np.log(mean_genes) + 0.4 * torch.randn(N, generator=g, device=dev)).round().clamp(200, 8000)
    rows, counts = [], []
    chunk = 4096
    for c0 in range(0, N, chunk):
        mc = m[c0:c0 + chunk]
        s = mc / w.sum()
        for _ in range(12):  # scale so that sum_g min(1, s*w_g) = m_c
            p = (s[:, None] * w[None, :]).clamp(max=1.0)
            s = s * mc / p.sum(dim=1)
        p = (s[:, None] * w[None, :]).clamp(max=1.0)
        hit = torch.rand(p.shape, generator=g, device=dev) < p
        counts.append(hit.sum(dim=1))
        rows.append(hit.nonzero()[:, 1].to(torch.int32))
    counts = torch.cat(counts)
    colptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    colptr[1:] = torch.cumsum(counts, 0)
    rowidx = torch.cat(rows)
    u = torch.rand(rowidx.numel(), generator=g, device=dev, dtype=torch.float32)
    x = (torch.log(u) / np.log(0.4)).floor().clamp(0, 32766).to(torch.float64) + 1.0
    return colptr, rowidx, x


def make_chain_runner(L, _lib, N, R, niter, X, U0, Z0, sig0, alpha, beta, rng_mode=None):
    """create a device chain and return (run(seed, flags) -> timing list, stats() -> dict, destroy)."""
    dp = lambda arr, t: arr.ctypes.data_as(C.POINTER(t))  # noqa: E731
    chain = C.c_void_p()
    mode = _lib.RNG_PHILOX if rng_mode is None else rng_mode
    rc = L.cs_chain_create(N, R, 0, niter, mode, C.byref(chain))
    assert rc == 0, L.cs_last_error()
    K0 = U0.shape[0]

    def run(seed, flags=0, sweeps=None):
        rc = L.cs_chain_init(chain, dp(X, C.c_double), K0, dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                             C.c_double(alpha), C.c_double(beta), C.c_uint32(seed), flags)
        assert rc == 0, L.cs_last_error()
        n = niter if sweeps is None else sweeps
        rc = L.cs_chain_run(chain, n, 1, 1)
        assert rc == 0, L.cs_last_error()
        ms = [C.c_double() for _ in range(5)]
        rc = L.cs_chain_timing(chain, *[C.byref(m) for m in ms])
        assert rc == 0, L.cs_last_error()
        return [m.value for m in ms]

    def stats():
        v = [C.c_int64() for _ in range(7)]
        L.cs_chain_stats(chain, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]))
        L.cs_chain_walk_stats(chain, C.byref(v[3]), C.byref(v[4]), C.byref(v[5]), C.byref(v[6]))
        return dict(events=int(v[1].value), cells_evaluated=int(v[0].value), walk_steps=int(v[3].value),
                    clusters_opened=int(v[5].value))

    return run, stats, (lambda: L.cs_chain_destroy(chain))


def leg_seeds(run, stats, N, niter, seeds):
    """the headline chain under other seeds: throughput must not hinge on one trajectory"""
    rows = []
    for sd in seeds:
        ms = run(sd)[0]
        st = stats()
        rows.append(dict(seed=sd, ms=ms, value=N * niter / (ms * 1e-3), events=st["events"], walk_steps=st["walk_steps"]))
    v = sorted(r["value"] for r in rows)
    return dict(unit="cell-updates/s", min=v[0], median=float(np.median(v)), max=v[-1], max_over_min=v[-1] / v[0],
                per_seed=rows)


def leg_fixtures(L, _lib, niter=200, r_mode_sweeps=20):
    """chains on the reference's real data (tests/golden/*.npz, extracted from the RDS files the
    reference ships): Philox mode all `niter` sweeps; R-compatible mode (bit-identical to the
    reference's own chain) on the first `r_mode_sweeps` sweeps"""
    out = {}
    for name in ("P5931", "BC_ductal2", "P5847_allcells"):
        path = os.path.join(ROOT, "tests", "golden", f"sampler_{name}.npz")
        if not os.path.exists(path):
            continue
        g = np.load(path)
        X = np.asfortranarray(g["X"])
        N, R = X.shape
        U0 = np.asfortranarray(g["U0"])
        Z0 = np.ascontiguousarray(g["Z0"], dtype=np.int32)
        sig0 = np.ascontiguousarray(g["sigmas0"])
        row = dict(cells=N, segments=R)
        run, stats, destroy = make_chain_runner(L, _lib, N, R, niter, X, U0, Z0, sig0, float(g["alpha"]), float(g["beta"]))
        run(200)
        ms = run(200)[0]
        st = stats()
        row["philox"] = dict(ms=ms, value=N * niter / (ms * 1e-3), unit="cell-updates/s", niter=niter,
                             moved_fraction=st["events"] / (N * niter), walk_steps=st["walk_steps"])
        destroy()
        run, stats, destroy = make_chain_runner(L, _lib, N, R, r_mode_sweeps, X, U0, Z0, sig0, float(g["alpha"]),
                                                float(g["beta"]), rng_mode=_lib.RNG_R)
        ms = run(int(g["seed"]))[0]
        row["r_compatible"] = dict(ms=ms, value=N * r_mode_sweeps / (ms * 1e-3), unit="cell-updates/s", niter=r_mode_sweeps,
                                   note="the reference's own stream and seed: the chain it ships, label for label")
        destroy()
        out[name] = row
    return out


def leg_kernels(torch, dev, L, hbm_peak, N, R, reps=20):
    """device-resident GB/s of the fixed-state log-likelihood matrix and the posterior tally"""
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    X = 0.3 + 2.2 * torch.rand(N, R, generator=g, device=dev, dtype=torch.float64)
    sig = torch.full((R,), 0.25, device=dev, dtype=torch.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for K in (2, 20):
        U = 0.3 + 2.2 * torch.rand(K, R, generator=g, device=dev, dtype=torch.float64)
        LL = torch.empty(N, K, device=dev, dtype=torch.float64)
        for _ in range(3):
            assert L.cs_loglik_matrix_dev(N, R, K, vp(X), vp(U), vp(sig), vp(LL), st) == 0, L.cs_last_error()
        e0.record()
        for _ in range(reps):
            L.cs_loglik_matrix_dev(N, R, K, vp(X), vp(U), vp(sig), vp(LL), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = N * R * 8 + N * K * 8
        out[f"loglik_matrix_K{K}"] = dict(ms=ms, gbs=by / (ms * 1e-3) / 1e9, frac_of_hbm_peak=by / (ms * 1e-3) / 1e9 / hbm_peak,
                                          algorithmic_bytes=by, gflops=5.0 * N * R * K / (ms * 1e-3) / 1e9,
                                          l2="X (N*R*8 bytes) exceeds the 126 MB L2" if N * R * 8 > 126e6 else "fits L2")
    T = 100
    zmaj = torch.empty(N, device=dev, dtype=torch.int32)
    by = T * N * 4 + N * 4
    # (a) a converged posterior: every cell sits in its home cluster, 5 % of its sweeps in a neighbour;
    # (b) the worst case: labels uniform over 20 clusters in every sweep (20 distinct labels per cell)
    home = torch.randint(1, 21, (N,), generator=g, device=dev, dtype=torch.int32)
    stray = torch.rand(T, N, generator=g, device=dev) < 0.05
    posterior = torch.where(stray, home[None, :] % 20 + 1, home[None, :]).contiguous()
    uniform = torch.randint(1, 21, (T, N), generator=g, device=dev, dtype=torch.int32)
    for name, Zall in (("tally_T100", posterior), ("tally_T100_uniform_labels", uniform)):
        for _ in range(3):
            assert L.cs_majority_vote_dev(T, N, vp(Zall), C.c_int64(N), C.c_int64(1), vp(zmaj), st) == 0, L.cs_last_error()
        e0.record()
        for _ in range(reps):
            L.cs_majority_vote_dev(T, N, vp(Zall), C.c_int64(N), C.c_int64(1), vp(zmaj), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = dict(ms=ms, gbs=by / (ms * 1e-3) / 1e9, frac_of_hbm_peak=by / (ms * 1e-3) / 1e9 / hbm_peak,
                         algorithmic_bytes=by, l2="40 MB: fits the 126 MB L2 (repeated launches re-read it from L2)")
    return out


def leg_allele(cs, niter=20):
    """the allele variant (coverage + BAF) on a synthetic case, both RNG modes"""
    rng = np.random.default_rng(0)
    N, R = 2000, 20
    G = np.array([[0.5 * ii, j / ii] for ii in range(1, 7) for j in range(ii + 1)])  # the maxcp = 6 genotype grid
    true = rng.integers(1, len(G) + 1, size=(3, R))
    true[0, :] = 4
    z = rng.integers(0, 3, N)
    Xc = G[true[z] - 1, 0] + 0.15 * rng.standard_normal((N, R))
    Xa = np.clip(G[true[z] - 1, 1] + 0.08 * rng.standard_normal((N, R)), 0, 1)
    out = dict(cells=N, segments=R, niter=niter)
    import contextlib, io
    for name in ("philox", "R"):
        with contextlib.redirect_stdout(io.StringIO()):
            cs.BayesNonparAlleleCluster(Xc, Xa, cna_states_WGS=true[1], niter=2, seed=200, rng=name)
            t0 = time.perf_counter()
            cs.BayesNonparAlleleCluster(Xc, Xa, cna_states_WGS=true[1], niter=niter, seed=200, rng=name)
            dt = time.perf_counter() - t0
        out[name] = dict(ms=dt * 1e3, value=N * niter / dt, unit="cell-updates/s (host API, end to end)")
    return out


def leg_config5(torch, dist, dev, L, cs, synth, a, world, rank, hbm_peak, barrier, max_over_ranks):
    """BASELINE.json configs[4]: `spots` cells in all, sharded over the ranks (strong scaling).
    Per rank: gene->bin aggregation of its columns, log-likelihood matrix of its cells against
    K = 20 clusters, sufficient statistics of its cells, ONE NCCL all-reduce of the statistics,
    tally of its cells over T = 100 kept sweeps.  No other collective on the path."""
    from clonalscope_b200 import distributed as D5
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lo, hi = D5.shard_range(a.spots, world, rank)
    n = hi - lo
    R, K, T, G = a.segs, 20, 100, a.genes
    bins, genes = synth.geometry(G=G)
    map_ptr, map_bin = synth.geometry_map(bins, genes)
    B = len(bins["chr"])
    dp = lambda arr, t: arr.ctypes.data_as(C.POINTER(t))  # noqa: E731
    colptr, rowidx, x = synth_counts_gpu(torch, dev, n, G, seed=5001 + rank, mean_genes=a.spot_genes)
    nnz_in = int(colptr[-1].item())
    plan = C.c_void_p()
    assert L.cs_bin_plan_create(G, B, dp(map_ptr, C.c_int32), dp(map_bin, C.c_int32), C.byref(plan)) == 0, L.cs_last_error()
    out_colptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    nnz_out = C.c_int64()
    assert L.cs_bin_counts_dev(plan, n, vp(colptr), vp(rowidx), vp(x), vp(out_colptr), None, None, 0, C.byref(nnz_out), st) == 0
    cap = nnz_out.value
    out_rowidx = torch.empty(cap, dtype=torch.int32, device=dev)
    out_x = torch.empty(cap, dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(77 + rank)
    X = 0.3 + 2.2 * torch.rand(n, R, generator=g, device=dev, dtype=torch.float64)
    U = 0.3 + 2.2 * torch.rand(K, R, generator=torch.Generator(device=dev).manual_seed(9), device=dev, dtype=torch.float64)
    sig = torch.full((R,), 0.25, device=dev, dtype=torch.float64)
    LL = torch.empty(n, K, device=dev, dtype=torch.float64)
    # posterior-like labels: every spot in its home cluster, 5 % of its sweeps in a neighbour
    home = torch.randint(1, K + 1, (n,), generator=g, device=dev, dtype=torch.int32)
    Zall = torch.where(torch.rand(T, n, generator=g, device=dev) < 0.05, home[None, :] % K + 1, home[None, :]).contiguous()
    del home
    zmaj = torch.empty(n, device=dev, dtype=torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]

    def step():
        ev[0].record()
        assert L.cs_bin_counts_dev(plan, n, vp(colptr), vp(rowidx), vp(x), vp(out_colptr), vp(out_rowidx), vp(out_x), cap,
                                   C.byref(nnz_out), st) == 0, L.cs_last_error()
        ev[1].record()
        assert L.cs_loglik_matrix_dev(n, R, K, vp(X), vp(U), vp(sig), vp(LL), st) == 0, L.cs_last_error()
        ev[2].record()
        assert L.cs_majority_vote_dev(T, n, vp(Zall), C.c_int64(n), C.c_int64(1), vp(zmaj), st) == 0, L.cs_last_error()
        ev[3].record()
        buf = D5.device_stats(X, zmaj, U, K)
        ev[4].record()
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        ev[5].record()
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(5)] + [ev[0].elapsed_time(ev[5])]

    for _ in range(3):
        step()
    rows = []
    for _ in range(5):
        barrier()
        rows.append(step())
    m = np.mean(np.array(rows), axis=0)
    tot = max_over_ranks(float(m[5]))
    # the all-reduce on its own: latency of one exchange of the statistics
    buf = torch.zeros(2 * K * R + K + 2 * R, dtype=torch.float64, device=dev)
    lat = None
    if world > 1:
        for _ in range(5):
            dist.all_reduce(buf)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_reduce(buf)
        e1.record()
        torch.cuda.synchronize()
        lat = max_over_ranks(e0.elapsed_time(e1) / 50 * 1e3)
    bin_bytes = 12 * nnz_in + 12 * cap + 2 * 8 * (n + 1) + 4 * (G + 1 + len(map_bin))
    L.cs_bin_plan_destroy(plan)
    return dict(spots=a.spots, spots_per_rank=n, genes=G, bins=B, segments=R, K=K, T=T, scaling="strong",
                value=a.spots / (tot * 1e-3), unit="spots/s (binning + log-likelihood matrix + tally + statistics + all-reduce)",
                ms=tot, ms_binning=float(m[0]), ms_loglik=float(m[1]), ms_tally=float(m[2]), ms_stats=float(m[3]),
                ms_allreduce_in_step=float(m[4]), allreduce_us=lat, allreduce_bytes=int(buf.numel() * 8),
                binning_gbs_per_rank=bin_bytes / (float(m[0]) * 1e-3) / 1e9,
                binning_frac_of_hbm_peak=bin_bytes / (float(m[0]) * 1e-3) / 1e9 / hbm_peak,
                note="rank 0's stage times; `ms` is the max over ranks of the whole step")



def run_ours(a):
    import torch
    import torch.distributed as dist
    import clonalscope_b200 as cs
    from clonalscope_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = cs.lib()
    assert L.cs_set_device(local) == 0
    hbm_peak, peak_src = peaks()
    N, R, niter = a.cells, a.segs, a.niter

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- sampler workload (each rank: an independent chain, seed + rank)
    s = synth.sampler_input(N, R=R)
    X = np.asfortranarray(s["X"])
    U0 = np.asfortranarray(s["U0"])
    sig0 = np.ascontiguousarray(s["sigmas0"])
    Z0 = (cs.loglik_matrix(X, s["U0"], sig0).argmax(axis=1) + 1).astype(np.int32)
    seed = 200 + rank
    dp = lambda arr, t: arr.ctypes.data_as(C.POINTER(t))  # noqa: E731
    run_chain, chain_stats, chain_destroy = make_chain_runner(L, _lib, N, R, niter, X, U0, Z0, sig0, 2.0, 2.0)

    for _ in range(a.warmup):
        run_chain(seed)
    barrier()
    run_ms = []
    with ClockSampler(local) as clk:
        for _ in range(a.steps):
            barrier()
            run_ms.append(run_chain(seed)[0])
        barrier()
        # kernel-group split (events around every launch group; a separate pass, not part of `value`)
        prof = run_chain(seed, 2)
    clocks = clk.summary()
    cstats = chain_stats()
    my_ms = float(np.mean(run_ms))
    step_ms = max_over_ranks(my_ms)
    value = world * N * niter / (step_ms * 1e-3)
    if world > 1:
        t = torch.tensor([my_ms, float(cstats["events"]), float(cstats["walk_steps"])], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        ranks = [dict(rank=i, seed=200 + i, ms_per_step=float(v[0]), events=int(v[1]), walk_steps=int(v[2])) for i, v in enumerate(allr)]
    else:
        ranks = [dict(rank=0, seed=seed, ms_per_step=my_ms, events=cstats["events"], walk_steps=cstats["walk_steps"])]

    seeds = None
    if world == 1 and not a.no_seeds:
        seeds = leg_seeds(run_chain, chain_stats, N, niter, list(range(200, 208)))

    # ---------------- end to end through the C ABI with host buffers
    Zall = np.zeros((niter, N), dtype=np.int32, order="F")
    sigma_all = np.zeros((R, niter), order="F")
    LLv, L0 = np.zeros(niter), C.c_double()
    kt_all = np.zeros(niter, dtype=np.int32)
    ucap = niter * 64
    u_off = np.zeros(niter + 1, dtype=np.int64)
    u_labels = np.zeros(ucap, dtype=np.int32)
    u_rows = np.zeros((ucap, R))
    opts = _lib.GibbsOpts()
    opts.rng_mode, opts.device = _lib.RNG_PHILOX, local

    def e2e_step():
        rc = L.cs_ncrp_gibbs(N, R, dp(X, C.c_double), 2, dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                             C.c_double(2.0), C.c_double(2.0), niter, C.c_uint32(seed), C.byref(opts),
                             dp(Zall, C.c_int), dp(sigma_all, C.c_double), dp(LLv, C.c_double), C.byref(L0),
                             dp(kt_all, C.c_int), C.c_int64(ucap), dp(u_off, C.c_int64), dp(u_labels, C.c_int),
                             dp(u_rows, C.c_double), None, None)
        assert rc == 0, L.cs_last_error()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / a.steps)
    h2d = X.nbytes + U0.nbytes + Z0.nbytes + sig0.nbytes
    d2h = Zall.nbytes + sigma_all.nbytes + LLv.nbytes + kt_all.nbytes + int(u_off[-1]) * (R * 8 + 4) + u_off.nbytes
    e2e = dict(value=world * N * niter / (e2e_ms * 1e-3), unit="cell-updates/s", ms_per_step=e2e_ms,
               h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h))
    chain_destroy()

    # ---------------- independent chains run concurrently on the GPU (supplementary: a chain's
    # sequential walk occupies one cluster of SMs, other chains' all-SM kernels fill the rest)
    conc = None
    if a.chains > 1:
        chs = []
        for k in range(a.chains):
            ch = C.c_void_p()
            assert L.cs_chain_create(N, R, 0, niter, _lib.RNG_PHILOX, C.byref(ch)) == 0, L.cs_last_error()
            chs.append(ch)

        def chains_step():
            for k, ch in enumerate(chs):
                rc = L.cs_chain_init(ch, dp(X, C.c_double), 2, dp(U0, C.c_double), dp(Z0, C.c_int), dp(sig0, C.c_double),
                                     C.c_double(2.0), C.c_double(2.0), C.c_uint32(seed + 1000 * (k + 1)), 0)
                assert rc == 0, L.cs_last_error()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in range(niter):  # sweeps enqueued round-robin: no chain's launches wait behind another's
                for ch in chs:
                    assert L.cs_chain_run(ch, 1, int(t == niter - 1), 0) == 0, L.cs_last_error()
            for ch in chs:
                assert L.cs_chain_sync(ch) == 0, L.cs_last_error()
            return (time.perf_counter() - t0) * 1e3

        chains_step()
        cms = max_over_ranks(chains_step())
        conc = dict(chains=a.chains, ms=cms, value=world * a.chains * N * niter / (cms * 1e-3), unit="cell-updates/s",
                    note="aggregate over independent chains sharing one GPU (host wall clock around "
                         "enqueue + sync); the headline `value` is ONE chain")
        for ch in chs:
            L.cs_chain_destroy(ch)

    # ---------------- gene -> bin aggregation (cells sharded across ranks, no collective)
    G, NB = a.genes, a.bin_cells
    bins, genes = synth.geometry(G=G)
    map_ptr, map_bin = synth.geometry_map(bins, genes)
    B = len(bins["chr"])
    colptr, rowidx, x = synth_counts_gpu(torch, dev, NB, G, seed=1001 + rank)
    nnz_in = int(colptr[-1].item())
    plan = C.c_void_p()
    rc = L.cs_bin_plan_create(G, B, dp(map_ptr, C.c_int32), dp(map_bin, C.c_int32), C.byref(plan))
    assert rc == 0, L.cs_last_error()
    out_colptr = torch.zeros(NB + 1, dtype=torch.int64, device=dev)
    nnz_out = C.c_int64()
    st = torch.cuda.current_stream().cuda_stream
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rc = L.cs_bin_counts_dev(plan, NB, vp(colptr), vp(rowidx), vp(x), vp(out_colptr), None, None, 0,
                             C.byref(nnz_out), C.c_void_p(st))
    assert rc == 0, L.cs_last_error()
    cap = nnz_out.value
    out_rowidx = torch.empty(cap, dtype=torch.int32, device=dev)
    out_x = torch.empty(cap, dtype=torch.float64, device=dev)
    bin_bytes = 12 * nnz_in + 12 * cap + 2 * 8 * (NB + 1) + 4 * (G + 1 + len(map_bin))

    def bin_step():
        rc = L.cs_bin_counts_dev(plan, NB, vp(colptr), vp(rowidx), vp(x), vp(out_colptr), vp(out_rowidx), vp(out_x),
                                 cap, C.byref(nnz_out), C.c_void_p(st))
        assert rc == 0, L.cs_last_error()

    for _ in range(max(3, a.warmup)):
        bin_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bin_ms, fill_ms, count_ms, dev_ms = [], [], [], []
    for _ in range(max(5, a.steps)):
        barrier()
        e0.record()
        bin_step()
        e1.record()
        torch.cuda.synchronize()
        bin_ms.append(e0.elapsed_time(e1))
        t = [C.c_double() for _ in range(3)]
        L.cs_bin_plan_timing(plan, *[C.byref(v) for v in t])
        count_ms.append(t[0].value)
        fill_ms.append(t[2].value)
        dev_ms.append(t[0].value + t[1].value + t[2].value)
    # device time of the call's GPU work (CUDA events on its stream around memset + kernels), max over ranks;
    # `call_ms` adds what the host puts around it (launch, the copy back of the total, the synchronise)
    bms = max_over_ranks(float(np.mean(dev_ms)))
    call_ms = max_over_ranks(float(np.mean(bin_ms)))
    fms = float(np.mean(fill_ms))
    binning = dict(value=world * bin_bytes / (bms * 1e-3) / 1e9, unit="GB/s (algorithmic bytes / device time, count+scan+fill)",
                   ms=bms, call_ms=call_ms, call_gbs=world * bin_bytes / (call_ms * 1e-3) / 1e9, cells=NB * world, genes=G, bins=B, nnz_in=nnz_in, nnz_out=cap,
                   algorithmic_bytes=bin_bytes, count_kernel_ms=float(np.mean(count_ms)), fill_kernel_ms=fms,
                   frac_of_hbm_peak=bin_bytes / (bms * 1e-3) / 1e9 / hbm_peak,
                   fill_kernel_gbs=(12 * nnz_in + 12 * cap) / (fms * 1e-3) / 1e9,
                   l2_note="inputs+outputs (~5 GB) exceed the 126 MB L2",
                   traffic=3105289000 + 2905120000 if NB == 100000 else None,
                   traffic_note="dram__bytes_read.sum + dram__bytes_write.sum of one bin_onepass_kernel launch at 100k cells "
                                "(profiles/r2_bin_onepass_full.txt): 1.08x the algorithmic bytes (rows read by the count and, "
                                "a column later, by the fill)")
    # binning end to end with host buffers (R layout), bounded to e2e_bin_cells
    nb2 = min(NB, a.e2e_bin_cells)
    cp_h = colptr[:nb2 + 1].cpu().numpy().astype(np.int32)
    ri_h = rowidx[:int(cp_h[-1])].cpu().numpy()
    x_h = x[:int(cp_h[-1])].cpu().numpy()
    from clonalscope_b200.api import bin_counts_csc
    bin_counts_csc(G, nb2, B, cp_h, ri_h, x_h, map_ptr, map_bin)
    reps = []
    for _ in range(3):
        t0 = time.perf_counter()
        ocp, ori, ox = bin_counts_csc(G, nb2, B, cp_h, ri_h, x_h, map_ptr, map_bin)
        reps.append(time.perf_counter() - t0)
    dt = float(np.median(reps))
    binning["e2e"] = dict(cells=nb2, ms=dt * 1e3, gbs=(12 * int(cp_h[-1]) + 12 * len(ori)) / dt / 1e9,
                          ms_reps=[round(r * 1e3, 1) for r in reps],
                          note="median of 3 calls through the host API: pageable inputs, freshly allocated outputs",
                          h2d_bytes=int(cp_h.nbytes + ri_h.nbytes + x_h.nbytes), d2h_bytes=int(ocp.nbytes + ori.nbytes + ox.nbytes))
    # ---------------- the caller of the aggregation (CreateSegtableNoWGS): bin x cluster pooling of the
    # device-resident result, then running mean + 4-state Viterbi of every (cluster, chromosome) sequence
    segtable = None
    if rank == 0 and world == 1 and not a.no_extra:
        Kc = 4
        labels = torch.randint(0, Kc, (NB,), device=dev, dtype=torch.int32)
        pooled = torch.zeros((Kc, B), dtype=torch.float64, device=dev)
        for _ in range(3):
            assert L.cs_rowsum_by_label_dev(B, NB, vp(out_colptr), vp(out_rowidx), vp(out_x), vp(labels), Kc, vp(pooled),
                                            C.c_void_p(st)) == 0, L.cs_last_error()
        e0.record()
        for _ in range(5):
            L.cs_rowsum_by_label_dev(B, NB, vp(out_colptr), vp(out_rowidx), vp(out_x), vp(labels), Kc, vp(pooled), C.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        pool_ms = e0.elapsed_time(e1) / 5
        prof_h = pooled.cpu().numpy()
        chrom_of_bin = np.asarray(bins["chr"])
        seqs = []
        for k in range(Kc):
            v = prof_h[k] / max(np.median(prof_h[k]), 1.0)
            for c in range(1, 23):
                seqs.append(v[chrom_of_bin == c])
        from clonalscope_b200.segtable import segment_hmm
        t_pi = 1e-6
        Pi = np.full((4, 4), t_pi)
        np.fill_diagonal(Pi, 1 - 3 * t_pi)
        segment_hmm(seqs, 100, [1.8, 1.5, 1.0, 0.5], [0.2] * 4, [0.1, 0.2, 0.5, 0.2], Pi)
        t0 = time.perf_counter()
        segment_hmm(seqs, 100, [1.8, 1.5, 1.0, 0.5], [0.2] * 4, [0.1, 0.2, 0.5, 0.2], Pi)
        hmm_ms = (time.perf_counter() - t0) * 1e3
        segtable = dict(pooling_ms=pool_ms, pooling_gbs=12.0 * cap / (pool_ms * 1e-3) / 1e9,
                        pooling_frac_of_hbm_peak=12.0 * cap / (pool_ms * 1e-3) / 1e9 / hbm_peak, clusters=Kc, cells=NB, bins=B,
                        hmm_ms=hmm_ms, hmm_sequences=len(seqs), hmm_bins=int(sum(len(q) for q in seqs)),
                        note="cs_rowsum_by_label_dev on the binning leg's device-resident output (12 B per stored count, "
                             "FP64 atomics); cs_segment_hmm through the host API (one launch for every cluster and chromosome)")
    L.cs_bin_plan_destroy(plan)

    # ---------------- EstRegionCov (a7) end to end with host buffers on the same count matrix
    region = None
    if rank == 0 and world == 1 and not a.no_region:
        import scipy.sparse as sp
        from clonalscope_b200.api import EstRegionCov
        rng = np.random.default_rng(77)
        mtx = sp.csc_matrix((x_h, ri_h, cp_h), shape=(G, nb2))
        feats = [f"g{i}" for i in range(G)]
        bars = [f"c{i}" for i in range(nb2)]
        bed = np.empty((G, 4), dtype=object)
        bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3] = [str(int(c)) for c in genes["chr"]], genes["start"], genes["end"], feats
        ct = np.empty((nb2, 2), dtype=object)
        ct[:, 0], ct[:, 1] = bars, rng.choice(["tumor", "normal"], size=nb2, p=[0.7, 0.3])
        seg = dict(chr=[], start=[], end=[], chrr=[])
        for c, size_c in enumerate(synth.CHR_SIZES):
            nseg = 14 if c < 14 else 13          # 300 segments in all
            edges = np.linspace(0, size_c, nseg + 1)
            for lo_, hi_ in zip(edges[:-1], edges[1:]):
                seg["chr"].append(str(c + 1)), seg["start"].append(float(lo_)), seg["end"].append(float(hi_))
                seg["chrr"].append(f"chr{c + 1}_{int(lo_)}_{int(hi_)}")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            EstRegionCov(mtx, bars, feats, bed, ct, seg_table_filtered=seg)
            rreps = []
            for _ in range(3):
                t0 = time.perf_counter()
                out_rc = EstRegionCov(mtx, bars, feats, bed, ct, seg_table_filtered=seg)
                rreps.append(time.perf_counter() - t0)
            rdt = float(np.median(rreps))
        region = dict(cells=nb2, genes=G, segments=len(out_rc["deltas_all"]), ms=rdt * 1e3,
                      ms_reps=[round(r * 1e3, 1) for r in rreps], cells_per_s=nb2 / rdt,
                      note="EstRegionCov through the host API (scipy CSC in, per-segment deltas out), median of 3 calls")
        if not a.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as O2
            ncpu = min(nb2, 100)
            dense = np.asarray(mtx[:, :ncpu].todense())
            t0 = time.perf_counter()
            O2.est_region_cov(dense, bars[:ncpu], feats, [bed[:, 0], bed[:, 1], bed[:, 2], bed[:, 3]],
                              (bars[:ncpu], list(ct[:ncpu, 1])), seg_table=seg)
            cdt = time.perf_counter() - t0
            region["cpu_cells_per_s"] = ncpu / cdt
            region["cpu_sample"] = f"dense restatement of the reference, {ncpu} cells, 1 core"

    # ---------------- CPU baseline on the box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        O.build()
        cs_n = min(N, a.cpu_cells)
        rate, dt = cpu_gibbs_rate(cs_n, R, a.cpu_sweeps, 1)
        t0 = time.perf_counter()
        nb3 = min(nb2, 5000)
        O.bin_counts(G, nb3, B, cp_h[:nb3 + 1].astype(np.int64), ri_h, x_h, map_ptr, map_bin)
        bdt = time.perf_counter() - t0
        cpu = dict(value=rate, unit="cell-updates/s", cores=1, kind="port",
                   sample=f"first {a.cpu_sweeps} sweeps of {cs_n} cells x {R} segments, R-compatible RNG "
                          f"(the reference is single-threaded R; this is its C restatement, {dt:.1f} s)",
                   host_cores_available=os.cpu_count(),
                   binning_gbs=24.0 * int(cp_h[nb3]) / bdt / 1e9,
                   binning_sample=f"sparse restatement, {nb3} cells, 1 core")

    fixtures = kernels = allele = None
    if world == 1 and not a.no_extra:
        fixtures = leg_fixtures(L, _lib)
        kernels = leg_kernels(torch, dev, L, hbm_peak, N, R)
        allele = leg_allele(cs)
    config5 = None
    if a.spots > 0:
        config5 = leg_config5(torch, dist if world > 1 else None, dev, L, cs, synth, a, world, rank, hbm_peak, barrier, max_over_ranks)
    if cpu is not None and not a.no_cpu:
        # the same algorithm as the GPU path (Philox mode of the oracle), one sweep, one core
        import oracle as O3
        t0 = time.perf_counter()
        O3.gibbs(s["X"][:cs_n], s["U0"], Z0[:cs_n], s["sigmas0"], 2.0, 2.0, 1, 200, rng_mode=O3.RNG_PHILOX, ucap_rows=512)
        cpu["philox_port_value"] = cs_n / (time.perf_counter() - t0)
        cpu["philox_port_note"] = "the oracle's Philox mode (the GPU path's own specification), 1 sweep, 1 core"

    if rank == 0:
        prep_ms = prof[1] / niter  # dominant kernel of the step
        walk_ms = prof[3] / niter
        alg_bytes = N * R * 8      # S*8 bytes per cell-update, N cell-updates per launch
        ach = alg_bytes / (prep_ms * 1e-3) / 1e9
        line = dict(
            metric=METRIC, value=value, unit="cell-updates/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
            ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic", config=workload_config(N, R, niter, NB, G, B),
            e2e=e2e, gpu_launches=int(a.steps * niter * KERNELS_PER_SWEEP),
            roofline=dict(bound="hbm", kernel="prepare_kernel", achieved=ach, peak=hbm_peak, unit="GB/s",
                          frac=ach / hbm_peak, traffic=PREPARE_TRAFFIC_BYTES, peak_source=peak_src,
                          note="one launch per sweep: reads X once (N*S*8 algorithmic bytes), scores every cell against "
                               "every live cluster in 64-bit fixed point, draws the proposal (Philox) and the speculative "
                               "assignment; instruction-bound (see profiles/), so the HBM fraction is low by construction",
                          ms_per_launch=prep_ms, share_of_step=prof[1] / max(prof[0], 1e-9),
                          fix_walk_kernel=dict(ms_per_launch=walk_ms, share_of_step=prof[3] / max(prof[0], 1e-9),
                                               achieved=alg_bytes / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else None,
                                               frac=alg_bytes / (walk_ms * 1e-3) / 1e9 / hbm_peak if walk_ms > 0 else None,
                                               traffic=WALK_TRAFFIC_BYTES,
                                               note="cooperative, all SMs: the sweep as the fixed point of a parallel map "
                                                    "(a few passes over the cells behind the last settled one); "
                                                    "latency-bound on grid barriers and gathers")),
            sweep_split_ms=dict(run=prof[0], prepare=prof[1], walk=prof[3], end_of_sweep=prof[4]),
            walk=cstats, ranks=ranks, binning=binning, clocks=clocks)
        if seeds:
            line["seeds"] = seeds
        if fixtures:
            line["fixtures"] = fixtures
        if kernels:
            line["kernels"] = kernels
        if segtable:
            line["segtable"] = segtable
        if allele:
            line["allele_sampler"] = allele
        if config5:
            line["config5"] = config5
        if conc:
            line["concurrent_chains"] = conc
        if region:
            line["region_cov"] = region
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's original stdout (see main())."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries write to stdout behind our back (NCCL prints its version banner there at N > 1): file
    # descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate of the
    # original stdout, so stdout carries exactly one line.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--segs", type=int, default=300)
    ap.add_argument("--niter", type=int, default=200)
    ap.add_argument("--no-region", action="store_true", help="skip the EstRegionCov leg")
    ap.add_argument("--chains", type=int, default=8, help="independent chains run concurrently (supplementary; 1 = skip)")
    ap.add_argument("--genes", type=int, default=30000)
    ap.add_argument("--bin-cells", type=int, default=100000)
    ap.add_argument("--e2e-bin-cells", type=int, default=20000)
    ap.add_argument("--cpu-cells", type=int, default=100000)
    ap.add_argument("--cpu-sweeps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-seeds", action="store_true", help="skip the seeds 200..207 leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the fixtures / kernels / allele legs")
    ap.add_argument("--spots", type=int, default=1000000, help="BASELINE config 5: cells sharded over the ranks (0 = skip)")
    ap.add_argument("--spot-genes", type=int, default=2000, help="mean expressed genes per spot in the config-5 leg")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
