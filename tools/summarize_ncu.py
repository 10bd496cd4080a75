"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep       > profiles/r1_x_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} {n:8d} {v / 1e3:10.3f} {v / n:10.1f} {v / tot:7.1%}")
    print(f"{'TOTAL':44s} {sum(n for n, _ in agg.values()):8d} {tot / 1e3:10.3f}")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none  ({path})")
    for row in rows[2:]:
        name = row[H.index("Kernel Name")] if "Kernel Name" in H else "?"
        print(f"kernel: {name}")
        for k in KEYS:
            if k in H:
                i = H.index(k)
                print(f"  {k:68s} {row[i]:>16s} {U[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        H = rows[1]
        cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[2:] if len(r) > max(cols) and r[0].startswith("0x")]
        agg = {H[i]: sum(float(r[i] or 0) for r in data) for i in cols}
        s = sum(agg.values()) or 1.0
        print("  warp stall sampling (share of samples):")
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
            print(f"    {k:28s} {v / s:6.1%}")
        ie = H.index("Instructions Executed")
        print(f"  SASS instructions: {len(data)}, warp instructions executed: {sum(float(r[ie]) for r in data):.4g}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
