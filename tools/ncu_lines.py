"""Per-source-line warp-stall samples of one profiled kernel.

  python tools/ncu_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [top]

The report's SASS page (ncu --set full --import-source on) gives samples per instruction; the
line table comes from `nvdisasm -g` of the cubin the kernel was built into
(`cuobjdump -xelf all libclonalscope_b200.so`).  Kernels must be built with -lineinfo.
"""
import collections
import csv
import re
import subprocess
import sys


def line_table(cubin, kernel):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    table, cur, inside = {}, ("?", 0), False
    for ln in dis:
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(cubin, kernel)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    H = rows[1]
    si, ii = H.index("# Samples"), H.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if r and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    per_line = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    total = 0.0
    for r in data:
        off = int(r[0], 16) - base
        key = table.get(off, (("?", 0), ""))[0]
        s = float(r[si] or 0)
        per_line[key][0] += s
        per_line[key][1] += float(r[ii] or 0)
        for i in stall_cols:
            v = float(r[i] or 0)
            if v:
                per_line[key][2][H[i]] += v
        total += s
    print(f"# {kernel}: {total:.0f} samples, {len(data)} SASS instructions")
    print(f"{'file:line':32s} {'samples':>8s} {'share':>7s} {'warp inst':>10s}  top stalls")
    for key, (s, n, st) in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
        tops = ", ".join(f"{k.replace('stall_', '')} {v / max(s, 1):.0%}" for k, v in st.most_common(3))
        print(f"{key[0] + ':' + str(key[1]):32s} {s:8.0f} {s / total:7.1%} {n:10.0f}  {tops}")


if __name__ == "__main__":
    main()
