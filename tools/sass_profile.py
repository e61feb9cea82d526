"""Per-source-line instruction counts of one kernel from an ncu report (developer tool).

    python tools/sass_profile.py <report.ncu-rep> <kernel substring> [cubin dir]

Joins the executed-instruction counts of `ncu --page source` (per SASS address) with the line
table of `nvdisasm -g` on the cubins extracted from mc_mpi_b200/libmcb200.so (built with
-lineinfo), and prints warp-instructions / stall samples per source line, heaviest first."""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_counts(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res, name, hdr, base = {}, None, None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            name, hdr, base = r[1], None, None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or name is None or kernel not in name or len(r) < len(hdr):
            continue
        a = int(r[hdr.index("Address")], 16)
        base = a if base is None else base
        res[a - base] = (int(r[hdr.index("Instructions Executed")]),
                         int(r[hdr.index("Thread Instructions Executed")]),
                         int(r[hdr.index("# Samples")]), r[hdr.index("Source")].strip())
    return res


def line_table(kernel, cubin_dir):
    tab, cur_fn, cur_line = {}, None, None
    for cubin in glob.glob(os.path.join(cubin_dir, "*.cubin")):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kernel in cur_fn:
                tab[int(m.group(1), 16)] = cur_line
    return tab


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    cubin_dir = sys.argv[3] if len(sys.argv) > 3 else None
    if cubin_dir is None:
        cubin_dir = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mc_mpi_b200", "libmcb200.so")],
                       cwd=cubin_dir, capture_output=True)
    counts = ncu_counts(rep, kernel)
    lines = line_table(kernel.replace("<", "").split("(")[0], cubin_dir)
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = sum(c[0] for c in counts.values())
    for off, (wi, ti, smp, _) in counts.items():
        key = lines.get(off, ("?", 0))
        a = agg[key]
        a[0] += wi
        a[1] += ti
        a[2] += smp
        a[3] += 1
    print(f"total warp-instructions {tot}")
    print(f"{'file:line':34s} {'warp-inst':>14s} {'share':>7s} {'lanes':>6s} {'samples':>9s} {'sass':>5s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "70"))]:
        print(f"{key[0] + ':' + str(key[1]):34s} {a[0]:14d} {100 * a[0] / tot:6.2f}% "
              f"{a[1] / max(a[0], 1):6.1f} {a[2]:9d} {a[3]:5d}")


if __name__ == "__main__":
    main()
