"""Summarise an `ncu --set full` report of a tracking kernel into profiles/r02_kernel_issue.json
(developer tool; bench.py reads that file for roofline.issue / roofline.traffic).

    python tools/ncu_summary.py <kernel key> <report.ncu-rep> <kernel regex> <events> <histories> [note]

`events` / `histories` = what the profiled launch processed (the tool that made the report prints
them); the report is read with `ncu -i ... --page raw --csv`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_kernel_issue.json")

key, rep, pattern, events, histories = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5])
note = sys.argv[6] if len(sys.argv) > 6 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
row = [r for r in rows[2:] if pattern in r[hdr.index("Kernel Name")]][-1]
g = lambda name: float(row[hdr.index(name)].replace(",", ""))
units = rows[1]
dram = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    v, u = g(name), units[hdr.index(name)]
    dram += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
inst = g("smsp__inst_executed.sum")
entry = {
    "kernel": row[hdr.index("Kernel Name")],
    "issue_active_pct": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "warp_inst_per_32_events": inst / (events / 32.0),
    "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warps_eligible_per_cycle": g("smsp__warps_eligible.avg.per_cycle_active"),
    "pipe_pct": {p: g(f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active")
                 for p in ("alu", "fma", "fp64", "xu", "lsu")},
    "dram_bytes": dram, "dram_bytes_per_history": dram / histories,
    "duration_ms": g("gpu__time_duration.sum") * {"msecond": 1, "usecond": 1e-3, "second": 1e3}.get(
        units[hdr.index("gpu__time_duration.sum")], 1),
    "events": events, "histories": histories,
    "source": os.path.relpath(os.path.abspath(rep), ROOT).replace("gpurun_out/", "profiles/"),
    "note": note,
}
data = {}
if os.path.isfile(OUT):
    with open(OUT) as f:
        data = json.load(f)
data[key] = entry
with open(OUT, "w") as f:
    json.dump(data, f, indent=1, sort_keys=True)
    f.write("\n")
print(json.dumps(entry, indent=1))
