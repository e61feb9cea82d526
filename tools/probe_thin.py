"""Retire / refill overhead without starvation (developer tool): ONE rank, ONE window, a slab of
`cells` cells spanning [0,1] -> every history is one short segment fed by in-kernel births.
usage: probe_thin.py cells particles [json opts]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.worker import LocalBox, totals  # noqa: E402

cells = int(sys.argv[1])
n = int(float(sys.argv[2]))
opts = json.loads(sys.argv[3]) if len(sys.argv) > 3 else {}
cfg = configs.SlabConfig(f"thin_{cells}", cells, n, 0.0)
box = LocalBox(cfg, 1, **opts)
box.set_option("max_run_ms", 120_000)
best = None
for rep in range(3):
    res = box.run()
    t = totals(res)
    if best is None or t["kernel_ms_max"] < best["kernel_ms_max"]:
        best = t
t = best
print(json.dumps({"cells": cells, **opts, "kernel_ms": round(t["kernel_ms_max"], 2),
                  "events_per_history": t["events"] / n,
                  "events_per_s": t["events"] / t["kernel_ms_max"] * 1e3,
                  "util": round(t["events"] / max(t["lane_slots"], 1), 4),
                  "idle_polls": t["idle_polls"]}), flush=True)
box.close()
