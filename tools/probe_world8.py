"""One-GPU proxy of an 8-GPU run (developer tool): 8 ranks share the GPU (74 CTAs each), cuts
given, per-rank lane utilisation printed.  usage: probe_world8.py [particles] [json opts] [cuts json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.worker import LocalBox, totals  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
opts = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
cuts = json.loads(sys.argv[3]) if len(sys.argv) > 3 else None
K = opts.pop("K", 8)
cfg = configs.reference_default(n)
box = LocalBox(cfg, K, cuts=cuts, max_ctas=592 // K, **opts)
box.set_option("max_run_ms", 120_000)
best = None
for rep in range(2):
    res = box.run()
    t = totals(res)
    if best is None or t["kernel_ms_max"] < best[0]["kernel_ms_max"]:
        best = (t, res)
t, res = best
print(json.dumps({"K": K, **opts, "cuts": cuts, "kernel_ms": round(t["kernel_ms_max"], 2),
                  "events_per_s": t["events"] / t["kernel_ms_max"] * 1e3,
                  "util": [round(r["events"] / max(r["lane_slots"], 1), 3) for r in res],
                  "occupancy": [round(__import__("mc_mpi_b200.worker", fromlist=["x"]).occupancy(r), 3) for r in res],
                  "events_share": [round(r["events"] / t["events"] * K, 3) for r in res],
                  "idle_polls": [r["idle_polls"] for r in res],
                  "blocked": t["blocked_passes"], "ring_cap": res[0]["ring_cap"]}), flush=True)
box.close()
