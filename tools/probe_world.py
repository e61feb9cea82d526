"""Device-side throughput of the persistent world kernel on ONE GPU (developer tool).
usage: python tools/probe_world.py [particles] [config] -> one JSON line per variant:
the layer path (track_kernel, bank in HBM) next to the world kernel (births in the kernel),
with windows / ranks sharing the GPU to see what the exchange costs per segment."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.layer import decompose_domain  # noqa: E402
from mc_mpi_b200.worker import LocalBox, totals  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
name = sys.argv[2] if len(sys.argv) > 2 else "single_gpu_slab_1000"
cfg = configs.BY_NAME[name]().with_particles(n)

g = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, n, cfg.particle_min_weight,
                     sigs=cfg.sigs, absorption_rates=cfg.absorption_rates)
best = None
for rep in range(3):
    g.create_particles(cfg.x_ini, 1.0 / n, n)
    c0 = g.counts()
    c = g.simulate(-1)
    ms, ev = c["track_ms"] - c0["track_ms"], c["events"] - c0["events"]
    if best is None or ms < best[0]:
        best = (ms, ev)
g.close()
print(json.dumps({"path": "layer", "config": cfg.name, "particles": n, "kernel_ms": round(best[0], 3),
                  "events_per_s": best[1] / best[0] * 1e3, "histories_per_s": n / best[0] * 1e3}), flush=True)

variants = [
    dict(K=1),
    dict(K=1, retire_batch=1), dict(K=1, retire_batch=4), dict(K=1, retire_batch=8),
    dict(K=1, block=512), dict(K=1, block=1024),
    dict(K=1, windows=2), dict(K=1, windows=4), dict(K=1, windows=8),
    dict(K=1, windows=8, retire_batch=2), dict(K=1, windows=8, retire_batch=8),
    dict(K=2, max_ctas=296), dict(K=4, max_ctas=148), dict(K=8, max_ctas=74),
]
if len(sys.argv) > 3:
    variants = [json.loads(sys.argv[3])]
for v in variants:
    opts = dict(v)
    K = opts.pop("K")
    try:
        box = LocalBox(cfg, K, **opts)
        box.set_option("max_run_ms", 120_000)
        best = None
        for rep in range(3):
            res = box.run()
            t = totals(res)
            if best is None or t["kernel_ms_max"] < best[0]["kernel_ms_max"]:
                best = (t, res)
        t, res = best
        print(json.dumps({"path": "world", **v, "config": cfg.name, "particles": n,
                          "kernel_ms": round(t["kernel_ms_max"], 3),
                          "events_per_s": t["events"] / t["kernel_ms_max"] * 1e3,
                          "histories_per_s": n / t["kernel_ms_max"] * 1e3,
                          "segments_per_history": (n + t["sent_left"] + t["sent_right"] + t["window_crossings"]) / n,
                          "windows": res[0]["windows"], "ctas": res[0]["ctas"], "block": res[0]["block"],
                          "stripes": res[0]["stripes"], "ring_cap": res[0]["ring_cap"],
                          "idle_polls": t["idle_polls"], "blocked_passes": t["blocked_passes"],
                          "bank_pushes": t["bank_pushes"]}), flush=True)
        box.close()
    except Exception as ex:  # keep sweeping
        print(json.dumps({"path": "world", **v, "error": str(ex)}), flush=True)
