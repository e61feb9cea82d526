"""Quick device-side throughput sweep of the tracking kernel variants (developer tool).
usage: python tools/probe.py [particles] [config] [K rank]   (K rank: track only that sub-slab,
the workload of one rank of a K-GPU run: births if it owns the source, else nothing)"""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.layer import decompose_domain  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
name = sys.argv[2] if len(sys.argv) > 2 else "single_gpu_slab_1000"
cfg = configs.BY_NAME[name]()
cfg = cfg.with_particles(n)
variants = [dict(tally_mode=1), dict(tally_mode=2)]
shapes = [dict(block=256, blocks_per_sm=4), dict(block=128, blocks_per_sm=8),
          dict(block=512, blocks_per_sm=2), dict(block=1024, blocks_per_sm=1),
          dict(block=256, blocks_per_sm=6), dict(block=256, blocks_per_sm=8),
          dict(block=128, blocks_per_sm=12), dict(block=256, blocks_per_sm=2)]
K, R = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1, 0)
if K > 1:
    variants = [dict(tally_mode=1, retire_batch=b) for b in (1, 2, 4, 8, 12)]
    shapes = shapes[:1] + shapes[4:5]
for v in variants:
    for s in (shapes if (v == variants[0] or K > 1) else shapes[:1]):
        g = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, K, R, cfg.nb_cells, cfg.nb_particles,
                             cfg.particle_min_weight, sigs=cfg.sigs, global_dx=K > 1,
                             absorption_rates=cfg.absorption_rates)
        for k, val in {**v, **s}.items():
            g.set_option(k, val)
        best = None
        for rep in range(3):
            g.create_particles(cfg.x_ini, 1.0 / cfg.nb_particles, cfg.nb_particles)
            c0 = g.counts()
            t = time.time()
            c = g.simulate(-1)
            wall = time.time() - t
            g.pop_left(), g.pop_right()
            ms = c["track_ms"] - c0["track_ms"]
            ev = c["events"] - c0["events"]
            if best is None or ms < best[0]:
                best = (ms, ev, wall)
        ms, ev, wall = best
        print(json.dumps({**v, **s, "config": cfg.name, "particles": n, "track_ms": round(ms, 3),
                          "wall_ms": round(wall * 1e3, 1), "events_per_s": ev / ms * 1e3,
                          "histories_per_s": n / ms * 1e3,
                          "roofline_frac_48B": ev / ms * 1e3 * 48 / 6532.5e9}), flush=True)
        g.close()
