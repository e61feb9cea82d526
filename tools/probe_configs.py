"""Throughput of the tracking kernel on the other BASELINE configurations (developer tool)."""
import json, sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from mc_mpi_b200 import configs
from mc_mpi_b200.layer import decompose_domain
cases = [("default_slab (config.yaml physics)", configs.reference_default(50_000_000)),
         ("single_gpu_slab_1000", configs.single_gpu_slab(50_000_000)),
         ("absorption_dominated", configs.absorption_dominated(50_000_000)),
         ("optically_thick", configs.optically_thick(4_000_000)),
         ("test_layer physics, 100 cells", configs.ref_test_layer().with_particles(100_000_000)),
         ("heterogeneous 8192 cells (shared tally)", configs.heterogeneous(8192, 2_000_000)),
         ("heterogeneous 65536 cells (L2 tally)", configs.heterogeneous(65536, 100_000))]
for name, cfg in cases:
    g = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, cfg.nb_particles,
                         cfg.particle_min_weight, sigs=cfg.sigs, absorption_rates=cfg.absorption_rates)
    best = None
    for rep in range(2):
        g.create_particles(cfg.x_ini, 1.0 / cfg.nb_particles, cfg.nb_particles)
        c0 = g.counts(); c = g.simulate(-1)
        ms = c["track_ms"] - c0["track_ms"]; ev = c["events"] - c0["events"]
        if best is None or ms < best[0]: best = (ms, ev, c["scatters"] - c0["scatters"])
    ms, ev, sc = best
    print(json.dumps({"config": name, "cells": cfg.nb_cells, "histories": cfg.nb_particles,
                      "track_ms": round(ms, 2), "events_per_history": round(ev / cfg.nb_particles, 1),
                      "scatters_per_history": round(sc / cfg.nb_particles, 2),
                      "events_per_s": ev / ms * 1e3, "histories_per_s": cfg.nb_particles / ms * 1e3,
                      "roofline_frac_48B": ev / ms * 1e3 * 48 / 6532.5e9}), flush=True)
    g.close()
