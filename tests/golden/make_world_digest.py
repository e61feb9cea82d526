"""Generates tests/golden/world_digest.json: what a whole run must produce, computed by the
ORACLE (oracle/mc_oracle.c, pinned against the compiled reference by tests/test_oracle_pin.py)
as ONE layer spanning the slab.  The exact tally is an integer sum, so it does not depend on
thread count, cuts, windows or GPU count: a K-GPU run must reproduce these digests bit for bit.

    python tests/golden/make_world_digest.py        # ~1 min on 8 cores

bench.py --gpus N checks its N-rank parity run against this file (no oracle needed on the box).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mc_mpi_b200 import configs  # noqa: E402
from util import make_oracle  # noqa: E402

CASES = {
    # name -> config: the bench's parity run uses the first one
    "default_slab_2e6": configs.reference_default(2_000_000),          # config.yaml physics, minw 1e-12
    "single_gpu_slab_2e6": configs.single_gpu_slab(2_000_000),          # bench workload physics, minw 0
    "default_slab_1e5": configs.reference_default(100_000),             # BASELINE configs[0]
    "absorption_dominated_2e5": configs.absorption_dominated(200_000),
    "optically_thick_2e4": configs.optically_thick(20_000),
}


def digest_of(cfg):
    o = make_oracle(cfg)
    o.simulate(-1, nthread=os.cpu_count() or 1)
    st = o.stats()
    exact = np.ascontiguousarray(o.tally_exact, dtype="<u4")
    cw = o.class_weights_exact
    out = {
        "config": cfg.name, "nb_cells": cfg.nb_cells, "nb_particles": cfg.nb_particles,
        "particle_min_weight": float(np.float32(cfg.particle_min_weight)),
        "seed": 5127801,
        "tally_exact_sha256": hashlib.sha256(exact.tobytes()).hexdigest(),
        "tally_lsb_log2": -120,
        "events": int(st["events"]), "scatters": int(st["scatters"]),
        "n_left": int(st["n_left"]), "n_right": int(st["n_right"]), "n_dead": int(st["n_dead"]),
        "w_left": float(cw[0]), "w_right": float(cw[1]), "w_dead": float(cw[2]),
        "w_absorbed": float(np.sum(o.tally_exact_f64)),
        # a few cells in clear, for diagnostics when the hash differs
        "cells_f64": {str(i): float(o.tally_exact_f64[i]) for i in (0, cfg.nb_cells // 2, cfg.nb_cells - 1)},
    }
    o.free()
    return out


if __name__ == "__main__":
    res = {name: digest_of(cfg) for name, cfg in CASES.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "world_digest.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
        f.write("\n")
    for k, v in res.items():
        print(k, v["events"], v["n_left"], v["n_right"], v["n_dead"], v["tally_exact_sha256"][:16])
