#!/usr/bin/env python
"""Regenerates tests/golden/ref_main_weights.npz from the reference's own executable
(tests/dropin/_bin/ref_main_cpu = unmodified src/main.cpp + workers + CPU Layer over minimpi):

    make -C tests/dropin cpu && python tests/golden/make_golden_main.py

Keys: sync_n{1,2,3} = per-cell absorbed weight (float32: the `weight` column of out/weights.csv
times rank 0's layer.dx), sync_n{1,2,3}_csv = that column as written.  Deterministic: nthread = 1
in tests/golden/config.yaml and the sync worker's exchange order is fixed.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(TESTS, "dropin", "minimpi"))
from minimpirun import launch  # noqa: E402

EXE = os.path.join(TESTS, "dropin", "_bin", "ref_main_cpu")
CONFIG = os.path.join(HERE, "config.yaml")


def main():
    out = {}
    for n in (1, 2, 3):
        with tempfile.TemporaryDirectory() as d:
            status, _ = launch(n, [EXE, CONFIG, "sync"], cwd=d, capture=True)
            assert status == 0
            rows = np.loadtxt(os.path.join(d, "out", "weights.csv"), delimiter=",", skiprows=1)
        m0 = int((rows[:, 0] == 0).sum())                       # rank 0's cells
        dx0 = (np.float32(m0 * (np.float32(1.0) / np.float32(1000.0))) - np.float32(0)) / np.float32(m0)
        col = rows[:, 2].astype(np.float32)
        out[f"sync_n{n}_csv"] = col
        out[f"sync_n{n}"] = (col * np.float32(dx0)).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_main_weights.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(v.astype(np.float64).sum()))


if __name__ == "__main__":
    main()
