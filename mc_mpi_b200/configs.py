"""The synthetic slab configurations of BASELINE.json / SURVEY.md section 8(d).

A config fixes the global slab (config.yaml keys, config.yaml:1-10) plus the
GLOBAL per-cell cross-section tables; tables equal to None mean the reference's
hard-coded ones (sigs = expf(-x_mid), absorption_rates = 0.5, src/layer.cpp:53-63).
Pure numpy: the same arrays are handed to the CUDA path and to the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

f32 = np.float32
X_INI = float(f32(np.sqrt(f32(2.0))) / f32(2.0))      # sqrtf(2)/2, src/test_layer.cpp:43
MINW_DEFAULT = float(f32(9.99999996e-13))             # config.yaml:4


@dataclass(frozen=True)
class SlabConfig:
    name: str
    nb_cells: int
    nb_particles: int
    particle_min_weight: float
    x_min: float = 0.0
    x_max: float = 1.0
    x_ini: float = X_INI
    sigs: np.ndarray | None = None              # global tables, nb_cells entries
    absorption_rates: np.ndarray | None = None
    events_per_history: float | None = None     # measured on the oracle

    def with_particles(self, n: int) -> "SlabConfig":
        return replace(self, nb_particles=int(n))


def default_sigs(nb_cells: int, x_min=0.0, x_max=1.0) -> np.ndarray:
    """sigs of a single layer spanning the slab, src/layer.cpp:56-60."""
    x_min, x_max = f32(x_min), f32(x_max)
    dx = f32(x_max - x_min) / f32(nb_cells)
    i = np.arange(nb_cells, dtype=np.float32)
    x_mid = ((x_min + i * dx).astype(np.float64) + 0.5 * float(dx)).astype(np.float32)
    # exp in double, rounded once to float.  glibc's expf is within 0.502 ulp, so
    # this equals the reference's table for the 100- and 1000-cell slabs
    # (tests/test_oracle_pin.py) but can differ by 1 ulp in <0.1 % of the
    # entries of a 1e6-cell table -- which is why a config that overrides the
    # tables always hands the SAME arrays to the CUDA path and to the oracle.
    return np.exp(-x_mid.astype(np.float64)).astype(np.float32)


def _lcg_reals(seed: int, n: int) -> np.ndarray:
    """n draws of the reference rnd_real stream (src/random.cpp:12-16), in Python ints."""
    g, c, mask = 6364136223846793005, 1442695040888963407, (1 << 63) - 1
    out = np.empty(n, dtype=np.float32)
    s = seed
    for k in range(n):
        s = (g * s + c) & mask
        out[k] = f32(s) * f32(2.0 ** -63)
    return out


def reference_default(nb_particles=100_000) -> SlabConfig:
    """cfg 1: the repository's config.yaml (`mpirun -n 5 ./main config.yaml sync`)."""
    return SlabConfig("default_slab", 1000, nb_particles, MINW_DEFAULT, events_per_history=582.8)


def ref_test_layer() -> SlabConfig:
    """src/test_layer.cpp:41-48: 100 cells, 100 particles, no weight cut-off."""
    return SlabConfig("test_layer", 100, 100, 0.0, events_per_history=60.35)


def single_gpu_slab(nb_particles=100_000_000) -> SlabConfig:
    """cfg 2: test_layer_perf / test_culayer physics (1000 cells, minw = 0) scaled to 1e8."""
    return SlabConfig("single_gpu_slab_1000", 1000, nb_particles, 0.0, events_per_history=582.1)


def absorption_dominated(nb_particles=100_000_000) -> SlabConfig:
    """cfg 2 variant: sigs x100, absorption 0.9, minw 1e-12 (E = 198.6)."""
    s = (default_sigs(1000) * f32(100.0)).astype(np.float32)
    a = np.full(1000, 0.9, dtype=np.float32)
    return SlabConfig("absorption_dominated", 1000, nb_particles, 1e-12, sigs=s,
                      absorption_rates=a, events_per_history=198.6)


def optically_thick(nb_particles=1_000_000) -> SlabConfig:
    """cfg 4: sigs x1000, absorption 0.01, minw 1e-12: 3559 events and 1756 scatters per
    history, every history ends by the weight cut-off."""
    s = (default_sigs(1000) * f32(1000.0)).astype(np.float32)
    a = np.full(1000, 0.01, dtype=np.float32)
    return SlabConfig("optically_thick", 1000, nb_particles, 1e-12, sigs=s,
                      absorption_rates=a, events_per_history=3558.7)


def heterogeneous(nb_cells=1_000_000, nb_particles=100_000) -> SlabConfig:
    """cfg 5: per-cell sigs = e^{-x_mid} (0.5 + u_i), absorption = 0.1 + 0.8 v_i with u, v
    alternating draws of the reference rnd_real stream seeded 30061994."""
    r = _lcg_reals(30061994, 2 * nb_cells)
    u, v = r[0::2], r[1::2]
    s = (default_sigs(nb_cells) * (f32(0.5) + u)).astype(np.float32)
    a = (f32(0.1) + f32(0.8) * v).astype(np.float32)
    return SlabConfig(f"heterogeneous_{nb_cells}", nb_cells, nb_particles, 1e-12, sigs=s,
                      absorption_rates=a, events_per_history=0.58 * nb_cells)


BY_NAME = {
    "default_slab": reference_default,
    "test_layer": ref_test_layer,
    "single_gpu_slab_1000": single_gpu_slab,
    "absorption_dominated": absorption_dominated,
    "optically_thick": optically_thick,
    "heterogeneous": heterogeneous,
}
