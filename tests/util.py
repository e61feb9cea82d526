"""Shared helpers of the parity tests (checker side: may use the oracle)."""
from __future__ import annotations

import numpy as np

from oracle.pyoracle import PARTICLE_DTYPE, OracleLayer, RefLayer  # noqa: F401


def sorted_particles(p: np.ndarray) -> np.ndarray:
    """canonical order for comparing particle multisets (outbox order is unspecified)."""
    p = np.ascontiguousarray(p, dtype=PARTICLE_DTYPE)
    return p[np.argsort(p["seed"], kind="stable")] if len(p) else p


def particles_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """bit-for-bit equality of two particle multisets."""
    a, b = sorted_particles(a), sorted_particles(b)
    return len(a) == len(b) and a.tobytes() == b.tobytes()


def apply_tables(layer, cfg, start=0):
    """overwrite the public sigs / absorption_rates vectors of an oracle / ref layer with the
    config's global tables (the way BASELINE configs 4-5 are expressed, layer.hpp:103-104)."""
    m = layer.m
    if cfg.sigs is not None:
        layer.sigs[:] = cfg.sigs[start:start + m]
    if cfg.absorption_rates is not None:
        layer.absorption_rates[:] = cfg.absorption_rates[start:start + m]


def make_oracle(cfg, world_size=1, world_rank=0, *, keep_border=True, cls=OracleLayer):
    from mc_mpi_b200.layer import split_cells
    lay = cls.decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, world_size, world_rank,
                               cfg.nb_cells, cfg.nb_particles, cfg.particle_min_weight)
    start, _ = split_cells(cfg.nb_cells, world_size, world_rank)
    apply_tables(lay, cfg, start)
    if cls is OracleLayer:
        lay.set_keep_border(keep_border)
    return lay


def run_chain(layers, nb_per_cycle, total, *, pop_left, pop_right, push, simulate, disabled,
              max_cycles=1_000_000):
    """The sync worker's loop without MPI (src/worker_sync.cpp:24-135): every layer simulates
    nb_per_cycle, escapees move to the neighbours, stop when all particles are disabled.
    Returns the number of cycles and the number of particle migrations."""
    K = len(layers)
    migrations = 0
    for cycle in range(max_cycles):
        for l in layers:
            simulate(l, nb_per_cycle)
        moved = []
        for r, l in enumerate(layers):
            moved.append((pop_left(l), pop_right(l)))
        for r, (pl, pr) in enumerate(moved):
            if r > 0 and len(pl):
                push(layers[r - 1], pl)
                migrations += len(pl)
            if r + 1 < K and len(pr):
                push(layers[r + 1], pr)
                migrations += len(pr)
        if sum(disabled(l) for l in layers) == total:
            return cycle + 1, migrations
    raise AssertionError("chain did not terminate")


def oracle_chain(cfg, K, nb_per_cycle, **kw):
    layers = [make_oracle(cfg, K, r, **kw) for r in range(K)]

    def pop(l, side):
        arr = (l.particles_left if side == 0 else l.particles_right).copy()
        (l.clear_left if side == 0 else l.clear_right)()
        return arr

    cycles, mig = run_chain(
        layers, nb_per_cycle, cfg.nb_particles,
        pop_left=lambda l: pop(l, 0), pop_right=lambda l: pop(l, 1),
        push=lambda l, p: l.push(p), simulate=lambda l, n: l.simulate(n),
        disabled=lambda l: l.nb_disabled)
    return layers, cycles, mig
