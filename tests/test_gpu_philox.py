"""The counter-based generator ("rng" = 1, Philox2x32-10): known answers, and the STATISTICAL
acceptance gate of SURVEY 8(d) -- its streams differ from the reference's LCG, so the tally is
compared with the oracle's through a per-cell z-score (variance estimated from independent
oracle batches), |z| < 5, and the left / right escape fractions within 4 sigma (binomial)."""
import numpy as np
import pytest

from mc_mpi_b200 import configs
from mc_mpi_b200.layer import Layer
from mc_mpi_b200.worker import LocalBox, totals
from util import OracleLayer, apply_tables

pytestmark = pytest.mark.gpu


def test_philox_known_answers(gpu, mcb_lib):
    # Random123 kat_vectors, philox2x32 10 rounds: counter, key -> output
    kat = [((0x00000000, 0x00000000), 0x00000000, (0xff1dae59, 0x6cd10df2)),
           ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
           ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]
    c0 = np.array([k[0][0] for k in kat], dtype=np.uint32)
    c1 = np.array([k[0][1] for k in kat], dtype=np.uint32)
    key = np.array([k[1] for k in kat], dtype=np.uint32)
    out = np.zeros((len(kat), 2), dtype=np.uint32)
    assert mcb_lib.mcb200_test_philox(0, c0.ctypes.data, c1.ctypes.data, key.ctypes.data,
                                      out.ctypes.data, len(kat)) == 0
    assert [tuple(int(v) for v in row) for row in out] == [k[2] for k in kat]


def oracle_batches(cfg, batches):
    """the oracle's tally of `cfg` as B independent batches (seed chains started from different
    seeds, each history with weight 1 / nb_particles): per-cell total and the variance of it"""
    n_b = cfg.nb_particles // batches
    per = []
    n_lr = np.zeros(2)
    for b in range(batches):
        o = OracleLayer.new(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight)
        apply_tables(o, cfg)
        o.set_keep_border(True)
        o.create_particles(cfg.x_ini, float(np.float32(1.0 / cfg.nb_particles)), n_b, 5127801 + 7919 * b)
        o.simulate(-1)
        per.append(o.tally_f64.copy())
        st = o.stats()
        n_lr += (st["n_left"], st["n_right"])
        o.free()
    per = np.array(per)
    total = per.sum(axis=0)
    var_total = batches * per.var(axis=0, ddof=1)   # variance of a sum of B i.i.d. batch tallies
    return total, var_total, n_lr / (n_b * batches)


def gate(tally, frac_lr, cfg, batches=16):
    total, var, p_lr = oracle_batches(cfg, batches)
    ok = var > 0
    # both sides are noisy estimates of the same expectation: Var(diff) = 2 Var(total)
    z = (tally[ok] - total[ok]) / np.sqrt(2.0 * var[ok])
    assert np.max(np.abs(z)) < 5.0, f"max |z| = {np.max(np.abs(z)):.2f}"
    assert abs(np.mean(z)) < 5.0 / np.sqrt(ok.sum()) * 3      # no systematic offset over the cells
    n = cfg.nb_particles
    for k in range(2):
        sigma = np.sqrt(2.0 * p_lr[k] * (1.0 - p_lr[k]) / n)
        assert abs(frac_lr[k] - p_lr[k]) <= 4.0 * sigma + 1e-12


@pytest.mark.parametrize("name", ["default_slab", "absorption_dominated"])
def test_philox_layer_statistical_gate(gpu, name):
    cfg = configs.BY_NAME[name]().with_particles(160_000)
    with Layer(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight, sigs=cfg.sigs,
               absorption_rates=cfg.absorption_rates) as g:
        g.set_option("rng", 1)
        g.create_particles(cfg.x_ini, float(np.float32(1.0 / cfg.nb_particles)), cfg.nb_particles, 12345)
        c = g.simulate(-1)
        assert c["nb_active"] == 0
        assert c["n_left"] + c["n_right"] + c["n_dead"] == cfg.nb_particles
        total = float(g.weights_absorbed_f64.sum()) + c["w_left"] + c["w_right"] + c["w_dead"]
        assert abs(total - 1.0) < 1e-5                       # weight conservation
        gate(g.weights_absorbed_f64, (c["n_left"] / cfg.nb_particles, c["n_right"] / cfg.nb_particles), cfg)
        # counter-based: a second run with the same key reproduces the first bit for bit,
        # whatever the launch shape
        x1, _ = g.weights_absorbed_exact()
    with Layer(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight, sigs=cfg.sigs,
               absorption_rates=cfg.absorption_rates) as g2:
        g2.set_option("rng", 1)
        g2.set_option("block", 512)
        g2.set_option("birth_chunk", 50_000)
        g2.create_particles(cfg.x_ini, float(np.float32(1.0 / cfg.nb_particles)), cfg.nb_particles, 12345)
        g2.simulate(-1)
        x2, _ = g2.weights_absorbed_exact()
        assert np.array_equal(x1, x2)


def test_philox_world_equals_philox_layer(gpu):
    """the persistent multi-rank kernel in Philox mode: 3 ranks x 2 windows reproduce the
    single-layer Philox run bit for bit (the generator is a pure function of history id and
    event number), and pass the statistical gate against the oracle"""
    cfg = configs.reference_default(160_000)
    with Layer(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight) as g:
        g.set_option("rng", 1)
        g.create_particles(cfg.x_ini, float(np.float32(1.0 / cfg.nb_particles)), cfg.nb_particles, 5127801)
        c = g.simulate(-1)
        x1, _ = g.weights_absorbed_exact()
    with LocalBox(cfg, 3, max_ctas=148, windows=2) as box:
        box.set_option("max_run_ms", 60_000)
        box.set_option("rng", 1)
        res = box.run()
        t = totals(res)
        assert (t["n_left"], t["n_right"], t["n_dead"]) == (c["n_left"], c["n_right"], c["n_dead"])
        assert t["events"] == c["events"] and t["scatters"] == c["scatters"]
        assert np.array_equal(box.gather_weights_absorbed_exact(), x1)
        gate(box.gather_weights_absorbed(), (t["n_left"] / cfg.nb_particles, t["n_right"] / cfg.nb_particles), cfg)
