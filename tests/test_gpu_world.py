"""The persistent exchange kernel behind mcb200_world_* (needs a B200).

A run on K ranks x V windows must reproduce the SINGLE-layer oracle bit for bit: all 128 bits
of every cell of the tally, events, scatters, histories absorbed at the global borders, dead.
On a one-GPU box the ranks share the device (max_ctas leaves room for all their kernels), so
the whole protocol -- rings, credits, back-pressure, banks, in-kernel births, the device-side
global count, the `done` broadcast -- runs exactly as it does across GPUs, minus NVLink."""
import numpy as np
import pytest

from mc_mpi_b200 import _abi, configs
from mc_mpi_b200.worker import LocalBox, _Rank, totals
from util import make_oracle

pytestmark = pytest.mark.gpu

_ORACLE = {}


def oracle_run(cfg):
    """the single-layer oracle result of a config (cached per test session)"""
    key = (cfg.name, cfg.nb_cells, cfg.nb_particles)
    if key not in _ORACLE:
        import os
        o = make_oracle(cfg)
        o.simulate(-1, nthread=os.cpu_count() or 1)
        st = o.stats()
        _ORACLE[key] = dict(st=st, exact=o.tally_exact.copy(), f64=o.tally_f64.copy(),
                            cw=tuple(o.class_weights_exact))
        o.free()
    return _ORACLE[key]


def assert_world_matches_oracle(box: LocalBox, results, cfg, runs=1):
    want = oracle_run(cfg)
    t = totals(results)
    st = want["st"]
    for r in results:
        assert r["error"] == 0
    assert (t["n_left"], t["n_right"], t["n_dead"]) == (st["n_left"], st["n_right"], st["n_dead"])
    assert t["events"] == st["events"] and t["scatters"] == st["scatters"]
    assert t["births"] == cfg.nb_particles
    if runs == 1:
        x = box.gather_weights_absorbed_exact()
        assert np.array_equal(x, want["exact"]), "exact tally differs from the oracle"
        # absorbed at a global border: ONE rank's exact sum, rounded once like the oracle's;
        # the dead are spread over the ranks (exact partial sums, one rounding each)
        assert (t["w_left"], t["w_right"]) == want["cw"][:2]
        assert abs(t["w_dead"] - want["cw"][2]) <= 1e-12 * max(want["cw"][2], 1e-300)
    wa = box.gather_weights_absorbed()
    nz = want["f64"] > 0
    assert np.max(np.abs(wa[nz] / runs - want["f64"][nz]) / want["f64"][nz]) < 1e-6
    return t


def safe(box: LocalBox, ms=60_000):
    box.set_option("max_run_ms", ms)      # never hang the GPU box, whatever goes wrong
    box.set_option("stall_ms", 10_000)
    return box


CASES = {
    # name: (config, ranks, options)
    "one_rank": (configs.reference_default(20_000), 1, {}),
    "test_layer": (configs.ref_test_layer(), 1, {}),
    "windows3": (configs.reference_default(20_000), 1, dict(windows=3)),
    "windows7_tiny_rings": (configs.reference_default(20_000), 1,
                            dict(windows=7, ring_cap=32, max_ctas=28, bank_cap=1 << 15)),
    "ranks3": (configs.reference_default(20_000), 3, dict(max_ctas=148)),
    "ranks4_tiny_rings": (configs.reference_default(30_000), 4,
                          dict(max_ctas=16, ring_cap=32, bank_cap=1 << 15)),
    "ranks2_windows2": (configs.reference_default(20_000), 2, dict(max_ctas=200, windows=2)),
    "ranks8": (configs.reference_default(20_000), 8, dict(max_ctas=64)),
    "thick_ranks3": (configs.optically_thick(2_000), 3, dict(max_ctas=148)),
    "absdom_ranks2": (configs.absorption_dominated(20_000), 2, dict(max_ctas=148)),
    "test_layer_ranks5": (configs.ref_test_layer(), 5, dict(max_ctas=32)),
    "hetero_8192_auto_windows": (configs.heterogeneous(8192, 256), 1, {}),
    "hetero_8192_xs_global": (configs.heterogeneous(8192, 256), 1, dict(xs_global=1)),
    "ranks3_xs_global": (configs.reference_default(20_000), 3, dict(max_ctas=148, xs_global=1)),
    "hetero_20000_ranks2": (configs.heterogeneous(20_000, 200), 2, dict(max_ctas=256)),
    # BASELINE config 5 at full width: 1e6 heterogeneous cells in ~590 windows of one CTA each
    "hetero_1e6_windows": (configs.heterogeneous(1_000_000, 64), 1, {}),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_world_parity(gpu, name):
    cfg, K, opts = CASES[name]
    with safe(LocalBox(cfg, K, **opts)) as box:
        res = box.run()
        t = assert_world_matches_oracle(box, res, cfg)
        if K > 1:
            assert t["sent_left"] + t["sent_right"] > 0        # the rank boundary was crossed
        if opts.get("windows", 0) > 1:
            assert t["window_crossings"] > 0


def test_uneven_cuts_give_the_same_bits(gpu):
    cfg = configs.reference_default(20_000)
    with safe(LocalBox(cfg, 3, cuts=[0, 650, 720, 1000], max_ctas=148)) as box:
        assert_world_matches_oracle(box, box.run(), cfg)


def test_backpressure_blocks_and_recovers(gpu):
    """a fast producer in front of a slow consumer with tiny rings: senders block, drain their
    own inbound stripes into the bank, and the run still reproduces the oracle bit for bit"""
    cfg = configs.reference_default(50_000)
    # the source (cell 707) sits 7 cells from rank 0, which has 700 cells of work per visitor
    with safe(LocalBox(cfg, 2, cuts=[0, 700, 1000], max_ctas=32, ring_cap=32,
                       bank_cap=1 << 16)) as box:
        res = box.run()
        t = assert_world_matches_oracle(box, res, cfg)
        assert t["blocked_passes"] > 0
        assert t["bank_pushes"] == t["bank_pops"]


def test_repeated_runs_accumulate(gpu):
    """the tally is cumulative over runs (like Layer::weights_absorbed over simulate calls), the
    counters are per run; rings and banks carry over cleanly"""
    cfg = configs.reference_default(10_000)
    with safe(LocalBox(cfg, 2, max_ctas=148, windows=2, ring_cap=32)) as box:
        for _ in range(3):
            res = box.run()
            assert totals(res)["births"] == cfg.nb_particles
        assert_world_matches_oracle(box, res, cfg, runs=3)
        box.reset_tally()
        assert_world_matches_oracle(box, box.run(), cfg)


def test_wide_slab_runs_in_windows(gpu):
    """BASELINE config 5 shape: more cells than one CTA-private tally holds -> windows, chosen
    automatically, every window with its tally in shared memory"""
    cfg = configs.heterogeneous(65_536, 64)
    with safe(LocalBox(cfg, 1), ms=120_000) as box:
        res = box.run()
        assert res[0]["windows"] > 1
        assert_world_matches_oracle(box, res, cfg)


def test_source_outside_the_slab_is_an_empty_run(gpu):
    cfg = configs.reference_default(1000)
    from dataclasses import replace
    cfg = replace(cfg, x_ini=1.5)
    with safe(LocalBox(cfg, 2, max_ctas=148)) as box:
        res = box.run()
        assert totals(res)["events"] == 0 and totals(res)["births"] == 0


def test_a_missing_peer_is_a_timeout_not_a_hang(gpu):
    """only the home rank is launched: its escapees fill the neighbour's rings, nothing
    moves any more, the host-side watch stops the run with MCB200_ERR_TIMEOUT"""
    cfg = configs.reference_default(200_000)
    with LocalBox(cfg, 2, max_ctas=64, ring_cap=32, bank_cap=1 << 12, inflight_limit=1 << 14) as box:
        box.set_option("stall_ms", 500)
        box.set_option("max_run_ms", 20_000)
        home = box.ranks[1]   # cell 707 of 1000 lives on rank 1 of 2
        for r in box.ranks:
            r.prepare(cfg.nb_particles)
        home.launch()
        with pytest.raises(_abi.McbError) as e:
            home.wait()
        assert e.value.code == _abi.ERR_TIMEOUT
        # the world is usable again afterwards
        box.set_option("stall_ms", 10_000)
        small = configs.reference_default(20_000)
        res = box.run(small.nb_particles)
        want = oracle_run(small)
        assert totals(res)["events"] == want["st"]["events"]


def test_geometry_mismatch_is_rejected(gpu):
    cfg = configs.reference_default(1000)
    a = _Rank(cfg, 0, 2, 0, None, max_ctas=64)
    b = _Rank(cfg, 1, 2, 0, None, max_ctas=128)
    try:
        with pytest.raises(_abi.McbError) as e:
            a.connect_local(b)
        assert e.value.code == _abi.ERR_INVALID
    finally:
        a.close()
        b.close()


def test_full_size_closure(gpu):
    """2e7 histories (1.2e10 events) through 2 ranks x 2 windows sharing the GPU: every history
    ends at a global border, the global count closes, weight is conserved"""
    cfg = configs.single_gpu_slab(20_000_000)
    with safe(LocalBox(cfg, 2, max_ctas=280, windows=2), ms=120_000) as box:
        res = box.run()
        t = totals(res)
        assert t["n_left"] + t["n_right"] + t["n_dead"] == cfg.nb_particles
        assert t["births"] == cfg.nb_particles
        total = float(box.gather_weights_absorbed().sum()) + t["w_left"] + t["w_right"] + t["w_dead"]
        assert abs(total - 1.0) < 1e-5
        assert abs(t["events"] / cfg.nb_particles - 582.15) < 0.5
