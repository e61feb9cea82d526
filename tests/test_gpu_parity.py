"""Parity of the CUDA tracking path with the oracle, through the C ABI (needs a B200).

The bar (BASELINE.json north_star): escapee counts bit-exact, per-cell tally within 1e-6
relative, weight conservation to FP tolerance.  Because the device reproduces the reference's
per-particle LCG stream and libm bit for bit, these tests ask for MORE: every escapee's final
state and the fixed-point tally are compared bit for bit."""
import os

import numpy as np
import pytest

from mc_mpi_b200 import configs
from mc_mpi_b200.layer import Layer, decompose_domain, split_cells
from util import PARTICLE_DTYPE, make_oracle, oracle_chain, particles_equal, run_chain

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_layer(cfg, K=1, r=0, keep_border=True, **kw):
    return decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, K, r, cfg.nb_cells, cfg.nb_particles,
                            cfg.particle_min_weight, sigs=cfg.sigs,
                            absorption_rates=cfg.absorption_rates, keep_border=keep_border, **kw)


def assert_layer_matches_oracle(g: Layer, o, *, tol=1e-6):
    c = g.counts()
    st = o.stats()
    # escapee / dead counts: bit-exact
    assert (c["n_left"], c["n_right"], c["n_dead"]) == (st["n_left"], st["n_right"], st["n_dead"])
    assert c["events"] == st["events"] and c["scatters"] == st["scatters"]
    assert c["nb_disabled"] == o.nb_disabled
    # the tally is an exact integer sum on both sides -> all 128 bits of every cell agree
    x, lsb = g.weights_absorbed_exact()
    assert lsb == -120
    assert np.array_equal(x, o.tally_exact)
    # and what the north star states: <= 1e-6 relative per cell vs the double tally
    t64 = o.tally_f64
    nz = t64 > 0
    assert np.max(np.abs(g.weights_absorbed_f64[nz] - t64[nz]) / t64[nz]) < tol
    # weight carried out / left in the dead: exact sums too
    cw = o.class_weights_exact
    assert (c["w_left"], c["w_right"], c["w_dead"]) == (cw[0], cw[1], cw[2])
    # weight conservation
    injected = c["n_left"] + c["n_right"] + c["n_dead"]
    total = float(np.sum(g.weights_absorbed_f64)) + c["w_left"] + c["w_right"] + c["w_dead"]
    return c, total, injected


CASES = {
    "test_layer": (configs.ref_test_layer(), 1, 0),
    "default_20k": (configs.reference_default(20_000), 1, 0),
    "default_K5_r3": (configs.reference_default(20_000), 5, 3),
    "default_K8_r5": (configs.reference_default(20_000), 8, 5),
    "thick_2k": (configs.optically_thick(2_000), 1, 0),
    "thick_K4_r2": (configs.optically_thick(1_000), 4, 2),
    "absdom_20k": (configs.absorption_dominated(20_000), 1, 0),
    "hetero_4096": (configs.heterogeneous(4096, 256), 1, 0),
    # BASELINE config 5 at full width: 1e6 heterogeneous cells (the tally no longer fits shared
    # memory -> global accumulator path); ~5.8e5 events per history, so few histories
    "hetero_1e6": (configs.heterogeneous(1_000_000, 48), 1, 0),
    "hetero_20000_K3_r2": (configs.heterogeneous(20_000, 300), 3, 2),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_layer_parity(gpu, name):
    cfg, K, r = CASES[name]
    o = make_oracle(cfg, K, r)
    o.simulate(-1)
    with gpu_layer(cfg, K, r) as g:
        g.simulate(-1)
        c, total, _ = assert_layer_matches_oracle(g, o)
        left, right = g.pop_left(), g.pop_right()
        # final state of EVERY escapee, bit for bit (order is unspecified -> sort by seed)
        want_left = np.concatenate([o.particles_left, o.absorbed_left])
        want_right = np.concatenate([o.particles_right, o.absorbed_right])
        assert particles_equal(left, want_left)
        assert particles_equal(right, want_right)
        if K == 1:
            # conservation to FP tolerance: `wmc -= dw` rounds once per event in the reference's
            # float arithmetic, so the drift grows with the events per history (582 on the
            # default slab, 5.8e5 on the 1e6-cell one); the GPU total equals the oracle's
            st = o.stats()
            total_oracle = float(np.sum(o.tally_exact_f64)) + sum(o.class_weights_exact)
            assert abs(total - total_oracle) < 1e-12
            assert abs(total - 1.0) < max(1e-5, 2e-9 * st["events"] / cfg.nb_particles)


def test_golden_file_test_layer(gpu, tmp_path):
    """the reference's TestLayer on the GPU path: WA.out byte-identical to
    data/test_layer_target_WA.out (src/test_layer.cpp:58-68)."""
    cfg = configs.ref_test_layer()
    with gpu_layer(cfg) as g:
        g.simulate(-1)
        out = tmp_path / "WA.out"
        g.dump_WA(out)
        assert out.read_bytes() == open(os.path.join(GOLD, "test_layer_target_WA.out"), "rb").read()
        # and the CUDA-vs-CPU criterion of the same test (:82-91): max abs diff < 1e-6
        o = make_oracle(cfg)
        o.simulate(-1)
        assert np.max(np.abs(g.weights_absorbed - o.weights_absorbed)) < 1e-6


def test_test_culayer_criterion(gpu):
    """src/test_culayer.cu:77-84: 1000 cells, GPU vs CPU tally, per-cell abs diff <= 1e-4
    (here 2e5 particles so the sequential oracle finishes in seconds)."""
    cfg = configs.single_gpu_slab(200_000)
    o = make_oracle(cfg)
    o.simulate(-1, nthread=os.cpu_count() or 1)
    with gpu_layer(cfg) as g:
        g.simulate(-1)
        assert np.max(np.abs(g.weights_absorbed - o.weights_absorbed)) <= 1e-4
        assert_layer_matches_oracle(g, o)


@pytest.mark.parametrize("mode", [1, 2])
def test_tally_strategies_are_bit_identical(gpu, mode):
    """CTA-private shared-memory tally vs global (L2) tally: same integers."""
    cfg = configs.reference_default(20_000)
    o = make_oracle(cfg)
    o.simulate(-1)
    with gpu_layer(cfg) as g:
        g.set_option("tally_mode", mode)
        g.simulate(-1)
        assert_layer_matches_oracle(g, o)


@pytest.mark.parametrize("block,bps", [(128, 2), (256, 4), (1024, 1), (64, 1)])
def test_launch_shapes(gpu, block, bps):
    cfg = configs.reference_default(5_000)
    o = make_oracle(cfg)
    o.simulate(-1)
    with gpu_layer(cfg) as g:
        g.set_option("block", block)
        g.set_option("blocks_per_sm", bps)
        g.simulate(-1)
        assert_layer_matches_oracle(g, o)


def test_partial_simulate_and_birth_chunks(gpu):
    """simulate(nb) in pieces (the workers' nb_particles_per_cycle, config.yaml:8) and births in
    small chunks give the same result as one call."""
    cfg = configs.reference_default(10_000)
    o = make_oracle(cfg)
    o.simulate(-1)
    with gpu_layer(cfg) as g:
        g.set_option("birth_chunk", 777)
        done = 0
        while g.nb_active() > 0:
            before = g.nb_active()
            g.simulate(500)
            assert g.nb_active() == max(before - 500, 0)
            done += 1
        assert done == 20
        assert_layer_matches_oracle(g, o)


def test_clone_is_a_deep_copy(gpu):
    """Layer is copy-constructed / returned by value in the reference (src/layer.cpp:41,
    include/mcmpi/worker.hpp:60): a clone taken mid-run carries bank, outboxes, tally and the
    unborn source, and both copies then finish identically to an uninterrupted run."""
    cfg = configs.reference_default(6_000)
    o = make_oracle(cfg, 5, 3)
    o.simulate(-1)
    a = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 5, 3, cfg.nb_cells, cfg.nb_particles,
                         cfg.particle_min_weight)
    a.simulate(2_500)
    b = a.clone()
    for g in (a, b):
        g.simulate(-1)
        assert np.array_equal(g.weights_absorbed_exact()[0], o.tally_exact)
        c, st = g.counts(), o.stats()
        assert (c["n_left"], c["n_right"], c["events"]) == (st["n_left"], st["n_right"], st["events"])
        assert particles_equal(g.pop_left(), o.particles_left)
        assert particles_equal(g.pop_right(), o.particles_right)
        g.close()


def test_push_pop_roundtrip_and_edge_cases(gpu):
    cfg = configs.reference_default(1000)
    start, m = split_cells(cfg.nb_cells, 4, 1)
    with Layer(0.25, 0.5, start, m, cfg.particle_min_weight, keep_border=True,
               left_border=False, right_border=False) as g:
        # empty simulate is a no-op
        c = g.simulate(-1)
        assert c["events"] == 0 and c["nb_active"] == 0
        assert len(g.pop_left()) == 0 and len(g.pop_right()) == 0
        # particles that are already outside / dead on arrival are classified without an event
        # exactly like simulate_particle (src/layer.cpp:195-217)
        p = np.zeros(5, dtype=PARTICLE_DTYPE)
        p["seed"] = [11, 12, 13, 14, 15]
        p["x"] = 0.3
        p["mu"] = 0.5
        p["wmc"] = [1e-3, 1e-3, 1e-13, 1e-3, 1e-13]
        p["index"] = [start - 1, start + m, start + 3, start - 7, start + m]
        g.push(p)
        c = g.simulate(-1)
        assert c["events"] == 0
        # index==start-1 -> left; ==start+m -> right (even when light); light inside -> dead;
        # far outside and heavy -> left (the trailing `return -1`)
        assert (c["n_left"], c["n_right"], c["n_dead"]) == (2, 2, 1)
        left, right = g.pop_left(), g.pop_right()
        assert sorted(left["seed"].tolist()) == [11, 14]
        assert sorted(right["seed"].tolist()) == [12, 15]
        assert particles_equal(left, p[[0, 3]]) and particles_equal(right, p[[1, 4]])
        # something that is not a Monte-Carlo weight (>= 2^7) is reported, not mis-tallied
        from mc_mpi_b200 import _abi
        p["wmc"] = 1000.0
        p["index"] = start + 3
        g.push(p)
        with pytest.raises(_abi.McbError) as ei:
            g.simulate(-1)
        assert ei.value.code == _abi.ERR_RANGE


def test_chain_of_layers_equals_reference_chain(gpu):
    """5 sub-slabs with the reference's own per-layer dx, cycled like worker_sync.cpp without
    MPI (all on one GPU): per-rank tallies, counts and cycle count equal the oracle chain."""
    cfg = configs.reference_default(6000)
    K, per_cycle = 5, 500
    o_layers, o_cycles, o_mig = oracle_chain(cfg, K, per_cycle)
    layers = [decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, K, r, cfg.nb_cells,
                               cfg.nb_particles, cfg.particle_min_weight) for r in range(K)]
    cycles, mig = run_chain(layers, per_cycle, cfg.nb_particles,
                            pop_left=lambda l: l.pop_left(), pop_right=lambda l: l.pop_right(),
                            push=lambda l, p: l.push(p), simulate=lambda l, n: l.simulate(n),
                            disabled=lambda l: l.nb_disabled)
    # the number of migrations is a property of the trajectories; the number of cycles is not
    # (which banked particles a layer picks first depends on the outbox order, unspecified here)
    assert mig == o_mig
    assert abs(cycles - o_cycles) <= max(4, o_cycles // 4)
    for g, o in zip(layers, o_layers):
        assert np.array_equal(g.weights_absorbed_exact()[0], o.tally_exact)
        assert g.nb_disabled == o.nb_disabled
        assert g.counts()["events"] == o.stats()["events"]
        g.close()


def test_global_dx_decomposition_equals_single_layer(gpu):
    """SURVEY hard part 3: with the ONE global dx, K sub-slabs reproduce the single-layer
    trajectories bit for bit (the reference's own K-rank runs do not)."""
    cfg = configs.reference_default(8000)
    with gpu_layer(cfg) as one:
        one.simulate(-1)
        q1, _ = one.weights_absorbed_exact()
        c1 = one.counts()
    for K in (2, 8):
        layers = [decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, K, r, cfg.nb_cells,
                                   cfg.nb_particles, cfg.particle_min_weight, global_dx=True)
                  for r in range(K)]
        run_chain(layers, 4000, cfg.nb_particles,
                  pop_left=lambda l: l.pop_left(), pop_right=lambda l: l.pop_right(),
                  push=lambda l, p: l.push(p), simulate=lambda l, n: l.simulate(n),
                  disabled=lambda l: l.nb_disabled)
        qK = np.concatenate([l.weights_absorbed_exact()[0] for l in layers])
        assert np.array_equal(qK, q1)
        assert sum(l.counts()["events"] for l in layers) == c1["events"]
        assert sum(l.counts()["scatters"] for l in layers) == c1["scatters"]
        assert layers[0].counts()["n_left"] == c1["n_left"]
        assert layers[-1].counts()["n_right"] == c1["n_right"]
        for l in layers:
            l.close()


@pytest.mark.parametrize("K,per_cycle", [(3, 3000), (8, 100_000)])
def test_direct_peer_exchange_on_one_gpu(gpu, K, per_cycle):
    """The fused path: each layer's tracking kernel stores its escapees straight into the
    neighbour layer's inbox (here all layers live on one GPU and are connected locally; across
    GPUs the same pointers are CUDA-IPC mappings over NVLink).  Double-buffered by parity,
    ingested after the cycle.  Must reproduce the single-layer run bit for bit."""
    cfg = configs.reference_default(20_000)
    with gpu_layer(cfg, keep_border=False) as one:
        one.simulate(-1)
        q1, _ = one.weights_absorbed_exact()
        c1 = one.counts()
    layers = [decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, K, r, cfg.nb_cells,
                               cfg.nb_particles, cfg.particle_min_weight, global_dx=True)
              for r in range(K)]
    for l in layers:
        l.inbox_create(per_cycle)
    for r, l in enumerate(layers):
        if r > 0:
            l.connect_local(0, layers[r - 1])
        if r + 1 < K:
            l.connect_local(1, layers[r + 1])
    cycles = 0
    while sum(l.nb_disabled for l in layers) < cfg.nb_particles:
        parity = cycles & 1
        for l in layers:
            l.set_exchange_parity(parity)
            l.simulate(per_cycle)                   # at most one launch: the inbox bounds it
        received = 0
        for r, l in enumerate(layers):
            if r > 0:
                received += l.ingest_inbox(0, parity)
            if r + 1 < K:
                received += l.ingest_inbox(1, parity)
        cycles += 1
        assert cycles < 10_000
    assert all(l.counts()["n_outbox_left"] == 0 and l.counts()["n_outbox_right"] == 0 for l in layers)
    qK = np.concatenate([l.weights_absorbed_exact()[0] for l in layers])
    assert np.array_equal(qK, q1)
    assert sum(l.counts()["events"] for l in layers) == c1["events"]
    assert sum(l.counts()["scatters"] for l in layers) == c1["scatters"]
    assert layers[0].counts()["n_left"] == c1["n_left"]
    assert layers[-1].counts()["n_right"] == c1["n_right"]
    for l in layers:
        l.disconnect_peers()
    for l in layers:
        l.close()


def test_full_size_properties(gpu):
    """BASELINE config 2 at full size (1e8 histories, 1000 cells): no oracle can follow, so
    size-independent properties: conservation, counts add up, the tally profile agrees with a
    1e6-history oracle run within Monte-Carlo noise, and the run is reproducible bit for bit."""
    cfg = configs.single_gpu_slab(100_000_000)
    qs = []
    for rep in range(2):
        with gpu_layer(cfg, keep_border=False) as g:
            if rep:
                g.set_option("block", 512)      # another schedule, same integers
                g.set_option("blocks_per_sm", 2)
            g.simulate(-1)
            c = g.counts()
            qs.append(g.weights_absorbed_exact()[0])
            assert c["n_left"] + c["n_right"] + c["n_dead"] == cfg.nb_particles
            assert c["nb_disabled"] == cfg.nb_particles and c["nb_active"] == 0
            total = float(np.sum(g.weights_absorbed_f64)) + c["w_left"] + c["w_right"] + c["w_dead"]
            assert abs(total - 1.0) < 1e-5
            assert abs(c["events"] / cfg.nb_particles - 582.1) < 0.3
            prof = g.weights_absorbed_f64
    assert np.array_equal(qs[0], qs[1])
    o = make_oracle(configs.single_gpu_slab(1_000_000))
    o.simulate(-1, nthread=os.cpu_count() or 1)
    rel = np.abs(prof - o.tally_f64) / o.tally_f64
    assert rel.max() < 0.02 and rel.mean() < 0.004   # ~1/sqrt(histories per cell) noise


def test_simulate_host_pipeline_equals_oracle(gpu, mcb_lib):
    """mcb200_layer_simulate_host (host particles copied in chunks under the tracking of the
    chunk before) gives what push + simulate gives: the oracle's result, bit for bit; the
    tally can be zeroed between runs (mcb200_layer_reset_tally)."""
    import ctypes as C
    cfg = configs.reference_default(60_000)
    o = make_oracle(cfg)
    o.simulate(-1)
    host = np.zeros(cfg.nb_particles, dtype=PARTICLE_DTYPE)
    dx = np.float32(np.float32(cfg.x_max - cfg.x_min) / np.float32(cfg.nb_cells))
    assert mcb_lib.mcb200_test_birth(0, cfg.x_ini, float(np.float32(1.0 / cfg.nb_particles)),
                                     float(dx), cfg.nb_particles, 5127801, host.ctypes.data) == 0
    with Layer(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight, keep_border=True) as g:
        for rep in range(2):
            # three chunk shapes: one chunk, several growing ones, a ragged tail
            g.set_option("host_chunk", (1 << 25, 1 << 14)[rep])
            g.simulate_host(host)
            x, _ = g.weights_absorbed_exact()
            assert np.array_equal(x, o.tally_exact)
            left, right = g.pop_left(), g.pop_right()
            assert particles_equal(left, np.concatenate([o.particles_left, o.absorbed_left]))
            assert particles_equal(right, np.concatenate([o.particles_right, o.absorbed_right]))
            g.reset_tally()
            assert not g.weights_absorbed_f64.any()
        c = g.counts()
        assert c["events"] == 2 * o.stats()["events"] and c["nb_active"] == 0
