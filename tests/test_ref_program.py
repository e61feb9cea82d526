"""The reference's executable itself as an oracle (CPU): src/main.cpp + the three workers + the
comm / timer / stats / yaml sources, compiled UNMODIFIED over tests/dropin/minimpi (a shared-memory
subset of MPI; none is installed in the image) with the reference's own CPU Layer ->
tests/dropin/_bin/ref_main_cpu.  Pins, against the real program:

  * the oracle's float tally and the sync worker's cycle structure (K = 1, byte-exact files),
  * mc_mpi_b200.main's writers (out/config.yaml, out/weights.csv formats),
  * the K-rank oracle chain used by the multi-GPU parity tests (K = 2, 3; sync / async / rma).

The same sources linked against this repository's Layer facade (ref_main_b200) run in
tests/test_gpu_zz_ref_program.py.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "dropin", "minimpi"))

from minimpirun import launch                    # noqa: E402
from util import OracleLayer, make_oracle, oracle_chain   # noqa: E402

EXE = os.path.join(HERE, "dropin", "_bin", "ref_main_cpu")
CONFIG = os.path.join(HERE, "golden", "config.yaml")
GOLD = os.path.join(HERE, "golden", "ref_main_weights.npz")

needs_exe = pytest.mark.skipif(not os.path.isfile(EXE),
                               reason="ref_main_cpu not built (needs /root/reference at build time)")


def run_program(tmp_path, n, mode, exe=EXE, timeout=300, env=None):
    """mpirun -n n ./main config.yaml mode, in tmp_path; returns (stdout of rank 0, weights.csv rows)."""
    status, out = launch(n, [exe, CONFIG, mode], timeout=timeout, cwd=str(tmp_path), capture=True,
                         env=env)
    assert status == 0, f"{mode} on {n} ranks exited {status}"
    rows = np.loadtxt(tmp_path / "out" / "weights.csv", delimiter=",", skiprows=1)
    return out, rows


def the_config():
    from mc_mpi_b200.main import load_config, slab_config
    opt = load_config(CONFIG)
    return opt, slab_config(opt)


@needs_exe
def test_single_rank_program_equals_the_oracle_byte_for_byte(tmp_path):
    from mc_mpi_b200.main import dump_config, dump_weights_absorbed
    out, rows = run_program(tmp_path, 1, "sync")
    assert len(out.split()) == 1 and float(out) > 0          # main.cpp:91, one wall-time line
    opt, cfg = the_config()
    lay = make_oracle(cfg, keep_border=False)
    cycles = 0
    while lay.nb_disabled != cfg.nb_particles:               # worker_sync.cpp:33-119
        lay.simulate(opt["nb_particles_per_cycle"])
        cycles += 1
    assert cycles == cfg.nb_particles // opt["nb_particles_per_cycle"]
    mine = tmp_path / "mine"
    mine.mkdir()
    dump_weights_absorbed(str(mine / "weights.csv"), np.asarray(lay.weights_absorbed), [0, cfg.nb_cells],
                          np.float32(lay.dx))
    assert (mine / "weights.csv").read_bytes() == (tmp_path / "out" / "weights.csv").read_bytes()
    dump_config(str(mine / "config.yaml"), opt, 1)
    theirs = (tmp_path / "out" / "config.yaml").read_text().splitlines()
    ours = (mine / "config.yaml").read_text().splitlines()
    # buffer_size is never set by options_from_config (worker.cpp:315-332): skip that line
    keep = lambda ls: [l for l in ls if not l.startswith("buffer_size")]
    assert keep(ours) == keep(theirs)
    stats = (tmp_path / "out" / "stats.csv").read_text().splitlines()
    assert stats[0].startswith("rank, starttime, endtime, time_comp, time_send, time_recv, time_idle")
    # nb_cycles column; the terminating cycle breaks out before the counter (worker_sync.cpp:110-131)
    assert sum(int(l.split(",")[-2]) for l in stats[1:]) == cycles - 1


@needs_exe
@pytest.mark.parametrize("mode", ["sync", "async", "rma"])
@pytest.mark.parametrize("n", [2, 3])
def test_multi_rank_program_against_the_oracle_chain(tmp_path, n, mode):
    """K ranks: decompose_domain gives every rank its own dx (layer.cpp:24-33); the tally of the
    K-rank oracle chain (exact sums) is what the program's float sums must round to."""
    _, rows = run_program(tmp_path, n, mode)
    opt, cfg = the_config()
    layers, _, _ = oracle_chain(cfg, n, opt["nb_particles_per_cycle"], keep_border=False)
    exact = np.concatenate([l.tally_exact_f64 for l in layers])
    dx0 = np.float32(layers[0].dx)
    got = rows[:, 2].astype(np.float32) * dx0                 # weights.csv holds w / layer.dx of rank 0
    assert rows.shape == (cfg.nb_cells, 3)
    sizes = [l.m for l in layers]
    assert np.array_equal(rows[:, 0], np.repeat(np.arange(n), sizes))
    assert np.abs(got - exact).max() <= 2e-6 * exact.max()
    assert abs(got.astype(np.float64).sum() - exact.sum()) <= 1e-6 * exact.sum()


@pytest.mark.parametrize("name", ["comm", "state_comm", "async_serialization", "rma_serialization"])
def test_reference_mpi_ctests_with_our_particle(tmp_path, name):
    """TestComm / TestStateComm / TestAsyncSerialization / TestRmaSerialization
    (CMakeLists.txt:324-333: `mpirun -n 2`), compiled from the reference's sources with THIS
    repository's `Particle` (include/mcb200/compat/particle.hpp): the 24-byte record survives the
    reference's AsyncComm / RmaComm byte for byte (its FNV hash check), and minimpi behaves like
    the MPI those tests were written for."""
    exe = os.path.join(HERE, "dropin", "_bin", f"ref_test_{name}")
    if not os.path.isfile(exe):
        pytest.skip(f"ref_test_{name} not built (needs /root/reference at build time)")
    status, _ = launch(2, [exe], timeout=120, cwd=str(tmp_path), capture=True)
    assert status == 0


def test_golden_weights_of_the_program_are_committed():
    """tests/golden/ref_main_weights.npz (make_golden_main.py): what the GPU build of the same
    program is compared with on the box, where the reference does not exist."""
    g = np.load(GOLD)
    for k in ("sync_n1", "sync_n2", "sync_n3"):
        assert g[k].shape == (1000,) and g[k].dtype == np.float32
    assert abs(float(g["sync_n1"].astype(np.float64).sum()) - 0.3327) < 1e-3


@needs_exe
def test_golden_weights_match_a_fresh_run(tmp_path):
    g = np.load(GOLD)
    _, rows = run_program(tmp_path, 2, "sync")
    assert np.array_equal(rows[:, 2].astype(np.float32), g["sync_n2_csv"])
