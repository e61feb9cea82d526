"""The config.yaml driver (mc_mpi_b200/main.py), SURVEY 8(f)-1: same keys, same one-line
output, same out/ files as `mpirun -n K ./main config.yaml sync` (src/main.cpp, src/worker.cpp)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CONFIG = os.path.join(HERE, "golden", "config.yaml")   # the reference's own config.yaml


def test_load_config_reads_the_reference_keys():
    from mc_mpi_b200.main import load_config, slab_config
    opt = load_config(CONFIG)
    # config.yaml:1-10
    assert opt["nb_cells"] == 1000 and opt["nb_particles"] == 100000
    assert opt["nb_particles_per_cycle"] == 500 and opt["nthread"] == 1
    assert opt["x_min"] == 0.0 and opt["x_max"] == 1.0
    assert np.float32(opt["x_ini"]) == np.float32(np.sqrt(np.float32(2.0)) / np.float32(2.0))
    assert np.float32(opt["particle_min_weight"]) == np.float32(9.99999996e-13)
    cfg = slab_config(opt)
    assert cfg.sigs is None and cfg.absorption_rates is None and cfg.nb_cells == 1000


def test_malformed_config_is_fatal_like_the_reference(tmp_path):
    from mc_mpi_b200.main import load_config
    p = tmp_path / "bad.yaml"
    p.write_text("nb_cells 1000\n")
    with pytest.raises(SystemExit):
        load_config(str(p))          # "Yaml File ... was not correctly formatted." + exit(1)
    with pytest.raises(SystemExit):
        load_config(str(tmp_path / "missing.yaml"))


def test_dump_formats(tmp_path):
    from mc_mpi_b200.main import dump_config, dump_weights_absorbed, load_config
    opt = load_config(CONFIG)
    dump_config(str(tmp_path / "config.yaml"), opt, 5)
    txt = (tmp_path / "config.yaml").read_text()
    # YamlDumper formats (src/yaml_dumper.cpp:13-24): %d, %.18e
    assert "# Read from config\nnb_cells: 1000\nx_min: 0.000000000000000000e+00\n" in txt
    assert "x_ini: 7.071067690849304199e-01\n" in txt and "world_size: 5\n" in txt
    w = np.linspace(1e-4, 2e-4, 10)
    dump_weights_absorbed(str(tmp_path / "weights.csv"), w, [0, 4, 10], np.float32(0.1))
    rows = (tmp_path / "weights.csv").read_text().splitlines()
    assert rows[0] == "proc, x, weight" and len(rows) == 11
    assert rows[1].startswith("0, ") and rows[5].startswith("1, ")     # displs switch at i == 4
    assert re.fullmatch(r"\d+, \d\.\d{18}e[+-]\d\d, \d\.\d{18}e[+-]\d\d", rows[3])


def test_dump_stats_has_the_reference_layout(tmp_path):
    """same header and field formats as the reference program's out/stats.csv
    (Worker::write_file, Timer::State::sprintf; an actual file is produced by ref_main_cpu in
    tests/test_ref_program.py)"""
    from mc_mpi_b200.main import dump_stats
    dump_stats(str(tmp_path / "stats.csv"),
               [[(10.0, 10.5, 0.25, 0.125, 0.0625, 0.0, 7)], [(10.0, 10.25, 0.1, 0.0, 0.0, 0.0, 3),
                                                               (10.25, 10.5, 0.2, 0.0, 0.0, 0.0, 4)]])
    rows = (tmp_path / "stats.csv").read_text().splitlines()
    assert rows[0] == "rank, starttime, endtime, time_comp, time_send, time_recv, time_idle, nb_cycles, "
    assert rows[1] == ("0, 1.000000000000000000e+01, 1.050000000000000000e+01, 2.500000000000000000e-01, "
                       "1.250000000000000000e-01, 6.250000000000000000e-02, 0.000000000000000000e+00, 7, ")
    assert [r.split(",")[0] for r in rows[1:]] == ["0", "1", "1"]


def test_driver_single_rank_flow_with_a_stand_in_layer(tmp_path, monkeypatch):
    """main()'s host flow for one rank -- config -> simulate(-1) -> one line + out/{config.yaml,
    weights.csv,stats.csv} + WA.out -- with the oracle standing in for the GPU layer (the real
    run is test_driver_end_to_end)"""
    sys.path.insert(0, HERE)
    import torch
    from mc_mpi_b200 import layer as mcb_layer, main as mcb_main
    from util import make_oracle

    class StandIn:
        def __init__(self, cfg):
            self.o = make_oracle(cfg, keep_border=False)

        def simulate(self, n):
            self.o.simulate(n)
            return {"track_ms": 0.0, "launches": 1}   # the counters main() reads (Layer.simulate)

        @property
        def weights_absorbed_f64(self):
            return self.o.tally_exact_f64

        def dump_WA(self, path):
            self.o.dump_WA(path)

    made = []

    def fake_decompose(x_min, x_max, x_ini, K, r, nb_cells, nb_particles, minw, **kw):
        from mc_mpi_b200 import configs
        made.append(StandIn(configs.SlabConfig("t", nb_cells, nb_particles, minw, x_min, x_max, x_ini)))
        return made[-1]

    monkeypatch.setattr(mcb_layer, "decompose_domain", fake_decompose)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.chdir(tmp_path)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    cfg_path = tmp_path / "config.yaml"
    cfg_path.write_text(open(CONFIG).read().replace("nb_particles: 100000", "nb_particles: 2000"))
    assert mcb_main.main([str(cfg_path)]) == 0
    assert sorted(os.listdir(tmp_path / "out")) == ["config.yaml", "stats.csv", "weights.csv"]
    assert (tmp_path / "WA.out").exists()
    rows = np.loadtxt(tmp_path / "out" / "weights.csv", delimiter=",", skiprows=1)
    dx = np.float32(1.0) / np.float32(1000)
    assert np.allclose(rows[:, 2] * dx, made[0].o.tally_exact_f64, rtol=1e-6, atol=0)
    stats = (tmp_path / "out" / "stats.csv").read_text().splitlines()
    assert len(stats) == 2 and stats[1].startswith("0, ") and stats[1].endswith(", 1, ")


@pytest.mark.gpu
def test_driver_end_to_end(gpu, tmp_path):
    """`python -m mc_mpi_b200.main config.yaml`: one wall-time line; out/weights.csv equals the
    oracle's tally (float view) for the reference's default workload."""
    sys.path.insert(0, HERE)
    from mc_mpi_b200 import configs
    from util import make_oracle
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-m", "mc_mpi_b200.main", CONFIG], cwd=tmp_path, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.strip().splitlines()
    assert len(lines) == 1 and float(lines[0]) > 0            # main.cpp:91
    for f in ("out/config.yaml", "out/weights.csv", "WA.out"):
        assert (tmp_path / f).is_file()
    got = np.loadtxt(tmp_path / "out" / "weights.csv", delimiter=",", skiprows=1)
    assert got.shape == (1000, 3) and np.all(got[:, 0] == 0)
    o = make_oracle(configs.reference_default(100_000))
    o.simulate(-1, nthread=os.cpu_count() or 1)
    dx = np.float32(o.dx)
    want = (o.tally_exact_f64.astype(np.float32) / dx).astype(np.float64)
    assert np.array_equal(got[:, 2], want)
    assert np.allclose(got[:, 1], float(dx) * (np.arange(1000) + 0.5), rtol=0, atol=1e-15)
    # phase timers of the native driver (Worker::write_file format, src/worker.cpp:63-181):
    # Computation is the DEVICE time of the tracking kernels (CUDA events), Idle the rest
    rows = (tmp_path / "out" / "stats.csv").read_text().splitlines()
    assert rows[0].startswith("rank, starttime, endtime, time_comp, time_send, time_recv, time_idle, nb_cycles")
    f = [v.strip() for v in rows[1].split(",")]
    start, end, comp, send, recv, idle, cycles = (float(v) for v in f[1:8])
    assert int(f[0]) == 0 and end > start
    assert 0.0 < comp <= (end - start) + 1e-6 and send == 0.0 and recv == 0.0
    assert abs((comp + idle) - (end - start)) < 1e-3 and cycles >= 1
    assert comp > 1e-4                      # 5.8e7 events take >= 0.2 ms of kernel time on a B200


@pytest.mark.gpu
def test_driver_two_ranks_device_timers(gpu, mcb_lib, tmp_path):
    """2 ranks (needs 2 GPUs): the persistent-kernel driver writes one stats row per rank whose
    Computation + Idle come from the device (kernel run time x lane occupancy / the rest), and the
    weights equal the single-layer oracle's."""
    if mcb_lib.mcb200_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, HERE)
    from mc_mpi_b200 import configs
    from util import make_oracle
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", "-m", "mc_mpi_b200.main", CONFIG, "nvl"]
    res = subprocess.run(cmd, cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    rows = (tmp_path / "out" / "stats.csv").read_text().splitlines()
    assert len(rows) == 3
    for r, row in enumerate(rows[1:]):
        f = [v.strip() for v in row.split(",")]
        start, end, comp, send, recv, idle = (float(v) for v in f[1:7])
        assert int(f[0]) == r and comp > 0 and idle >= 0 and comp <= (end - start) + 1e-6
    got = np.loadtxt(tmp_path / "out" / "weights.csv", delimiter=",", skiprows=1)
    o = make_oracle(configs.reference_default(100_000))
    o.simulate(-1, nthread=os.cpu_count() or 1)
    dx = np.float32(o.dx)
    want = (o.tally_exact_f64.astype(np.float32) / dx).astype(np.float64)
    assert np.array_equal(got[:, 2], want)
    assert sorted(set(got[:, 0])) == [0.0, 1.0]
