"""Drop-in boundary: the reference's OWN test programs, compiled unmodified from
/root/reference against this repository's surfaces (tests/dropin/Makefile, built by
__graft_entry__.build() where the reference sources exist), run on the GPU.

  ref_test_layer      = src/test_layer.cpp      + our `Layer` facade (include/mcb200/compat)
  ref_test_culayer    = src/test_culayer.cu + the reference's CPU layer.cpp + OUR `cusimulate`
  ref_test_layer_perf = src/test_layer_perf.cpp + our `Layer` facade
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "dropin", "_bin")
GOLD = os.path.join(HERE, "golden")


def _need(name):
    path = os.path.join(BIN, name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} not built (needs the reference sources at build time)")
    return path


def test_reference_TestLayer_on_our_layer(gpu, tmp_path):
    """TestLayer (src/test_layer.cpp:39-95): simulate(-1), dump_WA(), byte-compare WA.out with
    ../data/test_layer_target_WA.out; the exit code is the verdict."""
    exe = _need("ref_test_layer")
    (tmp_path / "data").mkdir()
    (tmp_path / "run").mkdir()
    shutil.copy(os.path.join(GOLD, "test_layer_target_WA.out"), tmp_path / "data")
    res = subprocess.run([exe], cwd=tmp_path / "run", stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout
    assert "ERROR" not in res.stdout


def test_reference_TestCuLayer_with_our_cusimulate(gpu, tmp_path):
    """TestCuLayer (src/test_culayer.cu:27-86): 1000 cells, 1e6 particles, the reference's CPU
    Layer::simulate vs cusimulate, per-cell abs diff <= 1e-4; prints both wall times."""
    exe = _need("ref_test_culayer")
    res = subprocess.run([exe], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout
    assert "CPU =" in res.stdout and "GPU =" in res.stdout
    print(res.stdout)


def test_reference_test_layer_perf_on_our_layer(gpu, tmp_path):
    """the reference's perf harness (src/test_layer_perf.cpp): 1000 cells, 1e6 particles; prints
    seconds and leaves WA.out, compared here with the oracle's tally."""
    exe = _need("ref_test_layer_perf")
    res = subprocess.run([exe, "1"], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout
    seconds = float(res.stdout.strip().splitlines()[-1])
    assert 0 < seconds < 60
    wa = np.loadtxt(tmp_path / "WA.out")
    gold = np.loadtxt(os.path.join(GOLD, "WA_1000_1000000.out"))
    rel = np.abs(wa[:, 1] - gold[:, 1]) / gold[:, 1]
    assert wa.shape == (1000, 2) and rel.max() < 2e-3   # 3 significant digits in WA.out


def test_whole_run_from_plain_c(gpu, tmp_path):
    """tests/dropin/world_native.c: K ranks created, connected, run (Worker::spin) and gathered
    (Worker::gather_weights_absorbed) from C99 through include/mcb200.h only -- no Python, no
    torch on the path.  Its output equals the single-layer oracle: counts exact, tally equal to
    the exact sums rounded once."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from mc_mpi_b200 import configs
    from util import make_oracle
    exe = _need("world_native")
    n = 30_000
    res = subprocess.run([exe, "3", str(n)], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.split("\n")
    counts = tuple(int(v) for v in lines[0].split())
    tally = np.array([float(v) for v in lines[1:1001]])
    o = make_oracle(configs.reference_default(n))
    o.simulate(-1, nthread=os.cpu_count() or 1)
    st = o.stats()
    assert counts == (st["events"], st["scatters"], st["n_left"], st["n_right"], st["n_dead"])
    assert np.array_equal(tally, o.tally_exact_f64)
