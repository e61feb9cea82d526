"""The reference's executable on the B200 Layer: src/main.cpp + worker_{sync,async,rma}.cpp + its
comm / timer / stats / yaml sources, compiled UNMODIFIED against include/mcb200/compat and linked
with libmcb200.so (tests/dropin/Makefile: ref_main_b200), launched as K ranks over
tests/dropin/minimpi (no MPI in the image).  Ranks take GPU `LOCAL_RANK % device_count`, so K ranks
also run on a one-GPU box.  Compared with tests/golden/ref_main_weights.npz = the same program
with the reference's CPU Layer (tests/golden/make_golden_main.py, tests/test_ref_program.py).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "dropin", "minimpi"))
from minimpirun import launch  # noqa: E402

EXE = os.path.join(HERE, "dropin", "_bin", "ref_main_b200")
CONFIG = os.path.join(HERE, "golden", "config.yaml")
GOLD = os.path.join(HERE, "golden", "ref_main_weights.npz")


@pytest.mark.parametrize("n,mode", [(1, "sync"), (2, "sync"), (3, "sync"), (2, "async"), (2, "rma")])
def test_reference_program_on_our_layer(gpu, tmp_path, n, mode):
    if not os.path.isfile(EXE):
        pytest.skip("ref_main_b200 not built (needs the reference sources at build time)")
    status, out = launch(n, [EXE, CONFIG, mode], timeout=300, cwd=str(tmp_path), capture=True)
    assert status == 0, f"{mode} on {n} ranks exited {status}"
    assert len(out.split()) == 1 and float(out) > 0            # main.cpp:91
    rows = np.loadtxt(tmp_path / "out" / "weights.csv", delimiter=",", skiprows=1)
    g = np.load(GOLD)
    want = g[f"sync_n{n}_csv"].astype(np.float64)              # the CPU program's float sums
    assert rows.shape == (want.size, 3)
    # the device tally is exact, the CPU one rounds ~100 float additions per cell: 2e-6 of the
    # largest cell (the bound tests/test_ref_program.py holds the CPU program to, against the
    # oracle's exact sums), and 1e-6 on the total
    assert np.abs(rows[:, 2] - want).max() <= 4e-6 * want.max()
    assert abs(rows[:, 2].sum() - want.sum()) <= 2e-6 * want.sum()
    cfg = (tmp_path / "out" / "config.yaml").read_text()
    assert f"world_size: {n}\n" in cfg and "nb_cells: 1000\n" in cfg
    assert (tmp_path / "out" / "stats.csv").read_text().startswith("rank, starttime, endtime")
