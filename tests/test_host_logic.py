"""Host-side logic that needs no GPU: the decomposition arithmetic, the cut helpers of the
multi-GPU driver, the slab configurations."""
import numpy as np
import pytest

from mc_mpi_b200 import configs
from mc_mpi_b200.layer import split_cells
from mc_mpi_b200.world import balanced_cuts, equal_cuts
from oracle.pyoracle import OracleLayer


@pytest.mark.parametrize("nb_cells,K", [(1000, 1), (1000, 5), (1000, 8), (1003, 7), (17, 16), (100, 3)])
def test_split_cells_is_the_reference_decomposition(nb_cells, K):
    # src/layer.cpp:24-27 through the pinned oracle
    total = 0
    for r in range(K):
        start, m = split_cells(nb_cells, K, r)
        o = OracleLayer.decompose_domain(0.0, 1.0, configs.X_INI, K, r, nb_cells, 10, 0.0)
        assert (start, m) == (o.index_start, o.m)
        total += m
    assert total == nb_cells
    cuts = equal_cuts(nb_cells, K)
    assert cuts[0] == 0 and cuts[-1] == nb_cells and len(cuts) == K + 1
    assert all(cuts[r] == split_cells(nb_cells, K, r)[0] for r in range(K))


def test_balanced_cuts_equalise_a_known_density():
    # cost density 2x higher in the right half: the cuts must crowd there
    cuts = equal_cuts(1000, 4)
    cost = [1.0, 1.0, 2.0, 2.0]
    new = balanced_cuts(cuts, cost, 1000)
    assert new[0] == 0 and new[-1] == 1000 and all(b > a for a, b in zip(new, new[1:]))
    dens = np.concatenate([np.full(250, c / 250) for c in cost])
    per_rank = [dens[new[r]:new[r + 1]].sum() for r in range(4)]
    assert max(per_rank) / min(per_rank) < 1.02
    # already balanced -> unchanged; minimum width is respected under absurd costs
    assert balanced_cuts(cuts, [1, 1, 1, 1], 1000) == cuts
    tiny = balanced_cuts(cuts, [1e9, 1, 1, 1], 1000, min_cells=8)
    assert all(b - a >= 8 for a, b in zip(tiny, tiny[1:]))


def test_configs_are_what_baseline_says():
    c = configs.reference_default()
    assert (c.nb_cells, c.nb_particles) == (1000, 100_000)            # config.yaml:5-6
    assert np.float32(c.particle_min_weight) == np.float32(9.99999996e-13)
    t = configs.optically_thick(10)
    assert np.allclose(t.sigs, configs.default_sigs(1000) * 1000) and np.all(t.absorption_rates == np.float32(0.01))
    a = configs.absorption_dominated(10)
    assert np.all(a.absorption_rates == np.float32(0.9))
    h = configs.heterogeneous(2048, 10)
    assert h.sigs.shape == (2048,) and 0.1 <= h.absorption_rates.min() and h.absorption_rates.max() <= 0.9
    # the table generator is the reference's rnd_real stream seeded 30061994
    from oracle import pyoracle
    _, r = pyoracle.rnd_real_stream(30061994, 8)
    assert np.array_equal(configs._lcg_reals(30061994, 8).view(np.uint32), r.view(np.uint32))
