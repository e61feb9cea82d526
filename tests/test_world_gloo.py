"""Host logic of the multi-GPU driver (mc_mpi_b200/world.py) on CPU: world_size-2 and -3
`gloo` process groups, with the ORACLE standing in for the GPU layer (checker-side adapter),
so that exchange, bookkeeping and termination are exercised without a GPU and compared with
the sequential emulation of the reference's sync loop."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class OracleAsLayer:
    """the oracle behind the interface SlabWorld expects from mc_mpi_b200.layer.Layer"""

    def __init__(self, cfg, K, r):
        sys.path.insert(0, HERE)
        from util import make_oracle
        self.o = make_oracle(cfg, K, r)
        self.device = 0

    def counts(self):
        st = self.o.stats()
        return {"n_outbox_left": len(self.o.particles_left), "n_outbox_right": len(self.o.particles_right),
                "nb_disabled": self.o.nb_disabled, "events": st["events"]}

    def simulate(self, nb):
        self.o.simulate(nb)
        return self.counts()

    def pop_left(self):
        a = self.o.particles_left.copy()
        self.o.clear_left()
        return a

    def pop_right(self):
        a = self.o.particles_right.copy()
        self.o.clear_right()
        return a

    def push(self, p):
        self.o.push(p)

    @property
    def weights_absorbed_f64(self):
        return self.o.tally_exact_f64


def _worker(rank, K, port, n, per_cycle, overlap, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=K)
    try:
        from mc_mpi_b200 import configs
        from mc_mpi_b200.world import SlabWorld
        cfg = configs.reference_default(n)
        w = SlabWorld(cfg, nb_particles_per_cycle=per_cycle, layer=OracleAsLayer(cfg, K, rank),
                      global_dx=False, overlap=overlap, statistics_cycle_time=1e-3)
        s = w.spin()
        wa = w.gather_weights_absorbed()
        stats = w.gather_stat_rows()
        q.put((rank, s["cycles"], s["migrations_out"], s["nb_disabled"], s["events"],
               None if wa is None else wa.tolist(), stats))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("K,n,per_cycle,overlap", [(2, 1500, 400, False), (3, 900, 100_000, False),
                                                   (3, 1200, 300, True)])
def test_slab_world_over_gloo(K, n, per_cycle, overlap):
    sys.path.insert(0, HERE)
    from mc_mpi_b200 import configs
    from util import oracle_chain
    cfg = configs.reference_default(n)
    layers, cycles, mig = oracle_chain(cfg, K, per_cycle)
    want_wa = np.concatenate([l.tally_exact_f64 for l in layers])

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, K, port, n, per_cycle, overlap, q))
             for r in range(K)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(K))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len({r[1] for r in res}) == 1                        # every rank left the loop together
    if not overlap:                                             # lock-step == the sync worker's cycles
        assert res[0][1] == cycles
    else:                                                       # transfers one cycle in flight
        assert res[0][1] >= cycles
    assert sum(r[2] for r in res) == mig                        # migrations
    assert [r[3] for r in res] == [l.nb_disabled for l in layers]
    assert sum(r[3] for r in res) == n                          # termination criterion
    assert [r[4] for r in res] == [l.stats()["events"] for l in layers]
    assert np.array_equal(np.array(res[0][5]), want_wa)         # gathered tally, rank 0
    assert all(r[5] is None for r in res[1:])
    # Timer::State rows (stats.csv): on rank 0 one list per rank, windows contiguous in time,
    # cycles adding up, phases inside the window
    stats = res[0][6]
    assert all(r[6] is None for r in res[1:]) and len(stats) == K
    for rows in stats:
        assert sum(row[6] for row in rows) == res[0][1]
        for row in rows:
            start, end, comp, send, recv, idle, _ = row
            assert end >= start and min(comp, send, recv) >= 0 and idle == 0
            assert comp + send + recv <= (end - start) + 1e-3
