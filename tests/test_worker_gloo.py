"""Host logic of the persistent-kernel driver (mc_mpi_b200/worker.py: Worker) on CPU: a
world_size-2 `gloo` process group with a stand-in for the native rank handle -- the exchange of
the IPC handles, the two barriers of a run, the gather of the exact tally and the comparison with
the committed oracle digest are exercised without a GPU.  The stand-in (checker side) computes
its rank's share of the result with the ORACLE; the real ranks are tested on the GPU box
(tests/test_gpu_world.py, test_gpu_multi.py, bench.py --gpus N)."""
import json
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class StandInRank:
    """what Worker needs from worker._Rank, answered by the oracle"""
    log = []

    def __init__(self, cfg, rank, world_size, device, cuts, **opts):
        from mc_mpi_b200.layer import split_cells
        self.cfg, self.rank, self.K = cfg, rank, world_size
        if cuts is None:
            self.lo, self.m = split_cells(cfg.nb_cells, world_size, rank)
        else:
            self.lo, self.m = cuts[rank], cuts[rank + 1] - cuts[rank]
        self.peers = {}
        self._exact = np.zeros((self.m, 4), dtype=np.uint32)

    def export(self):
        return bytes([self.rank]) * 64, (b"geom%d" % self.rank).ljust(56, b".")

    def connect_peer(self, peer, handle, geom):
        assert handle == bytes([peer]) * 64 and geom.startswith(b"geom%d" % peer)
        self.peers[peer] = True

    def disconnect(self):
        self.peers = {}

    def close(self):
        pass

    def set_option(self, k, v):
        pass

    def reset_tally(self):
        self._exact[:] = 0

    def prepare(self, n, seed):
        assert len(self.peers) == self.K - 1          # every other rank was connected first
        self._n, self._seed = n, seed
        StandInRank.log.append("prepare")

    def launch(self):
        StandInRank.log.append("launch")

    def wait(self):
        # the whole slab as ONE oracle layer (what K ranks must reproduce), this rank's slice
        from util import make_oracle
        o = make_oracle(self.cfg.with_particles(self._n))
        o.simulate(-1)
        st = o.stats()
        self._exact = o.tally_exact[self.lo:self.lo + self.m].copy()
        first, last = self.rank == 0, self.rank == self.K - 1
        cw = o.class_weights_exact
        r = {"events": st["events"] if first else 0, "scatters": st["scatters"] if first else 0,
             "n_left": st["n_left"] if first else 0, "n_right": st["n_right"] if last else 0,
             "n_dead": st["n_dead"] if first else 0, "sent_left": 7 * self.rank, "sent_right": 3,
             "error": 0, "w_left": cw[0] if first else 0.0, "w_right": cw[1] if last else 0.0,
             "w_dead": cw[2] if first else 0.0}
        o.free()
        return r

    def weights_absorbed_exact(self):
        return self._exact, -120

    @property
    def weights_absorbed_f64(self):
        d = self._exact.astype(np.float64)
        return (np.ldexp(d[:, 0], -120) + np.ldexp(d[:, 1], -88) + np.ldexp(d[:, 2], -56) +
                np.ldexp(d[:, 3], -24))


def _worker(rank, K, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=K)
    try:
        from mc_mpi_b200 import configs, worker
        worker._Rank = StandInRank                      # the native handle is not under test here
        with open(os.path.join(HERE, "golden", "world_digest.json")) as f:
            digest = json.load(f)
        cfg = configs.reference_default(100_000)
        w = worker.Worker(cfg, device=0, cuts=[0, 640, 1000])
        v = w.parity("default_slab_1e5", digest)      # (resets the tally before and after)
        w.spin(100_000)
        full = w.gather_weights_absorbed()
        table = w.all_ranks([rank + 1, 10 * rank], "table")
        mx = w.all_ranks([rank + 1, 10 * rank], "max")
        w.recut([0, 300, 1000])
        v2 = w.parity("default_slab_1e5", digest)
        w.close()
        q.put((rank, v, float(full.sum()), len(full), table, mx, v2, list(StandInRank.log)))
    finally:
        dist.destroy_process_group()


def test_worker_host_logic_over_gloo():
    K = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, K, port, q)) for r in range(K)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(K))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    with open(os.path.join(HERE, "golden", "world_digest.json")) as f:
        want = json.load(f)["default_slab_1e5"]
    for rank, v, total, n, table, mx, v2, log in res:
        for verdict in (v, v2):     # every rank gets the same verdict, for both sets of cuts
            assert verdict["checked"] and verdict["tally_bit_exact"] and verdict["counts_exact"]
            assert verdict["conservation_ok"] and verdict["kernel_error"] == 0
            assert verdict["events"] == want["events"] and verdict["ranks"] == K
        assert v["cuts"] == [0, 640, 1000] and v2["cuts"] == [0, 300, 1000]
        assert n == 1000 and abs(total - want["w_absorbed"]) < 1e-12     # slices concatenated
        assert table == [[1.0, 0.0], [2.0, 10.0]] and mx == [2.0, 10.0]
        assert log == ["prepare", "launch"] * 3          # one prepare / launch per run, in order
