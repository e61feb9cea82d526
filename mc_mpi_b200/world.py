"""Multi-GPU driver: the slab domain-decomposed one contiguous sub-slab per rank / GPU.

Stands in for the reference's worker loop (Worker::spin, src/worker_sync.cpp:24-135):
simulate -> ship escapees to the neighbour ranks -> stop when the disabled counts of all
ranks add up to nb_particles.  One process per GPU (torchrun); torch.distributed is the
plumbing (NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests).  The 1-D chain
topology is the reference's: rank r only ever talks to r-1 and r+1.

What is different from the MPI workers:
  * escapees never touch the host: the tracking kernel compacts them into device outboxes
    of 24-byte wire records, NCCL moves them GPU-to-GPU, `push_device` unpacks them into
    the neighbour's bank;
  * the exchange of cycle c is IN FLIGHT while cycle c+1 is being tracked (the async
    worker's idea, src/worker_async.cpp, without its polling): per cycle the cost is
    max(track, transfer) instead of their sum;
  * every sub-slab tracks with the ONE global dx and slices of the ONE global cross-section
    table (decompose_domain(global_dx=True)), so a K-GPU run reproduces the 1-GPU run bit
    for bit wherever the cuts are (the reference's K-rank runs differ from its 1-rank run
    by up to 5e-4 per cell, SURVEY hard part 3) -- which frees the cuts to be placed where
    the measured work balances (`balanced_cuts`);
  * the per-cycle bookkeeping (outbox sizes + disabled counts) is ONE small all-gather
    instead of 4 Sendrecv + Barrier + Allreduce (src/worker_sync.cpp:47-120).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from . import configs as _configs
from .layer import PARTICLE_DTYPE, Layer, decompose_domain, split_cells

RECORD = PARTICLE_DTYPE.itemsize  # 24 bytes on the wire


def equal_cuts(nb_cells: int, world_size: int):
    """the reference's split (src/layer.cpp:24-27) as K+1 cell boundaries"""
    return [split_cells(nb_cells, world_size, r)[0] for r in range(world_size)] + [nb_cells]


def balanced_cuts(cuts, cost_per_rank, nb_cells, min_cells=8):
    """New cell boundaries that equalise the measured cost, taking the cost density as
    uniform inside each current sub-slab.  With the global dx / global cross-section table
    the RESULT of a run does not depend on the cuts (bit for bit), so they are free to go
    where the measured work balances; the reference's equal split leaves the sub-slab that
    holds the source with ~18 % more events than the mean on the default slab."""
    K = len(cuts) - 1
    dens = np.concatenate([np.full(cuts[r + 1] - cuts[r], cost_per_rank[r] / (cuts[r + 1] - cuts[r]))
                           for r in range(K)])
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    new = [0]
    for r in range(1, K):
        c = int(np.searchsorted(cum, cum[-1] * r / K))
        c = max(new[-1] + min_cells, min(c, nb_cells - (K - r) * min_cells))
        new.append(c)
    return new + [nb_cells]


class _StatWindow:
    """one Timer::State: wall-clock start/end and the seconds spent per phase"""

    def __init__(self):
        self.start = time.time()
        self.comp = self.send = self.recv = self.idle = 0.0
        self.cycles = 0

    def add(self, comp=0.0, send=0.0, recv=0.0, idle=0.0):
        self.comp += comp
        self.send += send
        self.recv += recv
        self.idle += idle
        self.cycles += 1

    def close(self):
        return (self.start, time.time(), self.comp, self.send, self.recv, self.idle, self.cycles)


class SlabWorld:
    def __init__(self, cfg: _configs.SlabConfig, *, rank=None, world_size=None, device=None,
                 nb_particles_per_cycle=1 << 23, layer=None, group=None, global_dx=True,
                 cuts=None, ramp_from=None, overlap=False, transport="nccl",
                 statistics_cycle_time=None):
        self.cfg = cfg
        # Timer::State rows (include/timer/timer.hpp, src/worker_sync.cpp:122-139): one per
        # `statistics_cycle_time` seconds of spin() plus the final one; None = no rows
        self.statistics_cycle_time = statistics_cycle_time
        self.stat_rows = []
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.per_cycle = int(nb_particles_per_cycle)
        # births per cycle ramp up from `ramp_from` (doubling) and down again at the end, so
        # that filling / draining the K-stage pipeline costs small cycles, not full ones
        self.ramp_from = int(ramp_from) if ramp_from else None
        self.cuts = list(cuts) if cuts is not None else equal_cuts(cfg.nb_cells, self.world_size)
        if cuts is not None and not global_dx:
            raise ValueError("custom cuts need global_dx (the reference's per-layer dx depends "
                             "on the cuts)")
        self.device = device or 0
        self.global_dx = global_dx
        if layer is None:
            layer = self._make_layer()
        self.layer = layer
        # device-resident exchange when the layer lives on a GPU and the backend can move
        # device memory; host staging otherwise (gloo tests)
        self.on_device = isinstance(layer, Layer) and dist.get_backend(group) == "nccl"
        self.tdev = torch.device("cuda", layer.device) if self.on_device else torch.device("cpu")
        self.overlap = bool(overlap)
        # "nccl": outbox -> ncclSend/Recv -> bank.  "p2p": the tracking kernel stores escapees
        # straight into the neighbour GPU's inbox over NVLink (CUDA IPC mapping); per cycle
        # only the small all-gather of the disabled counts remains as communication call.
        self.transport = transport if self.on_device else "nccl"
        if self.transport == "p2p":
            self._connect_p2p()
        self._buf = {}
        self.cycles = 0
        self.migrations_out = 0
        self.t_simulate = self.t_exchange = 0.0
        self.trace = None   # set to [] to record one tuple per cycle (see spin)
        self.t_parts ={"counts": 0.0, "simulate": 0.0, "gather": 0.0, "finish": 0.0, "start": 0.0,
                        "start.pop": 0.0, "start.post": 0.0, "finish.wait": 0.0, "finish.push": 0.0}

    def _make_layer(self):
        cfg = self.cfg
        lo, hi = self.cuts[self.rank], self.cuts[self.rank + 1]
        return decompose_domain(
            cfg.x_min, cfg.x_max, cfg.x_ini, self.world_size, self.rank, cfg.nb_cells,
            cfg.nb_particles, cfg.particle_min_weight, device=self.device,
            global_dx=self.global_dx, sigs=cfg.sigs, absorption_rates=cfg.absorption_rates,
            cells=(lo, hi - lo) if self.global_dx else None,
            # the slab's borders are rank 0's left and the last rank's right edge, whatever the
            # coordinates: the reference's |x| < 1e-4 heuristic (src/layer.cpp:47-48) would take an
            # interior cut of a fine grid for a border, and miss a border of a slab not ending at 1
            left_border=self.rank == 0, right_border=self.rank == self.world_size - 1)

    def recut(self, cuts):
        """move the sub-slab boundaries (between runs: the layer is rebuilt, tallies reset)"""
        self.cuts = list(cuts)
        if self.transport == "p2p":
            # nobody may free an inbox that a neighbour still has mapped
            self.layer.disconnect_peers()
            dist.barrier(group=self.group)
        self.layer.close()
        self.layer = self._make_layer()
        if self.transport == "p2p":
            self._connect_p2p()

    def close(self):
        if self.transport == "p2p":
            self.layer.disconnect_peers()
            dist.barrier(group=self.group)
        self.layer.close()

    def _connect_p2p(self):
        """every rank allocates an inbox, the IPC handles go round with one all-gather, and each
        layer maps its two neighbours' inboxes (the RmaComm window set-up, src/rma_comm.cpp:49-121)"""
        from ._abi import IPC_HANDLE_BYTES, InboxGeom, McbError
        import ctypes
        K, r = self.world_size, self.rank
        n_h, n_g = IPC_HANDLE_BYTES, ctypes.sizeof(InboxGeom)
        ok = 1
        try:
            handle, geom = self.layer.inbox_create(self.per_cycle + self.per_cycle // 4)
        except McbError:
            ok, handle, geom = 0, bytes(n_h), bytes(n_g)
        blob = np.frombuffer(handle + geom + bytes([ok]), dtype=np.uint8)
        mine = torch.from_numpy(blob.copy()).to(self.tdev)
        table = torch.empty(K * mine.numel(), dtype=torch.uint8, device=self.tdev)
        dist.all_gather_into_tensor(table, mine, group=self.group)
        table = table.cpu().numpy().reshape(K, -1)
        ok = int(table[:, -1].min())
        if ok:
            try:
                for side, peer in ((0, r - 1), (1, r + 1)):
                    if 0 <= peer < K:
                        raw = table[peer].tobytes()
                        self.layer.connect_peer(side, raw[:n_h], raw[n_h:n_h + n_g])
            except McbError:
                ok = 0
        # every rank must agree: one rank without peer access puts everybody back on NCCL
        flag = torch.tensor([ok], dtype=torch.int32, device=self.tdev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.layer.disconnect_peers()
            self.transport = "nccl"
        dist.barrier(group=self.group)

    # -- buffers ---------------------------------------------------------------------
    def _buffer(self, name: str, n_records: int) -> torch.Tensor:
        need = max(int(n_records), 1) * RECORD
        b = self._buf.get(name)
        if b is None or b.numel() < need:
            b = torch.empty(int(need * 1.5) + RECORD, dtype=torch.uint8, device=self.tdev)
            self._buf[name] = b
        return b

    def _global(self, group_rank: int) -> int:
        return dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    # -- one cycle -------------------------------------------------------------------
    def _gather_counts(self, counts: dict) -> torch.Tensor:
        """[K, 3] table of (outbox left, outbox right, nb_disabled): replaces the sizes the
        MPI workers learn from MPI_Get_count and the Allreduce of src/worker_sync.cpp:112-120"""
        K = self.world_size
        mine = torch.tensor([counts["n_outbox_left"], counts["n_outbox_right"],
                             counts["nb_disabled"]], dtype=torch.int64, device=self.tdev)
        table = torch.empty(K * 3, dtype=torch.int64, device=self.tdev)
        dist.all_gather_into_tensor(table, mine, group=self.group)
        return table.view(K, 3).cpu()

    def _start_exchange(self, table: torch.Tensor, parity: int):
        """post the sends of particles_left / particles_right to rank-1 / rank+1 and the
        matching receives (src/worker_sync.cpp:47-108); returns what _finish_exchange needs"""
        K, r = self.world_size, self.rank
        n_send = {-1: int(table[r, 0]), +1: int(table[r, 1])}
        n_recv = {-1: int(table[r - 1, 1]) if r > 0 else 0,
                  +1: int(table[r + 1, 0]) if r + 1 < K else 0}
        ops, recv_bufs = [], {}
        for side, d in ((0, -1), (1, +1)):
            peer = r + d
            if not 0 <= peer < K:
                continue
            if n_send[d] > 0:
                sb = self._buffer(f"send{d}.{parity}", n_send[d])
                if self.on_device:
                    _t = time.perf_counter()
                    got = self.layer.pop_device(side, sb.data_ptr(), n_send[d])
                    self.t_parts["start.pop"] += time.perf_counter() - _t
                else:
                    arr = self.layer.pop_left() if side == 0 else self.layer.pop_right()
                    got = len(arr)
                    sb[: got * RECORD] = torch.from_numpy(
                        np.frombuffer(arr.tobytes(), dtype=np.uint8).copy())
                assert got == n_send[d]
                ops.append(dist.P2POp(dist.isend, sb[: got * RECORD], self._global(peer),
                                      group=self.group))
                self.migrations_out += got
            if n_recv[d] > 0:
                rb = self._buffer(f"recv{d}.{parity}", n_recv[d])
                recv_bufs[d] = (rb, n_recv[d])
                ops.append(dist.P2POp(dist.irecv, rb[: n_recv[d] * RECORD],
                                      self._global(peer), group=self.group))
        _t = time.perf_counter()
        reqs = dist.batch_isend_irecv(ops) if ops else []
        self.t_parts["start.post"] += time.perf_counter() - _t
        return reqs, recv_bufs

    def _finish_exchange(self, pending):
        """wait for the transfers and append what arrived to the bank"""
        if pending is None:
            return
        reqs, recv_bufs = pending
        _t = time.perf_counter()
        for req in reqs:
            req.wait()
        if reqs and self.on_device:
            torch.cuda.current_stream(self.tdev).synchronize()
        self.t_parts["finish.wait"] += time.perf_counter() - _t
        _t = time.perf_counter()
        for d, (rb, n) in recv_bufs.items():
            if self.on_device:
                self.layer.push_device(rb.data_ptr(), n)
            else:
                raw = rb[: n * RECORD].numpy().tobytes()
                self.layer.push(np.frombuffer(raw, dtype=PARTICLE_DTYPE))
        self.t_parts["finish.push"] += time.perf_counter() - _t

    def spin(self, max_cycles=10_000_000) -> dict:
        """Worker::spin: cycle until every source particle is disabled somewhere."""
        total = self.cfg.nb_particles
        births = min(self.ramp_from, self.per_cycle) if self.ramp_from else self.per_cycle
        pending = None
        p2p = self.transport == "p2p"
        seen = self.layer.counts() if p2p else None
        stat = _StatWindow() if self.statistics_cycle_time is not None else None
        while self.cycles < max_cycles:
            t0 = time.perf_counter()
            ta = t0
            if p2p:
                # escapees of this cycle are stored by the kernel into the neighbours' inboxes
                # of this parity; the all-gather below doubles as the barrier after which they
                # may be ingested
                self.layer.set_exchange_parity(self.cycles & 1)
            if self.ramp_from:
                # everything received so far + this cycle's share of the source
                st = self.layer.counts()
                unborn = st["n_unborn"]
                b = min(births, max(unborn // 2, self.ramp_from)) if unborn > 0 else 0
                ta = time.perf_counter()
                c = self.layer.simulate(st["n_bank"] + b)
                births = min(births * 2, self.per_cycle)
            else:
                c = self.layer.simulate(self.per_cycle)
            t1 = time.perf_counter()
            table = self._gather_counts(c)
            tb = time.perf_counter()
            if p2p:
                tc = tb
                for from_side, peer in ((0, self.rank - 1), (1, self.rank + 1)):
                    if 0 <= peer < self.world_size:
                        self.layer.ingest_inbox(from_side, self.cycles & 1)
                td = time.perf_counter()
                self.migrations_out += sum(
                    c[k] - seen[k] for k, border in (("n_left", self.layer.left_border),
                                                     ("n_right", self.layer.right_border))
                    if not border)
                seen = c
            else:
                # the previous cycle's transfers ran under this cycle's tracking
                self._finish_exchange(pending)
                tc = time.perf_counter()
                pending = self._start_exchange(table, self.cycles & 1)
                td = time.perf_counter()
                if not self.overlap:
                    self._finish_exchange(pending)
                    pending = None
            t2 = time.perf_counter()
            self.t_simulate += t1 - t0
            self.t_exchange += t2 - t1
            for k, v in (("counts", ta - t0), ("simulate", t1 - ta), ("gather", tb - t1),
                         ("finish", (tc - tb) + (t2 - td)), ("start", td - tc)):
                self.t_parts[k] += v
            if self.trace is not None:
                self.trace.append((self.cycles, round((t1 - ta) * 1e3, 3), round(c["track_ms"], 3),
                                   round((tb - t1) * 1e3, 3), round((td - tc) * 1e3, 3),
                                   round(((tc - tb) + (t2 - td)) * 1e3, 3),
                                   int(table[self.rank, 0]), int(table[self.rank, 1]),
                                   c["n_bank"], c["n_unborn"]))
            self.cycles += 1
            if stat is not None:
                # Computation / Recv (events) / Send (particles), as the sync worker tags them
                stat.add(comp=t1 - t0, recv=tb - t1, send=t2 - tb)
                if time.time() > stat.start + self.statistics_cycle_time:
                    self.stat_rows.append(stat.close())
                    stat = _StatWindow()
            if int(table[:, 2].sum()) == total:
                # every source particle is disabled somewhere: nothing can be in flight
                self._finish_exchange(pending)
                break
        else:
            raise RuntimeError("SlabWorld.spin: did not terminate")
        if stat is not None:
            self.stat_rows.append(stat.close())
        return self.summary()

    def gather_stat_rows(self):
        """Worker::write_file's gather (src/worker.cpp:63-181): every rank's rows on rank 0 as
        a list indexed by rank (None elsewhere)"""
        rows = [None] * self.world_size
        dist.all_gather_object(rows, self.stat_rows, group=self.group)
        return rows if self.rank == 0 else None

    # -- results ---------------------------------------------------------------------
    def summary(self) -> dict:
        c = self.layer.counts()
        return {"rank": self.rank, "cycles": self.cycles, "migrations_out": self.migrations_out,
                "t_simulate": self.t_simulate, "t_exchange": self.t_exchange, **c}

    def gather_weights_absorbed(self):
        """Worker::gather_weights_absorbed (src/worker.cpp:183-216): the disjoint per-rank
        slices concatenated on rank 0 (None elsewhere); float64 view of the exact tally."""
        K = self.world_size
        mine = torch.from_numpy(np.ascontiguousarray(self.layer.weights_absorbed_f64))
        sizes = [self.cuts[r + 1] - self.cuts[r] for r in range(K)]
        m_max = max(sizes)
        pad = torch.zeros(m_max, dtype=torch.float64)
        pad[: mine.numel()] = mine
        pad = pad.to(self.tdev)
        out = torch.empty(K * m_max, dtype=torch.float64, device=self.tdev)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        if self.rank != 0:
            return None
        out = out.cpu().view(K, m_max).numpy()
        return np.concatenate([out[r, : sizes[r]] for r in range(K)])
