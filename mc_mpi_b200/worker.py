"""Host-side mirror of the reference's Worker (include/mcmpi/worker.hpp) over the native world.

`Worker` = one rank of a run: the sub-slab of `decompose_domain` (src/layer.cpp:17-42) on one
GPU, `spin()` = Worker::spin (src/worker_sync.cpp:24-135), `gather_weights_absorbed()` =
src/worker.cpp:183-216.  Everything between `spin()`'s barrier and its return happens in
libmcb200.so: ONE resident kernel per rank tracks, exchanges escapees with the neighbour GPUs
through peer-mapped rings and stops on the device-side global count (csrc/mcb_world*.cu).
Python / torch.distributed only carry the IPC handles round at start-up, hold the barrier
before the launches and gather the final tally -- plumbing, never on the path.

`LocalBox` = all ranks inside this process (one host thread drives N GPUs, or N ranks share
one GPU for tests): mcb200_world_run.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from . import configs as _configs
from ._abi import WorldDesc, WorldGeom, WorldResult, check

SEED0 = 5127801  # src/layer.cpp:36


def _desc(cfg: _configs.SlabConfig, rank, world_size, device, cuts, opts, keep):
    d = WorldDesc()
    d.abi_version = _abi.ABI_VERSION
    d.device, d.rank, d.world_size = int(device), int(rank), int(world_size)
    d.x_min, d.x_max, d.x_ini = float(cfg.x_min), float(cfg.x_max), float(cfg.x_ini)
    d.nb_cells = int(cfg.nb_cells)
    d.particle_min_weight = float(cfg.particle_min_weight)
    if cuts is not None:
        c = np.ascontiguousarray(cuts, dtype=np.int32)
        if c.shape != (world_size + 1,):
            raise ValueError("cuts must have world_size + 1 entries")
        keep.append(c)
        d.cuts = c.ctypes.data
    for name, arr in (("sigs", cfg.sigs), ("absorption_rates", cfg.absorption_rates)):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=np.float32)
            if a.shape != (cfg.nb_cells,):
                raise ValueError(f"{name} must be a GLOBAL table of nb_cells entries")
            keep.append(a)
            setattr(d, name, a.ctypes.data)
    for k in ("windows", "block", "max_ctas", "ring_cap", "retire_batch", "xs_global", "bank_cap",
              "inflight_limit"):
        setattr(d, k, int(opts.pop(k, 0) or 0))
    if opts:
        raise TypeError(f"unknown world options: {sorted(opts)}")
    return d


class _Rank:
    """one mcb200_world handle"""

    def __init__(self, cfg, rank, world_size, device, cuts, **opts):
        self.cfg, self.rank, self.world_size, self.device = cfg, int(rank), int(world_size), int(device)
        keep = []
        d = _desc(cfg, rank, world_size, device, cuts, opts, keep)
        h = C.c_void_p()
        check(_abi.lib().mcb200_world_create(C.byref(d), C.byref(h)))
        self._h = h
        lo, m = C.c_int32(), C.c_int32()
        check(_abi.lib().mcb200_world_cells(self._h, C.byref(lo), C.byref(m)))
        self.lo, self.m = lo.value, m.value

    def close(self):
        if self._h is not None:
            _abi.lib().mcb200_world_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        check(_abi.lib().mcb200_world_set_option(self._h, key.encode(), int(value)))

    def export(self):
        handle = (C.c_uint8 * _abi.IPC_HANDLE_BYTES)()
        geom = WorldGeom()
        check(_abi.lib().mcb200_world_export(self._h, handle, C.byref(geom)))
        return bytes(handle), bytes(geom)

    def connect_peer(self, peer_rank: int, handle: bytes, geom: bytes):
        h = (C.c_uint8 * _abi.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        g = WorldGeom.from_buffer_copy(geom)
        check(_abi.lib().mcb200_world_connect_peer(self._h, int(peer_rank), h, C.byref(g)))

    def connect_local(self, other: "_Rank"):
        check(_abi.lib().mcb200_world_connect_local(self._h, other._h))

    def disconnect(self):
        check(_abi.lib().mcb200_world_disconnect(self._h))

    def prepare(self, nb_particles: int, seed: int = SEED0):
        check(_abi.lib().mcb200_world_prepare(self._h, int(nb_particles), int(seed)))

    def launch(self):
        check(_abi.lib().mcb200_world_launch(self._h))

    def wait(self) -> dict:
        r = WorldResult()
        rc = _abi.lib().mcb200_world_wait(self._h, C.byref(r))
        if rc != _abi.OK:
            err = _abi.McbError(rc, _abi.lib().mcb200_last_error().decode(errors="replace"))
            err.result = r.as_dict()   # the counters of the failed run, for diagnostics
            raise err
        return r.as_dict()

    def reset_tally(self):
        check(_abi.lib().mcb200_world_reset_tally(self._h))

    @property
    def weights_absorbed(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float32)
        check(_abi.lib().mcb200_world_tally(self._h, out.ctypes.data))
        return out

    @property
    def weights_absorbed_f64(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float64)
        check(_abi.lib().mcb200_world_tally_f64(self._h, out.ctypes.data))
        return out

    def weights_absorbed_exact(self):
        out = np.empty((self.m, 4), dtype=np.uint32)
        k = C.c_int32(0)
        check(_abi.lib().mcb200_world_tally_exact(self._h, out.ctypes.data, C.byref(k)))
        return out, k.value

    @property
    def stream_ptr(self) -> int:
        return int(_abi.lib().mcb200_world_stream(self._h) or 0)


class LocalBox:
    """All K ranks of a run inside this process: one host thread drives the GPUs in `devices`
    (ranks may share a device when `max_ctas` leaves room for all of their kernels at once --
    that is how the exchange protocol is tested on a single GPU)."""

    def __init__(self, cfg: _configs.SlabConfig, world_size=1, *, devices=None, cuts=None, **opts):
        K = int(world_size)
        devices = list(devices) if devices is not None else [0] * K
        self.cfg, self.K = cfg, K
        self.ranks = []
        try:
            for r in range(K):
                self.ranks.append(_Rank(cfg, r, K, devices[r], cuts, **dict(opts)))
            for a in self.ranks:
                for b in self.ranks:
                    if a is not b:
                        a.connect_local(b)
        except Exception:
            self.close()
            raise

    def set_option(self, key, value):
        for r in self.ranks:
            r.set_option(key, value)

    def run(self, nb_particles=None, seed=SEED0):
        """mcb200_world_run: prepare / launch / wait on every rank -> list of per-rank results"""
        n = self.cfg.nb_particles if nb_particles is None else int(nb_particles)
        hs = (C.c_void_p * self.K)(*[r._h for r in self.ranks])
        res = (WorldResult * self.K)()
        check(_abi.lib().mcb200_world_run(hs, self.K, n, int(seed), res))
        return [res[i].as_dict() for i in range(self.K)]

    def gather_weights_absorbed(self) -> np.ndarray:
        hs = (C.c_void_p * self.K)(*[r._h for r in self.ranks])
        out = np.empty(self.cfg.nb_cells, dtype=np.float64)
        check(_abi.lib().mcb200_world_gather_tally_f64(hs, self.K, out.ctypes.data))
        return out

    def gather_weights_absorbed_exact(self) -> np.ndarray:
        return np.concatenate([r.weights_absorbed_exact()[0] for r in self.ranks])

    def reset_tally(self):
        for r in self.ranks:
            r.reset_tally()

    def close(self):
        for r in self.ranks:
            try:
                r.disconnect()
            except Exception:
                pass
        for r in self.ranks:
            r.close()
        self.ranks = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def occupancy(r: dict) -> float:
    """fraction of a rank's lane-time that carried a live history during one run: lanes busy
    inside the event iterations x share of the time its warps had anything to track"""
    warps = r["ctas"] * r["block"] // 32
    idle = r["idle_warp_ns"] / max(r["kernel_ms"] * 1e6 * warps, 1.0)
    return (r["events"] / max(r["lane_slots"], 1)) * max(1.0 - idle, 0.0)


def totals(results) -> dict:
    """whole-world sums of the per-rank results of one run"""
    keys = ("events", "scatters", "n_left", "n_right", "n_dead", "births", "sent_left",
            "sent_right", "window_crossings", "idle_polls", "blocked_passes", "bank_pushes",
            "bank_pops", "lane_slots", "idle_warp_ns")
    out = {k: int(sum(r[k] for r in results)) for k in keys}
    for k in ("w_left", "w_right", "w_dead"):
        out[k] = float(sum(r[k] for r in results))
    out["kernel_ms_max"] = max(r["kernel_ms"] for r in results)
    return out


class Worker:
    """One rank per process (torchrun): Worker::Worker (src/worker.cpp:16-34) builds the
    rank's sub-slab; the IPC handles of the exchange blocks go round with one all-gather (the
    window set-up of RmaComm, src/rma_comm.cpp:49-121)."""

    def __init__(self, cfg: _configs.SlabConfig, *, device=0, group=None, cuts=None, **opts):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.cfg, self.group = cfg, group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world_size = 0, 1
        self.device = int(device)
        self.cuts = list(cuts) if cuts is not None else None
        self._opts = dict(opts)
        self.r = _Rank(cfg, self.rank, self.world_size, self.device, cuts, **dict(opts))
        self._connect()

    @property
    def tdev(self):
        # where the few plumbing tensors (IPC handles, counters) live for the collectives; the
        # CPU is only ever used by the gloo tests of this host logic (no compute happens here)
        if not self._torch.cuda.is_available():
            return self._torch.device("cpu")
        return self._torch.device("cuda", self.device)

    def _connect(self):
        if self.world_size == 1:
            return
        dist, torch = self._dist, self._torch
        handle, geom = self.r.export()
        blob = np.frombuffer(handle + geom, dtype=np.uint8)
        mine = torch.from_numpy(blob.copy()).to(self.tdev)
        table = torch.empty(self.world_size * mine.numel(), dtype=torch.uint8, device=self.tdev)
        dist.all_gather_into_tensor(table, mine, group=self.group)
        table = table.cpu().numpy().reshape(self.world_size, -1)
        n_h = _abi.IPC_HANDLE_BYTES
        for peer in range(self.world_size):
            if peer != self.rank:
                raw = table[peer].tobytes()
                self.r.connect_peer(peer, raw[:n_h], raw[n_h:])
        dist.barrier(group=self.group)

    def recut(self, cuts):
        """move the sub-slab boundaries (between runs; the tally restarts)"""
        if self.world_size > 1:
            self.r.disconnect()
            self._dist.barrier(group=self.group)   # nobody frees a block a peer still maps
        self.r.close()
        self.cuts = list(cuts) if cuts is not None else None
        self.r = _Rank(self.cfg, self.rank, self.world_size, self.device, self.cuts,
                       **dict(self._opts))
        self._connect()

    def spin(self, nb_particles=None, seed=SEED0) -> dict:
        """Worker::spin: the whole run.  barrier, prepare on every rank, barrier, launch, wait."""
        n = self.cfg.nb_particles if nb_particles is None else int(nb_particles)
        if self.world_size > 1:
            # every rank's previous kernel has ended before anybody wipes its rings: a consumer
            # publishes its last credits (stores into the PRODUCER's memory) after the producer
            # may already have seen `done`
            self._dist.barrier(group=self.group)
        self.r.prepare(n, seed)
        if self.world_size > 1:
            self._dist.barrier(group=self.group)
        self.r.launch()
        return self.r.wait()

    def gather_weights_absorbed(self, exact=False):
        """Worker::gather_weights_absorbed (src/worker.cpp:183-216): the disjoint slices of all
        ranks concatenated on every rank; float64 view of the exact tally, or its digits."""
        mine = (self.r.weights_absorbed_exact()[0].astype(np.int64) if exact
                else self.r.weights_absorbed_f64)
        if self.world_size == 1:
            return mine.astype(np.uint32) if exact else mine
        dist, torch = self._dist, self._torch
        K = self.world_size
        sizes = torch.zeros(K, dtype=torch.int64, device=self.tdev)
        sizes[self.rank] = self.r.m
        dist.all_reduce(sizes, group=self.group)
        sizes = [int(v) for v in sizes.tolist()]
        m_max = max(sizes)
        width = 4 if exact else 1
        pad = torch.zeros(m_max * width, dtype=torch.int64 if exact else torch.float64)
        pad[: mine.size] = torch.from_numpy(np.ascontiguousarray(mine).reshape(-1))
        pad = pad.to(self.tdev)
        out = torch.empty(K * m_max * width, dtype=pad.dtype, device=self.tdev)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        out = out.cpu().numpy().reshape(K, m_max * width)
        parts = [out[r, : sizes[r] * width] for r in range(K)]
        full = np.concatenate(parts)
        return full.reshape(-1, 4).astype(np.uint32) if exact else full

    def all_ranks(self, values, op="sum"):
        """element-wise sum / max / per-rank table of a few numbers over the ranks (collective)"""
        torch, dist = self._torch, self._dist
        v = torch.tensor([float(x) for x in values], dtype=torch.float64, device=self.tdev)
        if self.world_size == 1:
            return [v.tolist()] if op == "table" else v.tolist()
        if op == "table":
            out = torch.empty(self.world_size * v.numel(), dtype=torch.float64, device=self.tdev)
            dist.all_gather_into_tensor(out, v, group=self.group)
            return out.reshape(self.world_size, -1).tolist()
        dist.all_reduce(v, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM,
                        group=self.group)
        return v.tolist()

    def parity(self, case: str, digest: dict) -> dict:
        """Run the digest case `case` (tests/golden/world_digest.json: what the ORACLE produces for
        it as one layer) on this world and compare: SHA-256 over all 128 bits of every cell of the
        gathered tally, events, scatters, histories absorbed at the global borders / dead, weight
        conservation.  Collective; every rank gets the verdict.  The tally is reset before and
        after (it is cumulative over runs)."""
        import hashlib
        want = digest[case]
        if (want["nb_cells"], float(np.float32(want["particle_min_weight"]))) != (
                self.cfg.nb_cells, float(np.float32(self.cfg.particle_min_weight))):
            raise ValueError(f"digest case {case} is not this world's slab")
        self.r.reset_tally()
        res = self.spin(want["nb_particles"], want["seed"])
        exact = self.gather_weights_absorbed(exact=True)
        self.r.reset_tally()
        ev, sc, nl, nr, nd, sl, sr, err = (int(v) for v in self.all_ranks(
            [res[k] for k in ("events", "scatters", "n_left", "n_right", "n_dead", "sent_left",
                              "sent_right")] + [abs(res["error"])]))
        w_left, w_right, w_dead = self.all_ranks([res["w_left"], res["w_right"], res["w_dead"]])
        sha = hashlib.sha256(np.ascontiguousarray(exact, dtype="<u4").tobytes()).hexdigest()
        counts_exact = (ev, sc, nl, nr, nd) == tuple(
            want[k] for k in ("events", "scatters", "n_left", "n_right", "n_dead"))
        cells = exact.astype(np.float64)   # each cell's exact value rounded once to double
        w_abs = float(np.sum(np.ldexp(cells[:, 0], -120) + np.ldexp(cells[:, 1], -88) +
                             np.ldexp(cells[:, 2], -56) + np.ldexp(cells[:, 3], -24)))
        conservation = w_abs + w_left + w_right + w_dead
        return {"checked": True, "case": case, "tally_bit_exact": sha == want["tally_exact_sha256"],
                "counts_exact": bool(counts_exact), "conservation": conservation,
                "conservation_ok": abs(conservation - 1.0) < 1e-5, "kernel_error": err,
                "events": ev, "migrations_per_history": (sl + sr) / want["nb_particles"],
                "ranks": self.world_size, "cuts": self.cuts or "equal",
                "reference": "tests/golden/world_digest.json (oracle, one layer)"}

    def close(self):
        if self.r is None:
            return
        if self.world_size > 1:
            self.r.disconnect()
            self._dist.barrier(group=self.group)
        self.r.close()
        self.r = None
