#!/usr/bin/env python
"""Benchmark of the particle-tracking hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (N=1 default)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                     # the reference's CPU path

metric  particle histories/s, whole box.  A step = one complete simulation of the
        workload: nb_particles source histories born, tracked to termination, tallied.
N = 1   workload = BASELINE configs[1]: single-layer slab, 1000 cells, 1e8 histories.
N > 1   workload = BASELINE configs[2]: the same slab domain-decomposed one sub-slab per GPU,
        1.25e8 x N histories (1e9 at N = 8), escapees exchanged GPU-to-GPU  ("weak").
value   device-resident: source particles are born on the device, timed with CUDA events on
        the layer's stream (max over ranks for N > 1).
e2e     the same metric through the reference-facing host interface: source particles in
        pinned HOST memory as 24-byte `Particle` records (the reference's wire format),
        pushed H2D, tracked, tally + counters read back D2H, all inside the timed region.
roofline  algorithmic 48 B per event x events per launch / measured kernel time,
        against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline  the reference's own CPU Layer (oracle/_ref, all host cores) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_EVENT = 48          # 24 B particle state read + 24 B written
HISTORIES_1GPU = 100_000_000  # BASELINE configs[1]
HISTORIES_PER_GPU = 125_000_000  # BASELINE configs[2]: 1e9 on 8 GPUs


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload(n_gpus: int, particles: int | None):
    from mc_mpi_b200 import configs
    if n_gpus == 1:
        cfg = configs.single_gpu_slab(particles or HISTORIES_1GPU)
        desc = "single-GPU single-layer slab, 1000 cells (BASELINE configs[1])"
    else:
        cfg = configs.single_gpu_slab((particles or HISTORIES_PER_GPU * n_gpus))
        desc = f"{n_gpus}-GPU domain-decomposed slab, 1000 cells (BASELINE configs[2])"
    return cfg, desc


# ----------------------------------------------------------------------------- CPU arm --

def cpu_reference_run(cfg, sample: int, nthread: int):
    """the reference's own CPU implementation of the path on `sample` histories of the
    workload (oracle/_ref when it exists, else the oracle port) -> (histories/s, kind)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    cls = pyoracle.RefLayer if kind == "reference" else pyoracle.OracleLayer
    lay = cls.decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, sample,
                               cfg.particle_min_weight)
    if cfg.sigs is not None:
        lay.sigs[:] = cfg.sigs
    if cfg.absorption_rates is not None:
        lay.absorption_rates[:] = cfg.absorption_rates
    t = time.perf_counter()
    lay.simulate(-1, nthread)
    dt = time.perf_counter() - t
    assert lay.nb_disabled == sample
    lay.free()
    return sample / dt, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs and prints the reference arm
    cfg, desc = workload(args.gpus, args.particles)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    for _ in range(args.warmup):
        cpu_reference_run(cfg, max(sample // 10, 1000), cores)
    t = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, kind = cpu_reference_run(cfg, sample, cores)
        rates.append(r)
    dt = time.perf_counter() - t
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "particle histories/s (whole box)", "value": value,
        "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "nb_cells": cfg.nb_cells,
                   "particle_min_weight": cfg.particle_min_weight,
                   "histories_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} histories of the workload per step, "
                                   f"Layer::simulate(-1, {cores}) (OpenMP, one rank)"},
        "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm --

def run_gpu_arm(args):
    import numpy as np
    import torch

    from mc_mpi_b200 import _abi
    from mc_mpi_b200.layer import PARTICLE_DTYPE, Layer, decompose_domain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run "
                             "(one process per GPU)")
    if _abi.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg, desc = workload(world, args.particles)
    n_hist = cfg.nb_particles
    peak, peak_src = measured_hbm_peak()
    wmc = float(np.float32(1.0 / n_hist))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(x: float):
        if dist is None:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [round(v, 3) for v in t.tolist()]

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident arm ---------------------------------------------------------
    if world == 1:
        layer = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, n_hist,
                                 cfg.particle_min_weight, device=local_rank, sigs=cfg.sigs,
                                 absorption_rates=cfg.absorption_rates)

        def step():
            layer.create_particles(cfg.x_ini, wmc, n_hist)
            c = layer.simulate(-1)
            assert c["nb_active"] == 0
            return c
    else:
        from mc_mpi_b200.world import SlabWorld, balanced_cuts
        sw = SlabWorld(cfg, device=local_rank, nb_particles_per_cycle=args.per_cycle,
                       ramp_from=args.ramp_from if args.ramp_from > 0 else None,
                       overlap=args.overlap, transport=args.transport)
        layer = sw.layer

        def step():
            lay = sw.layer
            if lay.counts()["n_unborn"] > 0 or lay.index_start <= cell_ini < lay.index_start + lay.m:
                lay.create_particles(cfg.x_ini, wmc, n_hist)
            base = sum_over_ranks(float(lay.counts()["nb_disabled"]))
            sw.cfg = cfg.with_particles(int(base) + n_hist)  # disabled counts are cumulative
            sw.spin()
            return lay.counts()

        cell_ini = int(np.float32(np.float32(cfg.x_ini) - np.float32(cfg.x_min)) /
                       (np.float32(np.float32(cfg.x_max) - np.float32(cfg.x_min)) / np.float32(cfg.nb_cells)))

    for w in range(args.warmup):
        before = sw.layer.counts()["track_ms"] if world > 1 else 0.0
        step()
        if world > 1 and args.balance and w < args.warmup - 1:
            # measured load balancing: move the cuts so that every GPU gets the same tracking
            # time (results do not depend on the cuts: global dx + global cross-section table)
            cost = per_rank(sw.layer.counts()["track_ms"] - before)
            sw.cfg = cfg
            sw.recut(balanced_cuts(sw.cuts, cost, cfg.nb_cells))
    if world > 1:
        layer = sw.layer
    stream = torch.cuda.ExternalStream(layer.stream_ptr, device=torch.device("cuda", local_rank))
    barrier()
    if world > 1:   # the host-side split reported below covers the timed steps only
        sw.cycles = 0
        sw.t_simulate = sw.t_exchange = 0.0
        sw.t_parts = {k: 0.0 for k in sw.t_parts}
        if os.environ.get("MCB200_TRACE_DIR"):
            sw.trace = []
    c0 = layer.counts()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    c1 = layer.counts()
    events = sum_over_ranks(float(c1["events"] - c0["events"]))
    track_ms = c1["track_ms"] - c0["track_ms"]
    launches = c1["launches"] - c0["launches"]
    gpu_launches = int(sum_over_ranks(float(c1["gpu_launches"] - c0["gpu_launches"])))
    my_events = c1["events"] - c0["events"]
    # roofline of the dominant kernel (track_kernel), this rank: algorithmic bytes / kernel time
    achieved = my_events * BYTES_PER_EVENT / (track_ms * 1e-3) / 1e9 if track_ms > 0 else 0.0
    achieved_min = achieved
    if dist is not None:
        t = torch.tensor([achieved], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        achieved_min = float(t.item())
    value = n_hist * args.steps / (dev_ms * 1e-3)
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full`
    # capture (bytes per history x histories per launch)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "track_kernel_dram.json")) as f:
            per_hist = float(json.load(f)["dram_bytes_per_history"])
        traffic = per_hist * (sum_over_ranks(float(sum(c1[k] - c0[k] for k in ("n_left", "n_right", "n_dead"))))
                              / max(sum_over_ranks(float(launches)), 1.0))
    except Exception:
        pass
    world_info = None
    if world > 1:
        if sw.trace is not None:
            with open(os.path.join(os.environ["MCB200_TRACE_DIR"], f"trace_rank{rank}.json"), "w") as f:
                json.dump({"columns": ["cycle", "simulate_ms", "track_ms_cum", "gather_ms",
                                       "start_ms", "finish_ms", "n_out_left", "n_out_right",
                                       "n_bank_after", "n_unborn_after"], "rows": sw.trace}, f)
        world_info = {"cycles": sw.cycles,
                      "t_simulate_s_max": max_over_ranks(sw.t_simulate),
                      "t_exchange_s_max": max_over_ranks(sw.t_exchange),
                      "track_ms_max": max_over_ranks(track_ms),
                      "track_ms_sum": sum_over_ranks(track_ms),
                      "track_ms_per_rank": per_rank(track_ms),
                      "events_per_rank": per_rank(float(my_events)),
                      "segments_per_rank": per_rank(float(
                          sum(c1[k] - c0[k] for k in ("n_left", "n_right", "n_dead")))),
                      "cuts": sw.cuts, "balanced": bool(args.balance and args.warmup > 1),
                      "overlap": sw.overlap,
                      "host_split_s_per_rank": {k: per_rank(v) for k, v in sw.t_parts.items()},
                      "ramp_from": args.ramp_from,
                      "note": "host wall-clock split of SlabWorld.spin over warm-up + timed steps; "
                              "track_ms = tracking-kernel time of the timed steps"}

    # ---- end-to-end arm: host buffers through the reference-facing interface ------------
    e2e = None
    if world == 1 and not args.no_e2e:
        n_e2e = min(n_hist, args.e2e_particles)
        host = torch.empty(n_e2e * PARTICLE_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
        dx = float(layer.dx)
        _abi.check(_abi.lib().mcb200_test_birth(local_rank, cfg.x_ini, float(np.float32(1.0 / n_e2e)),
                                                dx, n_e2e, 5127801, host.data_ptr()))
        e_layer = Layer(cfg.x_min, cfg.x_max, 0, cfg.nb_cells, cfg.particle_min_weight,
                        device=local_rank, sigs=cfg.sigs, absorption_rates=cfg.absorption_rates)
        wa = np.empty(cfg.nb_cells, dtype=np.float32)

        def e2e_step():
            # Layer.particles (host, 24-byte records) -> GPU -> weights_absorbed + counters back
            _abi.check(_abi.lib().mcb200_layer_push(e_layer._h, host.data_ptr(), n_e2e))
            c = e_layer.simulate(-1)
            _abi.check(_abi.lib().mcb200_layer_weights_absorbed(e_layer._h, wa.ctypes.data))
            return c

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(args.steps):
            ce = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        assert ce["nb_active"] == 0 and float(wa.sum()) > 0
        import ctypes
        e2e = {"value": n_e2e * args.steps / dt, "unit": "histories/s",
               "h2d_bytes_per_step": n_e2e * PARTICLE_DTYPE.itemsize,
               "d2h_bytes_per_step": int(wa.nbytes + ctypes.sizeof(_abi.Counts)),
               "histories_per_step": n_e2e,
               "interface": "mcb200_layer_push(host Particle[]) + mcb200_layer_simulate(-1) + "
                            "mcb200_layer_weights_absorbed (what Layer::simulate / cusimulate do)"}
        e_layer.close()
    elif world > 1 and not args.no_e2e:
        # N > 1: the public call sequence of a run -- register the source
        # (Layer::create_particles takes scalars, no particle buffer), spin, gather the tally to
        # the host like Worker::dump (src/worker.cpp:36-61) -- timed by the host clock, the
        # read-back inside the timed region.  The host-BUFFER arm (24-byte Particle records
        # pushed over PCIe) is what the N = 1 line measures.
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            step()
            wa = sw.gather_weights_absorbed()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t)
        import ctypes
        e2e = {"value": n_hist * args.steps / dt, "unit": "histories/s",
               "h2d_bytes_per_step": int(ctypes.sizeof(_abi.LayerDesc)),
               "d2h_bytes_per_step": int(8 * cfg.nb_cells + world * ctypes.sizeof(_abi.Counts)),
               "histories_per_step": n_hist,
               "interface": "decompose_domain / create_particles (scalars) + SlabWorld.spin + "
                            "gather_weights_absorbed to the host; host-buffer arm: see N = 1"}
        if rank == 0:
            assert wa is not None and float(wa.sum()) > 0
    elif world > 1:
        e2e = {"value": None, "unit": "histories/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 0, "note": "--no-e2e"}

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            rate, kind = cpu_reference_run(cfg, args.cpu_sample, cores)
            cpu = {"value": rate, "unit": "histories/s", "cores": cores, "kind": kind,
                   "sample": f"{args.cpu_sample} histories of the same workload, "
                             f"Layer::simulate(-1, {cores}) (OpenMP, one rank)"}
        except Exception as ex:  # the checker is optional for the number, never for the tests
            cpu = {"value": None, "unit": "histories/s", "cores": cores, "kind": "unavailable",
                   "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": "particle histories/s (whole box)", "value": value, "unit": "histories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "nb_cells": cfg.nb_cells, "histories_per_step": n_hist,
                       "particle_min_weight": cfg.particle_min_weight,
                       "events_per_history": events / (n_hist * args.steps),
                       "l2_policy": "inputs larger than L2: the source bank is "
                                    f"{n_hist * 24 / 1e9:.1f} GB of particle state per step",
                       "parallelism": "1 GPU" if world == 1 else
                                      f"domain decomposition, {world} sub-slabs "
                                      f"({'cuts balanced on measured tracking time' if args.balance and args.warmup > 1 else 'equal cell counts'}), "
                                      f"<= {args.per_cycle} source histories per cycle, escapees "
                                      + ("stored by the tracking kernel into the neighbour GPU's "
                                         "inbox over NVLink (CUDA IPC)" if sw.transport == "p2p"
                                         else "shipped with ncclSend/Recv")},
            "events_per_s": events / (dev_ms * 1e-3),
            "wall_s": wall,
            "roofline": {"bound": "hbm", "achieved": achieved_min, "peak": peak, "unit": "GB/s",
                         "frac": achieved_min / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "model": f"{BYTES_PER_EVENT} B/event x events / track_kernel time "
                                  f"({launches} launches, {track_ms / max(launches, 1):.3f} ms avg"
                                  f"{', min over ranks' if world > 1 else ''})"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "world": world_info,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        sw.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["mcb200", "reference"], default="mcb200")
    ap.add_argument("--particles", type=int, default=None, help="histories per step (override)")
    ap.add_argument("--per-cycle", type=int, default=1 << 25, dest="per_cycle")
    ap.add_argument("--ramp-from", type=int, default=1 << 20, dest="ramp_from",
                    help="source histories of the first cycle (doubling up to --per-cycle); 0 = flat")
    ap.add_argument("--transport", choices=["nccl", "p2p"], default="p2p",
                    help="N > 1: nccl = outbox -> ncclSend/Recv -> bank; p2p = the tracking kernel "
                         "stores escapees straight into the neighbour GPU's inbox over NVLink")
    ap.add_argument("--overlap", action="store_true",
                    help="keep the exchange of cycle c in flight under the tracking of cycle c+1")
    ap.add_argument("--no-balance", action="store_false", dest="balance",
                    help="keep the reference's equal-cell-count decomposition")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, dest="cpu_sample")
    ap.add_argument("--e2e-particles", type=int, default=HISTORIES_1GPU, dest="e2e_particles")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
