"""Driver with the reference's command line (src/main.cpp:15-96):

    python -m mc_mpi_b200.main <config.yaml> [nvl]                         # one GPU
    python -m torch.distributed.run --nproc-per-node K -m mc_mpi_b200.main <config.yaml> nvl

reads the same flat `key: value` file (src/yaml_loader.cpp:4-41, keys of
options_from_config, src/worker.cpp:315-333), runs the slab domain-decomposed one
sub-slab per GPU, prints ONE line -- the wall time of the run, like main.cpp:87-91 --
and writes what Worker::dump writes (src/worker.cpp:36-61): out/config.yaml
(:218-241), out/weights.csv (`proc, x, weight`, :287-313) and ./WA.out (rank 0's
slice, src/layer.cpp:363-380), so that check.py and tex/report work unchanged.

The comm-mode argument of the reference (sync | async | rma) selects an MPI strategy;
here there is one transport, NVLink via NCCL ("nvl"); the MPI names are accepted and
mean the same thing.  Tuning keys that only concern the MPI workers (cycle_time,
statistics_cycle_time, nthread, buffer_size) are read, echoed and otherwise unused.
Extensions (optional keys): sigs_file / absorption_file -- whitespace-separated global
per-cell tables replacing the hard-coded ones (src/layer.cpp:53-63); balance: 1 --
place the cuts on measured work instead of equal cell counts.
"""
from __future__ import annotations

import os
import socket
import sys
import time

import numpy as np

from . import configs as _configs

KEYS_INT = ("nb_cells", "nb_particles", "nb_particles_per_cycle", "nthread")
KEYS_FLOAT = ("x_min", "x_max", "x_ini", "particle_min_weight", "cycle_time",
              "statistics_cycle_time")


def load_config(path: str) -> dict:
    """YamlLoader (src/yaml_loader.cpp:4-41): one `key: value` per line, no nesting."""
    try:
        lines = open(path).read().splitlines()
    except OSError:
        sys.stderr.write(f"Couldn't open Yaml File {path}.\n")
        raise SystemExit(1)
    raw = {}
    for line in lines:
        if not line.strip():
            continue
        if ":" not in line:
            sys.stderr.write(f"Yaml File {path} was not correctly formatted.\n")
            raise SystemExit(1)
        k, v = line.split(":", 1)
        raw[k.strip()] = v.strip()
    opt = {}
    for k in KEYS_INT:      # std::stoi: the integer prefix
        opt[k] = int(float(raw[k])) if k in raw else (1 if k == "nthread" else None)
    for k in KEYS_FLOAT:    # std::stod, then narrowed to real_t where the struct holds a float
        opt[k] = float(raw[k]) if k in raw else 0.0
    for k in ("nb_cells", "nb_particles"):
        if opt[k] is None:
            raise SystemExit(f"{path}: missing key {k}")
    if opt["nb_particles_per_cycle"] is None:
        opt["nb_particles_per_cycle"] = 1 << 24
    opt["buffer_size"] = int(float(raw.get("buffer_size", 0)))
    opt["sigs_file"] = raw.get("sigs_file")
    opt["absorption_file"] = raw.get("absorption_file")
    opt["balance"] = int(float(raw.get("balance", 0)))
    return opt


def dump_config(path: str, opt: dict, world_size: int) -> None:
    """Worker::dump_config (src/worker.cpp:218-241) through YamlDumper's formats."""
    with open(path, "w") as f:
        f.write("# Read from config\n")
        f.write(f"nb_cells: {opt['nb_cells']}\n")
        for k in ("x_min", "x_max", "x_ini", "particle_min_weight"):
            f.write(f"{k}: {float(np.float32(opt[k])):.18e}\n")
        f.write(f"nb_particles: {opt['nb_particles']}\n")
        f.write(f"buffer_size: {opt['buffer_size']}\n")
        f.write(f"cycle_time: {opt['cycle_time']:.18e}\n")
        f.write(f"nb_particles_per_cycle: {opt['nb_particles_per_cycle']}\n")
        f.write(f"nthread: {opt['nthread']}\n")
        f.write(f"statistics_cycle_time: {opt['statistics_cycle_time']:.18e}\n")
        f.write("\n# Other values\n")
        f.write(f"world_size: {world_size}\n")
        f.write(f"hostname: {socket.gethostname()}\n")


def dump_weights_absorbed(path: str, weights: np.ndarray, cuts, dx: np.float32) -> None:
    """Worker::dump_weights_absorbed (src/worker.cpp:287-313): proc, dx*(i+0.5), weight/dx."""
    w32 = weights.astype(np.float32)
    with open(path, "w") as f:
        f.write("proc, x, weight\n")
        proc = 0
        for i in range(len(w32)):
            while proc < len(cuts) - 2 and cuts[proc + 1] == i:
                proc += 1
            f.write(f"{proc}, {float(dx) * (i + 0.5):.18e}, {float(w32[i] / dx):.18e}\n")


def dump_WA(path: str, weights: np.ndarray, dx: np.float32, x_min: float = 0.0) -> None:
    """Layer::dump_WA (src/layer.cpp:363-380): "%.4e %.3e" of the cell centre and tally / dx,
    in the reference's float arithmetic."""
    w32 = weights.astype(np.float32)
    with open(path, "w") as f:
        for i in range(len(w32)):
            x_mid = np.float32(np.float32(x_min) + np.float32(i) * dx) + 0.5 * float(dx)
            f.write("%.4e %.3e\n" % (float(x_mid), float(np.float32(w32[i] / dx))))


def dump_stats(path: str, rows_by_rank) -> None:
    """Worker::write_file (src/worker.cpp:63-181) with Timer::State's formats (src/timer.cpp:41-53):
    `rank, starttime, endtime, time_comp, time_send, time_recv, time_idle, nb_cycles, ` -- one row
    per rank and statistics window.  nb_cycles counts the terminating cycle too (the reference
    breaks out just before its counter)."""
    with open(path, "w") as f:
        f.write("rank, starttime, endtime, time_comp, time_send, time_recv, time_idle, nb_cycles, \n")
        for rank, rows in enumerate(rows_by_rank):
            for row in rows:
                f.write(f"{rank}, " + "".join(f"{v:.18e}, " for v in row[:6]) + f"{int(row[6])}, \n")


def slab_config(opt: dict) -> _configs.SlabConfig:
    def table(path):
        if not path:
            return None
        t = np.loadtxt(path, dtype=np.float32).reshape(-1)
        if t.size != opt["nb_cells"]:
            raise SystemExit(f"{path}: expected {opt['nb_cells']} entries, found {t.size}")
        return t
    return _configs.SlabConfig(
        "config.yaml", opt["nb_cells"], opt["nb_particles"],
        float(np.float32(opt["particle_min_weight"])), float(np.float32(opt["x_min"])),
        float(np.float32(opt["x_max"])), float(np.float32(opt["x_ini"])),
        sigs=table(opt["sigs_file"]), absorption_rates=table(opt["absorption_file"]))


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 1 or len(argv) > 2 or (len(argv) == 2 and argv[1] not in ("nvl", "sync", "async", "rma")):
        sys.stderr.write("Usage: python -m mc_mpi_b200.main <config.yaml> [nvl]\n"
                         "  one process per GPU (torch.distributed.run); the reference's comm\n"
                         "  modes sync | async | rma are accepted and all mean NVLink/NCCL\n")
        return 1
    opt = load_config(argv[0])
    cfg = slab_config(opt)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch

    from .layer import decompose_domain
    torch.cuda.set_device(local)
    f32 = np.float32
    dx = f32(f32(cfg.x_max) - f32(cfg.x_min)) / f32(cfg.nb_cells)
    if world == 1:
        layer = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells,
                                 cfg.nb_particles, cfg.particle_min_weight, device=local,
                                 sigs=cfg.sigs, absorption_rates=cfg.absorption_rates)
        torch.cuda.synchronize()
        t0, w0 = time.perf_counter(), time.time()
        c = layer.simulate(-1)                  # Worker::spin
        elapsed = time.perf_counter() - t0
        weights, cuts, lay0 = layer.weights_absorbed_f64, [0, cfg.nb_cells], layer
        # phase timers (src/timer.cpp:57-102): Computation = device time of the tracking kernels
        # (CUDA events), Idle = the rest of the wall time (launch gaps, read-back); no exchange
        comp = min(c["track_ms"] * 1e-3, elapsed)
        stat_rows = [[(w0, w0 + elapsed, comp, 0.0, 0.0, elapsed - comp, c["launches"])]]
    else:
        import torch.distributed as dist

        from .worker import Worker, occupancy
        from .world import balanced_cuts
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # one persistent kernel per GPU for the whole run (csrc/mcb_world*.cu): the escapee
        # exchange and the termination count happen on the device
        wk = Worker(cfg, device=local)
        if opt["balance"]:
            # a short pilot run places the cuts where the measured lane occupancy balances
            pilot = max(min(cfg.nb_particles // 20, 20_000_000), 1000)
            r = wk.spin(pilot)
            cost = [row[0] for row in wk.all_ranks([occupancy(r)], "table")]
            equal = [k * (cfg.nb_cells // world) + min(k, cfg.nb_cells % world) for k in range(world + 1)]
            wk.recut(balanced_cuts(equal, cost, cfg.nb_cells))
        dist.barrier()
        torch.cuda.synchronize()
        t0, w0 = time.perf_counter(), time.time()
        r = wk.spin()                           # Worker::spin
        dist.barrier()
        elapsed = time.perf_counter() - t0
        weights = wk.gather_weights_absorbed()
        cuts = wk.cuts or [k * (cfg.nb_cells // world) + min(k, cfg.nb_cells % world)
                           for k in range(world + 1)]
        # phase timers from the DEVICE: the kernel's run time (CUDA events) splits into the time
        # its lanes carried a history (Computation; the escapee stores into the neighbour GPU --
        # the reference's Send -- are instructions of that kernel, the receive is a load from
        # local memory: both 0 here) and the time they waited for neighbours / the end (Idle)
        kern = r["kernel_ms"] * 1e-3
        busy = kern * occupancy(r)
        rows = wk.all_ranks([w0, w0 + elapsed, busy, 0.0, 0.0, max(elapsed - busy, 0.0), 1], "table")
        stat_rows = [[tuple(row)] for row in rows]
        lay0 = None
    if rank == 0:
        print(f"{elapsed:f}")                   # main.cpp:91
        os.makedirs("out", exist_ok=True)       # Worker::dump, src/worker.cpp:36-61
        dump_config(os.path.join("out", "config.yaml"), opt, world)
        dump_weights_absorbed(os.path.join("out", "weights.csv"), weights, cuts, dx)
        dump_stats(os.path.join("out", "stats.csv"), stat_rows)
        if lay0 is not None:
            lay0.dump_WA("WA.out")
        else:
            dump_WA("WA.out", weights[cuts[0]:cuts[1]], dx)   # rank 0's slice, src/layer.cpp:363-380
    if world > 1:
        import torch.distributed as dist
        wk.close()          # unmap the neighbours' exchange blocks before anybody frees one
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
