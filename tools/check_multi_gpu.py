"""Multi-GPU parity check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nproc-per-node K --master-addr 127.0.0.1 tools/check_multi_gpu.py [particles] [per_cycle] [config]
The K-GPU domain-decomposed run (global dx) must reproduce the 1-GPU run bit for bit: every
digit of the exact tally, events, scatters, escape counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.layer import decompose_domain  # noqa: E402
from mc_mpi_b200.world import SlabWorld  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
    per_cycle = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 19
    name = sys.argv[3] if len(sys.argv) > 3 else "default_slab"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, K = dist.get_rank(), dist.get_world_size()
    cfg = configs.BY_NAME[name]().with_particles(n)
    cuts = None
    if len(sys.argv) > 4 and sys.argv[4] == "uneven":
        # deliberately unequal sub-slabs + ramped source schedule: same bits expected
        wts = np.array([1.0 + 0.35 * ((3 * r) % 5 - 2) / 2 for r in range(K)])
        edges = np.concatenate([[0], np.cumsum(wts / wts.sum() * cfg.nb_cells)]).round().astype(int)
        edges[-1] = cfg.nb_cells
        cuts = edges.tolist()
    transport = sys.argv[5] if len(sys.argv) > 5 else "nccl"
    w = SlabWorld(cfg, device=local, nb_particles_per_cycle=per_cycle, cuts=cuts,
                  ramp_from=(per_cycle // 8 or 1) if cuts else None, transport=transport)
    s = w.spin()
    wa = w.gather_weights_absorbed()
    stats = torch.tensor([s["events"], s["scatters"], s["migrations_out"],
                          s["n_left"] if rank == 0 else 0, s["n_right"] if rank == K - 1 else 0,
                          s["n_dead"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(stats)
    if rank == 0:
        one = decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, n,
                               cfg.particle_min_weight, device=local, sigs=cfg.sigs,
                               absorption_rates=cfg.absorption_rates)
        c = one.simulate(-1)
        w1 = one.weights_absorbed_f64
        ok = np.array_equal(w1, wa)
        ev, sc, mig, nl, nr, nd = stats.tolist()
        same_counts = (ev, sc, nl, nr, nd) == (c["events"], c["scatters"], c["n_left"], c["n_right"], c["n_dead"])
        if not ok or not same_counts:
            bad = np.nonzero(w1 != wa)[0]
            print("[multi-gpu parity] counts K-GPU", (ev, sc, nl, nr, nd), "1-GPU",
                  (c["events"], c["scatters"], c["n_left"], c["n_right"], c["n_dead"]))
            print("[multi-gpu parity] differing cells:", len(bad), bad[:10],
                  "max rel", float(np.max(np.abs(w1 - wa) / w1)) if len(bad) else 0.0,
                  "sum K", float(wa.sum()), "sum 1", float(w1.sum()))
        ok &= same_counts
        print(f"[multi-gpu parity] cuts={w.cuts} transport={w.transport}")
        print(f"[multi-gpu parity] K={K} config={cfg.name} histories={n} cycles={s['cycles']} "
              f"migrations/history={mig / n:.3f} events={ev} tally_bit_exact={ok}", flush=True)
        one.close()
        if w.transport != transport:
            print(f"[multi-gpu parity] asked for transport={transport}, ran {w.transport}", flush=True)
            ok = False
    # every rank learns the verdict and closes the world (unmapping peers is collective) before
    # anybody exits: a failing rank 0 must not leave the others waiting in a barrier
    flag = torch.tensor([1 if (rank != 0 or ok) else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    w.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
