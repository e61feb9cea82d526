"""Multi-GPU parity of the persistent world run, under torchrun (one process per GPU):

    python -m torch.distributed.run --nproc-per-node K --master-addr 127.0.0.1 \
        tools/check_world.py [case] [cuts: equal|uneven] [windows]

The K-rank run must reproduce the ORACLE's single-layer result committed in
tests/golden/world_digest.json: SHA-256 of all 128 bits of every cell of the tally, events,
scatters, histories absorbed at the global borders / dead.  Every rank exits with the same code."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.worker import Worker  # noqa: E402

CASES = {
    "default_slab_2e6": lambda: configs.reference_default(2_000_000),
    "single_gpu_slab_2e6": lambda: configs.single_gpu_slab(2_000_000),
    "default_slab_1e5": lambda: configs.reference_default(100_000),
    "absorption_dominated_2e5": lambda: configs.absorption_dominated(200_000),
    "optically_thick_2e4": lambda: configs.optically_thick(20_000),
}


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "default_slab_2e6"
    cuts_kind = sys.argv[2] if len(sys.argv) > 2 else "equal"
    windows = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, K = dist.get_rank(), dist.get_world_size()
    with open(os.path.join(ROOT, "tests", "golden", "world_digest.json")) as f:
        digest = json.load(f)
    cfg = CASES[case]()
    cuts = None
    if cuts_kind == "uneven":
        wts = np.array([1.0 + 0.35 * ((3 * r) % 5 - 2) / 2 for r in range(K)])
        edges = np.concatenate([[0], np.cumsum(wts / wts.sum() * cfg.nb_cells)]).round().astype(int)
        edges[-1] = cfg.nb_cells
        cuts = edges.tolist()
    w = Worker(cfg, device=local, cuts=cuts, windows=windows)
    w.r.set_option("max_run_ms", 60_000)
    v = w.parity(case, digest)
    ok = v["tally_bit_exact"] and v["counts_exact"] and v["conservation_ok"] and v["kernel_error"] == 0
    if rank == 0:
        print(f"[world parity] K={K} cuts={w.cuts or 'equal'} windows={windows or 'auto'} {json.dumps(v)}",
              flush=True)
        print(f"[world parity] ok={ok}", flush=True)
    w.close()                      # every rank, before anybody exits
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
