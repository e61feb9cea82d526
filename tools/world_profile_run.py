"""A few whole runs of the persistent world kernel under torchrun with given cuts (developer
tool; rank 0 may be wrapped in ncu by tools/ncu_rank0.sh).
usage: world_profile_run.py <histories> <spins> [cuts json] [options json]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import configs  # noqa: E402
from mc_mpi_b200.worker import Worker, occupancy  # noqa: E402

n = int(float(sys.argv[1]))
spins = int(sys.argv[2])
cuts = json.loads(sys.argv[3]) if len(sys.argv) > 3 else None
opts = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = Worker(configs.reference_default(n), device=local, cuts=cuts, **opts)
w.r.set_option("max_run_ms", 60_000)
w.r.set_option("stall_ms", 20_000)
best = None
for k in range(spins):
    r = w.spin(n)
    print(f"rank {w.rank} spin {k}: kernel_ms={r['kernel_ms']:.2f} events={r['events']} "
          f"sent={r['sent_left'] + r['sent_right']} occupancy={occupancy(r):.3f}", flush=True)
    ms = w.all_ranks([r["kernel_ms"]], "max")[0]
    occ = [round(row[0], 3) for row in w.all_ranks([occupancy(r)], "table")]
    if k > 0 and (best is None or ms < best[0]):
        best = (ms, occ)
if w.rank == 0 and best:
    print(f"SUMMARY opts={opts} cuts={cuts} best_ms={best[0]:.2f} histories_per_s={n / best[0] * 1e3:.4g} "
          f"occupancy={best[1]}", flush=True)
dist.barrier()
w.close()
dist.destroy_process_group()
