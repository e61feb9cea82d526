"""A few whole runs of the persistent world kernel under torchrun with given cuts (developer
tool; rank 0 may be wrapped in ncu by tools/ncu_rank0.sh).
usage: world_profile_run.py <histories> <spins> [cuts json]"""
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
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = Worker(configs.reference_default(n), device=local, cuts=cuts)
w.r.set_option("max_run_ms", 60_000)
w.r.set_option("stall_ms", 20_000)
for k in range(spins):
    r = w.spin(n)
    print(f"rank {w.rank} spin {k}: kernel_ms={r['kernel_ms']:.2f} events={r['events']} "
          f"sent={r['sent_left'] + r['sent_right']} occupancy={occupancy(r):.3f}", flush=True)
dist.barrier()
w.close()
dist.destroy_process_group()
