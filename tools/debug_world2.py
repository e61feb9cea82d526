"""2-GPU diagnostics of the world kernel (developer tool).
   python tools/debug_world2.py local            # one process drives GPUs 0 and 1
   torchrun --nproc-per-node 2 tools/debug_world2.py ipc   # one process per GPU (CUDA IPC)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_mpi_b200 import _abi, configs  # noqa: E402
from mc_mpi_b200.worker import LocalBox, Worker, totals  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "local"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
cfg = configs.reference_default(n)
if mode == "local":
    box = LocalBox(cfg, 2, devices=[0, 1])
    box.set_option("max_run_ms", 20_000)
    box.set_option("stall_ms", 3_000)
    try:
        res = box.run()
        print("local ok", json.dumps(totals(res)))
    except _abi.McbError as e:
        print("local FAILED", e)
    box.close()
else:
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = Worker(cfg, device=local)
    w.r.set_option("max_run_ms", 20_000)
    w.r.set_option("stall_ms", int(sys.argv[3]) if len(sys.argv) > 3 else 3_000)
    stall = int(sys.argv[3]) if len(sys.argv) > 3 else 3_000
    for cuts in (None, [0, 540, 1000], [0, 560, 1000], [0, 720, 1000], [0, 500, 1000]):
        if cuts is not None:
            w.recut(cuts)
            w.r.set_option("max_run_ms", 20_000)
            w.r.set_option("stall_ms", stall)
        try:
            r = w.spin(n)
            print(f"rank {w.rank} cuts {cuts} ok ms={r['kernel_ms']:.2f} events={r['events']} "
                  f"util={r['events'] / max(r['lane_slots'], 1):.3f}", flush=True)
        except _abi.McbError as e:
            print(f"rank {w.rank} cuts {cuts} FAILED {e} :: {json.dumps(getattr(e, 'result', None))}", flush=True)
    dist.barrier()
    w.close()
    dist.destroy_process_group()
