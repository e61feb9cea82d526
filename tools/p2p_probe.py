"""NCCL point-to-point host-side cost probe (developer tool): torchrun --nproc-per-node K"""
import os, sys, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r, K = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
for mb in (1, 32, 256):
    n = mb << 20
    sl, sr = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    rl, rr = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    for mode in ("batch", "single", "a2a"):
        tp = tw = 0.0
        for it in range(12):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "batch":
                ops = []
                if r > 0: ops += [dist.P2POp(dist.isend, sl, r - 1), dist.P2POp(dist.irecv, rl, r - 1)]
                if r + 1 < K: ops += [dist.P2POp(dist.isend, sr, r + 1), dist.P2POp(dist.irecv, rr, r + 1)]
                reqs = dist.batch_isend_irecv(ops)
            elif mode == "single":
                reqs = []
                if r > 0: reqs += [dist.isend(sl, r - 1), dist.irecv(rl, r - 1)]
                if r + 1 < K: reqs += [dist.isend(sr, r + 1), dist.irecv(rr, r + 1)]
            else:
                inp = [torch.empty(0, dtype=torch.uint8, device=dev) for _ in range(K)]
                out = [torch.empty(0, dtype=torch.uint8, device=dev) for _ in range(K)]
                if r > 0: inp[r - 1], out[r - 1] = sl, rl
                if r + 1 < K: inp[r + 1], out[r + 1] = sr, rr
                reqs = [dist.all_to_all(out, inp, async_op=True)]
            t1 = time.perf_counter()
            for q in reqs: q.wait()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if it >= 2:
                tp += t1 - t0; tw += t2 - t1
        if r == min(1, K - 1):
            print(f"{mode:6s} {mb:4d} MB/dir: post {tp/10*1e3:7.3f} ms  wait {tw/10*1e3:7.3f} ms  -> {2*n/((tp+tw)/10)/1e9:6.1f} GB/s per GPU (out)", flush=True)
dist.destroy_process_group()
