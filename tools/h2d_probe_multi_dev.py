"""Developer: concurrent host->device bandwidth of all ranks (torchrun), page-locked source, 77 MB
copies: what the box gives N GPUs pulling at the same time (the sweep's upload pattern)."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total, chunk = 1 << 30, 77 << 20
host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dst = torch.empty(total, dtype=torch.uint8, device="cuda")
res = {}
for mode in ("alone", "together"):
    best = 1e9
    for rep in range(3):
        if mode == "together":
            dist.barrier(device_ids=[local])
        else:
            for r in range(dist.get_world_size()):  # one rank at a time
                dist.barrier(device_ids=[local])
                if r == dist.get_rank():
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for off in range(0, total, chunk):
                        n = min(chunk, total - off)
                        dst[off:off + n].copy_(host[off:off + n], non_blocking=True)
                    torch.cuda.synchronize()
                    best = min(best, time.perf_counter() - t0)
                dist.barrier(device_ids=[local])
            continue
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for off in range(0, total, chunk):
            n = min(chunk, total - off)
            dst[off:off + n].copy_(host[off:off + n], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[mode] = total / best / 1e9
out = [None] * dist.get_world_size()
dist.all_gather_object(out, res)
if dist.get_rank() == 0:
    for r, o in enumerate(out):
        print(f"rank {r}: alone {o['alone']:.1f} GB/s, all ranks together {o['together']:.1f} GB/s")
    print("sum together %.1f GB/s" % sum(o["together"] for o in out))
dist.destroy_process_group()
