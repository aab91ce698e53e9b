"""Developer: host->device copy bandwidth from page-locked memory on this box: one stream vs several,
chunk sizes, and a kernel-free zero-copy comparison is left to the library's own numbers."""
import time
import torch

assert torch.cuda.is_available()
dev = torch.device("cuda:0")
total = 1 << 30
host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dst = torch.empty(total, dtype=torch.uint8, device=dev)
for streams in (1, 2, 4):
    for chunk_mb in (8, 32, 77, 256):
        chunk = chunk_mb << 20
        ss = [torch.cuda.Stream() for _ in range(streams)]
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            off, i = 0, 0
            while off < total:
                n = min(chunk, total - off)
                with torch.cuda.stream(ss[i % streams]):
                    dst[off:off + n].copy_(host[off:off + n], non_blocking=True)
                off += n
                i += 1
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"streams {streams} chunk {chunk_mb:4d} MB: {total / best / 1e9:6.1f} GB/s")
