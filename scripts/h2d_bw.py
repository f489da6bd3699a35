"""Host->device bandwidth from page-locked memory: one stream vs two concurrent streams, several chunk sizes."""
import time
import torch

dev = torch.device("cuda:0")
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
s = [torch.cuda.Stream(), torch.cuda.Stream()]
for chunk_mb in (16, 64, 256, 2048):
    ch = chunk_mb << 20
    for nstreams in (1, 2):
        for rep in range(2):
            torch.cuda.synchronize()
            t = time.perf_counter()
            for i, off in enumerate(range(0, n, ch)):
                with torch.cuda.stream(s[i % nstreams]):
                    d[off:off + ch].copy_(h[off:off + ch], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
        print(f"chunk {chunk_mb:5d} MB, {nstreams} stream(s): {n / dt / 1e9:6.1f} GB/s")
