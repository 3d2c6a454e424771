"""Raw pinned-memory copy bandwidth of the box (device->host, host->device, and both at once in the proportion of the e2e step):
the ceiling `bench.py`'s `e2e` figure is measured against.  Usage: python tools/pcie_bw.py"""
import torch, time
dev = torch.device("cuda:0")
for mb in (8, 32, 88, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{direction} {mb} MB: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
# both directions at once on two streams
n = 88 << 20
d1 = torch.empty(n, dtype=torch.uint8, device=dev); h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(38 << 20, dtype=torch.uint8, device=dev); h2 = torch.empty(38 << 20, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 100
print(f"duplex 88 MB d2h + 38 MB h2d: {ms:.3f} ms per pair -> d2h {88 * 1.048576 / ms:.1f} GB/s")
