"""Size sweep of the env kernels (CUDA events, ring of buffers > L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import env  # noqa: E402

dev = torch.device("cuda:0")
env.lut(dev)
g = torch.Generator(device=dev).manual_seed(0)


def boards(n):
    e = torch.randint(1, 12, (n, 16), generator=g, device=dev, dtype=torch.int64)
    e[torch.rand((n, 16), generator=g, device=dev) < 0.30] = 0
    return (e << (torch.arange(16, device=dev) * 4)).sum(1)


for logn in (18, 20, 22, 24):
    n = 1 << logn
    ring = max(2, (1 << 30) // (n * 57) + 1)
    ring = min(ring, 16)
    ins = [boards(n) for _ in range(ring)]
    outs = [dict(succ=torch.empty((n, 4), dtype=torch.int64, device=dev), points=torch.empty((n, 4), dtype=torch.int32, device=dev),
                 legal=torch.empty(n, dtype=torch.uint8, device=dev), max_tile=None) for _ in range(ring)]
    acts = [torch.randint(0, 4, (n,), generator=g, device=dev, dtype=torch.uint8) for _ in range(ring)]
    souts = [dict(boards=torch.empty(n, dtype=torch.int64, device=dev), points=torch.empty(n, dtype=torch.int32, device=dev),
                  flags=torch.empty(n, dtype=torch.uint8, device=dev), shaping=torch.empty(n, dtype=torch.int64, device=dev))
             for _ in range(ring)]
    for name, fn, bytes_per in (
            ("expand4", lambda k: env.expand4(ins[k % ring], out=outs[k % ring]), 57),
            ("step+shaping", lambda k: env.step(ins[k % ring], acts[k % ring], seed=1, ctr=k, out=souts[k % ring]), 30)):
        for k in range(5):
            fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for k in range(reps):
            fn(k)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        # the same launches replayed from one CUDA graph (no host launch cost in the timed region)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn(0)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for k in range(reps):
                fn(k)
        graph.replay()
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        gus = e0.elapsed_time(e1) * 1e3 / reps
        print(f"{name:13s} n=2^{logn} ring={ring:2d}: {us:9.2f} us/launch eager, {gus:9.2f} us/launch graph  {n / gus * 1e6:.3e} units/s  {n * bytes_per / gus / 1e3:8.1f} GB/s")
    del ins, outs, acts, souts
    torch.cuda.empty_cache()
