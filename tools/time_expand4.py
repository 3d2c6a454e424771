"""expand4 at the C2 size under different ring sizes / graph lengths (why bench.py and time_env.py differ)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200")]
import torch
from g2048 import env
dev = torch.device("cuda:0"); env.lut(dev)
g = torch.Generator(device=dev).manual_seed(0)
n = 1 << 20
def boards():
    e = torch.randint(1, 12, (n, 16), generator=g, device=dev, dtype=torch.int64)
    e[torch.rand((n, 16), generator=g, device=dev) < 0.30] = 0
    return (e << (torch.arange(16, device=dev) * 4)).sum(1)
for ring, reps in ((16, 50), (12, 200), (12, 50), (16, 200), (24, 200), (6, 200)):
    ins = [boards() for _ in range(ring)]
    outs = [dict(succ=torch.empty((n, 4), dtype=torch.int64, device=dev), points=torch.empty((n, 4), dtype=torch.int32, device=dev),
                 legal=torch.empty(n, dtype=torch.uint8, device=dev), max_tile=None) for _ in range(ring)]
    fn = lambda k: env.expand4(ins[k % ring], out=outs[k % ring])
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for k in range(reps): fn(k)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        print(f"ring={ring} reps={reps}: {e0.elapsed_time(e1) * 1e3 / reps:.2f} us/launch")
    del ins, outs, graph
    torch.cuda.empty_cache()
