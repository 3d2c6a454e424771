"""C2 in its native form: g2048_step4 (all four moves of 2^20 boards, spawn + shaping) against g2048_step on the 4 Mi pairs.
CUDA events over graph replays, ring of buffers > L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import env  # noqa: E402

dev = torch.device("cuda:0")
env.lut(dev)
g = torch.Generator(device=dev).manual_seed(2048)
n = 1 << 20
ring = 6


def boards():
    e = torch.randint(1, 12, (n, 16), generator=g, device=dev, dtype=torch.int64)
    e[torch.rand((n, 16), generator=g, device=dev) < 0.30] = 0
    return (e << (torch.arange(16, device=dev) * 4)).sum(1)


ins = [boards() for _ in range(ring)]
reps = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 4, 2), generator=g, device=dev, dtype=torch.int64).to(torch.int32) for _ in range(ring)]
for shaping, replayed in ((True, False), (False, False), (True, True), (False, True)):
    outs = [dict(boards=torch.empty((n, 4), dtype=torch.int64, device=dev), points=torch.empty((n, 4), dtype=torch.int32, device=dev),
                 flags=torch.empty((n, 4), dtype=torch.uint8, device=dev), shaping=torch.empty((n, 4), dtype=torch.int64, device=dev) if shaping else None)
            for _ in range(ring)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(3):
            env.step4(ins[k % ring], seed=1, ctr=k, shaping=shaping, out=outs[k % ring], replay=reps[k % ring] if replayed else None)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for k in range(60):
                env.step4(ins[k % ring], seed=1, ctr=k, shaping=shaping, out=outs[k % ring], replay=reps[k % ring] if replayed else None)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    bpb = 8 + 4 * (8 + 4 + 1 + (8 if shaping else 0) + (8 if replayed else 0))
    print(f"step4 shaping={shaping} replayed_draws={replayed}: {us:.2f} us per 2^20 boards = {4 * n / us * 1e6:.4g} env-steps/s, {bpb} B/board = {n * bpb / us / 1e3:.1f} GB/s")
