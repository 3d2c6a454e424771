"""Sweep of HostStepper's chunk size / stream count on the C2 workload (host pinned buffers, PCIe both ways)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200")]
import torch
import bench
from g2048 import env

dev = torch.device("cuda:0")
env.lut(dev)
N = bench.N_TRANS
hb, ha = bench.c2_transitions(4242)
h_boards, h_actions = torch.from_numpy(hb).pin_memory(), torch.from_numpy(ha).pin_memory()
h_out = dict(boards=torch.empty(N, dtype=torch.int64).pin_memory(), points=torch.empty(N, dtype=torch.int32).pin_memory(),
             flags=torch.empty(N, dtype=torch.uint8).pin_memory(), shaping=torch.empty(N, dtype=torch.int64).pin_memory())
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    for streams in (2, 3, 4):
        st = env.HostStepper(N, device=dev, chunk=chunk, streams=streams)
        for w in range(3):
            st.step(h_boards, h_actions, h_out, seed=1, ctr=w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(20):
            st.step(h_boards, h_actions, h_out, seed=1, ctr=10 + k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"chunk=2^{chunk.bit_length() - 1} streams={streams}: {ms:.3f} ms/step  {N / ms * 1e3:.3e} env-steps/s  D2H {88.08 / ms:.1f} GB/s")
