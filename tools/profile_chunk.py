"""Kernel-level time breakdown of one fused update chunk (torch profiler, CUDA time per kernel)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200"), os.path.join(ROOT, "tests")]
from g2048 import update
from test_update_fused_gpu import _model, _boards, _samples
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
m = _model(196, 2, 1)
boards = _boards(n, 2)
old, actions, legal, adv, g_norm = _samples(n, 3)
packed = update.pack(m)
for _ in range(2):
    m.zero_grad(); update.loss_and_grads(m, boards, actions, legal, old, adv, g_norm, packed=packed)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        m.zero_grad(); update.loss_and_grads(m, boards, actions, legal, old, adv, g_norm, packed=packed)
    torch.cuda.synchronize()
rows = sorted(((e.device_time_total / 3e3, e.count // 3, e.key[:90]) for e in prof.key_averages()), reverse=True)
for ms, cnt, key in rows[:10]:
    print(f"{ms:8.3f} ms x{cnt:2d}  {key}")
