"""A few g2048_step4 launches on the C2 workload (ncu target).  Usage: python tools/run_step4.py [launches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import env  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2048)
n = 1 << 20
e = torch.randint(1, 12, (n, 16), generator=g, device=dev, dtype=torch.int64)
e[torch.rand((n, 16), generator=g, device=dev) < 0.30] = 0
boards = (e << (torch.arange(16, device=dev) * 4)).sum(1)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    out = env.step4(boards, seed=2048, ctr=1 + k)
torch.cuda.synchronize()
print("ok", int(out["points"].sum()))
