"""One small launch of the GameURM rollout kernels and update ops, meant to run under `compute-sanitizer --tool memcheck`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch
from g2048 import env, rollout, urm_ops
from g2048.policy import GameURM, GameURMConfig
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = GameURM(GameURMConfig(dropout=0.0)).to(dev)
up = rollout.pack_policy(m.eval())
b = env.reset(37, device=dev, seed=3)
rollout.rollout(up, b, 2, seed=3, precision="x3")
rollout.rollout(up, b, 2, seed=3, precision="fp16")
m.train()
l, v = urm_ops.forward(m, env.encode(b))
(l.sum() + v.sum()).backward()
torch.cuda.synchronize()
print("san urm ok")
