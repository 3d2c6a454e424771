"""Small standalone driver for profiling the fused rollout kernel (ncu target)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from g2048 import env, rollout  # noqa: E402
from g2048.policy import GameMLP, GameURM, GameURMConfig, MLPConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=148 * 128)
ap.add_argument("--steps", type=int, default=32)
ap.add_argument("--hidden", type=int, default=196)
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--urm", action="store_true", help="GameURM (config #5) instead of GameMLP")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
if a.urm:
    model = GameURM(GameURMConfig(dropout=0.0)).to(dev).eval()
else:
    model = GameMLP(MLPConfig(hidden_dim=a.hidden, num_layers=a.layers, dropout=0.0)).to(dev).eval()
pol = rollout.pack_policy(model)
boards = env.reset(a.envs, device=dev, seed=1, env0=0, ctr=0)
buf = rollout.RolloutBuffers.allocate(a.steps, a.envs, dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(a.reps):
    ev0.record()
    rollout.rollout(pol, boards, a.steps, seed=1, env0=0, ctr0=1 + r * a.steps, out=buf, precision=a.precision)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"[{a.precision}] rollout {a.envs} envs x {a.steps} steps: {ms:.3f} ms, {a.envs * a.steps / ms * 1e3:.4g} env-steps/s")
