"""GameURM rollout kernels against the torch fp32 model (max |dlogp|, |dV|) and their C5 timing (debug / timing tool)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from g2048 import env, rollout  # noqa: E402
from g2048.policy import GameURM, GameURMConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=262144)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--check-envs", type=int, default=1000)
ap.add_argument("--precisions", default="x3,fp16")
ap.add_argument("--no-time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = GameURM(GameURMConfig(dropout=0.0)).to(dev).eval()
pol = rollout.pack_policy(model)


def masked_lp(logits, legal):
    illegal = ((legal.reshape(-1).long()[:, None] >> torch.arange(4, device=logits.device)) & 1) == 0
    return torch.masked_fill(logits, illegal, float("-inf")).log_softmax(-1)


for prec in a.precisions.split(","):
    boards = env.reset(a.check_envs, device=dev, seed=3, env0=0, ctr=0)
    buf = rollout.rollout(pol, boards, 6, seed=3, env0=0, ctr0=1, precision=prec)
    torch.cuda.synchronize()
    with torch.no_grad():
        logits, v = model(env.encode(buf.boards.reshape(-1)))
    want = masked_lp(logits, buf.legal)
    got = buf.logp.reshape(-1, 4)
    fin = torch.isfinite(want)
    print(f"[{prec}] finite pattern equal: {bool(torch.equal(torch.isfinite(got), fin))}; max |dlogp| = "
          f"{float((got[fin] - want[fin]).abs().max()):.3e}, max |dV| = {float((buf.value.reshape(-1) - v.squeeze(1)).abs().max()):.3e}", flush=True)
    if a.no_time:
        continue
    boards = env.reset(a.envs, device=dev, seed=1, env0=0, ctr=0)
    buf = rollout.RolloutBuffers.allocate(a.steps, a.envs, dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(2):
        ev0.record()
        rollout.rollout(pol, boards, a.steps, seed=1, env0=0, ctr0=1 + r * a.steps, out=buf, precision=prec)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        print(f"[{prec}] rollout {a.envs} envs x {a.steps} steps: {ms:.2f} ms = {ms / a.steps:.2f} ms per step, {a.envs * a.steps / ms * 1e3:.4g} env-steps/s", flush=True)
