"""One small launch of every kernel, meant to run under `compute-sanitizer --tool memcheck`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import env, fused, ppo, rollout  # noqa: E402
from g2048.policy import GameMLP, GameURM, GameURMConfig, MLPConfig  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
n = (1 << 17) + 333
boards = env.reset(n, device=dev, seed=1)
acts = torch.randint(0, 4, (n,), device=dev, dtype=torch.uint8)
r = env.step(boards, acts, seed=1, ctr=1)                       # staged step kernel
env.step(boards[:1000], acts[:1000], seed=1, ctr=1, shaping=False)  # direct
ex = env.expand4(boards)
env.expand4(boards[:777], want_max_tile=True)
env.potentials(boards[:5000]); env.encode(boards[:5000])
env.potentials_ext(boards[:5000], ex["succ"][:5000, 0].contiguous())
env.augment(boards[:999], r["boards"][:999], acts[:999], ex["legal"][:999], torch.randn((999, 4), device=dev),
            torch.randint(0, 5, (999,), device=dev, dtype=torch.uint8))
mlp = GameMLP(MLPConfig(hidden_dim=196, num_layers=2, dropout=0.0)).to(dev)
pol = rollout.pack_policy(mlp)
for prec in ("fp32", "bf16"):
    b = env.reset(300, device=dev, seed=2)
    buf = rollout.rollout(pol, b, 5, seed=2, precision=prec)
up = rollout.pack_policy(GameURM(GameURMConfig(dropout=0.0)).to(dev))
rollout.rollout(up, env.reset(20, device=dev, seed=3), 2, seed=3)
adv = ppo.rtg_advantage(buf.points, buf.shaping, buf.flags, buf.value, gamma=0.99, w_points=0.1, w_mono=1.0, w_empt=0.0,
                        mu_c=0.0, stddev=1.0, want_raw=True)
logits, v = fused.mlp_forward(mlp.train(), env.encode(buf.boards.reshape(-1)))
loss, st = ppo.ppo_loss(logits, v, buf.logp.reshape(-1, 4), buf.actions.reshape(-1), buf.legal.reshape(-1),
                        adv["adv"].reshape(-1), adv["g_norm"].reshape(-1), flags=buf.flags.reshape(-1))
loss.backward()
torch.cuda.synchronize()
print("sanitize_smoke ok", float(loss))
