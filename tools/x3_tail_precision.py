"""Largest distance of the x3 rollout kernel's recorded log-probs / values / entropies from the torch fp32 policy (float64 reference
beside it) on the boards it recorded: python tools/x3_tail_precision.py [envs] [steps]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import env, rollout  # noqa: E402
from g2048.policy import GameMLP, MLPConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
for seed in (0, 1, 2):
    torch.manual_seed(seed)
    m = GameMLP(MLPConfig(hidden_dim=196, num_layers=2, dropout=0.0)).to(dev).eval()
    for p in m.parameters():
        p.data.mul_(1.0 + 0.5 * seed)                      # sharper policies with the seed
    boards = env.reset(B, device=dev, seed=seed)
    buf = rollout.rollout(rollout.pack_policy(m), boards, T, seed=seed, precision="x3")
    m64 = GameMLP(MLPConfig(hidden_dim=196, num_layers=2, dropout=0.0)).to(dev).double().eval()
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    worst = {"logp32": 0.0, "logp64": 0.0, "v32": 0.0, "ent64": 0.0}
    with torch.no_grad():
        for t in range(T):
            x = env.encode(buf.boards[t])
            valid = (buf.flags[t] & 0x80) != 0
            illegal = ((buf.legal[t].long()[:, None] >> torch.arange(4, device=dev)) & 1) == 0
            for name, model, xx in (("32", m, x), ("64", m64, x.double())):
                h_ = model.stem(xx)                     # (GameMLP.forward casts its input to float32, game.py:1150)
                for blk in model.backbone:
                    h_ = blk(h_)
                lg, v = model.action_head(h_), model.value_head(h_)
                ref = torch.masked_fill(lg, illegal, float("-inf")).log_softmax(-1)
                fin = torch.isfinite(ref) & valid[:, None]
                worst["logp" + name] = max(worst["logp" + name], float((buf.logp[t].double() - ref.double())[fin].abs().max()))
                if name == "32":
                    worst["v32"] = max(worst["v32"], float((buf.value[t] - v.squeeze(1))[valid].abs().max()))
                else:
                    pr = ref.exp()
                    ent = -(torch.where(fin, pr * ref, torch.zeros_like(pr))).sum(-1)
                    worst["ent64"] = max(worst["ent64"], float((buf.entropy[t].double() - ent)[valid].abs().max()))
    print(f"seed {seed}: max |dlogp| vs torch fp32 {worst['logp32']:.2e}, vs float64 {worst['logp64']:.2e}; |dV| vs fp32 {worst['v32']:.2e}; |dH| vs float64 {worst['ent64']:.2e}")
