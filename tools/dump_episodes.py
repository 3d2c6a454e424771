"""GPU side of the drop-in integration check: play a few games with play_games_batched (the module the
reference's train.py:30 imports) using the shipped checkpoint's weights and save the list[EpisodeData]
with CPU tensors -> gpurun_out/episodes_gpu.pt (committed as tests/golden/episodes_gpu.pt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import batched_rollout  # noqa: E402
from g2048 import policy  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "model_best.npz"))
m = policy.GameMLP(policy.MLPConfig(hidden_dim=int(g["hidden_dim"]), num_layers=int(g["num_layers"]), dropout=0.0))
m.load_state_dict(policy.load_state_dict_from_npz(g))
m = m.cuda().eval()
eps = batched_rollout.play_games_batched(m, num_games=6, max_steps=150, device=torch.device("cuda:0"), seed=11)
for ep in eps:
    for mv in ep["moves"]:
        mv["game_state"] = mv["game_state"].cpu()
        mv["points_possible"] = {k.value: v for k, v in mv["points_possible"].items()}   # enum of OUR module -> plain str
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
torch.save(eps, os.path.join(ROOT, "gpurun_out", "episodes_gpu.pt"))
print([len(e["moves"]) for e in eps], [e["total_points"] for e in eps])
