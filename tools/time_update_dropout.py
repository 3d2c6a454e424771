"""Update time of one C3-like train step with and without Dropout(0.1) in the fused update (Trainer, 65 536 envs x 128 steps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import trainer as tr  # noqa: E402

cfg = tr.TrainConfig(hidden_dim=196, num_layers=2, envs=65536, horizon=128, zero_heads=False)
t = tr.Trainer(cfg, torch.device("cuda:0"))
for p in (0.0, 0.1, 0.0, 0.1):
    cfg.dropout = p
    t.train_step()
    t.train_step()
    print(f"dropout {p}: update {t.times.update_ms:8.2f} ms ({t.times.update_ms / 2:.2f} per 4 Mi-sample chunk)  rollout {t.times.rollout_ms:6.2f} ms")
