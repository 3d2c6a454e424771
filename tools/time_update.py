"""Update time vs chunk size (samples per forward/backward chunk)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import trainer as tr  # noqa: E402

for chunk in (1 << 19, 1 << 20, 1 << 22, 1 << 23):
    cfg = tr.TrainConfig(hidden_dim=196, num_layers=2, envs=65536, horizon=128, zero_heads=False, chunk=chunk)
    t = tr.Trainer(cfg, torch.device("cuda:0"))
    t.train_step()
    t.train_step()
    print(f"chunk {chunk:9d}: update {t.times.update_ms:8.2f} ms  rollout {t.times.rollout_ms:6.2f} ms  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del t
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
