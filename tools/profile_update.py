"""Where does the update's time go at C3 scale?  torch.profiler over one train step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from g2048 import trainer as tr  # noqa: E402

cfg = tr.TrainConfig(hidden_dim=196, num_layers=2, envs=65536, horizon=int(os.environ.get("H", "128")), zero_heads=False)
t = tr.Trainer(cfg, torch.device("cuda:0"))
t.train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t.train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
print(t.times)
