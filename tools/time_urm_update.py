"""GameURM train step (rollout on the fused kernel + update): the update on this library's kernels (urm_ops) against the torch
mirror's own forward / backward (ATen + cuBLAS).  Usage: python tools/time_urm_update.py [envs] [horizon]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2048-ppo_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from g2048 import trainer as tr  # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
horizon = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
for mode in os.environ.get("URM_MODES", "ops,autograd,autograd_tf32").split(","):
    cfg = tr.TrainConfig(model_type="urm", envs=envs, horizon=horizon, zero_heads=False, warmup_steps=0,
                         urm_update="ops" if mode == "ops" else "autograd", update_matmul="tf32" if mode.endswith("tf32") else "fused")
    t = tr.Trainer(cfg, dev)
    t.train_step()
    ms = []
    for _ in range(3):
        t.train_step()
        ms.append((t.times.rollout_ms, t.times.update_ms))
    ro, up = min(m[0] for m in ms), min(m[1] for m in ms)
    n = envs * horizon
    print(f"[{mode}] {envs} envs x {horizon} steps = {n} samples: rollout {ro:.1f} ms, update {up:.1f} ms = {n / up * 1e3:.4g} samples/s", flush=True)
