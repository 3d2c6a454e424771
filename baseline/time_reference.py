"""Times the UNMODIFIED Python reference (staged under baseline/_ref by baseline/stage_reference.sh) on this host's cores:
BASELINE.md section 3, B1-B5.

  B1  game.Game2048.simulate_move   (game.py:121-160)   on the C2 board distribution, all four directions, every core
  B2  game.Game2048.step            (game.py:952-1030)  with game.random replaced by a replay RNG, every core
  B3  train.play_game_for_episode   (train.py:213-345)  GameMLP h=196, eval mode, single process (how the reference runs it)
  B4  train.calculate_advantage + train.model_optimize_step (train.py:651-904, 414-642) on B3's episodes, batch size 4 and one
      large batch, single process
  B5  config #1, the reference's own CLI on CPU: train.py train --model-type mlp -h 196 --batch-size=4 + README flags, a bounded
      number of steps (--train-steps; the config says 200), rates from the run's own JSONL log

Prints one JSON object.  Runs in its own process (bench.py launches it with subprocess) so that the worker pool never
forks a CUDA context; nothing of this repo's engine is imported: only numpy, torch and the reference's own modules.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "batched_rollout" not in sys.modules:            # train.py:30 imports a module the reference does not ship
        stub = types.ModuleType("batched_rollout")
        stub.play_games_batched = None
        sys.modules["batched_rollout"] = stub
    import game  # noqa: F401
    return game


def _boards(n, seed):
    """C2 distribution: exponents 1..11, each cell emptied with probability 0.30."""
    import numpy as np
    rng = np.random.default_rng(seed)
    e = rng.integers(1, 12, (n, 16))
    e[rng.random((n, 16)) < 0.30] = 0
    return e.reshape(n, 4, 4).tolist()


class _Replay:
    """Stand-in for the `random` module inside game.py, fed (u0, u1) pairs like the kernels' replay tensor."""

    def __init__(self, pairs):
        self.pairs, self.i, self.cur = pairs, 0, None

    def choice(self, seq):
        self.cur = self.pairs[self.i % len(self.pairs)]
        self.i += 1
        return seq[(int(self.cur[0]) * len(seq)) >> 32]

    def random(self):
        return int(self.cur[1]) / 2 ** 32


def _worker_simulate(args):
    seed, seconds = args
    game = _import_reference()
    boards = _boards(2048, seed)
    dirs = list(game.Direction)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for g in boards:
            for d in dirs:
                game.Game2048.simulate_move(g, d)
        n += len(boards) * 4
    return n, time.perf_counter() - t0


def _worker_step(args):
    seed, seconds = args
    import numpy as np
    game = _import_reference()
    boards = _boards(512, seed)
    rng = np.random.default_rng(seed + 1)
    game.random = _Replay(rng.integers(0, 2 ** 32, (4096, 2)).tolist())
    dirs = list(game.Direction)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for g in boards:
            for d in dirs:
                env = game.Game2048([row[:] for row in g])
                env.step(d)
        n += len(boards) * 4
    return n, time.perf_counter() - t0


def _play(seconds):
    import torch
    game = _import_reference()
    import train
    torch.manual_seed(0)
    model = game.GameMLP(game.MLPConfig(hidden_dim=196)).eval()
    t0, n, episodes = time.perf_counter(), 0, []
    while time.perf_counter() - t0 < seconds:
        ep = train.play_game_for_episode(model, max_steps=None, device=torch.device("cpu"))
        n += len(ep["moves"])
        episodes.append(ep)
    return n, time.perf_counter() - t0, torch.get_num_threads(), model, episodes


class _Sgd:
    """the optimizer interface model_optimize_step expects (train.py:1232-1281) over plain SGD"""

    def __init__(self, model):
        import torch
        self.opt = torch.optim.SGD(model.parameters(), lr=1e-3)

    def step(self):
        self.opt.step()

    def zero_grad(self):
        self.opt.zero_grad(set_to_none=True)

    def scheduler_step(self):
        pass


def _advantage_and_update(model, episodes):
    """B4: calculate_advantage (README weights), then model_optimize_step at batch size 4 and with one large batch."""
    import train
    t0 = time.perf_counter()
    eps, _aug, _m1, _m2, _mu = train.calculate_advantage(episodes, 0.99, 0.0, 0.10, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0,
                                                         rtg_beta=0.99, rtg_m2=1.0, rtg_mu=0.0, rtg_step=1, upsample_ratio=0.0)
    t_adv = time.perf_counter() - t0
    n = sum(len(e["moves"]) for e in eps)
    out = {"samples": n, "calculate_advantage_samples_per_sec": n / t_adv}
    for name, bs in (("batch4", 4), ("one_batch", max(n, 1))):
        t0 = time.perf_counter()
        train.model_optimize_step(model=model, episodes=eps, optimizer=_Sgd(model), lr_scheduler=None, kl_strength=0.02,
                                  critic_strength=0.2, device=None, batch_size=bs, epochs=1)
        out[f"model_optimize_step_{name}_samples_per_sec"] = n / (time.perf_counter() - t0)
    return out


def _config1(train_steps):
    """B5: the reference's CLI, unmodified, in its own process on CPU (README shaping flags); rates from its JSONL log."""
    import datetime
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "batched_rollout.py"), "w") as f:      # the import slot of train.py:30; never called on this path
            f.write("def play_games_batched(*a, **k):\n    raise RuntimeError('not used by the CPU configuration')\n")
        cmd = [sys.executable, os.path.join(REF, "train.py"), "train", "--model-type", "mlp", "-h", "196", "--batch-size=4",
               f"--steps={train_steps}", "--lr", "0.001", "--critic-lr", "1e-4", "--gamma", "0.99", "--entropy", "0.02", "--points", "0.10",
               "--mono", "1.0", "--critic", "0.2", "--rtg-beta", "0.99", "--warmup-steps", "10", "--upsample-ratio", "0.25",
               "--emptiness", "0", "--smoothness", "0", "--tile-bonus", "0", "--corner", "0", "--log-dir", os.path.join(tmp, "logs")]
        env = dict(os.environ, PYTHONPATH=tmp + os.pathsep + REF, CUDA_VISIBLE_DEVICES="")
        t0 = time.perf_counter()
        res = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
        wall = time.perf_counter() - t0
        if res.returncode != 0:
            return {"unavailable": (res.stderr or res.stdout).strip().splitlines()[-1][:200]}
        logs = [os.path.join(tmp, "logs", f) for f in os.listdir(os.path.join(tmp, "logs")) if f.endswith(".jsonl")]
        rows = [json.loads(line) for line in open(logs[0])]
    ts = [datetime.datetime.fromisoformat(r["timestamp"]).timestamp() for r in rows]
    span = ts[-1] - ts[0]                                                    # first logged step .. last: start-up excluded
    played = sum(r["samples"] - r.get("augmented_samples", 0) for r in rows[1:])
    return {"train_steps": len(rows), "wall_s_with_startup": wall,
            "train_steps_per_sec": (len(rows) - 1) / span if span > 0 else None,
            "env_steps_per_sec": played / span if span > 0 else None,
            "command": "train.py train --model-type mlp -h 196 --batch-size=4 --steps=%d + README shaping flags, CPU" % train_steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0, help="per leg")
    ap.add_argument("--train-steps", type=int, default=20, help="steps of the config #1 run (B5); 0 = skip B4 and B5")
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "game.py")):
        print(json.dumps({"unavailable": "baseline/_ref/game.py is not staged (run baseline/stage_reference.sh where /root/reference exists)"}))
        return
    cores = os.cpu_count() or 1
    out = {"cores": cores, "kind": "reference", "source": "baseline/_ref (unmodified game.py / train.py)"}
    with mp.get_context("spawn").Pool(cores) as pool:
        r = pool.map(_worker_simulate, [(100 + i, a.seconds) for i in range(cores)])
        out["simulate_move_transitions_per_sec"] = sum(n / dt for n, dt in r)
        r = pool.map(_worker_step, [(200 + i, a.seconds) for i in range(cores)])
        out["step_env_steps_per_sec"] = sum(n / dt for n, dt in r)
    out["simulate_move_per_core"] = out["simulate_move_transitions_per_sec"] / cores
    out["step_per_core"] = out["step_env_steps_per_sec"] / cores
    n, dt, threads, model, episodes = _play(a.seconds)
    out["play_game_for_episode_env_steps_per_sec"] = n / dt
    out["play_game_torch_threads"] = threads
    if a.train_steps > 0:
        out["advantage_and_update"] = _advantage_and_update(model, episodes)          # B4
        out["config1_train_cli"] = _config1(a.train_steps)                            # B5
    out["sample"] = (f"{a.seconds:.0f} s per leg: simulate_move and Game2048.step (replayed spawn draws) on C2-distribution boards x 4 "
                     f"directions over {cores} processes; play_game_for_episode with GameMLP h=196 in one process; calculate_advantage + "
                     f"model_optimize_step on those episodes; {a.train_steps} steps of the config #1 command")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
