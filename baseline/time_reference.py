"""Times the UNMODIFIED Python reference (staged under baseline/_ref by baseline/stage_reference.sh) on this host's cores:
BASELINE.md section 3, B1-B3.

  B1  game.Game2048.simulate_move   (game.py:121-160)   on the C2 board distribution, all four directions, every core
  B2  game.Game2048.step            (game.py:952-1030)  with game.random replaced by a replay RNG, every core
  B3  train.play_game_for_episode   (train.py:213-345)  GameMLP h=196, eval mode, single process (how the reference runs it)

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
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        ep = train.play_game_for_episode(model, max_steps=None, device=torch.device("cpu"))
        n += len(ep["moves"])
    return n, time.perf_counter() - t0, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0, help="per leg")
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
    n, dt, threads = _play(a.seconds)
    out["play_game_for_episode_env_steps_per_sec"] = n / dt
    out["play_game_torch_threads"] = threads
    out["sample"] = (f"{a.seconds:.0f} s per leg: simulate_move and Game2048.step (replayed spawn draws) on C2-distribution boards x 4 "
                     f"directions over {cores} processes; play_game_for_episode with GameMLP h=196 in one process")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
