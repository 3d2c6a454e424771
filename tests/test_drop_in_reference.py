"""Drop-in boundary against the REAL reference consumers (CPU, build container only: needs the
reference checkout, skipped elsewhere).  tests/golden/episodes_gpu.pt is the list[EpisodeData] that
our `batched_rollout.play_games_batched` produced on a B200 (tools/dump_episodes.py).  Here the
unmodified reference `train.py` is imported with OUR `batched_rollout` module in its import slot
(train.py:30) and its own calculate_advantage / model_optimize_step consume those episodes."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("G2048_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, ROOT)            # /batched_rollout.py = the module train.py:30 imports
    sys.path.insert(0, REF)
    import batched_rollout
    import game
    import train
    assert train.play_games_batched is batched_rollout.play_games_batched     # the reference bound OUR function
    return game, train


@pytest.fixture()
def episodes(ref):
    game, _ = ref
    eps = torch.load(os.path.join(ROOT, "tests", "golden", "episodes_gpu.pt"), weights_only=False)
    for ep in eps:
        for mv in ep["moves"]:
            mv["points_possible"] = {game.Direction(k): v for k, v in mv["points_possible"].items()}
    return eps


def test_reference_env_replays_our_episodes(ref, episodes):
    """Every recorded move is a legal reference move whose simulate_move result + one spawned tile is our
    result_state, with the same points and the same potentials the reference computes itself."""
    game, _ = ref
    dirs = [game.Direction.UP, game.Direction.DOWN, game.Direction.LEFT, game.Direction.RIGHT]
    for ep in episodes:
        for mv in ep["moves"]:
            g = game.Game2048([r[:] for r in mv["state_before"]])
            d = dirs[mv["selected_direction"]]
            assert [x not in g.current_valid_directions() for x in dirs] == mv["action_mask"]
            assert g.preview_move_rewards() == mv["points_possible"]
            pre, pts, mt = game.Game2048.simulate_move(mv["state_before"], d)
            assert pts == mv["points_earned"] and mt == mv["max_tile_created"]
            diff = [(a, b) for ra, rb in zip(pre, mv["result_state"]) for a, b in zip(ra, rb) if a != b]
            assert len(diff) == 1 and diff[0][0] == 0 and diff[0][1] in (1, 2)
            assert mv["monotonicity_before"] == game.Game2048.monotonicity(mv["state_before"])
            assert mv["emptiness_before"] == game.Game2048.emptiness(mv["state_before"])
            assert mv["smoothness_delta"] == game.Game2048.smoothness_score(pre) - game.Game2048.smoothness_score(mv["state_before"])
            assert mv["corner_delta"] == game.Game2048.corner_bonus(pre) - game.Game2048.corner_bonus(mv["state_before"])
            assert mv["adjacency_delta"] == game.Game2048.adjacency_bonus(pre) - game.Game2048.adjacency_bonus(mv["state_before"])
            assert mv["chain_delta"] == game.Game2048.monotonic_chain_score(pre) - game.Game2048.monotonic_chain_score(mv["state_before"])
            anchor = game.Game2048._choose_anchor_corner(mv["state_before"])
            assert mv["topological_delta"] == game.Game2048.topological_score(pre, anchor) - game.Game2048.topological_score(mv["state_before"], anchor)
            torch.testing.assert_close(mv["game_state"], game.Game2048([r[:] for r in mv["state_before"]]).to_model_format())


def test_reference_advantage_and_optimize_step_consume_our_episodes(ref, episodes, golden):
    game, train = ref
    eps, aug, m1, m2, mu = train.calculate_advantage(
        copy.deepcopy(episodes), 0.99, 0.0, 0.10, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0,
        rtg_beta=0.99, rtg_m2=1.0, rtg_mu=0.0, rtg_step=1, upsample_ratio=0.25)          # README flags
    n = sum(len(ep["moves"]) for ep in eps)
    assert n == sum(len(ep["moves"]) for ep in episodes) and len(aug) > 0
    assert all(np.isfinite(m["advantage"]) for ep in eps for m in ep["moves"])
    # the reference's augmentation (mirror / rotate + remaps) accepts our records too
    eps.append({"moves": aug, "total_points": 0, "total_steps": len(aug), "augmented": True,
                "final_state": aug[-1]["result_state"]})
    g = golden("model_best")
    model = game.GameMLP(game.MLPConfig(hidden_dim=int(g["hidden_dim"]), num_layers=int(g["num_layers"]), dropout=0.0))
    from g2048.policy import load_state_dict_from_npz
    model.load_state_dict(load_state_dict_from_npz(g))

    class Opt:
        steps = 0

        def step(self):
            Opt.steps += 1

        def zero_grad(self):
            for p in model.parameters():
                p.grad = None

        def scheduler_step(self):
            pass

    stats = train.model_optimize_step(model=model, episodes=eps, optimizer=Opt(), lr_scheduler=None, kl_strength=0.02,
                                      critic_strength=0.2, device=None, batch_size=256, epochs=1)
    assert Opt.steps == (n + len(aug) + 255) // 256
    assert all(np.isfinite(stats[k]) for k in ("loss", "policy_loss", "value_loss", "entropy", "kl_average"))
    # the rollout policy IS this model: the recorded log-probs are its own, so the first-epoch KL is ~0
    assert abs(stats["kl_average"]) < 1e-4


def test_reference_timing_script_runs_b1_to_b5():
    """baseline/time_reference.py (bench.py's `cpu_baseline.python_reference`): every leg of BASELINE.md section 3 on the staged,
    unmodified reference, with tiny budgets -- the script itself is what is checked here, not the numbers."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "baseline", "_ref", "game.py")):
        pytest.skip("baseline/_ref is not staged (baseline/stage_reference.sh needs /root/reference)")
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    out = subprocess.run([sys.executable, os.path.join(root, "baseline", "time_reference.py"), "--seconds", "0.5", "--train-steps", "2"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("simulate_move_transitions_per_sec", "step_env_steps_per_sec", "play_game_for_episode_env_steps_per_sec"):
        assert d[k] > 0, k
    b4, b5 = d["advantage_and_update"], d["config1_train_cli"]
    assert b4["samples"] > 0 and b4["model_optimize_step_batch4_samples_per_sec"] > 0 and b4["model_optimize_step_one_batch_samples_per_sec"] > 0
    assert "unavailable" not in b5 and b5["train_steps"] == 2 and b5["env_steps_per_sec"] > 0
