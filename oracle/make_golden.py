"""Generate tests/golden/*.npz from the UNMODIFIED reference (game.py / train.py).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python oracle/make_golden.py

The reference draws tile spawns through the module-global `random`
(game.py:5,937,939).  We replace `game.random` by ReplayRandom, which turns one
u32 pair (u0,u1) per spawn into `choice(seq) = seq[mulhi32(u0,len(seq))]` and
`random() = u1 / 2**32` -- the same mapping the CUDA kernels and the C oracle
use -- so identical draws reach both sides.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("G2048_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

# train.py imports a module the reference does not ship (train.py:30).
stub = types.ModuleType("batched_rollout")
stub.play_games_batched = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
sys.modules["batched_rollout"] = stub
sys.path.insert(0, REF)

import game as G  # noqa: E402
import train as TR  # noqa: E402
from oracle import oracle as O  # noqa: E402

DIRS = [G.Direction.UP, G.Direction.DOWN, G.Direction.LEFT, G.Direction.RIGHT]  # train.py:266


class ReplayRandom:
    """Stand-in for the `random` module inside game.py."""

    def __init__(self, pairs):
        self._it = iter(pairs)
        self._cur = None
        self.consumed = 0

    def choice(self, seq):
        self._cur = next(self._it)
        self.consumed += 1
        return seq[(int(self._cur[0]) * len(seq)) >> 32]

    def random(self):
        return int(self._cur[1]) / 2**32


def pack(grid):
    b = 0
    for r in range(4):
        for c in range(4):
            b |= min(int(grid[r][c]), 15) << (4 * (4 * r + c))
    return b


def cells(grid):
    return [int(grid[r][c]) for r in range(4) for c in range(4)]


def unpack(b):
    return [[(int(b) >> (4 * (4 * r + c))) & 15 for c in range(4)] for r in range(4)]


def legal_mask(grid):
    g = G.Game2048([row[:] for row in grid]) if any(any(r) for r in grid) else G.Game2048()
    valid = g.current_valid_directions()
    return sum(1 << i for i, d in enumerate(DIRS) if d in valid)


# --------------------------------------------------------------------------- boards

def random_boards(rng: np.random.Generator):
    boards = []
    # C2-style: exps 1..11, 30 % empties
    for _ in range(2500):
        e = rng.integers(1, 12, size=16)
        e[rng.random(16) < 0.30] = 0
        boards.append(e)
    # sparse / dense / high exponents (incl. 15 -> nibble overflow on merge)
    for p_empty, hi in ((0.7, 6), (0.0, 4), (0.1, 16), (0.5, 16), (0.0, 3), (0.9, 16)):
        for _ in range(500):
            e = rng.integers(1, hi, size=16)
            e[rng.random(16) < p_empty] = 0
            boards.append(e)
    # few distinct values -> many merges and ties for the max tile
    for _ in range(1000):
        e = rng.choice([0, 1, 2, 3], size=16, p=[0.2, 0.4, 0.3, 0.1])
        boards.append(e)
    special = [
        [0] * 16,
        [1, 2, 1, 2, 2, 1, 2, 1, 1, 2, 1, 2, 2, 1, 2, 1],             # checkerboard, terminal
        [1, 0, 0, 0] + [0] * 12,
        [1, 2, 3, 4, 8, 7, 6, 5, 9, 10, 11, 12, 0, 0, 0, 13],           # SURVEY board B
        [5, 5, 1, 0, 0, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 5],               # SURVEY board C
        [0, 5, 1, 0, 0, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 5],               # SURVEY board D
        [15] * 16,
        [1] * 16,
        [15, 15, 15, 15, 14, 14, 14, 14, 0, 0, 0, 0, 1, 1, 2, 2],
        [3, 3, 0, 3, 2, 1, 1, 2, 1, 0, 0, 1, 1, 1, 1, 1],
    ]
    boards.extend(np.array(s) for s in special)
    return np.array(boards, dtype=np.int64)


def gen_env_step(rng):
    boards = random_boards(rng)
    n = boards.shape[0]
    rec = {k: [] for k in (
        "board", "action", "draw", "out_cells", "points", "done", "invalid", "mono_before", "mono_after",
        "empt_before", "empt_after", "max_tile_created", "max_exp_before", "max_exp_after",
        "smooth_delta", "corner_delta", "smooth_before", "smooth_after", "corner_before", "corner_after",
        "legal_before", "legal_after", "pre_spawn_cells")}
    for i in range(n):
        grid = boards[i].reshape(4, 4).tolist()
        for a in range(4):
            draw = rng.integers(0, 2**32, size=2, dtype=np.uint64)
            # force the 0.9 threshold neighbourhood now and then
            if rng.random() < 0.05:
                draw[1] = 3865470566 + int(rng.integers(0, 2))
            rr = ReplayRandom([(int(draw[0]), int(draw[1]))])
            G.random = rr
            g = G.Game2048()
            g.grid = [row[:] for row in grid]
            lb = legal_mask(grid)
            pre, _, _ = G.Game2048.simulate_move(grid, DIRS[a])
            new_state, pts, done, info = g.step(DIRS[a])
            rec["board"].append(pack(grid))
            rec["action"].append(a)
            rec["draw"].append([int(draw[0]), int(draw[1])])
            rec["out_cells"].append(cells(new_state))
            rec["points"].append(pts)
            rec["done"].append(int(done))
            rec["invalid"].append(int(info["invalid_move"]))
            rec["mono_before"].append(int(info["monotonicity_before"]))
            rec["mono_after"].append(int(info["monotonicity_after"]))
            rec["empt_before"].append(int(info["emptiness_before"]))
            rec["empt_after"].append(int(info["emptiness_after"]))
            rec["max_tile_created"].append(int(info["max_tile_created"]))
            rec["max_exp_before"].append(int(info.get("max_exponent_before", 0)))
            rec["max_exp_after"].append(int(info.get("max_exponent_after", 0)))
            rec["smooth_delta"].append(int(info["smoothness_delta"]))
            rec["corner_delta"].append(int(info["corner_delta"]))
            if info["invalid_move"]:
                sb = sa = cb = ca = 0
                pre = grid
            else:
                sb = int(G.Game2048.smoothness_score(grid))
                sa = int(G.Game2048.smoothness_score(pre))
                cb = int(G.Game2048.corner_bonus(grid))
                ca = int(G.Game2048.corner_bonus(pre))
            rec["smooth_before"].append(sb)
            rec["smooth_after"].append(sa)
            rec["corner_before"].append(cb)
            rec["corner_after"].append(ca)
            rec["legal_before"].append(lb)
            rec["legal_after"].append(legal_mask(new_state))
            rec["pre_spawn_cells"].append(cells(pre))
    out = {
        "board": np.array(rec["board"], dtype=np.uint64),
        "action": np.array(rec["action"], dtype=np.uint8),
        "draw": np.array(rec["draw"], dtype=np.uint32),
        "out_cells": np.array(rec["out_cells"], dtype=np.uint8),
        "pre_spawn_cells": np.array(rec["pre_spawn_cells"], dtype=np.uint8),
    }
    for k in rec:
        if k not in out:
            out[k] = np.array(rec[k], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "env_step.npz"), **out)
    print("env_step:", len(rec["board"]), "transitions,", int(out["invalid"].sum()), "invalid,",
          int(out["done"].sum()), "done")


def gen_potentials(rng):
    boards = random_boards(rng)
    vals = []
    for e in boards:
        grid = e.reshape(4, 4).tolist()
        vals.append([
            int(G.Game2048.monotonicity(grid)), int(G.Game2048.emptiness(grid)),
            int(G.Game2048.smoothness_score(grid)), int(G.Game2048.corner_bonus(grid)),
            max(max(r) for r in grid), legal_mask(grid)])
    np.savez_compressed(os.path.join(OUT, "potentials.npz"),
                        board=np.array([pack(e.reshape(4, 4)) for e in boards], dtype=np.uint64),
                        values=np.array(vals, dtype=np.int32))
    print("potentials:", len(vals))


def gen_best_game():
    d = json.load(open(os.path.join(REF, "docs", "data", "best_game.json")))
    name2a = {"UP": 0, "DOWN": 1, "LEFT": 2, "RIGHT": 3}
    lg = lambda v: 0 if v == 0 else int(v).bit_length() - 1
    before, after, act, pts = [], [], [], []
    for m in d["moves"]:
        before.append(pack([[lg(v) for v in row] for row in m["state_before"]]))
        after.append(pack([[lg(v) for v in row] for row in m["state_after"]]))
        act.append(name2a[m["action"]])
        pts.append(m["points_earned"])
    np.savez_compressed(os.path.join(OUT, "best_game.npz"), before=np.array(before, dtype=np.uint64),
                        after=np.array(after, dtype=np.uint64), action=np.array(act, dtype=np.uint8),
                        points=np.array(pts, dtype=np.int32), score=np.int64(d["score"]))
    print("best_game:", len(act), "moves, score", d["score"])


def load_best_model():
    ck = torch.load(os.path.join(REF, "docs", "data", "best_model.pt"), map_location="cpu", weights_only=False)
    cfg = ck["config"]
    model = G.GameMLP(G.MLPConfig(**cfg))
    model.load_state_dict(ck["model_state_dict"])
    model.eval()
    return model, cfg


def gen_model(rng):
    model, cfg = load_best_model()
    boards = random_boards(rng)[:512]
    x = torch.stack([G.Game2048(b.reshape(4, 4).tolist() if b.any() else None).to_model_format()
                     for b in boards])
    with torch.no_grad():
        logits, v = model(x)
    sd = {k.replace(".", "__"): t.numpy() for k, t in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "model_best.npz"),
                        board=np.array([pack(b.reshape(4, 4)) for b in boards], dtype=np.uint64),
                        inputs=x.numpy(), logits=logits.numpy(), value=v.numpy(),
                        hidden_dim=np.int64(cfg["hidden_dim"]), num_layers=np.int64(cfg["num_layers"]),
                        **{"sd__" + k: a for k, a in sd.items()})
    print("model_best: h=%d L=%d, %d inputs" % (cfg["hidden_dim"], cfg["num_layers"], len(boards)))
    return model


def gen_rollout(model, seed=2048, n_games=6, max_steps=(None, None, None, 120, 60, None)):
    """Reference play_game_for_episode (train.py:213-345) with Philox-derived spawn draws."""
    torch.manual_seed(1234)
    eps_out = []
    flat = {k: [] for k in ("env", "t", "board", "action", "mask", "logp", "value", "entropy", "points",
                            "mono_before", "mono_after", "empt_before", "empt_after", "max_tile_created",
                            "smooth_delta", "corner_delta", "result", "done", "points_possible")}
    episodes = []
    for env in range(n_games):
        def pairs(env=env):
            d = O.philox(seed, env, 0)
            yield (d[0], d[1])
            yield (d[2], d[3])
            t = 0
            while True:
                d = O.philox(seed, env, 1 + t)
                yield (d[0], d[1])
                t += 1
        G.random = ReplayRandom(pairs())
        ep = TR.play_game_for_episode(model, max_steps=max_steps[env], device=None)
        episodes.append(ep)
        n = len(ep["moves"])
        for t, m in enumerate(ep["moves"]):
            done = int(t == n - 1 and not G.Game2048.state_has_next_step(m["result_state"]))
            flat["env"].append(env)
            flat["t"].append(t)
            flat["board"].append(pack(m["state_before"]))
            flat["action"].append(m["selected_direction"])
            flat["mask"].append(sum(1 << i for i, ill in enumerate(m["action_mask"]) if not ill))
            flat["logp"].append(m["policy_logprobs"])
            flat["value"].append(m["predicted_future_value"])
            flat["entropy"].append(m["entropy"])
            flat["points"].append(m["points_earned"])
            flat["mono_before"].append(m["monotonicity_before"])
            flat["mono_after"].append(m["monotonicity_after"])     # already fixed up (train.py:318-322)
            flat["empt_before"].append(m["emptiness_before"])
            flat["empt_after"].append(m["emptiness_after"])
            flat["max_tile_created"].append(m["max_tile_created"])
            flat["smooth_delta"].append(m["smoothness_delta"])
            flat["corner_delta"].append(m["corner_delta"])
            flat["result"].append(pack(m["result_state"]))
            flat["done"].append(done)
            flat["points_possible"].append([m["points_possible"][d] for d in DIRS])
        eps_out.append((n, ep["total_points"], ep["total_steps"], pack(ep["final_state"])))
    arr = dict(
        seed=np.uint64(seed), env=np.array(flat["env"], dtype=np.int32), t=np.array(flat["t"], dtype=np.int32),
        board=np.array(flat["board"], dtype=np.uint64), action=np.array(flat["action"], dtype=np.uint8),
        legal=np.array(flat["mask"], dtype=np.uint8), logp=np.array(flat["logp"], dtype=np.float32),
        value=np.array(flat["value"], dtype=np.float32), entropy=np.array(flat["entropy"], dtype=np.float32),
        result=np.array(flat["result"], dtype=np.uint64), done=np.array(flat["done"], dtype=np.uint8),
        points_possible=np.array(flat["points_possible"], dtype=np.int32),
        ep_len=np.array([e[0] for e in eps_out], dtype=np.int32),
        ep_points=np.array([e[1] for e in eps_out], dtype=np.int64),
        ep_total_steps=np.array([e[2] for e in eps_out], dtype=np.int32),
        ep_final=np.array([e[3] for e in eps_out], dtype=np.uint64),
        ep_max_steps=np.array([m or 0 for m in max_steps], dtype=np.int32),
    )
    for k in ("points", "mono_before", "mono_after", "empt_before", "empt_after", "max_tile_created",
              "smooth_delta", "corner_delta"):
        arr[k] = np.array(flat[k], dtype=np.float64).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "rollout.npz"), **arr)
    print("rollout:", [e[0] for e in eps_out], "moves per game; points", [e[1] for e in eps_out])
    return episodes


def gen_advantage(episodes):
    """train.calculate_advantage (train.py:651-904) on the rollout episodes, two weight sets."""
    import copy
    out = {}
    cases = {
        "readme": dict(gamma=0.99, points=0.10, mono=1.0, empt=0.0, beta=0.99, step=1, mu=0.0, m2=1.0),
        "warm": dict(gamma=0.97, points=0.05, mono=0.5, empt=0.25, beta=0.9, step=7, mu=12.5, m2=900.0),
    }
    kept = None
    for name, c in cases.items():
        eps = copy.deepcopy(episodes)
        eps, aug, m1, m2, mu = TR.calculate_advantage(
            eps, c["gamma"], c["mu"], c["points"], 0.0, 0.0, 0.0, 0.0, 0.0, c["mono"], c["empt"], 0.0, 0.0,
            rtg_beta=c["beta"], rtg_m2=c["m2"], rtg_mu=c["mu"], rtg_step=c["step"], upsample_ratio=0.0)
        mv = [m for ep in eps for m in ep["moves"]]
        out[name + "__reward"] = np.array([m["reward"] for m in mv], dtype=np.float64)
        out[name + "__g_raw"] = np.array([m["future_reward_raw"] for m in mv], dtype=np.float64)
        out[name + "__g_norm"] = np.array([m["future_reward"] for m in mv], dtype=np.float64)
        out[name + "__adv"] = np.array([m["advantage"] for m in mv], dtype=np.float64)
        out[name + "__cfg"] = np.array([c["gamma"], c["points"], c["mono"], c["empt"], c["beta"], c["step"],
                                        c["mu"], c["m2"]], dtype=np.float64)
        out[name + "__moments_out"] = np.array([mu, m2], dtype=np.float64)
        if name == "readme":
            kept = eps
    np.savez_compressed(os.path.join(OUT, "advantage.npz"), **out)
    print("advantage:", {k: v.shape for k, v in out.items() if k.endswith("adv")})
    return kept


class Recorder:
    """Optimizer stand-in: snapshots the (already clipped) gradients at .step()."""

    def __init__(self, model):
        self.model = model
        self.grads = None

    def step(self):
        self.grads = {n: p.grad.detach().clone() for n, p in self.model.named_parameters()}

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None

    def scheduler_step(self):
        pass


def gen_loss(episodes_with_adv):
    """train.model_optimize_step (train.py:414-642), one full-batch step, dropout = 0."""
    base, cfg = load_best_model()
    model = G.GameMLP(G.MLPConfig(hidden_dim=cfg["hidden_dim"], num_layers=cfg["num_layers"], dropout=0.0))
    model.load_state_dict(base.state_dict())
    with torch.no_grad():          # policy B != rollout policy so that the PPO ratio leaves [0.8, 1.2]
        model.action_head.weight.mul_(1.3)
        model.value_head.bias.add_(0.3)
    n = sum(len(ep["moves"]) for ep in episodes_with_adv)
    out = {}
    for name, (ent, crit) in {"readme": (0.02, 0.2), "alt": (0.1, 1.0)}.items():
        rec = Recorder(model)
        stats = TR.model_optimize_step(model=model, episodes=episodes_with_adv, optimizer=rec, lr_scheduler=None,
                                       kl_strength=ent, critic_strength=crit, device=None, batch_size=n, epochs=1)
        out[name + "__stats"] = np.array([stats["loss"], stats["policy_loss"], stats["value_loss"],
                                          stats["entropy"], stats["grad_norm"], stats["entropy_loss"]],
                                         dtype=np.float64)
        out[name + "__coef"] = np.array([ent, crit], dtype=np.float64)
        for k, g in rec.grads.items():
            out[name + "__grad__" + k.replace(".", "__")] = g.numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), n=np.int64(n), **out)
    print("loss:", n, "samples; stats", out["readme__stats"])


class SgdStack:
    """The optimizer interface model_optimize_step expects (train.py:1232-1281 MultiOptimizer) over plain SGD."""

    def __init__(self, model, lr):
        self.opt = torch.optim.SGD(model.parameters(), lr=lr)

    def step(self):
        self.opt.step()

    def zero_grad(self):
        self.opt.zero_grad(set_to_none=True)

    def scheduler_step(self):
        pass


def gen_optimize(episodes_with_adv):
    """train.model_optimize_step (train.py:414-642) as the train loop calls it: shuffled minibatches, two epochs,
    a real optimizer (SGD lr 0.005), dropout = 0.  Records the returned statistics and the final weights."""
    base, cfg = load_best_model()
    model = G.GameMLP(G.MLPConfig(hidden_dim=cfg["hidden_dim"], num_layers=cfg["num_layers"], dropout=0.0))
    model.load_state_dict(base.state_dict())
    with torch.no_grad():
        model.action_head.weight.mul_(1.3)
        model.value_head.bias.add_(0.3)
    keys = ["loss", "policy_loss", "entropy_loss", "value_loss", "grad_norm", "entropy", "kl_total", "kl_average", "kl_max", "lr"]
    torch.manual_seed(777)
    stats = TR.model_optimize_step(model=model, episodes=episodes_with_adv, optimizer=SgdStack(model, 0.005), lr_scheduler=None,
                                   kl_strength=0.02, critic_strength=0.2, device=None, batch_size=128, epochs=2)
    out = {"stats": np.array([stats[k] for k in keys], dtype=np.float64), "seed": np.int64(777), "lr": np.float64(0.005),
           "batch_size": np.int64(128), "epochs": np.int64(2), "coef": np.array([0.02, 0.2])}
    for k, t in model.state_dict().items():
        out["final__" + k.replace(".", "__")] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "optimize.npz"), **out)
    print("optimize:", dict(zip(keys, out["stats"].round(6))))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "optimize":     # only tests/golden/optimize.npz (the other fixtures stay untouched)
        save, np.savez_compressed = np.savez_compressed, lambda *a, **k: None
        eps_adv = gen_advantage(gen_rollout(gen_model(np.random.default_rng(11))))
        np.savez_compressed = save
        gen_optimize(eps_adv)
        return
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2048)
    gen_best_game()
    gen_potentials(np.random.default_rng(7))
    gen_env_step(rng)
    model = gen_model(np.random.default_rng(11))
    episodes = gen_rollout(model)
    eps_adv = gen_advantage(episodes)
    gen_loss(eps_adv)
    gen_optimize(eps_adv)
    gen_urm(np.random.default_rng(13))



def gen_urm(rng):
    """Reference GameURM (game.py:1355-1458), default config, seeded init, eval mode."""
    torch.manual_seed(64)
    cfg = G.GameURMConfig()
    model = G.GameURM(cfg).eval()
    boards = random_boards(rng)[:256]
    x = torch.stack([G.Game2048(b.reshape(4, 4).tolist() if b.any() else None).to_model_format() for b in boards])
    with torch.no_grad():
        logits, v = model(x)
    sd = {k.replace(".", "__"): t.numpy() for k, t in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "model_urm.npz"),
                        board=np.array([pack(b.reshape(4, 4)) for b in boards], dtype=np.uint64),
                        inputs=x.numpy(), logits=logits.numpy(), value=v.numpy(),
                        **{"sd__" + k: a for k, a in sd.items()})
    print("model_urm:", {k: tuple(v.shape) for k, v in model.state_dict().items()})


def gen_potentials_ext(rng):
    """adjacency_bonus / monotonic_chain_score / topological_score(+anchor) of (board, pre-spawn successor)
    pairs exactly as Game2048.step evaluates them (game.py:981-1001)."""
    boards = random_boards(rng)
    rows_b, rows_a, vals = [], [], []
    for i, e in enumerate(boards):
        grid = e.reshape(4, 4).tolist()
        if max(max(r) for r in grid) > 14:
            continue
        for a in range(4):
            if (i + a) % 2:
                continue
            after, _, _ = G.Game2048.simulate_move(grid, DIRS[a])
            anchor = G.Game2048._choose_anchor_corner(grid)
            vals.append([G.Game2048.adjacency_bonus(grid), G.Game2048.adjacency_bonus(after),
                         G.Game2048.monotonic_chain_score(grid), G.Game2048.monotonic_chain_score(after),
                         G.Game2048.topological_score(grid, anchor), G.Game2048.topological_score(after, anchor),
                         4 * anchor[0] + anchor[1]])
            rows_b.append(pack(grid))
            rows_a.append(pack(after))
    np.savez_compressed(os.path.join(OUT, "potentials_ext.npz"), before=np.array(rows_b, dtype=np.uint64),
                        after=np.array(rows_a, dtype=np.uint64), values=np.array(vals, dtype=np.float64))
    print("potentials_ext:", len(vals))


def gen_augment(rng):
    """Reference mirror_grid / rotate_grid (game.py:508-590) on random boards: ops 0..4 =
    mirror horizontal, mirror vertical, rotate 90 / 180 / 270 clockwise."""
    boards = random_boards(rng)[:2000]
    ops = [lambda g: G.Game2048.mirror_grid(g, "horizontal"), lambda g: G.Game2048.mirror_grid(g, "vertical"),
           lambda g: G.Game2048.rotate_grid(g, 90), lambda g: G.Game2048.rotate_grid(g, 180),
           lambda g: G.Game2048.rotate_grid(g, 270)]
    out = np.array([[pack(op(b.reshape(4, 4).tolist())) for op in ops] for b in boards], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "augment.npz"),
                        board=np.array([pack(b.reshape(4, 4)) for b in boards], dtype=np.uint64), transformed=out)
    print("augment:", out.shape)


def gen_onnx():
    """The reference's shipped browser model (docs/data/model.onnx, written by train.export_model_to_onnx, train.py:33-78):
    its structure without the weights -> tests/golden/onnx_structure.json, and the outputs of tests/onnx_mini.py's numpy
    evaluator on it for the inputs of model_best.npz -> tests/golden/onnx_reference.npz.  Checked here, where torch can run the same weights: evaluator == torch forward."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import onnx_mini

    m = onnx_mini.load(os.path.join(REF, "docs", "data", "model.onnx"))
    json.dump(onnx_mini.structure(m), open(os.path.join(OUT, "onnx_structure.json"), "w"), indent=1)
    cfg = json.load(open(os.path.join(REF, "docs", "data", "model_config.json")))
    model = G.GameMLP(G.MLPConfig(**cfg)).eval()
    init = m["initializers"]
    mm = iter(sorted(k for k in init if k.startswith("onnx::MatMul_")))
    sd = {}
    for k in model.state_dict():
        sd[k] = torch.from_numpy(init[k].copy()) if k in init else torch.from_numpy(init[next(mm)].T.copy())
    model.load_state_dict(sd)
    x = np.load(os.path.join(OUT, "model_best.npz"))["inputs"]
    out = onnx_mini.run(m, {"board_state": x})
    with torch.no_grad():
        lg, v = model(torch.from_numpy(x))
    assert np.abs(out["action_logits"] - lg.numpy()).max() < 2e-5 and np.abs(out["value"] - v.numpy()).max() < 2e-5
    best = np.load(os.path.join(OUT, "model_best.npz"))        # the shipped file holds best_model.pt's weights: not stored twice
    assert all(np.array_equal(best["sd__" + k.replace(".", "__")], t.numpy()) for k, t in sd.items())
    np.savez_compressed(os.path.join(OUT, "onnx_reference.npz"), action_logits=out["action_logits"], value=out["value"],
                        config=json.dumps(cfg))
    print("onnx:", len(m["nodes"]), "nodes,", len(init), "initializers; evaluator == torch forward")


if __name__ == "__main__":
    if "--ext-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        gen_potentials_ext(np.random.default_rng(19))
    elif "--augment-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        gen_augment(np.random.default_rng(17))
    elif "--onnx-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        gen_onnx()
    elif "--urm-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        gen_urm(np.random.default_rng(13))
    else:
        main()
