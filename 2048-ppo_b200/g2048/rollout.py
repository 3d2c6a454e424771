"""Batched actor-critic rollout on the fused CUDA kernel + the `play_games_batched` drop-in.

Reference interfaces (file:line in RobotSail/2048-PPO):
  play_game_for_episode(model, max_steps, device) -> EpisodeData      train.py:213-345
  batched_rollout.play_games_batched(model, num_games, max_steps, device) -> list[EpisodeData]
      imported at train.py:30 and called at train.py:1677-1679, 2034 -- the module is missing
      from the reference; this file (re-exported by /batched_rollout.py) fills that slot.
  StepData / EpisodeData schemas                                        train.py:123-177
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, env
from .env import DIRECTIONS, _ptr, _stream


class _RolloutStruct(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("T", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("auto_reset", C.c_int32),
        ("seed", C.c_uint64), ("env0", C.c_uint64), ("ctr0", C.c_uint64),
        ("packed_weights", C.c_void_p), ("lut", C.c_void_p), ("boards", C.c_void_p), ("alive", C.c_void_p),
        ("forced_actions", C.c_void_p), ("rec_boards", C.c_void_p), ("rec_actions", C.c_void_p),
        ("rec_legal", C.c_void_p), ("rec_logp", C.c_void_p), ("rec_value", C.c_void_p), ("rec_points", C.c_void_p),
        ("rec_shaping", C.c_void_p), ("rec_flags", C.c_void_p), ("rec_entropy", C.c_void_p),
        ("tensor_cores", C.c_int32), ("reserved_", C.c_int32), ("sched_workspace", C.c_void_p),
    ]


_lib.lib().g2048_mlp_packed_floats.restype = C.c_int64
_lib.lib().g2048_mlp_packed_floats.argtypes = [C.c_int32, C.c_int32]
_lib.register("g2048_mlp_pack", [C.c_int32, C.c_int32] + [C.c_void_p] * 12)
_lib.register("g2048_rollout_mlp", [C.POINTER(_RolloutStruct), C.c_void_p])
_lib.lib().g2048_urm_packed_floats.restype = C.c_int64
_lib.lib().g2048_urm_packed_floats.argtypes = [C.c_int32] * 4
_lib.register("g2048_urm_pack", [C.c_int32] + [C.c_void_p] * 16)
_lib.register("g2048_rollout_urm", [C.POINTER(_RolloutStruct), C.c_int32, C.c_void_p])


TC_MAX_LAYERS = 6      # residual blocks the bf16 tensor-core kernel keeps LayerNorm parameters for (fp32 kernel: 8)
X3_MAX_LAYERS = 4      # same for the split-fp16 (fp32-grade) tensor-core kernel
TC_MIN_ENVS = 16384   # "auto": envs per GPU from which the rollout GEMMs run on the tensor cores
_PRECISION = {"fp32": 0, "bf16": 1, "x3": 2}     # G2048_ROLLOUT_* of include/g2048.h


def _dp(t):
    return t.data_ptr() if t is not None else None


@dataclass
class PackedPolicy:
    hidden: int
    layers: int
    weights: torch.Tensor   # float32, kernel layout
    kind: str = "mlp"       # "mlp" (GameMLP) or "urm" (GameURM)
    loops: int = 0          # GameURMConfig.num_loops


def _pack_urm(model, device) -> PackedPolicy:
    """GameURM (ours or the reference's, game.py:1355-1458) -> kernel layout."""
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = model.config
    h, L = sd["stem.0.weight"].shape[0], len({k.split(".")[1] for k in sd if k.startswith("layers.")})
    inter = sd["layers.0.mlp.down_proj.weight"].shape[1]
    dev = env.init(device if device is not None else (sd["stem.0.weight"].device if sd["stem.0.weight"].is_cuda else None))
    n = int(_lib.lib().g2048_urm_packed_floats(h, L, cfg.num_heads, inter))
    if n < 0 or cfg.conv_kernel != 2 or abs(cfg.rms_norm_eps - 1e-5) > 1e-12:
        raise ValueError("the fused URM kernel is built for hidden_dim=64, num_heads=4, inter=120, conv_kernel=2, "
                         f"rms_norm_eps=1e-5, 1..2 layers (got h={h}, heads={cfg.num_heads}, inter={inter}, L={L})")
    t = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in sd.items()}
    arr = lambda fmt: C.cast((C.c_void_p * L)(*[t[fmt.format(i)].data_ptr() for i in range(L)]), C.c_void_p)
    with torch.cuda.device(dev):
        out = torch.empty(n, dtype=torch.float32, device=dev)
        _lib.call("g2048_urm_pack", L, _dp(t["stem.0.weight"]), _dp(t["stem.1.weight"]), _dp(t["stem.1.bias"]),
                  _dp(t["init_hidden"]), arr("layers.{}.attn.qkv_proj.weight"), arr("layers.{}.attn.o_proj.weight"),
                  arr("layers.{}.mlp.gate_up_proj.weight"), arr("layers.{}.mlp.dwconv.weight"),
                  arr("layers.{}.mlp.dwconv.bias"), arr("layers.{}.mlp.down_proj.weight"),
                  _dp(t["action_head.weight"]), _dp(t["action_head.bias"]), _dp(t["value_head.weight"]),
                  _dp(t["value_head.bias"]), _dp(out), _stream())
        torch.cuda.current_stream().synchronize()
    return PackedPolicy(h, L, out, kind="urm", loops=int(cfg.num_loops))


def pack_policy(model, device=None) -> PackedPolicy:
    """GameMLP (ours or the reference's: same state_dict keys, game.py:1064-1085) -> kernel layout."""
    if "init_hidden" in dict(model.named_parameters()):
        return _pack_urm(model, device)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    h = sd["stem.0.weight"].shape[0]
    L = len({k.split(".")[1] for k in sd if k.startswith("backbone.")})
    dev = env.init(device if device is not None else (sd["stem.0.weight"].device
                                                       if sd["stem.0.weight"].is_cuda else None))
    n = int(_lib.lib().g2048_mlp_packed_floats(h, L))
    if n < 0:
        raise ValueError(f"the fused rollout kernel supports hidden_dim <= 208 and <= 8 layers (got h={h}, L={L})")
    f = lambda k: sd[k].to(device=dev, dtype=torch.float32).contiguous()
    t = {k: f(k) for k in sd}
    arr = lambda keys: (C.c_void_p * max(L, 1))(*[t[k].data_ptr() for k in keys])
    bw = arr([f"backbone.{i}.mlp.0.weight" for i in range(L)])
    bg = arr([f"backbone.{i}.mlp.1.weight" for i in range(L)])
    bb = arr([f"backbone.{i}.mlp.1.bias" for i in range(L)])
    with torch.cuda.device(dev):
        out = torch.empty(n, dtype=torch.float32, device=dev)
        _lib.call("g2048_mlp_pack", h, L, _dp(t["stem.0.weight"]), _dp(t["stem.1.weight"]), _dp(t["stem.1.bias"]),
                  C.cast(bw, C.c_void_p), C.cast(bg, C.c_void_p), C.cast(bb, C.c_void_p),
                  _dp(t["action_head.weight"]), _dp(t["action_head.bias"]), _dp(t["value_head.weight"]),
                  _dp(t["value_head.bias"]), _dp(out), _stream())
        torch.cuda.current_stream().synchronize()   # `t` (temporaries) must outlive the pack kernel
    return PackedPolicy(h, L, out)


@dataclass
class RolloutBuffers:
    """Time-major [T,B] records of one rollout call (include/g2048.h G2048Rollout)."""
    boards: torch.Tensor     # int64  state_before
    actions: torch.Tensor    # uint8
    legal: torch.Tensor      # uint8  legal-direction bits of state_before
    logp: torch.Tensor       # float32 [T,B,4]
    value: torch.Tensor      # float32
    points: torch.Tensor     # int32
    shaping: torch.Tensor    # int64  packed G2048_SH_* words
    flags: torch.Tensor      # uint8  step flags | 0x80 (valid)
    entropy: torch.Tensor    # float32

    @staticmethod
    def allocate(T: int, B: int, device) -> "RolloutBuffers":
        e = lambda dt, *s: torch.empty((T, B, *s), dtype=dt, device=device)
        return RolloutBuffers(e(torch.int64), e(torch.uint8), e(torch.uint8), e(torch.float32, 4), e(torch.float32),
                              e(torch.int32), e(torch.int64), e(torch.uint8), e(torch.float32))

    @property
    def shape(self):
        return tuple(self.flags.shape)


_SCHED: dict[int, torch.Tensor] = {}


def _sched_workspace(dev, B: int) -> torch.Tensor:
    """Per-device scratch for the x3 kernel's horizon segments (include/g2048.h G2048Rollout.sched_workspace); launches on
    one device are stream-ordered, so one buffer per device serves them all."""
    n = (B + 127) // 128
    t = _SCHED.get(dev.index)
    if t is None or t.numel() < n:
        t = torch.zeros(max(n, 1024), dtype=torch.int32, device=dev)
        _SCHED[dev.index] = t
    return t


def resolve_precision(precision: str, B: int, layers: int) -> str:
    """What "auto" means for B envs per GPU: the fp32-grade tensor-core kernel at large env batch, else fp32 FFMA."""
    if precision != "auto":
        return precision
    return "x3" if B >= TC_MIN_ENVS and layers <= X3_MAX_LAYERS else "fp32"


def rollout(policy: PackedPolicy, boards: torch.Tensor, T: int, *, seed: int, env0: int = 0, ctr0: int = 1,
            auto_reset: bool = True, alive: torch.Tensor | None = None, forced_actions: torch.Tensor | None = None,
            out: RolloutBuffers | None = None, precision: str = "auto") -> RolloutBuffers:
    """Play T steps of every board in `boards` (updated in place) with one fused kernel launch.

    precision: "fp32" = FFMA GEMMs, "x3" = tcgen05 tensor-core GEMMs on split-fp16 operands (x = hi + lo, three
    products, fp32 accumulation: fp32-grade like "fp32", log-probs / values within 2e-5 of the torch fp32 policy),
    "bf16" = tcgen05 GEMMs on bf16-rounded operands (log-probs ~1e-2 off the fp32 policy: a labelled variant, never
    chosen automatically), "auto" = "x3" from TC_MIN_ENVS envs up (tensor cores only at large env batch)."""
    if precision not in ("auto", "fp32", "bf16", "fp16", "x3"):
        raise ValueError(f"precision must be auto, fp32, x3 or bf16 (GameURM: auto, x3 or fp16), got {precision!r}")
    if policy.kind == "urm":
        # GameURM: the fp32-grade split-fp16 kernel unless the single-fp16-operand variant (1e-2 off the fp32 model) is asked for by name
        if precision == "fp32":
            raise ValueError("GameURM has no FFMA rollout kernel: precision is auto / x3 (fp32 grade) or fp16 (a labelled variant)")
        precision = "bf16" if precision in ("bf16", "fp16") else "x3"
    elif precision == "fp16":
        raise ValueError("precision fp16 names the GameURM variant; GameMLP takes auto, fp32, x3 or bf16")
    boards = env._req(boards, torch.int64, "boards")
    B = boards.numel()
    dev = env.init(boards.device)
    with torch.cuda.device(dev):
        buf = out if out is not None else RolloutBuffers.allocate(T, B, dev)
        assert buf.shape == (T, B)
        if forced_actions is not None:
            forced_actions = env._req(forced_actions, torch.uint8, "forced_actions")
            assert forced_actions.shape == (T, B)
        if alive is not None:
            alive = env._req(alive, torch.uint8, "alive")
        s = _RolloutStruct(B, T, policy.hidden, policy.layers, int(auto_reset), seed, env0, ctr0,
                           _dp(policy.weights), _dp(env.lut(dev)), _dp(boards), _dp(alive), _dp(forced_actions),
                           _dp(buf.boards), _dp(buf.actions), _dp(buf.legal), _dp(buf.logp), _dp(buf.value),
                           _dp(buf.points), _dp(buf.shaping), _dp(buf.flags), _dp(buf.entropy),
                           _PRECISION[resolve_precision(precision, B, policy.layers)], 0, _dp(_sched_workspace(dev, B)))
        if policy.kind == "urm":
            _lib.call("g2048_rollout_urm", C.byref(s), policy.loops, _stream())
        else:
            _lib.call("g2048_rollout_mlp", C.byref(s), _stream())
    return buf


def _check_overflow(flags: torch.Tensor) -> None:
    """A merge of two 32768 tiles would need exponent 16, which the packed 4-bit cells cannot hold (the kernels saturate
    the cell and raise G2048_FLAG_OVERFLOW; everything else of that transition is unspecified): never silent."""
    if bool(((flags & env.FLAG_OVERFLOW) != 0).any()):
        raise OverflowError("a rollout merged two 32768 tiles: the packed board format stops at exponent 15 (game.py allows 16)")


# ----------------------------------------------------------------------------- drop-in

_GAMES_PLAYED = itertools.count()
_CHUNK = 256


def _episode_dicts(bufs: list[RolloutBuffers], final_boards: torch.Tensor, device, game_state_on_device: bool = False) -> list[dict]:
    """Device records -> list[EpisodeData] exactly as train.py:299-345 builds them (host-side slow path)."""
    cat = lambda name: torch.cat([getattr(b, name) for b in bufs], dim=0)
    boards, flags = cat("boards"), cat("flags")
    T, B = flags.shape
    _check_overflow(flags)
    game_state = env.encode(boards.reshape(-1)).reshape(T, B, 48)
    if not game_state_on_device:
        game_state = game_state.cpu()
    ex = env.expand4(boards.reshape(-1))
    pts_possible = ex["points"].reshape(T, B, 4).cpu().numpy()
    pre_spawn = torch.gather(ex["succ"], 1, cat("actions").reshape(-1, 1).long()).reshape(-1)
    ext = env.potentials_ext(boards.reshape(-1), pre_spawn).reshape(T, B, 7).cpu().numpy()   # game.py:983-1001
    h = {k: cat(k).cpu().numpy() for k in ("actions", "legal", "logp", "value", "points", "shaping", "entropy")}
    boards_h, flags_h, final_h = boards.cpu().numpy(), flags.cpu().numpy(), final_boards.cpu().numpy()
    sh = {k: v.reshape(T, B) for k, v in env.decode_shaping(h["shaping"].reshape(-1)).items()}
    episodes = []
    for b in range(B):
        n = int((flags_h[:, b] & 0x80 != 0).sum())
        moves = []
        total_points = 0
        for t in range(n):
            done = bool(flags_h[t, b] & env.FLAG_DONE)
            lm = int(h["legal"][t, b])
            result = boards_h[t + 1, b] if t + 1 < n else final_h[b]
            pp = {d: int(pts_possible[t, b, i]) for i, d in enumerate(DIRECTIONS)}
            moves.append({
                "predicted_future_value": float(h["value"][t, b]),
                "selected_direction": int(h["actions"][t, b]),
                "game_state": game_state[t, b],
                "state_before": env.unpack_board(boards_h[t, b]),
                "result_state": env.unpack_board(result),
                "max_points_possible": max(pp.values()),
                "points_earned": int(h["points"][t, b]),
                "points_possible": pp,
                "action_mask": [not bool((lm >> i) & 1) for i in range(4)],
                "smoothness_delta": float(sh["smooth_after"][t, b] - sh["smooth_before"][t, b]),
                "max_tile_created": int(sh["max_tile_created"][t, b]),
                "max_exponent_before": int(sh["max_exp_before"][t, b]),
                "max_exponent_after": int(sh["max_exp_after"][t, b]),
                "corner_delta": float(sh["corner_after"][t, b] - sh["corner_before"][t, b]),
                "adjacency_delta": float(ext[t, b, 1] - ext[t, b, 0]),
                "chain_delta": float(ext[t, b, 3] - ext[t, b, 2]),
                "topological_delta": float(ext[t, b, 5] - ext[t, b, 4]),
                "monotonicity_after": int(sh["mono_after"][t, b]) if not done else 0.0,   # train.py:318-319
                "monotonicity_before": int(sh["mono_before"][t, b]),
                "emptiness_before": int(sh["empt_before"][t, b]),
                "emptiness_after": int(sh["empt_after"][t, b]) if not done else 0.0,      # train.py:322
                "entropy": float(h["entropy"][t, b]),
                "policy_logprobs": [float(x) for x in h["logp"][t, b]],
            })
            total_points += int(h["points"][t, b])
        ended = n > 0 and bool(flags_h[n - 1, b] & env.FLAG_DONE)
        episodes.append({"moves": moves, "total_points": total_points,
                         "total_steps": n - 1 if ended else n,                          # train.py:334-343
                         "final_state": env.unpack_board(final_h[b])})
    return episodes


@torch.no_grad()
def play_games_batched(model, num_games: int, max_steps: int | None = None, device=None, *, seed: int | None = None,
                       game_state_on_device: bool = False):
    """Play `num_games` games to the end (or `max_steps` moves each) -> list[EpisodeData].

    `game_state` tensors are returned on the HOST by default: the reference's symmetry augmentation builds the
    `game_state` of its copies with `Game2048(...).to_model_format()` on the CPU (train.py:840, 869) and `collate_fn` stacks
    both kinds in one minibatch (train.py:396) before `model_optimize_step` moves the batch to the device (train.py:468-474),
    so device tensors here would make the unmodified `train.py --gpu --upsample-ratio 0.25` fail in `torch.stack`
    (play_game_for_episode's own device tensors, train.py:257-258, 303, hit exactly that).  game_state_on_device=True keeps
    them on `device`."""
    if device is None or torch.device(device).type != "cuda":
        raise RuntimeError("play_games_batched runs on a CUDA device only (no CPU fallback); pass --gpu / device='cuda'")
    dev = env.init(device)
    policy = pack_policy(model, dev)
    if seed is None:      # like the reference's unseeded `random`: a fresh stream per process unless the caller seeds torch
        seed = int(os.environ["G2048_SEED"]) if "G2048_SEED" in os.environ else torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
    env0 = next(_GAMES_PLAYED) * (1 << 32)
    boards = env.reset(num_games, device=dev, seed=seed, env0=env0, ctr=0)
    alive = torch.ones(num_games, dtype=torch.uint8, device=dev)
    bufs, played = [], 0
    limit = max_steps if max_steps and max_steps > 0 else None
    while True:
        T = _CHUNK if limit is None else min(_CHUNK, limit - played)
        if T <= 0:
            break
        bufs.append(rollout(policy, boards, T, seed=seed, env0=env0, ctr0=1 + played, auto_reset=False, alive=alive))
        played += T
        if not bool(alive.any()):
            break
    return _episode_dicts(bufs, boards, dev, game_state_on_device)


@torch.no_grad()
def evaluate(model, eval_games: int = 100, max_steps: int | None = None, device=None, *, seed: int = 0) -> dict:
    """The reference's periodic evaluation (train.py:1840-1875): play `eval_games` seeded games to the end with the
    sampled policy and report the metrics it logs ("eval/max_score", "eval/avg_score", "eval/median_score",
    "eval/pct_512|1024|2048").  Game i uses the Philox stream (seed, env id i) where the reference seeds Python's
    `random` with i.  Everything stays on the device; no episode dictionaries are built."""
    if device is None or torch.device(device).type != "cuda":
        raise RuntimeError("evaluate runs on a CUDA device only (no CPU fallback)")
    dev = env.init(device)
    policy = pack_policy(model, dev)
    boards = env.reset(eval_games, device=dev, seed=seed, env0=0, ctr=0)
    alive = torch.ones(eval_games, dtype=torch.uint8, device=dev)
    scores = torch.zeros(eval_games, dtype=torch.int64, device=dev)
    played = 0
    limit = max_steps if max_steps and max_steps > 0 else None
    while True:
        T = _CHUNK if limit is None else min(_CHUNK, limit - played)
        if T <= 0:
            break
        buf = rollout(policy, boards, T, seed=seed, env0=0, ctr0=1 + played, auto_reset=False, alive=alive)
        _check_overflow(buf.flags)
        scores += (buf.points.long() * ((buf.flags & 0x80) != 0)).sum(0)
        played += T
        if not bool(alive.any()):
            break
    cells = (boards.unsqueeze(1) >> (4 * torch.arange(16, device=dev))) & 15          # final boards -> exponents
    max_tile = torch.where(cells.max(1).values > 0, 1 << cells.max(1).values, torch.zeros_like(scores))
    sc = scores.cpu().tolist()
    pct = lambda t: float((max_tile >= t).sum()) / eval_games * 100
    return {"eval/max_score": max(sc), "eval/avg_score": sum(sc) / len(sc), "eval/median_score": sorted(sc)[len(sc) // 2],
            "eval/pct_512": pct(512), "eval/pct_1024": pct(1024), "eval/pct_2048": pct(2048), "scores": sc}


def save_best_checkpoint(path, model, eval_avg_score: float, train_step: int) -> None:
    """best_model.pt in the reference's format (train.py:1887-1897): loads with the reference's own code."""
    cfg = model.config.model_dump() if hasattr(model.config, "model_dump") else dict(model.config)
    torch.save({"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "config": cfg,
                "eval_avg_score": eval_avg_score, "train_step": train_step}, path)
