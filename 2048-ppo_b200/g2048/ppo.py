"""Advantage and loss of the reference's update step (train.py:414-772) on fused CUDA kernels.

  rtg_advantage(...)  <-> train.calculate_advantage (first half, train.py:651-772, 898-904)
  ppo_loss(...)       <-> the loss section of train.model_optimize_step (train.py:497-554),
                          a torch.autograd.Function whose backward is the kernel's analytic
                          gradient w.r.t. logits and value; the MLP trunk stays torch autograd.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from .env import _ptr, _req, _stream, init

_lib.register("g2048_rtg_advantage",
              [C.c_void_p] * 4 + [C.c_int32, C.c_int64] + [C.c_double] * 6 + [C.c_void_p] * 6 + [C.c_void_p])
_lib.register("g2048_rtg_advantage_bootstrap",
              [C.c_void_p] * 5 + [C.c_int32, C.c_int64] + [C.c_double] * 6 + [C.c_void_p] * 6 + [C.c_void_p])
_lib.register("g2048_ppo_loss",
              [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_int64] + [C.c_float] * 4 + [C.c_void_p] * 5)
_lib.register("g2048_masked_kl", [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 4)
_lib.lib().g2048_reduce_workspace_bytes.restype = C.c_int64

_WS: dict[int, torch.Tensor] = {}


def _workspace(dev: torch.device) -> torch.Tensor:
    if dev.index not in _WS:
        _WS[dev.index] = torch.empty(int(_lib.lib().g2048_reduce_workspace_bytes()), dtype=torch.uint8, device=dev)
    return _WS[dev.index]


@dataclass
class RtgMoments:
    """EMA moments of the return-to-go (train.py:1550-1552: mu = 0, m2 = 1 at start)."""
    mu: float = 0.0
    m2: float = 1.0
    step: int = 1          # 1-indexed train step (train.py:1705)

    def corrected(self, beta: float):
        eps = 1e-8
        bias = max(1 - beta ** max(self.step, 1), eps)      # train.py:746
        mu_c = self.mu / bias                                 # train.py:749
        var = max(self.m2 / bias - mu_c ** 2, eps)            # train.py:752-753
        return mu_c, var ** 0.5                               # train.py:754

    def update(self, beta: float, s1: float, s2: float, n: float) -> None:
        """train.py:738-739 + 898-901 from the (possibly all-reduced) sums {sum G, sum G^2, N}."""
        if n <= 0:
            return
        mean = s1 / n
        var = 0.0 if n <= 1 else max(s2 / n - mean * mean, 0.0)
        self.mu = beta * self.mu + (1 - beta) * mean
        self.m2 = beta * self.m2 + (1 - beta) * (var + mean * mean)
        self.step += 1


def rtg_advantage(points, shaping, flags, value, *, gamma, w_points, w_mono, w_empt, mu_c, stddev,
                  want_raw: bool = False, bootstrap: torch.Tensor | None = None) -> dict:
    """Time-major [T,B] rollout buffers -> g_norm, adv (float32 [T,B]) and stats (float64[3] on
    the device: sum G, sum G^2, count).  want_raw also returns reward and the raw return.
    bootstrap: optional float32 [B], the raw return-to-go expected after the last slot of every column (games that
    continue past a fixed-horizon buffer); None = 0, the reference's truncation (train.py:724-728)."""
    points = _req(points, torch.int32, "points")
    shaping = _req(shaping, torch.int64, "shaping")
    flags = _req(flags, torch.uint8, "flags")
    value = _req(value, torch.float32, "value")
    T, B = points.shape
    dev = init(points.device)
    with torch.cuda.device(dev):
        g_norm = torch.empty((T, B), dtype=torch.float32, device=dev)
        adv = torch.empty_like(g_norm)
        reward = torch.empty_like(g_norm) if want_raw else None
        g_raw = torch.empty_like(g_norm) if want_raw else None
        stats = torch.empty(3, dtype=torch.float64, device=dev)
        if bootstrap is not None:
            bootstrap = _req(bootstrap.reshape(-1), torch.float32, "bootstrap")
            assert bootstrap.numel() == B
        _lib.call("g2048_rtg_advantage_bootstrap", _ptr(points), _ptr(shaping), _ptr(flags), _ptr(value), _ptr(bootstrap), T, B,
                  float(gamma), float(w_points), float(w_mono), float(w_empt), float(mu_c), float(stddev),
                  _ptr(reward), _ptr(g_raw), _ptr(g_norm), _ptr(adv), _ptr(stats), _ptr(_workspace(dev)), _stream())
    return dict(g_norm=g_norm, adv=adv, reward=reward, g_raw=g_raw, stats=stats)


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, value, old_logp, actions, legal, flags, adv, g_norm, clip_eps, c_v, beta_ent, n_total):
        logits_c = _req(logits.detach(), torch.float32, "logits")
        value_c = _req(value.detach().reshape(-1), torch.float32, "value")
        n = logits_c.shape[0]
        assert logits_c.shape == (n, 4)
        old_logp = _req(old_logp, torch.float32, "old_logp")
        stride = 4 if old_logp.numel() == 4 * n else 1
        assert old_logp.numel() == stride * n
        dev = init(logits_c.device)
        with torch.cuda.device(dev):
            dlogits = torch.empty_like(logits_c)
            dvalue = torch.empty_like(value_c)
            stats = torch.empty(4, dtype=torch.float64, device=dev)
            inv_n = 1.0 / float(n_total if n_total else max(n, 1))
            _lib.call("g2048_ppo_loss", _ptr(logits_c), _ptr(value_c), _ptr(old_logp), stride,
                      _ptr(_req(actions, torch.uint8, "actions")), _ptr(_req(legal, torch.uint8, "legal")),
                      _ptr(None if flags is None else _req(flags, torch.uint8, "flags")),
                      _ptr(_req(adv, torch.float32, "adv")), _ptr(_req(g_norm, torch.float32, "g_norm")), n,
                      float(clip_eps), float(c_v), float(beta_ent), inv_n, _ptr(dlogits), _ptr(dvalue),
                      _ptr(stats), _ptr(_workspace(dev)), _stream())
        ctx.save_for_backward(dlogits, dvalue)
        ctx.value_shape = value.shape
        s = stats * inv_n                                   # means of ppo, smooth-l1, entropy (all 0 for an empty batch)
        loss = (-(s[0] - c_v * s[1] + beta_ent * s[2])).to(torch.float32)
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        dlogits, dvalue = ctx.saved_tensors
        return (dlogits * grad_loss, (dvalue * grad_loss).reshape(ctx.value_shape)) + (None,) * 10


def ppo_loss(logits, value, old_logp, actions, legal, adv, g_norm, *, flags=None, clip_eps=0.2,
             critic_strength=1.0, entropy_strength=0.1, n_total: int | None = None):
    """Returns (loss, stats) with stats = float64[4] device tensor {sum ppo, sum smooth_l1,
    sum entropy, count}.  `n_total`: the divisor of the mean (defaults to this batch's size; pass the
    global sample count when the batch is a shard or a chunk of a larger minibatch)."""
    return _PPOLoss.apply(logits, value, old_logp, actions, legal, flags, adv, g_norm, clip_eps,
                          critic_strength, entropy_strength, n_total)


def masked_kl(old_logits, new_logits, legal, *, flags=None, want_per_sample: bool = False):
    """The KL(old || new) statistic of model_optimize_step (train.py:577-597): per sample the sum over the legal moves of
    p_old (log p_old - log p_new) with both distributions the masked softmax of their logits [n, 4].
    Returns (stats, kl): stats = float64[3] device tensor {sum KL, count, max KL} over the valid slots (deterministic
    reduction), kl = float32 [n] per-sample values or None."""
    old_logits = _req(old_logits, torch.float32, "old_logits")
    new_logits = _req(new_logits, torch.float32, "new_logits")
    n = old_logits.shape[0]
    assert old_logits.shape == (n, 4) and new_logits.shape == (n, 4)
    dev = init(old_logits.device)
    with torch.cuda.device(dev):
        stats = torch.empty(3, dtype=torch.float64, device=dev)
        kl = torch.empty(n, dtype=torch.float32, device=dev) if want_per_sample else None
        _lib.call("g2048_masked_kl", _ptr(old_logits), _ptr(new_logits), _ptr(_req(legal, torch.uint8, "legal")),
                  _ptr(None if flags is None else _req(flags, torch.uint8, "flags")), n, _ptr(kl), _ptr(stats),
                  _ptr(_workspace(dev)), _stream())
    return stats, kl


def sample_augmentation(valid: torch.Tensor, upsample_ratio: float, generator: torch.Generator | None = None):
    """The sampling policy of the reference's symmetry augmentation (train.py:776-863) on the device:
    int(N * ratio) of the N recorded steps are drawn without replacement; each drawn step independently yields a
    mirrored copy with probability 1/2 (axis uniform) and a rotated copy with probability 1/2 (90/180/270 uniform).
    `valid`: bool [n] mask of recorded steps.  Returns (source indices int64 [m], ops uint8 [m]) for env.augment.
    (The reference draws from Python's `random`; the distribution is the same, the draws are torch's.)"""
    from .env import MIRROR_H, ROT90
    idx = torch.nonzero(valid.reshape(-1), as_tuple=False).reshape(-1)
    n = idx.numel()
    k = min(int(n * upsample_ratio), n)
    dev = valid.device
    if k <= 0:
        return idx[:0], torch.empty(0, dtype=torch.uint8, device=dev)
    pick = idx[torch.randperm(n, device=dev, generator=generator)[:k]]
    u = torch.rand((4, k), device=dev, generator=generator)
    mirror, rotate = u[0] < 0.5, u[2] < 0.5
    m_op = (MIRROR_H + (u[1] >= 0.5).to(torch.uint8)).to(torch.uint8)                       # horizontal | vertical
    r_op = (ROT90 + torch.clamp((u[3] * 3).to(torch.uint8), max=2)).to(torch.uint8)         # 90 | 180 | 270
    src = torch.cat([pick[mirror], pick[rotate]])
    ops = torch.cat([m_op[mirror], r_op[rotate]])
    return src, ops


def loss_stats(stats: torch.Tensor, critic_strength: float, entropy_strength: float) -> dict:
    """The scalars model_optimize_step logs (train.py:528,541,546,554,618)."""
    s1, s2, s3, n = (float(x) for x in stats.tolist())
    n = max(n, 1.0)
    return {"loss": -(s1 - critic_strength * s2 + entropy_strength * s3) / n, "policy_loss": -s1 / n,
            "value_loss": critic_strength * s2 / n, "entropy": s3 / n, "entropy_loss": -entropy_strength * s3 / n}
