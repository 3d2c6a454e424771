"""Shared test infrastructure: fixture layout helpers and a plain-torch restatement of the
reference's loss section (train.py:497-554) used as the fp32 checker for the fused loss kernel."""
import numpy as np
import torch


def shaping_words(mono_b, mono_a, empt_b, empt_a, max_tile=None):
    w = (np.asarray(mono_b, dtype=np.uint64) | np.asarray(mono_a, dtype=np.uint64) << np.uint64(6) |
         np.asarray(empt_b, dtype=np.uint64) << np.uint64(12) | np.asarray(empt_a, dtype=np.uint64) << np.uint64(17))
    if max_tile is not None:
        w |= np.asarray(max_tile, dtype=np.uint64) << np.uint64(22)
    return w.view(np.int64)


def rollout_as_tb(g):
    """Golden rollout.npz -> time-major [T,B] arrays (one game per column) + flat<->tb index maps."""
    n_env = len(g["ep_len"])
    T = int(g["ep_len"].max())
    z = lambda dt: np.zeros((T, n_env), dtype=dt)
    a = dict(points=z(np.int32), shaping=z(np.int64), flags=z(np.uint8), value=z(np.float32),
             board=z(np.int64), action=z(np.uint8), legal=z(np.uint8), logp=np.zeros((T, n_env, 4), np.float32))
    t, env = g["t"], g["env"]
    a["points"][t, env] = g["points"]
    a["shaping"][t, env] = shaping_words(g["mono_before"], g["mono_after"], g["empt_before"], g["empt_after"],
                                         g["max_tile_created"])
    a["flags"][t, env] = 0x80 | (g["done"].astype(np.uint8) << 4)
    a["value"][t, env] = g["value"]
    a["board"][t, env] = g["board"].view(np.int64)
    a["action"][t, env] = g["action"]
    a["legal"][t, env] = g["legal"]
    a["logp"][t, env] = g["logp"]
    return a, (t, env)


def ref_ppo_loss_torch(logits, value, old_logp4, actions, legal_mask_bits, adv, g_norm, clip_eps=0.2,
                       critic_strength=1.0, entropy_strength=0.1):
    """train.py:497-554 in plain torch (autograd-able).  Returns (loss, dict of means)."""
    bits = torch.arange(4, device=logits.device)
    illegal = ((legal_mask_bits.long()[:, None] >> bits) & 1) == 0          # action_mask: True = illegal
    masked = torch.masked_fill(logits, illegal, float("-inf"))
    new_lp = masked.log_softmax(dim=-1)
    idx = actions.long().unsqueeze(1)
    nl = torch.gather(new_lp, -1, idx)
    ol = torch.gather(old_logp4, -1, idx)
    ratio = (nl - ol).squeeze(1).clamp(-20, 20).exp()
    clipped = ratio.clamp(1 - clip_eps, 1 + clip_eps)
    ppo = torch.minimum(adv * ratio, adv * clipped)
    mlc = masked.clamp(-20, 20)
    mlp = torch.log_softmax(mlc, dim=-1)
    ent = -torch.masked.sum(mlp * mlp.exp(), dim=-1, mask=~illegal)
    vl = torch.nn.functional.smooth_l1_loss(value.view(-1), g_norm, reduction="none")
    loss = -(ppo - critic_strength * vl + entropy_strength * ent).mean()
    return loss, dict(ppo=ppo.mean(), vl=vl.mean(), ent=ent.mean())
