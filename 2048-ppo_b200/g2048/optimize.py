"""Drop-in for the reference's `train.model_optimize_step` (train.py:414-642) on the fused GPU kernels.

Same signature, same statistics dictionary, same minibatch composition: the reference shuffles with
`DataLoader(shuffle=True)` (train.py:438-443), which draws two int64 values from torch's global RNG per
epoch (the iterator's base seed, then the RandomSampler seed) and cuts `torch.randperm(n, generator)` into
batches; `_epoch_order` repeats exactly that, so under the same `torch.manual_seed` every optimizer step
sees the same samples as the reference's.  Per minibatch: forward + PPO-clip / critic / entropy loss +
backward through `g2048.update.loss_and_grads` (one tcgen05 kernel + tensor-core weight gradients),
`clip_grad_norm_(1.0)`, `optimizer.step()`, `optimizer.zero_grad()`, then KL(old || new) from a forward-only
pass of the same kernel (train.py:575-598).

The model must be a GameMLP (the reference's or ours) on a CUDA device.  Like the reference, the update runs in
train() mode: the blocks' Dropout(MLPConfig.dropout, 0.1 by default) is active in the loss forward AND in the
KL forward (train.py:483, 579), with Philox masks keyed from torch's global generator (g2048.update.dropout_mask
restates them; torch's own dropout stream is not reproducible across implementations).  No CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ppo, update


def _epoch_order(n: int) -> torch.Tensor:
    """Sample order of one pass of DataLoader(dataset, shuffle=True) with no generator argument."""
    torch.empty((), dtype=torch.int64).random_()                        # _BaseDataLoaderIter base seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())     # RandomSampler.__iter__
    gen = torch.Generator()
    gen.manual_seed(seed)
    return torch.randperm(n, generator=gen)


def episodes_to_batch(episodes, device) -> dict:
    """MyDataset + collate_fn (train.py:360-411) for the whole dataset at once: device tensors boards (packed
    from the exponent channel of `game_state`), actions, legal bits (complement of `action_mask`), advantage,
    future_reward, old log-probs [n, 4]."""
    moves = [m for ep in episodes for m in ep["moves"]]
    n = len(moves)
    if n == 0:
        raise ValueError("model_optimize_step: no moves in the episodes")
    x = torch.stack([m["game_state"].detach().to("cpu", torch.float32) for m in moves]).numpy()
    exps = np.rint(x[:, 0::3]).astype(np.uint64)                        # to_model_format: [exp, r/3, c/3] per cell
    if exps.size and int(exps.max()) > 15:
        raise ValueError("model_optimize_step: a board holds exponent 16 (a 65536 tile); the packed 4-bit board format of the "
                         "kernels stops at 15 (32768) -- game.py's own ceiling is 16 (game.py:59-60)")
    boards = np.zeros(n, dtype=np.uint64)
    for i in range(16):
        boards |= exps[:, i] << np.uint64(4 * i)
    mask = np.array([m["action_mask"] for m in moves], dtype=bool)      # True = illegal (train.py:268)
    legal = ((~mask).astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(axis=1).astype(np.uint8)
    f32 = lambda key: torch.tensor([m[key] for m in moves], dtype=torch.float32)
    return dict(
        boards=torch.from_numpy(boards.view(np.int64)).to(device),
        actions=torch.tensor([m["selected_direction"] for m in moves], dtype=torch.uint8).to(device),
        legal=torch.from_numpy(legal).to(device),
        adv=f32("advantage").to(device), g_norm=f32("future_reward").to(device),
        logp=f32("policy_logprobs").reshape(n, 4).to(device))


def optimize_batch(model, batch: dict, optimizer, lr_scheduler=None, kl_strength: float = 0.1, critic_strength: float = 1.0,
                   batch_size: int = 32, epochs: int = 1) -> dict:
    """model_optimize_step on an already collated dataset (see episodes_to_batch)."""
    n = batch["boards"].numel()
    tot = dict(loss=0.0, policy_loss=0.0, entropy_loss=0.0, value_loss=0.0, grad_norm=0.0, entropy=0.0, kl_total=0.0,
               kl_average=0.0)
    max_kl, num_batches, current_lr = 0.0, 0, 0.0
    for _ in range(epochs):
        order = _epoch_order(n).to(batch["boards"].device)
        for i in range(0, n, batch_size):
            idx = order[i:i + batch_size]
            b = {k: v[idx] for k, v in batch.items()}
            nb = idx.numel()
            model.train()                                                        # train.py:483
            old_logits = torch.empty((nb, 4), dtype=torch.float32, device=idx.device)     # for the KL statistic
            stats = update.loss_and_grads(model, b["boards"], b["actions"], b["legal"], b["logp"], b["adv"], b["g_norm"],
                                          clip_eps=0.2, critic_strength=critic_strength, entropy_strength=kl_strength,
                                          logits_out=old_logits)
            grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)   # train.py:561
            optimizer.step()
            optimizer.zero_grad()
            current_lr = lr_scheduler.get_last_lr()[0] if lr_scheduler is not None else 0.0
            with torch.no_grad():
                new_logits, _ = update.forward(model, b["boards"])
                kl_sum, _, kl_max = ppo.masked_kl(old_logits, new_logits, b["legal"])[0].tolist()   # train.py:586-597
            s_ppo, s_vl, s_ent, _ = (float(v) for v in stats.tolist())
            tot["loss"] += -(s_ppo - critic_strength * s_vl + kl_strength * s_ent) / nb
            tot["policy_loss"] += -s_ppo / nb
            tot["entropy_loss"] += -kl_strength * s_ent / nb
            tot["value_loss"] += critic_strength * s_vl / nb
            tot["grad_norm"] += float(grad_norm)
            tot["entropy"] += s_ent / nb
            tot["kl_total"] += kl_sum
            tot["kl_average"] += kl_sum / nb
            max_kl = max(max_kl, kl_max)
            num_batches += 1
    optimizer.scheduler_step()                                                   # train.py:625
    out = {k: v / num_batches for k, v in tot.items()}
    out["kl_max"] = max_kl
    out["lr"] = current_lr
    return out


def model_optimize_step(model, episodes, optimizer, lr_scheduler=None, kl_strength: float = 0.1,
                        critic_strength: float = 1.0, device=None, batch_size: int = 32, epochs: int = 1) -> dict:
    """train.model_optimize_step (train.py:414-642): same arguments, same returned statistics."""
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    if dev.type != "cuda":
        raise ValueError("g2048.optimize.model_optimize_step runs on CUDA devices only (there is no CPU path)")
    if next(model.parameters()).device != dev:
        model.to(dev)
    if not update.supported(model):
        raise ValueError("g2048.optimize.model_optimize_step: GameMLP with hidden % 4 == 0 in [16, 208] and 1-2 blocks required")
    return optimize_batch(model, episodes_to_batch(episodes, dev), optimizer, lr_scheduler, kl_strength, critic_strength,
                          batch_size, epochs)
