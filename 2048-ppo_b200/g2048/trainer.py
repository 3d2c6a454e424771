"""Rollout + update driver for the hot path (the caller of the kernels, kept thin).

Mirrors what train.py's loop does per train step (train.py:1669-1737) with the batched engine:
fused rollout -> rewards-to-go / advantage -> PPO-clip + critic + entropy update, with the
reference's optimiser stack (Muon for 2-D weights + AdamW for 1-D, cosine schedule with warm-up,
train.py:1587-1612).  Orchestration only: the reference's CLI, logging, evaluation and
checkpointing are out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import dp, env, fused, ppo, rollout, update, urm_ops
from .policy import GameMLP, GameURM, GameURMConfig, MLPConfig


@dataclass
class TrainConfig:
    model_type: str = "mlp"        # "mlp" = GameMLP (the reference's only trainable model) or "urm" = GameURM: rollout on the
                                   # fused URM kernel, update through torch autograd on the policy mirror (truncated loops under
                                   # no_grad as in game.py:1400-1413) with the fused PPO-loss kernel -- SURVEY 8(f) N4, host half
    urm: GameURMConfig = field(default_factory=lambda: GameURMConfig(dropout=0.0))
    urm_chunk: int = 1 << 17       # samples per autograd chunk of the URM update (~0.4 MB of saved activations per sample: 50 GB;
                                   # a chunk is ~300 small launches, so short chunks leave the GPU waiting for the host)
    urm_update: str = "ops"        # "ops": every block op of the URM update on this library's kernels, forward and backward
                                   # (g2048/urm_ops.py: tcgen05 projections + hand-written attention / ConvSwiGLU / norm kernels);
                                   # "autograd": the torch mirror's own forward (ATen / cuBLAS) -- the comparator of the tests
    hidden_dim: int = 196
    num_layers: int = 2
    envs: int = 65536              # global env count (sharded over ranks)
    horizon: int = 512             # steps per rollout
    gamma: float = 0.99            # README flags (README.md:11-13)
    rtg_beta: float = 0.99
    points_weight: float = 0.10
    mono_weight: float = 1.0
    emptiness_weight: float = 0.0
    entropy_strength: float = 0.02
    critic_strength: float = 0.2
    clip_eps: float = 0.2          # train.py:518
    lr: float = 1e-3
    critic_lr: float = 1e-4
    weight_decay: float = 0.01
    warmup_steps: int = 10
    total_steps: int = 20000
    epochs: int = 1
    minibatches: int = 1           # optimizer steps per epoch (1 = full batch)
    chunk: int = 1 << 22           # samples per forward/backward chunk (measured: 4M is 7.6 % faster than 1M; ~26 GB live)
    seed: int = 2048
    zero_heads: bool = True        # train.py:1559-1567
    kl_stats: bool = False         # the KL(old || new) statistic the reference logs after every optimizer step (train.py:577-597):
                                   # a second, forward-only pass over the minibatch with the updated weights (GameMLP, fused update)
    dropout: float = 0.0           # Dropout of the residual blocks in the update forward (the reference's MLPConfig default is 0.1)
    bootstrap: bool = True         # the envs persist across train steps (auto-reset), so a buffer usually ends in the middle of
                                   # a game: start the return-to-go scan of such a column from the critic's value of the
                                   # carried-over board, V * sd + mu_c, instead of 0 (which biases the last ~1/(1-gamma) targets
                                   # low).  False = the reference's truncation, as for an episode cut by max_steps.
    upsample_ratio: float = 0.0    # symmetry augmentation (train.py:774-881; the README recipe uses 0.25): that share of the
                                   # recorded steps is drawn and mirrored / rotated copies join the update batch
    rollout_precision: str = "auto"   # "fp32" FFMA, "x3" split-fp16 tcgen05 (fp32 grade), "bf16" tcgen05 (a labelled variant, ~1e-2);
                                      # "auto" = x3 from rollout.TC_MIN_ENVS envs per GPU up, else fp32
    update_matmul: str = "fused"      # the update's forward/backward: "fused" = one tcgen05 kernel for forward + loss + backward-data
                                      # and x3 tensor-core weight gradients (g2048.update; split-fp16, ~1e-5), "x3" = torch autograd graph
                                      # with the x3 GEMM kernels (g2048.linear), "fp32" = autograd + cuBLAS SGEMM (reference
                                      # precision), "tf32" = autograd + cuBLAS TF32


def cosine_with_warmup(warmup: int, total: int):
    def f(step: int) -> float:
        if step < warmup:
            return step / max(1, warmup)
        prog = (step - warmup) / max(1, total - warmup)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))
    return f


class MultiOptimizer:  # train.py:1232-1281
    def __init__(self, *pairs):
        self.optimizers = [p[0] for p in pairs]
        self.schedulers = [p[1] for p in pairs]

    def step(self):
        for o in self.optimizers:
            o.step()

    def zero_grad(self):
        for o in self.optimizers:
            o.zero_grad(set_to_none=True)

    def scheduler_step(self):
        for s in self.schedulers:
            if s is not None:
                s.step()

    def get_lr(self):
        return [o.param_groups[0]["lr"] for o in self.optimizers]


def make_optimizer(model, cfg: TrainConfig) -> MultiOptimizer:
    o2, o1, v2, v1 = model.get_param_groups(cfg.critic_lr, cfg.lr)
    adamw = torch.optim.AdamW([o1, v1], betas=(0.9, 0.999), weight_decay=cfg.weight_decay)
    muon = torch.optim.Muon([o2, v2], adjust_lr_fn="match_rms_adamw", weight_decay=cfg.weight_decay)
    sched = lambda opt: torch.optim.lr_scheduler.LambdaLR(opt, cosine_with_warmup(cfg.warmup_steps, cfg.total_steps))
    return MultiOptimizer((muon, sched(muon)), (adamw, sched(adamw)))


@dataclass
class StepTimes:
    rollout_ms: float = 0.0
    advantage_ms: float = 0.0
    update_ms: float = 0.0
    allreduce_ms: float = 0.0       # the 3-scalar return-moment all-reduce (after the update)
    optimizer_ms: float = 0.0
    grad_allreduce_ms: float = 0.0  # the flat-gradient all-reduce of the last optimizer step (0 on one GPU)


class Trainer:
    def __init__(self, cfg: TrainConfig, device: torch.device, model: GameMLP | None = None):
        self.cfg, self.device = cfg, env.init(device)
        self.rank, self.world = dp.world()
        self.lo, self.hi = dp.shard_range(cfg.envs, self.rank, self.world)
        torch.manual_seed(cfg.seed)   # identical initial weights on every rank
        if cfg.urm_update not in ("ops", "autograd"):
            raise ValueError(f"urm_update must be 'ops' or 'autograd', got {cfg.urm_update!r}")
        if cfg.model_type not in ("mlp", "urm"):
            raise ValueError(f"model_type must be 'mlp' or 'urm', got {cfg.model_type!r}")
        if model is not None:
            self.model = model
        elif cfg.model_type == "urm":
            self.model = GameURM(cfg.urm)
        else:
            self.model = GameMLP(MLPConfig(hidden_dim=cfg.hidden_dim, num_layers=cfg.num_layers, dropout=cfg.dropout))
        self.is_mlp = isinstance(self.model, GameMLP) or hasattr(self.model, "backbone")
        self.model.to(self.device)
        if cfg.zero_heads and model is None:
            with torch.no_grad():
                for t in (self.model.action_head.weight, self.model.action_head.bias,
                          self.model.value_head.weight, self.model.value_head.bias):
                    t.zero_()
        self.opt = make_optimizer(self.model, cfg)
        self.bucket = dp.FlatGradBucket(self.model.parameters())
        self.moments = ppo.RtgMoments()
        self.B = self.hi - self.lo
        self.boards = env.reset(self.B, device=self.device, seed=cfg.seed, env0=self.lo, ctr=0)
        self.ctr = 1
        self.buf = rollout.RolloutBuffers.allocate(cfg.horizon, self.B, self.device)
        self.times = StepTimes()
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        self._aug_gen = torch.Generator(device=self.device)
        self._aug_gen.manual_seed(cfg.seed * 1000003 + self.rank)
        self.n_update_samples = 0
        self._ev_ar = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self._dropout_calls = 0

    def _dropout_seed(self) -> int:
        """A fresh Philox key per fused-update call (identical on every rank: the masks are keyed by the sample's index
        in its chunk, and ranks hold different samples)."""
        self._dropout_calls += 1
        return (self.cfg.seed * 0x9E3779B1 + self._dropout_calls * 0x85EBCA77 + self.rank * 0xC2B2AE3D) & 0xFFFFFFFFFFFFFFFF

    # -- phases ---------------------------------------------------------------------------
    def collect(self) -> rollout.RolloutBuffers:
        self.model.eval()
        pol = rollout.pack_policy(self.model, self.device)
        rollout.rollout(pol, self.boards, self.cfg.horizon, seed=self.cfg.seed, env0=self.lo, ctr0=self.ctr,
                        auto_reset=True, out=self.buf, precision=self.cfg.rollout_precision)
        self.ctr += self.cfg.horizon
        return self.buf

    def advantages(self, buf) -> dict:
        c = self.cfg
        mu_c, sd = self.moments.corrected(c.rtg_beta)
        boot = None
        if c.bootstrap:
            with torch.no_grad():
                self.model.eval()
                if self.is_mlp and update.supported(self.model):
                    _, v = update.forward(self.model, self.boards)
                elif not self.is_mlp and c.urm_update == "ops" and urm_ops.supported(self.model):
                    v = torch.cat([urm_ops.forward(self.model, env.encode(self.boards[i:i + c.urm_chunk]))[1]
                                   for i in range(0, self.boards.numel(), c.urm_chunk)])
                else:
                    _, v = self.model(env.encode(self.boards))
                boot = v.reshape(-1).float() * (sd + 1e-8) + mu_c        # the critic predicts the NORMALISED return (train.py:760-772)
        return ppo.rtg_advantage(buf.points, buf.shaping, buf.flags, buf.value, gamma=c.gamma,
                                 w_points=c.points_weight, w_mono=c.mono_weight, w_empt=c.emptiness_weight,
                                 mu_c=mu_c, stddev=sd, bootstrap=boot)

    def update(self, buf, adv) -> dict:
        c = self.cfg
        n_local = buf.flags.numel()
        flat = lambda t, *s: t.reshape(n_local, *s)
        boards, actions, legal, flags = flat(buf.boards), flat(buf.actions), flat(buf.legal), flat(buf.flags)
        logp, a, g = flat(buf.logp, 4), flat(adv["adv"]), flat(adv["g_norm"])
        if c.upsample_ratio > 0:
            # augmented copies keep the ORIGINAL advantage / return / value of their source step (train.py:829,858) and
            # are appended like the reference's "augmented" pseudo-episode (train.py:1710-1718)
            src, ops = ppo.sample_augmentation((flags & 0x80) != 0, c.upsample_ratio, self._aug_gen)
            if src.numel() > 0:
                aug = env.augment(boards[src], boards[src], actions[src], legal[src], logp[src], ops)
                boards, actions, legal = torch.cat([boards, aug["before"]]), torch.cat([actions, aug["action"]]), torch.cat([legal, aug["legal"]])
                logp, a, g, flags = torch.cat([logp, aug["logp"]]), torch.cat([a, a[src]]), torch.cat([g, g[src]]), torch.cat([flags, flags[src]])
                n_local = boards.numel()
        self.n_update_samples = n_local
        c_mb = c.minibatches
        # minibatch m of this rank = samples [cut[m], cut[m+1]) of its (shuffled) shard: every rank runs exactly
        # c.minibatches optimizer steps (matching all-reduces even when a shard is short or empty) and the loss divisor
        # of a step is the all-reduced sample count of that minibatch, so the summed gradient is the global mean
        # (train.py:554) whatever the shard sizes are
        cut = [(n_local * m) // c_mb for m in range(c_mb + 1)]
        counts = torch.tensor([float(cut[m + 1] - cut[m]) for m in range(c_mb)], dtype=torch.float64, device=self.device)
        n_mb_global = [int(x) for x in dp.allreduce_stats(counts).tolist()]
        self.model.train()
        last = None
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = c.update_matmul == "tf32"
        for _ in range(c.epochs):
            order = None if c_mb == 1 else torch.randperm(n_local, device=self.device)
            for m in range(c_mb):
                m0, m1 = cut[m], cut[m + 1]
                self.bucket.zero()
                tot = torch.zeros(4, dtype=torch.float64, device=self.device)
                use_fused = self.is_mlp and c.update_matmul == "fused" and update.supported(self.model)
                packed = update.pack(self.model) if use_fused else None
                chunk = c.chunk if self.is_mlp else c.urm_chunk
                want_kl = use_fused and c.kl_stats
                old_logits = torch.empty((m1 - m0, 4), dtype=torch.float32, device=self.device) if want_kl else None
                for c0 in range(m0, m1, chunk):
                    sl = slice(c0, min(m1, c0 + chunk)) if order is None else order[c0:min(m1, c0 + chunk)]
                    if use_fused:
                        tot += update.loss_and_grads(self.model, boards[sl], actions[sl], legal[sl], logp[sl], a[sl], g[sl],
                                                     flags=flags[sl], clip_eps=c.clip_eps, critic_strength=c.critic_strength,
                                                     entropy_strength=c.entropy_strength, n_total=max(n_mb_global[m], 1), packed=packed,
                                                     dropout_p=c.dropout, dropout_seed=self._dropout_seed(),
                                                     logits_out=old_logits[c0 - m0:min(m1, c0 + chunk) - m0] if want_kl else None)
                        continue
                    if self.is_mlp:
                        logits, v = fused.mlp_forward(self.model, env.encode(boards[sl]),
                                                      matmul="x3" if c.update_matmul in ("x3", "fused") else "cublas")
                    elif c.urm_update == "ops":
                        logits, v = urm_ops.forward(self.model, env.encode(boards[sl]))
                    else:
                        logits, v = self.model(env.encode(boards[sl]))
                    loss, stats = ppo.ppo_loss(logits, v, logp[sl], actions[sl], legal[sl], a[sl], g[sl], flags=flags[sl],
                                               clip_eps=c.clip_eps, critic_strength=c.critic_strength,
                                               entropy_strength=c.entropy_strength, n_total=max(n_mb_global[m], 1))
                    loss.backward()
                    tot += stats
                e_ar = self._ev_ar
                e_ar[0].record()
                self.bucket.allreduce()
                e_ar[1].record()
                gn = self.bucket.clip_norm_(1.0)                                    # train.py:561
                self.opt.step()
                kl = None
                if want_kl:                       # train.py:577-597: forward of the UPDATED model (train() mode: a fresh dropout mask)
                    packed = update.pack(self.model)
                    kl = torch.zeros(3, dtype=torch.float64, device=self.device)
                    for c0 in range(m0, m1, chunk):
                        c1 = min(m1, c0 + chunk)
                        sl = slice(c0, c1) if order is None else order[c0:c1]
                        new_logits, _ = update.forward(self.model, boards[sl], packed, dropout_p=c.dropout, dropout_seed=self._dropout_seed())
                        ks, _ = ppo.masked_kl(old_logits[c0 - m0:c1 - m0], new_logits, legal[sl], flags=flags[sl])
                        kl[:2] += ks[:2]
                        kl[2] = torch.maximum(kl[2], ks[2])
                last = (tot, gn, kl)
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        self.opt.scheduler_step()                                                   # train.py:625
        tot = dp.allreduce_stats(last[0].clone())
        out = ppo.loss_stats(tot, c.critic_strength, c.entropy_strength)
        out["grad_norm"] = float(last[1])
        if last[2] is not None:
            kl = last[2]
            if dp.world()[1] > 1 and dp.COLLECTIVES:
                mx = kl[2:3].clone()
                torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
                dp.allreduce_stats(kl)
                kl[2] = mx[0]
            s_kl, n_kl, m_kl = kl.tolist()
            out.update(kl_total=s_kl, kl_average=s_kl / max(n_kl, 1.0), kl_max=m_kl)
        return out

    def finish_moments(self, adv) -> None:
        s = dp.allreduce_stats(adv["stats"].clone())
        s1, s2, n = s.tolist()
        self.moments.update(self.cfg.rtg_beta, s1, s2, n)

    # -- one train step -------------------------------------------------------------------
    def train_step(self) -> dict:
        e = self._ev
        e[0].record()
        buf = self.collect()
        e[1].record()
        adv = self.advantages(buf)
        e[2].record()
        stats = self.update(buf, adv)
        e[3].record()
        self.finish_moments(adv)
        e[4].record()
        torch.cuda.synchronize(self.device)
        self.times = StepTimes(e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]),
                               e[3].elapsed_time(e[4]), 0.0, self._ev_ar[0].elapsed_time(self._ev_ar[1]))
        stats["env_steps"] = buf.flags.numel() * self.world
        stats["mean_points_per_step"] = float(buf.points.float().mean())
        return stats
