"""The policy update's forward + loss + backward as hand-written kernels (no autograd graph).

`loss_and_grads(model, ...)` is what train.model_optimize_step does between `optimizer.zero_grad()` and
`clip_grad_norm_` (train.py:491-556) for a GameMLP (the blocks' Dropout active in train() mode, as in the reference,
with Philox masks -- see dropout_mask): it ADDS d loss / d parameter into every `p.grad` and returns the loss sums.  One fused tcgen05 kernel (g2048_update_mlp_fwd_bwd) runs the forward,
the PPO-clip / critic / entropy terms and the backward-data chain per 128-sample tile; the weight gradients
are split-fp16 tcgen05 reductions over samples (g2048_x3_wgrad_images).  There is no fallback: unsupported model
shapes raise (use g2048.fused.mlp_forward + g2048.ppo.ppo_loss with autograd for those).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, env, linear
from .env import _ptr, _req, _stream, init

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class _UpdateMlp(C.Structure):   # include/g2048.h: G2048UpdateMlp
    _fields_ = [("n", i64), ("hidden", i32), ("layers", i32), ("decouple_critic", i32), ("backward", i32),
                ("boards", vp), ("actions", vp), ("legal", vp), ("flags", vp), ("old_logp", vp),
                ("old_logp_stride", i32), ("reserved_", i32), ("adv", vp), ("g_norm", vp),
                ("clip_eps", f32), ("critic_strength", f32), ("entropy_strength", f32), ("inv_n", f32),
                ("packed", vp), ("workspace", vp), ("h_out", vp), ("dz_out", vp), ("dhead", vp), ("logits", vp),
                ("value", vp), ("ln_grad", vp), ("head_bias_grad", vp), ("stats", vp),
                ("dropout_p", f32), ("reserved2_", i32), ("dropout_seed", C.c_uint64), ("dropout_sample0", C.c_uint64)]


_lib.register("g2048_update_mlp_pack", [i32, i32] + [vp] * 12)
_lib.register("g2048_update_mlp_fwd_bwd", [C.POINTER(_UpdateMlp), vp])
for _name in ("g2048_update_mlp_pack_bytes", "g2048_update_mlp_workspace_bytes"):
    getattr(_lib.lib(), _name).restype = i64
    getattr(_lib.lib(), _name).argtypes = [i32, i32]
_lib.lib().g2048_update_mlp_padded.restype = i32
_lib.lib().g2048_update_mlp_padded.argtypes = [i32]


def padded_width(h: int) -> int:
    """Column count of the kernel's operand tiles / images for hidden size h (64, 128, 192 or 208)."""
    return int(_lib.lib().g2048_update_mlp_padded(h))


def loss_scale(n_total: int) -> float:
    """Power-of-two factor the gradient chain is multiplied with inside the kernels, so that its fp16 operand terms sit at
    O(1) like the activations: the per-sample gradients carry 1 / n_total (train.py:554 is a mean), which for 3e7 samples
    is below the fp16 range.  Everything the kernels emit is divided by it again (exactly) when it is added to p.grad."""
    return float(2 ** max(0, int(n_total).bit_length() - 1))

_WS: dict[tuple[int, int, int], torch.Tensor] = {}


def supported(model) -> bool:
    h, L = model.stem[0].weight.shape[0], len(model.backbone)
    ps = {float(blk.mlp[3].p) for blk in model.backbone}
    return 16 <= h <= 208 and h % 4 == 0 and 1 <= L <= 2 and len(ps) == 1 and 0.0 <= min(ps) < 1.0 and model.stem[0].weight.shape[1] == 48


def model_dropout_p(model) -> float:
    """The dropout probability the reference's forward would apply right now (nn.Dropout is the identity in eval mode)."""
    return float(model.backbone[0].mlp[3].p) if model.training and len(model.backbone) else 0.0


def fresh_dropout_seed() -> int:
    """A 64-bit Philox key drawn from torch's global CPU generator (so torch.manual_seed makes the masks reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def dropout_mask(n: int, h: int, L: int, p: float, seed: int, sample0: int = 0):
    """The keep mask the fused kernel applies, as a bool array [L, n, h] (include/g2048.h G2048UpdateMlp.dropout_p):
    numpy restatement of Philox4x32-10 for tests and for anyone who needs to reproduce a step."""
    import numpy as np
    groups = (h + 7) // 8
    thr = int(p * 65536.0 + 0.5)
    i = (np.arange(n, dtype=np.uint64) + np.uint64(sample0))[None, :, None]
    c0 = np.broadcast_to((i & np.uint64(0xFFFFFFFF)), (L, n, groups)).copy()
    c1 = np.broadcast_to((i >> np.uint64(32)), (L, n, groups)).copy()
    c2 = np.broadcast_to(np.arange(1, L + 1, dtype=np.uint64)[:, None, None], (L, n, groups)).copy()
    c3 = np.broadcast_to(np.arange(groups, dtype=np.uint64)[None, None, :], (L, n, groups)).copy()
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    lanes = np.stack([c0 & np.uint64(0xFFFF), c0 >> np.uint64(16), c1 & np.uint64(0xFFFF), c1 >> np.uint64(16),
                      c2 & np.uint64(0xFFFF), c2 >> np.uint64(16), c3 & np.uint64(0xFFFF), c3 >> np.uint64(16)], axis=-1)
    return (lanes >= np.uint64(thr)).reshape(L, n, groups * 8)[:, :, :h]


def _shape(model):
    return model.stem[0].weight.shape[0], len(model.backbone)


def _workspace(dev, h, L):
    key = (dev.index, h, L)
    if key not in _WS:
        _WS[key] = torch.empty(int(_lib.lib().g2048_update_mlp_workspace_bytes(h, L)), dtype=torch.uint8, device=dev)
    return _WS[key]


def pack(model) -> torch.Tensor:
    """Kernel-side image of the model's current parameters (re-pack after every optimizer step)."""
    if not supported(model):
        raise ValueError("g2048.update: GameMLP with hidden % 4 == 0 in [16, 208] and 1-2 blocks required")
    h, L = _shape(model)
    dev = init(model.stem[0].weight.device)
    f = lambda t: _req(t.detach(), torch.float32, "parameter")
    keep = [f(model.stem[0].weight), f(model.stem[1].weight), f(model.stem[1].bias)]
    bw = [f(b.mlp[0].weight) for b in model.backbone]
    bg = [f(b.mlp[1].weight) for b in model.backbone]
    bb = [f(b.mlp[1].bias) for b in model.backbone]
    heads = [f(model.action_head.weight), f(model.action_head.bias), f(model.value_head.weight), f(model.value_head.bias)]
    arr = lambda ts: (vp * len(ts))(*[t.data_ptr() for t in ts])
    with torch.cuda.device(dev):
        packed = torch.empty(int(_lib.lib().g2048_update_mlp_pack_bytes(h, L)), dtype=torch.uint8, device=dev)
        _lib.call("g2048_update_mlp_pack", h, L, _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]),
                  C.cast(arr(bw), vp), C.cast(arr(bg), vp), C.cast(arr(bb), vp),
                  _ptr(heads[0]), _ptr(heads[1]), _ptr(heads[2]), _ptr(heads[3]), _ptr(packed), _stream())
    return packed


def forward(model, boards: torch.Tensor, packed: torch.Tensor | None = None, *, dropout_p: float | None = None,
            dropout_seed: int | None = None, dropout_sample0: int = 0):
    """(logits [n,4], value [n,1]) of GameMLP on packed boards through the fused kernel (forward only).
    dropout_p: None = what the model's mode implies (MLPConfig.dropout in train(), 0 in eval())."""
    h, L = _shape(model)
    boards = _req(boards.reshape(-1), torch.int64, "boards")
    n = boards.numel()
    dev = init(boards.device)
    packed = pack(model) if packed is None else packed
    with torch.cuda.device(dev):
        logits = torch.empty((n, 4), dtype=torch.float32, device=dev)
        value = torch.empty((n, 1), dtype=torch.float32, device=dev)
        dp_ = model_dropout_p(model) if dropout_p is None else float(dropout_p)
        u = _UpdateMlp(n=n, hidden=h, layers=L, decouple_critic=int(model.decouple_critic), backward=0,
                       boards=boards.data_ptr(), packed=packed.data_ptr(), workspace=_workspace(dev, h, L).data_ptr(),
                       logits=logits.data_ptr(), value=value.data_ptr(), dropout_p=dp_,
                       dropout_seed=(fresh_dropout_seed() if dropout_seed is None else dropout_seed) if dp_ > 0 else 0,
                       dropout_sample0=dropout_sample0)
        _lib.call("g2048_update_mlp_fwd_bwd", C.byref(u), _stream())
    return logits, value


def _acc(p: torch.nn.Parameter, g: torch.Tensor, alpha: float = 1.0) -> None:
    g = g.reshape(p.shape)
    if p.grad is None:
        p.grad = g * alpha
    else:
        p.grad.add_(g, alpha=alpha)


def loss_and_grads(model, boards, actions, legal, old_logp, adv, g_norm, *, flags=None, clip_eps=0.2,
                   critic_strength=1.0, entropy_strength=0.1, n_total: int | None = None,
                   packed: torch.Tensor | None = None, keep: dict | None = None,
                   logits_out: torch.Tensor | None = None, dropout_p: float | None = None, dropout_seed: int | None = None,
                   dropout_sample0: int = 0) -> torch.Tensor:
    """Adds the gradients of the minibatch-mean loss (train.py:554) over these samples into `p.grad` of every
    parameter and returns float64[4] = {sum ppo, sum smooth_l1, sum entropy, count} (device tensor).
    `n_total`: divisor of the mean when this call is a chunk / shard of a larger minibatch.
    `keep`: optional dict that receives the intermediate tensors (tests); `logits_out`: optional float32 [n, 4]
    CUDA tensor that receives the forward's action logits.
    dropout_p: None = what the model's mode implies (MLPConfig.dropout when model.training -- the reference updates in
    train() mode, train.py:483 -- else 0); dropout_seed: Philox key of the masks (None: drawn from torch's global
    generator); dropout_sample0: index of the first sample in the mask's counter space (chunks of one minibatch).
    `dropout_mask` restates the mask."""
    h, L = _shape(model)
    dp_ = model_dropout_p(model) if dropout_p is None else float(dropout_p)
    if dp_ > 0 and dropout_seed is None:
        dropout_seed = fresh_dropout_seed()
    boards = _req(boards.reshape(-1), torch.int64, "boards")
    n = boards.numel()
    dev = init(boards.device)
    packed = pack(model) if packed is None else packed
    old_logp = _req(old_logp, torch.float32, "old_logp")
    stride = 4 if old_logp.numel() == 4 * n else 1
    assert old_logp.numel() == stride * n
    n_div = int(n_total if n_total else max(n, 1))
    scale = loss_scale(n_div)
    with torch.cuda.device(dev):
        hp = padded_width(h)
        per_layer = (n + 127) // 128 * 128 * hp
        h_out = torch.empty((L + 1, per_layer), dtype=torch.float32, device=dev)    # fp16 hi|lo operand images (4 B / value), see untile()
        dz_out = torch.empty((L + 1, per_layer), dtype=torch.float32, device=dev)
        dhead = torch.empty((n, 8), dtype=torch.float32, device=dev)
        ln_grad = torch.empty((L + 1, 2, h), dtype=torch.float32, device=dev)
        hb_grad = torch.empty(5, dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float64, device=dev)
        u = _UpdateMlp(n=n, hidden=h, layers=L, decouple_critic=int(model.decouple_critic), backward=1,
                       boards=boards.data_ptr(), actions=_req(actions, torch.uint8, "actions").data_ptr(),
                       legal=_req(legal, torch.uint8, "legal").data_ptr(),
                       flags=None if flags is None else _req(flags, torch.uint8, "flags").data_ptr(),
                       old_logp=old_logp.data_ptr(), old_logp_stride=stride,
                       adv=_req(adv, torch.float32, "adv").data_ptr(), g_norm=_req(g_norm, torch.float32, "g_norm").data_ptr(),
                       clip_eps=clip_eps, critic_strength=critic_strength, entropy_strength=entropy_strength,
                       inv_n=scale / float(n_div),
                       packed=packed.data_ptr(), workspace=_workspace(dev, h, L).data_ptr(),
                       h_out=h_out.data_ptr(), dz_out=dz_out.data_ptr(), dhead=dhead.data_ptr(),
                       ln_grad=ln_grad.data_ptr(), head_bias_grad=hb_grad.data_ptr(), stats=stats.data_ptr(),
                       dropout_p=dp_, dropout_seed=(dropout_seed or 0) if dp_ > 0 else 0, dropout_sample0=dropout_sample0)
        if logits_out is not None:
            assert logits_out.shape == (n, 4) and logits_out.is_contiguous()
            u.logits = _req(logits_out, torch.float32, "logits_out").data_ptr()
        _lib.call("g2048_update_mlp_fwd_bwd", C.byref(u), _stream())
        if n > 0:
            # weight gradients: reductions over samples on the tensor cores (fp16 terms; every operand carries `scale`)
            inv = 1.0 / scale
            wg = lambda dy, x, nn, kk, **kw: linear.wgrad_tiled(dy, x, n, nn, kk, fp16=True, **kw)
            _acc(model.stem[0].weight, wg(dz_out[0], boards, h, 48, dy_hp=hp, x_hp=-1), inv)   # X = the packed boards
            for l, blk in enumerate(model.backbone):
                _acc(blk.mlp[0].weight, wg(dz_out[l + 1], h_out[l], h, h, dy_hp=hp, x_hp=hp), inv)
            dwh = wg(dhead, h_out[L], 8, h, x_hp=hp)
            _acc(model.action_head.weight, dwh[:4], inv)
            _acc(model.value_head.weight, dwh[4:5], inv)
            lns = [model.stem[1]] + [blk.mlp[1] for blk in model.backbone]
            for l, ln in enumerate(lns):
                _acc(ln.weight, ln_grad[l, 0], inv)
                _acc(ln.bias, ln_grad[l, 1], inv)
            _acc(model.action_head.bias, hb_grad[:4], inv)
            _acc(model.value_head.bias, hb_grad[4:5], inv)
    if keep is not None:
        inv = 1.0 / scale
        keep.update(h_out=torch.stack([untile(t, n, h) for t in h_out]), dz_out=torch.stack([untile(t, n, h) for t in dz_out]) * inv,
                    dhead=dhead * inv, ln_grad=ln_grad * inv, head_bias_grad=hb_grad * inv, loss_scale=scale)
    return stats


def untile(t: torch.Tensor, n: int, h: int, hp: int | None = None, dtype=torch.float16) -> torch.Tensor:
    """[n, h] fp32 copy of a tensor the fused kernel wrote as an fp16 hi|lo operand image: per tile of 128 samples
    [hi | lo][16-feature block][sample 0..127][32 B with the 16-byte halves swapped on (sample >> 2) & 1]
    (csrc/g2048_update_x3.cu storer).  Returns hi + lo (22 mantissa bits of the value).  hp: padded column count of the
    image (default: the update kernel's); dtype: the term format (the weight-gradient kernel also takes bf16 images)."""
    hp = padded_width(h) if hp is None else hp
    tiles = (n + 127) // 128
    x = t.contiguous().view(torch.uint8)[: tiles * 128 * hp * 4].view(dtype).view(tiles, 2, hp // 16, 128, 2, 8)
    swap = ((torch.arange(128, device=t.device) >> 2) & 1).view(1, 1, 1, 128, 1, 1).bool()
    x = torch.where(swap, x.flip(-2), x)
    v = x[:, 0].float() + x[:, 1].float()                            # [tiles, blocks, 128, 2, 8]
    return v.permute(0, 2, 1, 3, 4).reshape(tiles * 128, hp)[:n, :h].contiguous()
