"""GameURM update on this library's kernels, forward and hand-written backward.

`forward(model, x48)` is `GameURM.forward` (game.py:1418-1458; train-mode semantics: the first `num_truncated_loops` loops under
`no_grad`, game.py:1400-1413) with every block op behind the C ABI:

    projections  qkv / o / gate / up / down    g2048_x3_gemm + g2048_x3_wgrad      (linear.linear; tcgen05, split operands)
    attention    16 tokens, 4 heads            g2048_urm_attn_fwd / _bwd           (game.py:1296-1317)
    ConvSwiGLU   silu(g) * u -> conv -> silu   g2048_urm_swiglu_fwd / _bwd         (game.py:1264-1276)
    norms        rms_norm(hidden + branch)     g2048_urm_norm_fwd / _bwd           (game.py:1223-1229, 1345-1350)

torch autograd is the tape that strings the ops together (and runs the 3 -> 64 stem and the two heads, 0.2 % of the FLOPs).
Default GameURMConfig shapes only; dropout must be 0 (the reference's CLI refuses GameURM training altogether, train.py:1523-1532).
There is no fallback: unsupported configurations raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .env import _ptr, _stream, init
from .linear import linear

_V = C.c_void_p
_lib.register("g2048_urm_attn_fwd", [_V, _V, C.c_int64, _V])
_lib.register("g2048_urm_attn_bwd", [_V, _V, _V, C.c_int64, _V])
_lib.register("g2048_urm_swiglu_fwd", [_V, _V, _V, _V, _V, C.c_int64, _V])
_lib.register("g2048_urm_swiglu_bwd", [_V] * 10 + [C.c_int64, _V])
_lib.register("g2048_urm_norm_fwd", [_V, _V, _V, _V, C.c_int64, C.c_float, _V])
_lib.register("g2048_urm_norm_bwd", [_V, _V, _V, _V, C.c_int64, _V])
_lib.lib().g2048_urm_swiglu_workspace_floats.restype = C.c_int64
_lib.lib().g2048_urm_swiglu_workspace_floats.argtypes = []

SEQ, HIDDEN, HEADS, INTER = 16, 64, 4, 120
_WS: dict[int, torch.Tensor] = {}


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float32):
        raise ValueError(f"{name} must be a float32 CUDA tensor (got {t.dtype}, {t.device})")
    return t.contiguous()


class _Attention(torch.autograd.Function):
    """qkv [B,16,192] -> attention output [B,16,64] (softmax(q k^T / 4) v per head, heads concatenated)."""

    @staticmethod
    def forward(ctx, qkv):
        qkv = _f32(qkv, "qkv")
        b = qkv.shape[0]
        assert qkv.shape[1:] == (SEQ, 3 * HIDDEN)
        ctx.save_for_backward(qkv)
        dev = init(qkv.device)
        with torch.cuda.device(dev):
            out = torch.empty((b, SEQ, HIDDEN), dtype=torch.float32, device=dev)
            _lib.call("g2048_urm_attn_fwd", _ptr(qkv), _ptr(out), b, _stream())
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        dout = _f32(dout, "grad_output")
        dev = init(qkv.device)
        with torch.cuda.device(dev):
            dqkv = torch.empty_like(qkv)
            _lib.call("g2048_urm_attn_bwd", _ptr(qkv), _ptr(dout), _ptr(dqkv), qkv.shape[0], _stream())
        return dqkv


class _ConvSwiGLU(torch.autograd.Function):
    """gate, up [B,16,120], dwconv.weight [120,1,2], dwconv.bias [120] -> silu(conv(silu(gate) * up)) [B,16,120]."""

    @staticmethod
    def forward(ctx, gate, up, conv_w, conv_b):
        gate, up, conv_w, conv_b = _f32(gate, "gate"), _f32(up, "up"), _f32(conv_w, "dwconv.weight"), _f32(conv_b, "dwconv.bias")
        b = gate.shape[0]
        assert gate.shape[1:] == (SEQ, INTER) and up.shape == gate.shape and conv_w.numel() == 2 * INTER and conv_b.numel() == INTER
        ctx.save_for_backward(gate, up, conv_w, conv_b)
        dev = init(gate.device)
        with torch.cuda.device(dev):
            y = torch.empty_like(gate)
            _lib.call("g2048_urm_swiglu_fwd", _ptr(gate), _ptr(up), _ptr(conv_w), _ptr(conv_b), _ptr(y), b, _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        gate, up, conv_w, conv_b = ctx.saved_tensors
        dy = _f32(dy, "grad_output")
        dev = init(gate.device)
        with torch.cuda.device(dev):
            if dev.index not in _WS:
                _WS[dev.index] = torch.empty(int(_lib.lib().g2048_urm_swiglu_workspace_floats()), dtype=torch.float32, device=dev)
            dgate, dup = torch.empty_like(gate), torch.empty_like(up)
            dw, db = torch.empty_like(conv_w), torch.empty_like(conv_b)
            _lib.call("g2048_urm_swiglu_bwd", _ptr(gate), _ptr(up), _ptr(conv_w), _ptr(conv_b), _ptr(dy), _ptr(dgate), _ptr(dup),
                      _ptr(dw), _ptr(db), _ptr(_WS[dev.index]), gate.shape[0], _stream())
        return dgate, dup, dw, db


class _ResidualNorm(torch.autograd.Function):
    """rms_norm(x + r) over rows of 64; the same gradient flows to x and r."""

    @staticmethod
    def forward(ctx, x, r, eps):
        x, r = _f32(x, "x"), _f32(r, "r")
        assert x.shape == r.shape and x.shape[-1] == HIDDEN
        rows = x.numel() // HIDDEN
        dev = init(x.device)
        with torch.cuda.device(dev):
            y = torch.empty_like(x)
            rs = torch.empty(rows, dtype=torch.float32, device=dev)
            _lib.call("g2048_urm_norm_fwd", _ptr(x), _ptr(r), _ptr(y), _ptr(rs), rows, float(eps), _stream())
        ctx.save_for_backward(y, rs)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, rs = ctx.saved_tensors
        dy = _f32(dy, "grad_output")
        dev = init(y.device)
        with torch.cuda.device(dev):
            ds = torch.empty_like(y)
            _lib.call("g2048_urm_norm_bwd", _ptr(y), _ptr(rs), _ptr(dy), _ptr(ds), rs.numel(), _stream())
        return ds, ds, None


attention = _Attention.apply
conv_swiglu = _ConvSwiGLU.apply
residual_norm = _ResidualNorm.apply


def supported(model) -> bool:
    cfg = getattr(model, "config", None)
    return (cfg is not None and hasattr(model, "init_hidden") and cfg.hidden_dim == HIDDEN and cfg.num_heads == HEADS
            and model.layers[0].mlp.down_proj.weight.shape[1] == INTER and cfg.conv_kernel == 2
            and (cfg.dropout == 0.0 or not model.training))


def block(layer, h: torch.Tensor, b: int) -> torch.Tensor:
    """One GameURMBlock (game.py:1320-1352) on [b * 16, 64] hidden rows."""
    eps = layer.norm_eps
    qkv = linear(h, layer.attn.qkv_proj.weight)
    o = attention(qkv.view(b, SEQ, 3 * HIDDEN))
    h = residual_norm(h, linear(o.view(b * SEQ, HIDDEN), layer.attn.o_proj.weight), eps)
    w = layer.mlp.gate_up_proj.weight
    gate, up = linear(h, w[:INTER]), linear(h, w[INTER:])              # chunk(2, -1) of the fused projection (game.py:1265)
    x = conv_swiglu(gate.view(b, SEQ, INTER), up.view(b, SEQ, INTER), layer.mlp.dwconv.weight, layer.mlp.dwconv.bias)
    return residual_norm(h, linear(x.view(b * SEQ, INTER), layer.mlp.down_proj.weight), eps)


def forward(model, inputs: torch.Tensor):
    """GameURM.forward(inputs [B,48]) -> (logits [B,4], value [B,1]) on the kernels above."""
    if not supported(model):
        raise ValueError("urm_ops.forward: default GameURMConfig shapes (hidden 64, 4 heads, inter 120, conv kernel 2) and dropout 0 only")
    if inputs.ndim == 1:
        inputs = inputs.unsqueeze(0)
    b = inputs.shape[0]
    emb = model.stem(inputs.view(b, SEQ, 3)).reshape(b * SEQ, HIDDEN)
    h = model.init_hidden.expand(b, -1, -1).reshape(b * SEQ, HIDDEN)
    trunc = model.config.num_truncated_loops

    def loop(h):
        h = h + emb
        for layer in model.layers:
            h = block(layer, h, b)
        return h

    if trunc > 0:
        with torch.no_grad():
            for _ in range(trunc):
                h = loop(h)
    for _ in range(model.config.num_loops - trunc):
        h = loop(h)
    pooled = h.view(b, SEQ, HIDDEN).mean(dim=1)
    return model.action_head(pooled), model.value_head(pooled)
