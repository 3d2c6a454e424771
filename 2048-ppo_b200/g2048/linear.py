"""fp32-grade Linear layers of the policy update on tcgen05 (csrc/g2048_linear.cu).

`linear(x, weight)` is `torch.nn.functional.linear(x, weight)` (no bias: GameMLP's stem and block
Linears have none, game.py:1069, 1039) for 2-D fp32 CUDA tensors, forward and backward, with every
GEMM running as split-bf16 ("x3") tensor-core products behind the C ABI:

    forward   y  = x  W^T      g2048_x3_gemm  (image of W)
    dgrad     dx = dy W        g2048_x3_gemm  (image of W^T)
    wgrad     dW = dy^T x      g2048_x3_wgrad

There is no fallback: unsupported shapes raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .env import _ptr, _stream, init

_lib.register("g2048_x3_pack", [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p])
_lib.register("g2048_x3_gemm", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p])
_lib.register("g2048_x3_wgrad", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p])
_lib.register("g2048_x3_wgrad_tiled", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p])
_lib.register("g2048_x3_wgrad_images", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_void_p])
_lib.lib().g2048_x3_image_bytes.restype = C.c_int64
_lib.lib().g2048_x3_image_bytes.argtypes = [C.c_int32, C.c_int32]
_lib.lib().g2048_x3_wgrad_workspace_bytes.restype = C.c_int64
_lib.lib().g2048_x3_wgrad_workspace_bytes.argtypes = []

MAX_FEATURES = 208
_WS: dict[int, torch.Tensor] = {}


def supported(n_out: int, n_in: int) -> bool:
    return all(4 <= f <= MAX_FEATURES and f % 4 == 0 for f in (n_out, n_in))


def _check(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
        raise ValueError(f"{name} must be a 2-D float32 CUDA tensor (got {t.dtype}, {t.device}, {t.dim()}-D)")
    return t.contiguous()


def pack_weight(weight: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """Operand image of `weight` [out, in] (or of its transpose) for `gemm`."""
    w = _check(weight.detach(), "weight")
    r, c = w.shape
    rows, cols = (c, r) if transpose else (r, c)
    dev = init(w.device)
    with torch.cuda.device(dev):
        img = torch.empty(int(_lib.lib().g2048_x3_image_bytes(rows, cols)), dtype=torch.uint8, device=dev)
        _lib.call("g2048_x3_pack", _ptr(w), r, c, int(transpose), _ptr(img), _stream())
    return img


def gemm(a: torch.Tensor, image: torch.Tensor, n: int) -> torch.Tensor:
    """a [M, K] times the packed [n, K] operand, transposed: returns [M, n]."""
    a = _check(a, "a")
    m, k = a.shape
    dev = init(a.device)
    with torch.cuda.device(dev):
        out = torch.empty((m, n), dtype=torch.float32, device=dev)
        _lib.call("g2048_x3_gemm", _ptr(a), _ptr(image), _ptr(out), m, n, k, _stream())
    return out


def wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dy [M, N], x [M, K] -> dy^T x [N, K]."""
    dy, x = _check(dy, "dy"), _check(x, "x")
    m, n = dy.shape
    k = x.shape[1]
    if x.shape[0] != m:
        raise ValueError("dy and x disagree on the sample count")
    dev = init(dy.device)
    with torch.cuda.device(dev):
        if dev.index not in _WS:
            _WS[dev.index] = torch.empty(int(_lib.lib().g2048_x3_wgrad_workspace_bytes()), dtype=torch.uint8, device=dev)
        out = torch.empty((n, k), dtype=torch.float32, device=dev)
        _lib.call("g2048_x3_wgrad", _ptr(dy), _ptr(x), _ptr(out), _ptr(_WS[dev.index]), m, n, k, _stream())
    return out


def wgrad_tiled(dy: torch.Tensor, x: torch.Tensor, m: int, n: int, k: int, dy_hp: int = 0, x_hp: int = 0, fp16: bool = False) -> torch.Tensor:
    """dy^T x over m samples where either operand may be a hi|lo operand image written by the fused update kernel (fp16 = True:
    the terms are fp16 -- the fused update's images and loss-scaled gradients -- else bf16)
    (`*_hp` = its padded column count; the kernel bulk-copies it straight into its operand ring), 0 = row-major fp32
    [m, features] (loaded, split and stored by the loader warps); x_hp = -1: `x` is int64 packed boards [m] and k == 48 --
    the model input [exponent, row/3, col/3] per cell is formed in the loader (no g2048_encode pass)."""
    dev = init(dy.device)
    with torch.cuda.device(dev):
        if dev.index not in _WS:
            _WS[dev.index] = torch.empty(int(_lib.lib().g2048_x3_wgrad_workspace_bytes()), dtype=torch.uint8, device=dev)
        out = torch.empty((n, k), dtype=torch.float32, device=dev)
        _lib.call("g2048_x3_wgrad_images", _ptr(dy), _ptr(x), _ptr(out), _ptr(_WS[dev.index]), m, n, k, dy_hp, x_hp, int(fp16), _stream())
    return out


class _LinearX3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        n, k = weight.shape
        if x.shape[1] != k or not supported(n, k):
            raise ValueError(f"linear: unsupported shapes x {tuple(x.shape)}, weight {tuple(weight.shape)}")
        x = _check(x, "x")
        ctx.save_for_backward(x, weight)
        return gemm(x, pack_weight(weight), n)

    @staticmethod
    def backward(ctx, gout):
        x, weight = ctx.saved_tensors
        gout = _check(gout, "grad_output")
        dx = gemm(gout, pack_weight(weight, transpose=True), weight.shape[1]) if ctx.needs_input_grad[0] else None
        dw = wgrad(gout, x) if ctx.needs_input_grad[1] else None
        return dx, dw


def linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """x @ weight.T for fp32 CUDA tensors via the split-bf16 tcgen05 kernels (differentiable)."""
    return _LinearX3.apply(x, weight)
