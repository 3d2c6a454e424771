"""Update-path forward of GameMLP with the elementwise chain of every block fused into one
hand-written kernel pair (csrc/g2048_update.cu):

    x = res + ReLU(LayerNorm(z))        game.py:1038-1046 (ResidualBlock), 1069-1073 (stem, res = 0)

torch keeps the graph, the Linear GEMMs (cuBLAS) and the parameters; the fused op is a
torch.autograd.Function.  Dropout must be off (p = 0 or eval): the reference's p = 0.1 dropout is not
reproducible across implementations anyway -- with p > 0 use model(x) instead.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .env import _ptr, _stream, init

_lib.register("g2048_ln_relu_res_fwd", [C.c_void_p] * 7 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p])
_lib.register("g2048_ln_relu_res_bwd", [C.c_void_p] * 10 + [C.c_int64, C.c_int32, C.c_void_p])
_lib.lib().g2048_ln_workspace_floats.restype = C.c_int64
_lib.lib().g2048_ln_workspace_floats.argtypes = [C.c_int32]

_WS: dict[tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device, h: int) -> torch.Tensor:
    key = (dev.index, h)
    if key not in _WS:
        _WS[key] = torch.empty(int(_lib.lib().g2048_ln_workspace_floats(h)), dtype=torch.float32, device=dev)
    return _WS[key]


def supported(h: int) -> bool:
    return 4 <= h <= 256 and h % 4 == 0


class _LNReLURes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gamma, beta, res, eps):
        z = z.contiguous()
        n, h = z.shape
        dev = init(z.device)
        with torch.cuda.device(dev):
            y = torch.empty_like(z)
            mean = torch.empty(n, dtype=torch.float32, device=dev)
            rstd = torch.empty(n, dtype=torch.float32, device=dev)
            r = None if res is None else res.contiguous()
            _lib.call("g2048_ln_relu_res_fwd", _ptr(z), _ptr(gamma.contiguous()), _ptr(beta.contiguous()), _ptr(r),
                      _ptr(y), _ptr(mean), _ptr(rstd), n, h, float(eps), _stream())
        ctx.save_for_backward(z, gamma, beta, mean, rstd)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, gout):
        z, gamma, beta, mean, rstd = ctx.saved_tensors
        n, h = z.shape
        gout = gout.contiguous()
        dev = z.device
        with torch.cuda.device(dev):
            dz = torch.empty_like(z)
            dgamma = torch.empty_like(gamma)
            dbeta = torch.empty_like(beta)
            _lib.call("g2048_ln_relu_res_bwd", _ptr(z), _ptr(gamma.contiguous()), _ptr(beta.contiguous()), _ptr(mean),
                      _ptr(rstd), _ptr(gout), _ptr(dz), _ptr(dgamma), _ptr(dbeta), _ptr(_workspace(dev, h)), n, h,
                      _stream())
        return dz, dgamma, dbeta, (gout if ctx.has_res else None), None


def ln_relu_res(z, gamma, beta, res=None, eps: float = 1e-5):
    """res + relu(layer_norm(z, gamma, beta)) as one fused CUDA op (float32, [n,h], h % 4 == 0, h <= 256)."""
    return _LNReLURes.apply(z, gamma, beta, res, eps)


def mlp_forward(model, x48: torch.Tensor, matmul: str = "cublas"):
    """GameMLP.forward (game.py:1145-1220) with fused block epilogues; same parameters, same outputs.
    matmul = "x3": the stem / block Linears (forward, dgrad, wgrad) run on the split-bf16 tcgen05 kernels
    of g2048.linear instead of cuBLAS; the two tiny heads (N = 4, 1) stay in torch."""
    h = model.stem[0].weight.shape[0]
    if not supported(h) or any(blk.mlp[3].p > 0 and model.training for blk in model.backbone):
        if matmul == "x3":
            raise ValueError("mlp_forward(matmul='x3') needs dropout off and hidden % 4 == 0, hidden <= 208")
        return model(x48)
    if matmul == "x3":
        from . import linear as _lx
        if not _lx.supported(h, h):
            raise ValueError(f"mlp_forward(matmul='x3'): hidden {h} unsupported")
        lin = _lx.linear
    else:
        lin = torch.nn.functional.linear
    ln = model.stem[1]
    x = ln_relu_res(lin(x48, model.stem[0].weight), ln.weight, ln.bias, None, ln.eps)
    for blk in model.backbone:
        ln = blk.mlp[1]
        x = ln_relu_res(lin(x, blk.mlp[0].weight), ln.weight, ln.bias, x, ln.eps)
    logits = model.action_head(x)
    value = model.value_head(x.detach() if model.decouple_critic else x)
    return logits, value
