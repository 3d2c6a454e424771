"""Fused y = res + ReLU(LayerNorm(z)) kernels (csrc/g2048_update.cu) against plain PyTorch fp32
autograd of the same expression, and the fused GameMLP forward against the module."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,with_res", [(1, 196, True), (37, 196, True), (4096, 196, False), (100003, 196, True),
                                          (513, 64, True), (129, 256, False), (1000, 192, True), (64, 4, True)])
def test_ln_relu_res_forward_backward(n, h, with_res):
    from g2048 import fused
    g = torch.Generator(device="cuda").manual_seed(n + h)
    z = (torch.randn((n, h), generator=g, device="cuda") * 3 + 0.5).requires_grad_(True)
    gamma = (torch.randn(h, generator=g, device="cuda") * 0.5 + 1).requires_grad_(True)
    beta = (torch.randn(h, generator=g, device="cuda") * 0.3).requires_grad_(True)
    res = torch.randn((n, h), generator=g, device="cuda").requires_grad_(True) if with_res else None
    w = torch.randn((n, h), generator=g, device="cuda")
    y = fused.ln_relu_res(z, gamma, beta, res)
    (y * w).sum().backward()
    got = [y.detach(), z.grad, gamma.grad, beta.grad] + ([res.grad] if with_res else [])
    z2, g2, b2 = (t.detach().double().requires_grad_(True) for t in (z, gamma, beta))
    r2 = res.detach().double().requires_grad_(True) if with_res else None
    ref = F.relu(F.layer_norm(z2, (h,), g2, b2, 1e-5))
    if with_res:
        ref = r2 + ref
    (ref * w.double()).sum().backward()
    want = [ref.detach(), z2.grad, g2.grad, b2.grad] + ([r2.grad] if with_res else [])
    names = ["y", "dz", "dgamma", "dbeta", "dres"]
    for name, a, b in zip(names, got, want):
        scale = float(b.abs().max()) + 1e-12
        # elements whose pre-activation sits within rounding of zero may flip the ReLU gate
        bad = ((a.double() - b).abs() > 1e-4 * scale + 1e-5 * b.abs()).float().mean().item()
        assert bad < (2e-4 if name in ("y", "dz") else 1e-9), (name, bad)


def test_fused_mlp_forward_matches_module(golden):
    from g2048 import fused, policy
    g = golden("model_best")
    m = policy.GameMLP(policy.MLPConfig(hidden_dim=int(g["hidden_dim"]), num_layers=int(g["num_layers"]), dropout=0.0))
    m.load_state_dict(policy.load_state_dict_from_npz(g))
    m = m.cuda().train()
    x = torch.from_numpy(g["inputs"]).cuda()
    logits, v = fused.mlp_forward(m, x)
    np.testing.assert_allclose(logits.detach().cpu().numpy(), g["logits"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(v.detach().cpu().numpy(), g["value"], rtol=1e-5, atol=2e-5)
    # gradients against the unfused module
    t = torch.randn_like(logits)
    (logits * t).sum().backward(retain_graph=True)
    (v.sum() * 0.3).backward()
    fused_grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    l2, v2 = m(x)
    ((l2 * t).sum() + v2.sum() * 0.3).backward()
    for k, p in m.named_parameters():
        scale = float(p.grad.abs().max()) + 1e-12
        np.testing.assert_allclose(fused_grads[k].cpu().numpy(), p.grad.cpu().numpy(), rtol=1e-3, atol=1e-4 * scale, err_msg=k)
