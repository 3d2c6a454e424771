"""GameURM update ops through the C ABI: every hand-written forward / backward against torch autograd on the
reference formulation (game.py:1223-1317) in float64, and the whole `urm_ops.forward` against the GameURM mirror's own forward
(outputs and every parameter gradient)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("B", [1, 7, 300])
def test_attention_forward_backward_match_float64_autograd(B):
    from g2048 import urm_ops
    torch.manual_seed(B)
    qkv = torch.randn(B, 16, 192, device="cuda") * 1.5
    dout = torch.randn(B, 16, 64, device="cuda")
    x = qkv.clone().requires_grad_(True)
    out = urm_ops.attention(x)
    out.backward(dout)
    ref_in = qkv.double().requires_grad_(True)
    t = ref_in.view(B, 16, 3, 4, 16).permute(2, 0, 3, 1, 4)                    # game.py:1306-1307
    ref = F.scaled_dot_product_attention(t[0], t[1], t[2]).transpose(1, 2).reshape(B, 16, 64)
    ref.backward(dout.double())
    assert _rel(out, ref) < 2e-6
    assert _rel(x.grad, ref_in.grad) < 2e-6


@pytest.mark.parametrize("B", [1, 5, 1000])
def test_conv_swiglu_forward_backward_match_float64_autograd(B):
    from g2048 import urm_ops
    torch.manual_seed(B + 1)
    conv = torch.nn.Conv1d(120, 120, kernel_size=2, padding=1, groups=120, bias=True).cuda()      # game.py:1253-1260
    gate = torch.randn(B, 16, 120, device="cuda") * 2
    up = torch.randn(B, 16, 120, device="cuda")
    dy = torch.randn(B, 16, 120, device="cuda")
    g, u = gate.clone().requires_grad_(True), up.clone().requires_grad_(True)
    w, b = conv.weight.detach().clone().requires_grad_(True), conv.bias.detach().clone().requires_grad_(True)
    y = urm_ops.conv_swiglu(g, u, w, b)
    y.backward(dy)
    conv64 = torch.nn.Conv1d(120, 120, kernel_size=2, padding=1, groups=120, bias=True).cuda().double()
    conv64.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
    g64, u64 = gate.double().requires_grad_(True), up.double().requires_grad_(True)
    h = F.silu(g64) * u64                                                                          # game.py:1266-1276
    ref = F.silu(conv64(h.transpose(1, 2))[..., :16]).transpose(1, 2)
    ref.backward(dy.double())
    assert _rel(y, ref) < 2e-6
    assert _rel(g.grad, g64.grad) < 5e-6 and _rel(u.grad, u64.grad) < 5e-6
    assert _rel(w.grad, conv64.weight.grad) < 5e-6 and _rel(b.grad, conv64.bias.grad) < 5e-6


@pytest.mark.parametrize("rows", [1, 33, 4800])
def test_residual_norm_forward_backward_match_float64_autograd(rows):
    from g2048 import urm_ops
    from g2048.policy import rms_norm
    torch.manual_seed(rows)
    x, r, dy = (torch.randn(rows, 64, device="cuda") for _ in range(3))
    a, b = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    y = urm_ops.residual_norm(a, b, 1e-5)
    y.backward(dy)
    a64, b64 = x.double().requires_grad_(True), r.double().requires_grad_(True)
    ref = rms_norm(a64 + b64, 1e-5)
    ref.backward(dy.double())
    assert _rel(y, ref) < 2e-6
    assert _rel(a.grad, a64.grad) < 5e-6 and torch.equal(a.grad, b.grad)


def test_urm_ops_forward_and_gradients_match_the_model_mirror():
    """The whole update forward on the kernels against GameURM.forward in train mode (first loop under no_grad): outputs and the
    gradient of every parameter.  The projections run on split-bf16 tensor-core GEMMs (16 mantissa bits), hence 2e-4 / 2e-3."""
    from g2048 import env, policy, urm_ops
    torch.manual_seed(3)
    model = policy.GameURM(policy.GameURMConfig(dropout=0.0)).cuda().train()
    boards = env.reset(500, device=0, seed=9, env0=0, ctr=0)
    x = env.encode(boards)
    wl, wv = torch.randn(500, 4, device="cuda"), torch.randn(500, 1, device="cuda")
    logits, v = urm_ops.forward(model, x)
    ((logits * wl).sum() + (v * wv).sum()).backward()
    got = {k: p.grad.clone() if p.grad is not None else None for k, p in model.named_parameters()}
    model.zero_grad()
    ref_l, ref_v = model(x)
    ((ref_l * wl).sum() + (ref_v * wv).sum()).backward()
    assert _rel(logits, ref_l) < 2e-4 and _rel(v, ref_v) < 2e-4
    for k, p in model.named_parameters():
        if p.grad is None:
            assert got[k] is None or not got[k].any(), k          # init_hidden only feeds the truncated loop
            continue
        assert got[k] is not None, k
        assert _rel(got[k], p.grad) < 2e-3, (k, _rel(got[k], p.grad))
