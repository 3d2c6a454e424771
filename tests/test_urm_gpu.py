"""GameURM fused rollout kernels (BASELINE config #5) through the C ABI: the integer env path bit-exact against the oracle;
log-probs / values of the default kernel (split-fp16 operands, "x3") against the fp32 reference model -- the reference-generated
fixture tests/golden/model_urm.npz and the torch mirror -- at rtol 1e-5 / atol 2e-5 (the GameMLP bar); the single-fp16-operand
variant ("fp16", never chosen by default) against a torch emulation of its arithmetic and, loosely, against the fp32 model."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

RESET_TWEAK = 0x9E3779B97F4A7C15


def load_urm(golden):
    from g2048 import policy
    g = golden("model_urm")
    m = policy.GameURM(policy.GameURMConfig(dropout=0.0)).eval()
    m.load_state_dict(policy.load_state_dict_from_npz(g))
    return m.cuda(), g


def urm_fp16_emulation(model, x48):
    from g2048.policy import rms_norm
    r = lambda t: t.half().float()
    b = x48.shape[0]
    cfg = model.config
    emb = model.stem(x48.view(b, 16, 3))
    h = model.init_hidden.expand(b, -1, -1).clone()
    for _ in range(cfg.num_loops):
        h = h + emb
        for layer in model.layers:
            a = layer.attn
            qkv = (r(h) @ r(a.qkv_proj.weight).T).view(b, 16, 3, a.num_heads, a.head_dim).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], r(qkv[1]), r(qkv[2])
            p = torch.softmax((q @ k.transpose(-1, -2)) * 0.25, dim=-1)
            o = (p @ v).transpose(1, 2).reshape(b, 16, 64)
            h = rms_norm(h + r(o) @ r(a.o_proj.weight).T, 1e-5)
            m = layer.mlp
            gate, up = (r(h) @ r(m.gate_up_proj.weight).T).chunk(2, dim=-1)
            x = F.silu(gate) * up
            c = m.dwconv(x.transpose(1, 2))[..., :16]
            xc = F.silu(c).transpose(1, 2)
            h = rms_norm(h + r(xc) @ r(m.down_proj.weight).T, 1e-5)
    pooled = h.mean(1)
    return model.action_head(pooled), model.value_head(pooled).squeeze(1)


def masked_lp(logits, legal):
    illegal = ((legal.reshape(-1).long()[:, None] >> torch.arange(4, device=logits.device)) & 1) == 0
    return torch.masked_fill(logits, illegal, float("-inf")).log_softmax(-1)


ATOL_X3, RTOL_X3 = 2e-5, 1e-5      # fp32 grade: the bar of the GameMLP rollout kernels (tests/test_rollout_gpu.py)


@pytest.mark.parametrize("precision", ["x3", "fp16"])
def test_urm_rollout_first_step_matches_reference_fixture(golden, precision):
    from g2048 import rollout
    model, g = load_urm(golden)
    boards_np = g["board"]
    keep = np.array([O.potentials_batch(boards_np[i:i + 1])[0, 5] != 0 for i in range(len(boards_np))])
    boards_np = boards_np[keep]
    ref_logits, ref_v = torch.from_numpy(g["logits"][keep]).cuda(), torch.from_numpy(g["value"][keep]).cuda().squeeze(1)
    boards = torch.from_numpy(boards_np.view(np.int64).copy()).cuda()
    B = boards.numel()
    buf = rollout.rollout(rollout.pack_policy(model), boards.clone(), 1, seed=5, env0=0, ctr0=1, auto_reset=False,
                          alive=torch.ones(B, dtype=torch.uint8, device="cuda"), precision=precision)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(buf.boards[0].cpu().numpy().view(np.uint64), boards_np)
    want = masked_lp(ref_logits, buf.legal[0])
    got = buf.logp[0]
    fin = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got), fin)
    err_lp = float((got[fin] - want[fin]).abs().max())
    err_v = float((buf.value[0] - ref_v).abs().max())
    print(f"URM tensor-core rollout ({precision}) vs the reference's fp32 outputs: max |dlogp| = {err_lp:.2e}, max |dV| = {err_v:.2e}")
    if precision == "x3":                       # the reference's own fp32 outputs, at the GameMLP bar
        torch.testing.assert_close(got[fin], want[fin], rtol=RTOL_X3, atol=ATOL_X3)
        torch.testing.assert_close(buf.value[0], ref_v, rtol=RTOL_X3, atol=ATOL_X3)
        return
    assert err_lp < 0.02 and err_v < 0.02      # single fp16 operands; bf16 operands (round 1): 0.10 / 0.08
    from g2048 import env
    with torch.no_grad():
        el, ev = urm_fp16_emulation(model, env.encode(boards))
    emu = masked_lp(el, buf.legal[0])
    assert float((got[fin] - emu[fin]).abs().max()) < 2e-2 and float((buf.value[0] - ev).abs().max()) < 2e-2
    assert float((got[fin] - emu[fin]).abs().mean()) < 2e-3 and float((buf.value[0] - ev).abs().mean()) < 2e-3


@pytest.mark.parametrize("precision", ["x3", "fp16"])
@pytest.mark.parametrize("B,T,layers", [(100, 6, 2), (8, 3, 1), (1000, 2, 2), (2500, 3, 2), (20004, 2, 2)])
def test_urm_rollout_env_path_bit_exact(B, T, layers, precision):
    from g2048 import env, policy, rollout
    torch.manual_seed(B)
    model = policy.GameURM(policy.GameURMConfig(num_layers=layers, dropout=0.0)).cuda().eval()
    seed, env0 = 31, 77
    boards = env.reset(B, device=0, seed=seed, env0=env0, ctr=0)
    start = boards.clone()
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=seed, env0=env0, ctr0=1, auto_reset=True, precision=precision)
    torch.cuda.synchronize()
    b = start.cpu().numpy().view(np.uint64)
    for t in range(T):
        np.testing.assert_array_equal(buf.boards[t].cpu().numpy().view(np.uint64), b)
        nb, info = O.step_batch(b, buf.actions[t].cpu().numpy(), seed=seed, env0=env0, ctr=1 + t)
        assert (info["invalid"] == 0).all()
        np.testing.assert_array_equal(buf.legal[t].cpu().numpy(), info["legal_before"])
        np.testing.assert_array_equal(buf.points[t].cpu().numpy(), info["points"])
        np.testing.assert_array_equal(buf.flags[t].cpu().numpy(), 0x80 | info["legal_after"] | (info["done"] << 4))
        d = info["done"].astype(bool)
        if d.any():
            nb = np.where(d, O.reset_batch(B, seed=seed ^ RESET_TWEAK, env0=env0, ctr=1 + t), nb)
        b = nb
    np.testing.assert_array_equal(boards.cpu().numpy().view(np.uint64), b)      # the boards handed back = the next step's
    got = buf.logp.reshape(-1, 4)
    if precision == "x3":                       # every recorded log-prob / value against the torch fp32 model on the recorded boards
        with torch.no_grad():
            logits, v = model(env.encode(buf.boards.reshape(-1)))
        want = masked_lp(logits, buf.legal)
        fin = torch.isfinite(want)
        assert torch.equal(torch.isfinite(got), fin)
        torch.testing.assert_close(got[fin], want[fin], rtol=RTOL_X3, atol=ATOL_X3)
        torch.testing.assert_close(buf.value.reshape(-1), v.squeeze(1), rtol=RTOL_X3, atol=ATOL_X3)
        return
    with torch.no_grad():
        el, ev = urm_fp16_emulation(model, env.encode(buf.boards.reshape(-1)))
    emu = masked_lp(el, buf.legal)
    fin = torch.isfinite(emu)
    assert torch.equal(torch.isfinite(got), fin)
    # 8 block applications of fp16-operand GEMMs: a rounding flip can move a single output by ~1e-2
    # (outlier bound), while the bulk agrees to a few 1e-4 (mean bound)
    assert float((got[fin] - emu[fin]).abs().max()) < 2e-2 and float((buf.value.reshape(-1) - ev).abs().max()) < 2e-2
    assert float((got[fin] - emu[fin]).abs().mean()) < 2e-3 and float((buf.value.reshape(-1) - ev).abs().mean()) < 2e-3


def test_trainer_urm_train_steps_match_autograd_on_the_same_batch():
    """SURVEY 8(f) N4, host half: TrainConfig(model_type="urm") rolls out on the fused URM kernel and updates
    through torch autograd on the GameURM mirror (first loop under no_grad, game.py:1400-1413) with the fused
    PPO-loss kernel.  One update is replayed by hand on the recorded batch with the plain-torch loss restatement
    (tests/helpers.py) and must give the same clipped gradient; the weights must move and the loss stay finite."""
    from g2048 import env, trainer as tr
    from helpers import ref_ppo_loss_torch
    dev = torch.device("cuda:0")
    cfg = tr.TrainConfig(model_type="urm", envs=256, horizon=8, zero_heads=False, urm_chunk=512, warmup_steps=0)
    t = tr.Trainer(cfg, dev)
    assert not t.is_mlp and sum(p.numel() for p in t.model.parameters()) == 81237
    buf = t.collect()
    adv = t.advantages(buf)
    n = buf.flags.numel()
    # by hand: whole batch, plain torch
    ref = type(t.model)(cfg.urm).to(dev)
    ref.load_state_dict(t.model.state_dict())
    ref.train()
    logits, v = ref(env.encode(buf.boards.reshape(n)))
    valid = (buf.flags.reshape(n) & 0x80) != 0
    loss, _ = ref_ppo_loss_torch(logits[valid], v[valid], buf.logp.reshape(n, 4)[valid], buf.actions.reshape(n)[valid],
                                 buf.legal.reshape(n)[valid], adv["adv"].reshape(n)[valid], adv["g_norm"].reshape(n)[valid],
                                 clip_eps=cfg.clip_eps, critic_strength=cfg.critic_strength, entropy_strength=cfg.entropy_strength)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    before = [p.detach().clone() for p in t.model.parameters()]
    stats = t.update(buf, adv)                      # 4 chunks of 512 samples, gradient accumulated
    assert abs(stats["loss"] - float(loss.detach())) <= 1e-4 * max(1.0, abs(float(loss.detach())))
    for (name, p), q in zip(t.model.named_parameters(), ref.parameters()):
        if q.grad is None:                          # init_hidden only feeds the truncated (no_grad) loop
            assert name == "init_hidden" and (p.grad is None or not p.grad.any())
            continue
        assert p.grad is not None, name
        torch.testing.assert_close(p.grad, q.grad, rtol=2e-3, atol=2e-5, msg=name)
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, t.model.parameters()))
    s2 = t.train_step()
    assert all(np.isfinite(s2[k]) for k in ("loss", "policy_loss", "value_loss", "entropy", "grad_norm"))


def test_urm_precision_names():
    """GameURM has the fp32-grade kernel (auto / x3) and the labelled fp16 variant; there is no FFMA kernel to fall back to."""
    from g2048 import env, policy, rollout
    model = policy.GameURM(policy.GameURMConfig(dropout=0.0)).cuda().eval()
    pol = rollout.pack_policy(model)
    boards = env.reset(16, device=0, seed=1, env0=0, ctr=0)
    with pytest.raises(ValueError):
        rollout.rollout(pol, boards.clone(), 1, seed=1, precision="fp32")
    a = rollout.rollout(pol, boards.clone(), 2, seed=1, precision="auto")
    b = rollout.rollout(pol, boards.clone(), 2, seed=1, precision="x3")
    torch.cuda.synchronize()
    assert torch.equal(a.logp, b.logp) and torch.equal(a.boards, b.boards)
    mlp = rollout.pack_policy(policy.GameMLP(policy.MLPConfig(hidden_dim=64, num_layers=1, dropout=0.0)).cuda().eval())
    with pytest.raises(ValueError):
        rollout.rollout(mlp, boards.clone(), 1, seed=1, precision="fp16")


@pytest.mark.parametrize("scale", [4.0, 16.0, 40.0])
def test_urm_rollout_stays_finite_on_ill_conditioned_weights(scale):
    """Weights scaled until the fp32 reference itself is ill conditioned (torch fp32 and fp64 differ by 1e-3 at x4, 0.3 at x16) and
    activations leave the fp16 range of the split operands: the kernel must stay finite (saturating split, capped SiLU exponents) with
    the reference's pattern of legal actions, and at x4 stay within a small multiple of the reference's own fp32 noise."""
    from g2048 import env, policy, rollout
    torch.manual_seed(1)
    model = policy.GameURM(policy.GameURMConfig(dropout=0.0)).cuda().eval()
    with torch.no_grad():
        for layer in model.layers:
            layer.mlp.gate_up_proj.weight.mul_(scale)
            layer.mlp.dwconv.weight.mul_(scale)
            layer.attn.qkv_proj.weight.mul_(scale ** 0.5)
    boards = env.reset(2000, device=0, seed=3)
    buf = rollout.rollout(rollout.pack_policy(model), boards, 3, seed=3, precision="x3")
    torch.cuda.synchronize()
    with torch.no_grad():
        logits, v = model(env.encode(buf.boards.reshape(-1)))
    want = masked_lp(logits, buf.legal)
    got = buf.logp.reshape(-1, 4)
    assert not torch.isnan(got).any() and torch.isfinite(buf.value).all()
    assert torch.equal(torch.isfinite(got), torch.isfinite(want))
    if scale == 4.0:
        fin = torch.isfinite(want)
        assert float((got[fin] - want[fin]).abs().max()) < 2e-2 and float((buf.value.reshape(-1) - v.squeeze(1)).abs().max()) < 2e-2
