"""GPU parity of the rewards-to-go/advantage scan and the fused PPO loss (through the C ABI):
against the reference-generated fixtures, the float64 oracle and a plain-torch fp32 restatement.
Tolerance (north star): 1e-5 relative for losses and advantages."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import ref_ppo_loss_torch, rollout_as_tb, shaping_words  # noqa: E402
from oracle import oracle as O  # noqa: E402

RTOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_rtg_advantage_matches_reference_fixture(golden):
    from g2048 import ppo
    g, adv = golden("rollout"), golden("advantage")
    a, (t, env) = rollout_as_tb(g)
    for name in ("readme", "warm"):
        gamma, wp, wm, we, beta, step, mu, m2 = adv[name + "__cfg"].tolist()
        mom = ppo.RtgMoments(mu=mu, m2=m2, step=int(step))
        mu_c, sd = mom.corrected(beta)
        r = ppo.rtg_advantage(cu(a["points"]), cu(a["shaping"]), cu(a["flags"]), cu(a["value"]), gamma=gamma,
                              w_points=wp, w_mono=wm, w_empt=we, mu_c=mu_c, stddev=sd, want_raw=True)
        for key, ref in (("reward", "reward"), ("g_raw", "g_raw"), ("g_norm", "g_norm"), ("adv", "adv")):
            got = r[key].cpu().numpy()[t, env]
            np.testing.assert_allclose(got, adv[name + "__" + ref].astype(np.float32), rtol=RTOL, atol=1e-6,
                                       err_msg=name + key)
        s1, s2, n = r["stats"].tolist()
        assert n == len(t)
        mom.update(beta, s1, s2, n)
        np.testing.assert_allclose([mom.mu, mom.m2], adv[name + "__moments_out"], rtol=1e-9)
        # invalid slots (games shorter than T) are zero-filled
        inv = a["flags"] == 0
        assert (r["adv"].cpu().numpy()[inv] == 0).all()


@pytest.mark.parametrize("T,B", [(1, 1), (7, 33), (512, 4096), (64, 70001)])
def test_rtg_advantage_matches_oracle_random(T, B):
    from g2048 import ppo
    rng = np.random.default_rng(T * 1000 + B)
    points = (rng.integers(0, 64, (T, B)) * 4 * (rng.random((T, B)) < 0.4)).astype(np.int32)
    mono_b, mono_a = rng.integers(0, 49, (T, B)), rng.integers(0, 49, (T, B))
    empt_b, empt_a = rng.integers(0, 17, (T, B)), rng.integers(0, 17, (T, B))
    done = (rng.random((T, B)) < 0.02).astype(np.uint8)
    valid = np.ones((T, B), dtype=np.uint8)
    # some columns stop early (no auto-reset): everything after their first done is invalid
    for b in range(0, B, 5):
        d = np.nonzero(done[:, b])[0]
        if len(d):
            valid[d[0] + 1:, b] = 0
    value = rng.normal(size=(T, B)).astype(np.float32)
    flags = (valid << 7) | (done << 4)
    want = O.rtg_adv(points, mono_b, mono_a, empt_b, empt_a, done, valid, value, 0.99, 0.1, 1.0, 0.25, 0.99, 3, 2.5, 400.0)
    mom = ppo.RtgMoments(mu=2.5, m2=400.0, step=3)
    mu_c, sd = mom.corrected(0.99)
    r = ppo.rtg_advantage(cu(points), cu(shaping_words(mono_b, mono_a, empt_b, empt_a)), cu(flags.astype(np.uint8)),
                          cu(value), gamma=0.99, w_points=0.1, w_mono=1.0, w_empt=0.25, mu_c=mu_c, stddev=sd,
                          want_raw=True)
    for k in ("reward", "g_raw", "g_norm", "adv"):
        np.testing.assert_allclose(r[k].cpu().numpy(), want[k], rtol=RTOL, atol=1e-6, err_msg=k)
    s1, s2, n = r["stats"].tolist()
    assert n == want["stats"]["n"]
    np.testing.assert_allclose([s1, s2], [want["stats"]["sum"], want["stats"]["sumsq"]], rtol=1e-12)
    mom.update(0.99, s1, s2, n)
    np.testing.assert_allclose([mom.mu, mom.m2], [want["rtg_mu"], want["rtg_m2"]], rtol=1e-9)


def test_rtg_bootstrap_continues_games_past_the_buffer():
    """bootstrap[b] seeds the scan of column b (a game that continues past the buffer); a DONE move or an invalid slot at the
    end of the column cuts it off, and NULL is the reference's truncation with 0."""
    from g2048 import ppo
    rng = np.random.default_rng(5)
    T, B, gamma = 37, 301, 0.97
    points = (rng.integers(0, 64, (T, B)) * 4).astype(np.int32)
    mono_b, mono_a = rng.integers(0, 49, (T, B)), rng.integers(0, 49, (T, B))
    empt_b, empt_a = rng.integers(0, 17, (T, B)), rng.integers(0, 17, (T, B))
    done = (rng.random((T, B)) < 0.03).astype(np.uint8)
    valid = np.ones((T, B), dtype=np.uint8)
    valid[-3:, ::7] = 0                                   # some columns end with invalid slots
    done[-1, 1::7] = 1                                    # some end exactly on a DONE move
    flags = (valid << 7) | (done << 4)
    boot = rng.normal(size=B).astype(np.float32) * 50
    value = rng.normal(size=(T, B)).astype(np.float32)
    kw = dict(gamma=gamma, w_points=0.1, w_mono=1.0, w_empt=0.5, mu_c=3.0, stddev=20.0, want_raw=True)
    args = (cu(points), cu(shaping_words(mono_b, mono_a, empt_b, empt_a)), cu(flags.astype(np.uint8)), cu(value))
    got = ppo.rtg_advantage(*args, bootstrap=cu(boot), **kw)
    base = ppo.rtg_advantage(*args, **kw)
    # float64 restatement of the scan with a seeded return
    G = np.zeros((T, B))
    g = boot.astype(np.float64).copy()
    for t in range(T - 1, -1, -1):
        d, v = done[t].astype(bool), valid[t].astype(bool)
        g = np.where(d | ~v, 0.0, g)
        r = points[t] * 0.1 + (gamma * np.where(d, 0, mono_a[t]) - mono_b[t]) + 0.5 * (gamma * np.where(d, 0, empt_a[t]) - empt_b[t])
        g = np.where(v, r + gamma * g, 0.0)
        G[t] = g
    np.testing.assert_allclose(got["g_raw"].cpu().numpy(), G, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(got["g_norm"].cpu().numpy(), np.where(valid, (G - 3.0) / (20.0 + 1e-8), 0.0), rtol=1e-5, atol=1e-5)
    # the bootstrap only reaches the slots after the last DONE / invalid slot of its column
    diff = (got["g_raw"] - base["g_raw"]).cpu().numpy()
    cut = np.zeros(B, dtype=bool)
    for t in range(T - 1, -1, -1):
        cut |= done[t].astype(bool) | ~valid[t].astype(bool)
        want = np.where(cut, 0.0, boot * gamma ** (T - t))
        np.testing.assert_allclose(diff[t], want, rtol=1e-4, atol=1e-3)


def test_rtg_linearity_property_full_size():
    """C3-sized scan (512 x 65536): returns are linear in the reward weights."""
    from g2048 import ppo
    T, B = 512, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    points = (torch.randint(0, 64, (T, B), generator=g, device="cuda", dtype=torch.int32) * 4)
    shaping = torch.randint(0, 1 << 22, (T, B), generator=g, device="cuda", dtype=torch.int64)
    flags = (0x80 | ((torch.rand((T, B), generator=g, device="cuda") < 0.01).to(torch.uint8) << 4)).to(torch.uint8)
    value = torch.zeros((T, B), device="cuda")
    kw = dict(gamma=0.99, mu_c=0.0, stddev=1.0 - 1e-8, want_raw=True)
    a = ppo.rtg_advantage(points, shaping, flags, value, w_points=1.0, w_mono=0.0, w_empt=0.0, **kw)["g_raw"]
    b = ppo.rtg_advantage(points, shaping, flags, value, w_points=0.0, w_mono=1.0, w_empt=0.0, **kw)["g_raw"]
    c = ppo.rtg_advantage(points, shaping, flags, value, w_points=0.5, w_mono=2.0, w_empt=0.0, **kw)["g_raw"]
    torch.testing.assert_close(c, 0.5 * a + 2.0 * b, rtol=1e-5, atol=1e-3)
    # a done move cuts the return: G_t == reward_t there
    r = ppo.rtg_advantage(points, shaping, flags, value, w_points=1.0, w_mono=0.0, w_empt=0.0, **kw)
    d = (flags & 0x10) != 0
    torch.testing.assert_close(r["g_raw"][d], r["reward"][d])


def _random_loss_inputs(n, seed, extreme=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((n, 4), generator=g) * (12.0 if extreme else 2.0)
    value = torch.randn(n, generator=g) * (3.0 if extreme else 1.0)
    legal = torch.randint(1, 16, (n,), generator=g, dtype=torch.uint8)
    # action = a random legal direction
    bits = ((legal.long()[:, None] >> torch.arange(4)) & 1).float()
    actions = torch.multinomial(bits, 1, generator=g).squeeze(1).to(torch.uint8)
    old = torch.masked_fill(logits + torch.randn((n, 4), generator=g) * (1.5 if extreme else 0.2), bits == 0,
                            float("-inf")).log_softmax(-1)
    adv = torch.randn(n, generator=g) * 2
    g_norm = value + torch.randn(n, generator=g) * (2.0 if extreme else 0.5)
    return logits, value, old, actions, legal, adv, g_norm


@pytest.mark.parametrize("n,extreme", [(1, False), (257, False), (100003, False), (100003, True)])
def test_ppo_loss_matches_torch_restatement(n, extreme):
    from g2048 import ppo
    logits, value, old, actions, legal, adv, g_norm = _random_loss_inputs(n, n, extreme)
    lg, vg = logits.cuda().requires_grad_(True), value.cuda().requires_grad_(True)
    loss, stats = ppo.ppo_loss(lg, vg, old.cuda(), actions.cuda(), legal.cuda(), adv.cuda(), g_norm.cuda(),
                               clip_eps=0.2, critic_strength=0.2, entropy_strength=0.02)
    (loss * 1.7).backward()
    lr, vr = logits.double().requires_grad_(True), value.double().requires_grad_(True)
    ref, parts = ref_ppo_loss_torch(lr, vr, old.double(), actions, legal, adv.double(), g_norm.double(), 0.2, 0.2, 0.02)
    (ref * 1.7).backward()
    np.testing.assert_allclose(float(loss), float(ref), rtol=RTOL, atol=1e-7)
    s = stats.cpu().numpy()
    np.testing.assert_allclose(s[:3] / n, [float(parts["ppo"]), float(parts["vl"]), float(parts["ent"])], rtol=RTOL, atol=1e-7)
    assert s[3] == n
    scale = float(lr.grad.abs().max())
    np.testing.assert_allclose(lg.grad.cpu().numpy(), lr.grad.float().numpy(), rtol=1e-4, atol=max(1e-5 * scale, 1e-9))
    np.testing.assert_allclose(vg.grad.cpu().numpy(), vr.grad.float().numpy(), rtol=1e-4, atol=1e-7)


def test_ppo_loss_stride1_flags_and_global_n():
    from g2048 import ppo
    n = 5000
    logits, value, old, actions, legal, adv, g_norm = _random_loss_inputs(n, 3)
    chosen = torch.gather(old, 1, actions.long()[:, None]).squeeze(1)
    flags = torch.full((n,), 0x80, dtype=torch.uint8)
    flags[::3] = 0
    keep = flags != 0
    a = ppo.ppo_loss(logits.cuda(), value.cuda(), chosen.cuda(), actions.cuda(), legal.cuda(), adv.cuda(),
                     g_norm.cuda(), flags=flags.cuda(), n_total=int(keep.sum()))
    b = ppo.ppo_loss(logits[keep].cuda(), value[keep].cuda(), old[keep].cuda(), actions[keep].cuda(),
                     legal[keep].cuda(), adv[keep].cuda(), g_norm[keep].cuda())
    np.testing.assert_allclose(float(a[0]), float(b[0]), rtol=1e-6)
    assert a[1][3].item() == int(keep.sum())


def test_ppo_loss_and_grads_match_reference_optimize_step(golden):
    """End to end against train.model_optimize_step's recorded stats and clipped gradients:
    policy mirror forward (torch) -> fused loss kernel -> torch backward -> clip_grad_norm_."""
    from g2048 import env, ppo
    from test_train_cpu import policy_b
    gr, ga, gl = golden("rollout"), golden("advantage"), golden("loss")
    m = policy_b(golden).cuda()
    m.train()
    x = env.encode(cu(gr["board"].view(np.int64)))
    adv = cu(ga["readme__adv"].astype(np.float32))
    gn = cu(ga["readme__g_norm"].astype(np.float32))
    for name in ("readme", "alt"):
        ent, crit = gl[name + "__coef"].tolist()
        m.zero_grad()
        logits, v = m(x)
        loss, stats = ppo.ppo_loss(logits, v, cu(gr["logp"]), cu(gr["action"]), cu(gr["legal"]), adv, gn,
                                   clip_eps=0.2, critic_strength=crit, entropy_strength=ent)
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        want = gl[name + "__stats"]
        st = ppo.loss_stats(stats, crit, ent)
        np.testing.assert_allclose(float(loss), want[0], rtol=RTOL)
        np.testing.assert_allclose([st["loss"], st["policy_loss"], st["value_loss"], st["entropy"], st["entropy_loss"]],
                                   [want[0], want[1], want[2], want[3], want[5]], rtol=RTOL)
        np.testing.assert_allclose(float(gnorm), want[4], rtol=1e-4)
        for k, p in m.named_parameters():
            ref = gl[name + "__grad__" + k.replace(".", "__")]
            np.testing.assert_allclose(p.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-5 * max(1.0, np.abs(ref).max()),
                                       err_msg=k)


def test_trainer_with_symmetry_augmentation():
    """README recipe (--upsample-ratio 0.25): the update batch grows by the mirrored / rotated copies and the step
    still produces finite statistics through the fused update kernel."""
    from g2048 import trainer as tr
    cfg = tr.TrainConfig(hidden_dim=64, num_layers=2, envs=512, horizon=16, zero_heads=False, upsample_ratio=0.25)
    t = tr.Trainer(cfg, torch.device("cuda:0"))
    s = t.train_step()
    n = 512 * 16
    assert n * 1.20 < t.n_update_samples < n * 1.30              # + 0.25 * (1/2 + 1/2) on average
    assert all(np.isfinite(s[k]) for k in ("loss", "policy_loss", "value_loss", "entropy", "grad_norm"))
    cfg0 = tr.TrainConfig(hidden_dim=64, num_layers=2, envs=512, horizon=16, zero_heads=False)
    t0 = tr.Trainer(cfg0, torch.device("cuda:0"))
    t0.train_step()
    assert t0.n_update_samples == n


def _ref_masked_kl(old_logits, new_logits, legal):
    """train.py:586-597 in float64: sum over the legal moves of p_old (log p_old - log p_new)."""
    ok = ((legal.long()[:, None] >> torch.arange(4)) & 1) == 1
    lo = torch.masked_fill(old_logits.double(), ~ok, float("-inf")).log_softmax(-1)
    ln = torch.masked_fill(new_logits.double(), ~ok, float("-inf")).log_softmax(-1)
    return torch.where(ok, lo.exp() * (lo - ln), torch.zeros_like(lo)).sum(-1)


@pytest.mark.parametrize("n", [1, 255, 100003])
def test_masked_kl_matches_torch_restatement(n):
    from g2048 import ppo
    logits, _, _, _, legal, _, _ = _random_loss_inputs(n, n + 5)
    g = torch.Generator().manual_seed(n)
    new = logits + 0.3 * torch.randn((n, 4), generator=g)
    flags = torch.full((n,), 0x80, dtype=torch.uint8)
    flags[1::5] = 0
    ref = _ref_masked_kl(logits, new, legal)
    stats, kl = ppo.masked_kl(logits.cuda(), new.cuda(), legal.cuda(), want_per_sample=True)
    np.testing.assert_allclose(kl.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=2e-6)
    s = stats.tolist()
    np.testing.assert_allclose(s[0], float(ref.sum()), rtol=1e-5, atol=1e-6 * n)
    assert s[1] == n
    np.testing.assert_allclose(s[2], float(ref.max()), rtol=1e-4, atol=2e-6)
    # invalid slots count nothing
    keep = flags != 0
    stats_f, kl_f = ppo.masked_kl(logits.cuda(), new.cuda(), legal.cuda(), flags=flags.cuda(), want_per_sample=True)
    assert float(kl_f.cpu()[~keep].abs().sum()) == 0.0 and stats_f[1].item() == int(keep.sum())
    if int(keep.sum()):
        np.testing.assert_allclose(stats_f[0].item(), float(ref[keep].sum()), rtol=1e-5, atol=1e-6 * n)
        np.testing.assert_allclose(stats_f[2].item(), float(ref[keep].max()), rtol=1e-4, atol=2e-6)
    # identical distributions: exactly zero; empty batch: zeros
    z = ppo.masked_kl(logits.cuda(), logits.cuda(), legal.cuda())[0].tolist()
    assert z[0] == 0.0 and z[2] == 0.0
    e = ppo.masked_kl(torch.empty((0, 4), device="cuda"), torch.empty((0, 4), device="cuda"), torch.empty(0, dtype=torch.uint8, device="cuda"))
    assert e[0].tolist() == [0.0, 0.0, 0.0]


def test_trainer_kl_statistic():
    """TrainConfig.kl_stats: the statistic of train.py:577-597 after every optimizer step, from a forward-only pass of the
    fused kernel with the updated weights; it must not change the training itself."""
    import copy
    from g2048 import ppo, trainer as tr, update
    dev = torch.device("cuda:0")
    kw = dict(hidden_dim=64, num_layers=2, envs=512, horizon=16, zero_heads=False, minibatches=1, seed=5, warmup_steps=0)
    t = tr.Trainer(tr.TrainConfig(kl_stats=True, **kw), dev)
    before = copy.deepcopy(t.model)
    buf = t.collect()
    adv = t.advantages(buf)
    s = t.update(buf, adv)
    assert 0.0 <= s["kl_average"] <= s["kl_max"] and np.isfinite(s["kl_total"]) and s["kl_average"] > 0.0
    # by hand: old logits from the weights before the step, new logits from the weights after it, on the recorded boards
    boards, legal, flags = buf.boards.reshape(-1), buf.legal.reshape(-1), buf.flags.reshape(-1)
    lo, _ = update.forward(before.eval(), boards)
    ln, _ = update.forward(t.model.eval(), boards)
    keep = (flags.cpu() & 0x80) != 0
    ref = _ref_masked_kl(lo.cpu(), ln.cpu(), legal.cpu())[keep]
    # (a per-sample KL of ~1e-4 is a difference of nearly equal fp32 log-probs: ~1e-7 absolute)
    np.testing.assert_allclose(s["kl_total"], float(ref.sum()), rtol=1e-3)
    np.testing.assert_allclose(s["kl_average"], float(ref.mean()), rtol=1e-3)
    np.testing.assert_allclose(s["kl_max"], float(ref.max()), rtol=1e-3, atol=2e-6)
    # the same step without the statistic: identical weights afterwards
    t0 = tr.Trainer(tr.TrainConfig(**kw), dev)
    b0 = t0.collect()
    s0 = t0.update(b0, t0.advantages(b0))
    assert "kl_average" not in s0 and s0["loss"] == s["loss"]
    for (k, v), (_, v0) in zip(t.model.state_dict().items(), t0.model.state_dict().items()):
        assert torch.equal(v, v0), k
