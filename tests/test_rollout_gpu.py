"""GPU parity of the fused rollout kernel (through the C ABI): replay of the reference's own
play_game_for_episode trajectories, oracle env simulation, torch fp32 policy, determinism,
shard invariance, sampling statistics and the play_games_batched drop-in schema."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rollout_as_tb  # noqa: E402
from oracle import oracle as O  # noqa: E402

RESET_TWEAK = 0x9E3779B97F4A7C15
SH_KEYS = ("mono_before", "mono_after", "empt_before", "empt_after", "max_tile_created", "smooth_before",
           "smooth_after", "corner_before", "corner_after", "max_exp_before", "max_exp_after")


def best_model(golden):
    from g2048 import policy
    g = golden("model_best")
    m = policy.GameMLP(policy.MLPConfig(hidden_dim=int(g["hidden_dim"]), num_layers=int(g["num_layers"]), dropout=0.0))
    m.load_state_dict(policy.load_state_dict_from_npz(g))
    return m.cuda().eval()


def random_model(h=196, L=2, seed=0):
    from g2048 import policy
    torch.manual_seed(seed)
    return policy.GameMLP(policy.MLPConfig(hidden_dim=h, num_layers=L, dropout=0.0)).cuda().eval()


def torch_policy_outputs(model, boards, legal):
    from g2048 import env
    with torch.no_grad():
        logits, v = model(env.encode(boards.reshape(-1)))
    illegal = ((legal.reshape(-1).long()[:, None] >> torch.arange(4, device=boards.device)) & 1) == 0
    lp = torch.masked_fill(logits, illegal, float("-inf")).log_softmax(-1)
    p = lp.exp()
    ent = -(torch.where(p > 0, p * lp, torch.zeros_like(p))).sum(-1)
    return lp, v.squeeze(1), ent


@pytest.mark.parametrize("precision", ["fp32", "x3"])
def test_replays_reference_play_game_for_episode(golden, precision):
    """Same weights, same Philox spawn draws, the reference's sampled actions forced: every StepData
    field the reference recorded must come back (ints bit-exact, floats to 1e-5) -- from the fp32 FFMA kernel and
    from the split-fp16 tensor-core kernel (the default at large env batch)."""
    from g2048 import env, rollout
    g = golden("rollout")
    a, (t, e) = rollout_as_tb(g)
    T, B = a["flags"].shape
    seed = int(g["seed"])
    model = best_model(golden)
    boards = env.reset(B, device=0, seed=seed, env0=0, ctr=0)
    forced = torch.from_numpy(a["action"]).cuda()
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=seed, env0=0, ctr0=1, auto_reset=False,
                          alive=torch.ones(B, dtype=torch.uint8, device="cuda"), forced_actions=forced, precision=precision)
    h = lambda x: x.cpu().numpy()
    np.testing.assert_array_equal(h(buf.boards)[t, e].view(np.uint64), g["board"])
    np.testing.assert_array_equal(h(buf.legal)[t, e], g["legal"])
    np.testing.assert_array_equal(h(buf.points)[t, e], g["points"])
    fl = h(buf.flags)[t, e]
    assert (fl & 0x80).all()
    np.testing.assert_array_equal((fl >> 4) & 1, g["done"])
    sh = env.decode_shaping(h(buf.shaping)[t, e])
    done = g["done"].astype(bool)
    np.testing.assert_array_equal(sh["mono_before"], g["mono_before"])
    np.testing.assert_array_equal(np.where(done, 0, sh["mono_after"]), g["mono_after"])      # train.py:318
    np.testing.assert_array_equal(sh["empt_before"], g["empt_before"])
    np.testing.assert_array_equal(np.where(done, 0, sh["empt_after"]), g["empt_after"])      # train.py:322
    np.testing.assert_array_equal(sh["max_tile_created"], g["max_tile_created"])
    np.testing.assert_array_equal(sh["smooth_after"] - sh["smooth_before"], g["smooth_delta"])
    np.testing.assert_array_equal(sh["corner_after"] - sh["corner_before"], g["corner_delta"])
    lp, ref = h(buf.logp)[t, e], g["logp"]
    np.testing.assert_array_equal(np.isinf(lp), np.isinf(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(lp[fin], ref[fin], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(h(buf.value)[t, e], g["value"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(h(buf.entropy)[t, e], g["entropy"], rtol=1e-4, atol=2e-5)
    # games that ended by `done` go idle: their later slots are invalid; final boards match
    full = h(buf.flags)
    for env_i, n in enumerate(g["ep_len"]):
        if g["done"][(g["env"] == env_i)][-1]:
            assert (full[n:, env_i] == 0).all()
            assert h(boards).view(np.uint64)[env_i] == g["ep_final"][env_i]


@pytest.mark.parametrize("precision", ["fp32", "x3"])
@pytest.mark.parametrize("h,L,B,T", [(196, 2, 1000, 48), (64, 1, 130, 40), (128, 3, 257, 16), (192, 2, 128, 8), (196, 0, 128, 4),
                                     (200, 4, 300, 6)])
def test_rollout_matches_oracle_env_and_torch_policy(h, L, B, T, precision):
    from g2048 import env, rollout
    model = random_model(h, L, seed=h + L)
    seed, env0 = 99, 12345
    boards = env.reset(B, device=0, seed=seed, env0=env0, ctr=0)
    start = boards.clone()
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=seed, env0=env0, ctr0=1, auto_reset=True, precision=precision)
    b = start.cpu().numpy().view(np.uint64)
    n_done = 0
    for t in range(T):
        np.testing.assert_array_equal(buf.boards[t].cpu().numpy().view(np.uint64), b)
        acts = buf.actions[t].cpu().numpy()
        nb, info = O.step_batch(b, acts, seed=seed, env0=env0, ctr=1 + t)
        assert (info["invalid"] == 0).all()                       # sampled actions are always legal
        np.testing.assert_array_equal(buf.legal[t].cpu().numpy(), info["legal_before"])
        np.testing.assert_array_equal(buf.points[t].cpu().numpy(), info["points"])
        fl = buf.flags[t].cpu().numpy()
        np.testing.assert_array_equal(fl, 0x80 | info["legal_after"] | (info["done"] << 4))
        sh = env.decode_shaping(buf.shaping[t].cpu().numpy())
        for k in SH_KEYS:
            np.testing.assert_array_equal(sh[k], info[k], err_msg=k)
        d = info["done"].astype(bool)
        n_done += int(d.sum())
        if d.any():   # auto-reset: fresh board from the reset key stream at the same counter
            fresh = O.reset_batch(B, seed=seed ^ RESET_TWEAK, env0=env0, ctr=1 + t)
            nb = np.where(d, fresh, nb)
        b = nb
    np.testing.assert_array_equal(boards.cpu().numpy().view(np.uint64), b)
    lp, v, ent = torch_policy_outputs(model, buf.boards, buf.legal)
    got = buf.logp.reshape(-1, 4)
    fin = torch.isfinite(lp)
    assert torch.equal(torch.isfinite(got), fin)
    torch.testing.assert_close(got[fin], lp[fin], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(buf.value.reshape(-1), v, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(buf.entropy.reshape(-1), ent, rtol=1e-4, atol=2e-5)


def test_rollout_is_deterministic_and_shard_invariant():
    from g2048 import env, rollout
    model = random_model(196, 2, seed=5)
    pol = rollout.pack_policy(model)
    B, T, seed = 640, 32, 4

    def run(lo, hi):
        boards = env.reset(hi - lo, device=0, seed=seed, env0=lo, ctr=0)
        return rollout.rollout(pol, boards, T, seed=seed, env0=lo, ctr0=1, auto_reset=True), boards

    whole, wb = run(0, B)
    again, _ = run(0, B)
    left, lb = run(0, 300)
    right, rb = run(300, B)
    for name in ("boards", "actions", "legal", "points", "shaping", "flags"):
        assert torch.equal(getattr(whole, name), getattr(again, name)), name
        assert torch.equal(getattr(whole, name), torch.cat([getattr(left, name), getattr(right, name)], dim=1)), name
    assert torch.equal(wb, torch.cat([lb, rb]))
    assert torch.equal(whole.logp, again.logp) and torch.equal(whole.value, again.value)


def test_sampling_follows_the_policy_distribution():
    from g2048 import env, rollout
    model = random_model(196, 2, seed=11)
    B, T = 8192, 16
    boards = env.reset(B, device=0, seed=1, env0=0, ctr=0)
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=1, env0=0, ctr0=1, auto_reset=True)
    p = buf.logp.reshape(-1, 4).exp().double()
    a = buf.actions.reshape(-1).long()
    assert bool((torch.gather(p, 1, a[:, None]) > 0).all())          # never an illegal action
    expected = p.sum(0)
    observed = torch.bincount(a, minlength=4).double()
    var = (p * (1 - p)).sum(0)
    z = (observed - expected) / var.sqrt()
    assert float(z.abs().max()) < 5.0, (observed, expected, z)


def test_play_games_batched_drop_in_schema(golden):
    import batched_rollout
    from g2048 import env
    model = best_model(golden)
    eps = batched_rollout.play_games_batched(model, num_games=5, max_steps=None, device=torch.device("cuda:0"), seed=3)
    assert len(eps) == 5
    for ep in eps:
        moves = ep["moves"]
        assert set(ep) == {"moves", "total_points", "total_steps", "final_state"}
        assert ep["total_steps"] == len(moves) - 1                   # ended by `done` (train.py:334-343)
        assert ep["total_points"] == sum(m["points_earned"] for m in moves)
        assert not env.Game2048.state_has_next_step(ep["final_state"])
        for i, m in enumerate(moves):
            assert m["game_state"].shape == (48,) and not m["game_state"].is_cuda      # host tensors: see play_games_batched
            assert len(m["policy_logprobs"]) == 4 and len(m["action_mask"]) == 4
            assert {"adjacency_delta", "chain_delta", "topological_delta", "smoothness_delta", "corner_delta"} <= set(m)
            assert not m["action_mask"][m["selected_direction"]]
            assert np.isfinite(m["policy_logprobs"][m["selected_direction"]])
            if i + 1 < len(moves):
                assert m["result_state"] == moves[i + 1]["state_before"]
        assert moves[-1]["result_state"] == ep["final_state"]
        assert moves[-1]["monotonicity_after"] == 0.0 and moves[-1]["emptiness_after"] == 0.0
        # replay through the oracle with the recorded actions
        b = np.array([O.pack_grid(moves[0]["state_before"])], dtype=np.uint64)
        pre = O.expand4_batch(b)[0][0, moves[0]["selected_direction"]]
        cells = lambda x: np.array(O.unpack_board(x)).reshape(-1)
        diff = cells(pre) != np.array(moves[0]["result_state"]).reshape(-1)
        assert diff.sum() == 1
    capped = batched_rollout.play_games_batched(model, num_games=3, max_steps=10, device="cuda:0", seed=3, game_state_on_device=True)
    assert all(len(ep["moves"]) == 10 and ep["total_steps"] == 10 for ep in capped)
    assert all(m["game_state"].is_cuda for ep in capped for m in ep["moves"])
    with pytest.raises(RuntimeError):
        batched_rollout.play_games_batched(model, num_games=1, device=None)


def test_trainer_step_runs_and_learns_signal():
    """One rollout + advantage + update step of the thin driver on a small config."""
    from g2048 import trainer as tr
    cfg = tr.TrainConfig(hidden_dim=64, num_layers=1, envs=512, horizon=32, chunk=4096, minibatches=2, epochs=1)
    t = tr.Trainer(cfg, torch.device("cuda:0"))
    w0 = t.model.action_head.weight.clone()
    s1 = t.train_step()
    s2 = t.train_step()
    for s in (s1, s2):
        assert np.isfinite(s["loss"]) and np.isfinite(s["grad_norm"]) and s["env_steps"] == 512 * 32
    assert not torch.equal(w0, t.model.action_head.weight)        # the optimizer moved the policy head
    assert t.moments.step == 3 and t.moments.m2 != 1.0
    assert abs(s1["entropy"]) > 0.5                                # zero-initialised heads: near-uniform policy


def bf16_emulated_policy(model, boards, legal):
    """The tensor-core kernel's arithmetic in plain torch: bf16-rounded GEMM operands, fp32
    accumulation, fp32 LayerNorm / residual stream / heads."""
    from g2048 import env
    import torch.nn.functional as F
    r = lambda t: t.bfloat16().float()
    sd = model.state_dict()
    x48 = env.encode(boards.reshape(-1))
    w = sd["stem.0.weight"]
    pos = x48.clone()
    pos[:, 0::3] = 0
    z = x48[:, 0::3] @ r(w[:, 0::3]).T + pos @ w.T           # exponents through bf16 weights, positions in fp32
    h = w.shape[0]
    x = F.relu(F.layer_norm(z, (h,), sd["stem.1.weight"], sd["stem.1.bias"], 1e-5))
    L = len(model.backbone)
    for l in range(L):
        z = r(x) @ r(sd[f"backbone.{l}.mlp.0.weight"]).T
        x = x + F.relu(F.layer_norm(z, (h,), sd[f"backbone.{l}.mlp.1.weight"], sd[f"backbone.{l}.mlp.1.bias"], 1e-5))
    logits = x @ sd["action_head.weight"].T + sd["action_head.bias"]
    v = (x @ sd["value_head.weight"].T + sd["value_head.bias"]).squeeze(1)
    illegal = ((legal.reshape(-1).long()[:, None] >> torch.arange(4, device=boards.device)) & 1) == 0
    return torch.masked_fill(logits, illegal, float("-inf")).log_softmax(-1), v


@pytest.mark.parametrize("h,L,B,T", [(196, 2, 1000, 24), (64, 1, 130, 16), (192, 3, 300, 8), (196, 0, 128, 4)])
def test_tensor_core_rollout(h, L, B, T):
    """bf16 tcgen05 variant: the integer env path stays bit-exact; log-probs / values match a torch
    emulation of the same bf16-operand arithmetic tightly and the fp32 policy to ~1e-2."""
    from g2048 import env, rollout
    model = random_model(h, L, seed=h + L + 1)
    seed, env0 = 77, 999
    boards = env.reset(B, device=0, seed=seed, env0=env0, ctr=0)
    start = boards.clone()
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=seed, env0=env0, ctr0=1, auto_reset=True,
                          precision="bf16")
    torch.cuda.synchronize()
    b = start.cpu().numpy().view(np.uint64)
    for t in range(T):
        np.testing.assert_array_equal(buf.boards[t].cpu().numpy().view(np.uint64), b)
        nb, info = O.step_batch(b, buf.actions[t].cpu().numpy(), seed=seed, env0=env0, ctr=1 + t)
        assert (info["invalid"] == 0).all()
        np.testing.assert_array_equal(buf.legal[t].cpu().numpy(), info["legal_before"])
        np.testing.assert_array_equal(buf.points[t].cpu().numpy(), info["points"])
        np.testing.assert_array_equal(buf.flags[t].cpu().numpy(), 0x80 | info["legal_after"] | (info["done"] << 4))
        sh = env.decode_shaping(buf.shaping[t].cpu().numpy())
        for k in SH_KEYS:
            np.testing.assert_array_equal(sh[k], info[k], err_msg=k)
        d = info["done"].astype(bool)
        if d.any():
            nb = np.where(d, O.reset_batch(B, seed=seed ^ RESET_TWEAK, env0=env0, ctr=1 + t), nb)
        b = nb
    got = buf.logp.reshape(-1, 4)
    lp, v = bf16_emulated_policy(model, buf.boards, buf.legal)
    fin = torch.isfinite(lp)
    assert torch.equal(torch.isfinite(got), fin)
    # (a 1e-7 difference before a bf16 rounding can flip one operand ulp: 1e-2, not fp32 tolerances)
    torch.testing.assert_close(got[fin], lp[fin], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(buf.value.reshape(-1), v, rtol=1e-2, atol=1e-2)
    assert float((got[fin] - lp[fin]).abs().mean()) < 5e-4
    lp32, v32, _ = torch_policy_outputs(model, buf.boards, buf.legal)
    err_lp = float((got[fin] - lp32[fin]).abs().max())
    err_v = float((buf.value.reshape(-1) - v32).abs().max())
    print(f"bf16 tensor-core rollout vs fp32 policy: max |dlogp| = {err_lp:.4f}, max |dV| = {err_v:.4f}")
    assert err_lp < 0.1 and err_v < 0.1


@pytest.mark.parametrize("precision", ["fp32", "bf16", "x3"])
def test_full_size_rollout_properties(precision):
    """C3-sized env batch (65 536 envs): size-independent properties instead of a CPU replay --
    the rollout kernel's inlined env agrees with the standalone g2048_step kernel fed the recorded
    actions and the same Philox counters (boards, points, shaping, flags), every sampled action is
    legal, auto-reset boards hold exactly two tiles, and a second run is bit-identical."""
    from g2048 import env, rollout
    model = random_model(196, 2, seed=3)
    pol = rollout.pack_policy(model)
    B, T, seed = 65536, 12, 2048
    boards = env.reset(B, device=0, seed=seed, env0=0, ctr=0)
    start = boards.clone()
    buf = rollout.rollout(pol, boards, T, seed=seed, env0=0, ctr0=1, auto_reset=True, precision=precision)
    again = rollout.rollout(pol, start.clone(), T, seed=seed, env0=0, ctr0=1, auto_reset=True, precision=precision)
    for name in ("boards", "actions", "points", "shaping", "flags", "logp", "value"):
        assert torch.equal(getattr(buf, name), getattr(again, name)), name
    assert torch.equal(buf.boards[0], start)
    assert bool(((buf.legal.long() >> buf.actions.long()) & 1).all())
    assert bool((buf.flags & 0x80).all())
    for t in range(T):
        r = env.step(buf.boards[t], buf.actions[t], seed=seed, env0=0, ctr=1 + t)
        assert torch.equal(r["points"], buf.points[t]) and torch.equal(r["shaping"], buf.shaping[t])
        assert torch.equal(r["flags"] | 0x80, buf.flags[t])
        nxt = buf.boards[t + 1] if t + 1 < T else boards
        done = (r["flags"] & 0x10) != 0
        assert torch.equal(r["boards"][~done], nxt[~done])
        if bool(done.any()):   # auto-reset: a fresh board with exactly two tiles
            fresh = nxt[done]
            tiles = sum(((fresh >> (4 * k)) & 15) != 0 for k in range(16))
            assert bool((tiles == 2).all())


def test_c3_horizon_rollout_replayed_through_the_oracle():
    """The C3 horizon (512 steps, auto-reset) on 4096 envs with the kernel `auto` selects at C3's env count: every one of the
    2 097 152 transitions is replayed through the oracle env (boards, legal masks, points, flags, every shaping integer,
    the reset boards), and every recorded log-prob / value / entropy is compared with the torch fp32 policy."""
    from g2048 import env, rollout
    model = random_model(196, 2, seed=17)
    B, T, seed, env0 = 4096, 512, 2048, 1 << 20
    assert rollout.resolve_precision("auto", 65536, 2) == "x3"
    boards = env.reset(B, device=0, seed=seed, env0=env0, ctr=0)
    start = boards.clone()
    buf = rollout.rollout(rollout.pack_policy(model), boards, T, seed=seed, env0=env0, ctr0=1, auto_reset=True, precision="x3")
    h = {k: getattr(buf, k).cpu().numpy() for k in ("boards", "actions", "legal", "points", "flags", "shaping")}
    b = start.cpu().numpy().view(np.uint64)
    resets = 0
    for t in range(T):
        np.testing.assert_array_equal(h["boards"][t].view(np.uint64), b, err_msg=f"t={t}")
        nb, info = O.step_batch(b, h["actions"][t], seed=seed, env0=env0, ctr=1 + t)
        assert (info["invalid"] == 0).all()
        np.testing.assert_array_equal(h["legal"][t], info["legal_before"])
        np.testing.assert_array_equal(h["points"][t], info["points"])
        np.testing.assert_array_equal(h["flags"][t], 0x80 | info["legal_after"] | (info["done"] << 4))
        sh = env.decode_shaping(h["shaping"][t])
        for k in SH_KEYS:
            np.testing.assert_array_equal(sh[k], info[k], err_msg=k)
        d = info["done"].astype(bool)
        if d.any():
            resets += int(d.sum())
            nb = np.where(d, O.reset_batch(B, seed=seed ^ RESET_TWEAK, env0=env0, ctr=1 + t), nb)
        b = nb
    np.testing.assert_array_equal(boards.cpu().numpy().view(np.uint64), b)
    assert resets > 0, "a 512-step horizon must see finished games"
    lp, v, ent = torch_policy_outputs(model, buf.boards, buf.legal)
    got = buf.logp.reshape(-1, 4)
    fin = torch.isfinite(lp)
    assert torch.equal(torch.isfinite(got), fin)
    torch.testing.assert_close(got[fin], lp[fin], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(buf.value.reshape(-1), v, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(buf.entropy.reshape(-1), ent, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("auto_reset", [True, False])
def test_horizon_segments_give_identical_rollouts(auto_reset):
    """More tiles than SMs and a horizon of >= 64 steps: the x3 kernel cuts each tile's horizon into segments dealt round-robin
    over the SMs (board state handed over through global memory behind a flag).  Every record -- floats included -- must equal
    the rollout of the same envs launched in two halves that fit the SMs (no segmentation), with and without auto-reset."""
    from g2048 import env, rollout
    model = random_model(196, 2, seed=23)
    pol = rollout.pack_policy(model)
    B, T, seed = 20000, 72, 77                      # 157 tiles on 148 SMs
    names = ("boards", "actions", "legal", "points", "shaping", "flags", "logp", "value", "entropy")

    def run(lo, hi):
        boards = env.reset(hi - lo, device=0, seed=seed, env0=lo, ctr=0)
        alive = None if auto_reset else torch.ones(hi - lo, dtype=torch.uint8, device="cuda")
        buf = rollout.rollout(pol, boards, T, seed=seed, env0=lo, ctr0=1, auto_reset=auto_reset, alive=alive, precision="x3")
        return buf, boards, alive

    whole, wb, wa = run(0, B)
    left, lb, la = run(0, B // 2)
    right, rb, ra = run(B // 2, B)
    for name in names:
        assert torch.equal(getattr(whole, name), torch.cat([getattr(left, name), getattr(right, name)], dim=1)), name
    assert torch.equal(wb, torch.cat([lb, rb]))
    if not auto_reset:
        assert torch.equal(wa, torch.cat([la, ra]))
    assert bool((whole.flags & 0x80).any())


def test_evaluate_reports_the_reference_metrics(golden, tmp_path):
    """rollout.evaluate = the eval block of train.py:1840-1875 without the episode dictionaries: its scores are the
    total_points play_games_batched reports for the same seeded games, and the checkpoint round-trips."""
    import torch
    from g2048 import rollout
    from test_train_cpu import load_policy
    m, _ = load_policy(golden)
    m = m.cuda().eval()
    ev = rollout.evaluate(m, eval_games=64, device="cuda", seed=3)
    assert set(ev) >= {"eval/max_score", "eval/avg_score", "eval/median_score", "eval/pct_512", "eval/pct_1024", "eval/pct_2048"}
    assert ev["eval/max_score"] >= ev["eval/median_score"] > 0 and ev["eval/avg_score"] > 500      # the shipped checkpoint plays
    assert 0.0 <= ev["eval/pct_2048"] <= ev["eval/pct_1024"] <= ev["eval/pct_512"] <= 100.0
    again = rollout.evaluate(m, eval_games=64, device="cuda", seed=3)
    assert again["scores"] == ev["scores"]                                                        # seeded => reproducible
    path = tmp_path / "best_model.pt"
    rollout.save_best_checkpoint(path, m, ev["eval/avg_score"], 7)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"model_state_dict", "config", "eval_avg_score", "train_step"} and ck["train_step"] == 7
    assert all(torch.equal(ck["model_state_dict"][k], v.cpu()) for k, v in m.state_dict().items())


def test_tensor_core_rollout_forced_actions_and_idle_envs():
    """Without auto-reset games end and their envs go idle (zero records); with forced actions the env path of
    the tensor-core kernel (whose per-env tail is split over the four threads of a row) must write the same
    integer records as the fp32 kernel, idle rows included."""
    from g2048 import env, rollout
    model = random_model(64, 1, seed=5)
    pol = rollout.pack_policy(model)
    B, T, seed = 200, 320, 11
    start = env.reset(B, device=0, seed=seed, env0=3, ctr=0)
    ones = lambda: torch.ones(B, dtype=torch.uint8, device="cuda")
    a1 = ones()
    ref = rollout.rollout(pol, start.clone(), T, seed=seed, env0=3, ctr0=1, auto_reset=False, alive=a1, precision="fp32")
    assert int(a1.sum()) < B, "the horizon should outlive some games"
    assert bool((ref.flags[-1] == 0).any()) and bool((ref.flags[0] & 0x80).all())
    idle = ref.flags == 0
    for precision in ("bf16", "x3"):
        a2 = ones()
        tcb = rollout.rollout(pol, start.clone(), T, seed=seed, env0=3, ctr0=1, auto_reset=False, alive=a2, precision=precision,
                              forced_actions=ref.actions.clone())
        assert torch.equal(a1, a2)
        for name in ("boards", "actions", "legal", "points", "shaping", "flags"):
            assert torch.equal(getattr(ref, name), getattr(tcb, name)), (precision, name)
        assert bool((tcb.value[idle] == 0).all()) and bool((tcb.logp[idle] == 0).all())
