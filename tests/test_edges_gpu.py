"""Edge cases through the C ABI: empty batches on every entry point, ragged sizes around the tile /
staging thresholds, error reporting for bad arguments."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402


def test_empty_batches_are_no_ops():
    from g2048 import env, fused, ppo, rollout
    from g2048.policy import GameMLP, MLPConfig
    dev = torch.device("cuda:0")
    e64 = torch.empty(0, dtype=torch.int64, device=dev)
    e8 = torch.empty(0, dtype=torch.uint8, device=dev)
    assert env.reset(0, device=dev).numel() == 0
    r = env.step(e64, e8)
    assert r["boards"].numel() == 0 and r["shaping"].numel() == 0
    assert env.expand4(e64)["succ"].shape == (0, 4)
    r4 = env.step4(e64)
    assert r4["boards"].shape == (0, 4) and r4["shaping"].shape == (0, 4)
    assert env.potentials(e64).shape == (0, 6) and env.encode(e64).shape == (0, 48)
    assert env.potentials_ext(e64, e64).shape == (0, 7)
    a = env.augment(e64, e64, e8, e8, torch.empty((0, 4), device=dev), e8)
    assert a["before"].numel() == 0
    z = torch.empty((0, 8), device=dev)
    adv = ppo.rtg_advantage(torch.empty((0, 4), dtype=torch.int32, device=dev), torch.empty((0, 4), dtype=torch.int64, device=dev),
                            torch.empty((0, 4), dtype=torch.uint8, device=dev), torch.empty((0, 4), device=dev),
                            gamma=0.99, w_points=1, w_mono=1, w_empt=0, mu_c=0, stddev=1)
    assert adv["stats"].tolist() == [0.0, 0.0, 0.0]
    loss, stats = ppo.ppo_loss(torch.empty((0, 4), device=dev), torch.empty(0, device=dev), torch.empty((0, 4), device=dev),
                               e8, e8, torch.empty(0, device=dev), torch.empty(0, device=dev), n_total=1)
    assert float(loss) == 0.0 and stats.tolist() == [0.0] * 4
    assert fused.ln_relu_res(z, torch.ones(8, device=dev), torch.zeros(8, device=dev)).shape == (0, 8)
    pol = rollout.pack_policy(GameMLP(MLPConfig(hidden_dim=64, num_layers=1)).to(dev))
    for prec in ("fp32", "bf16", "x3"):
        assert rollout.rollout(pol, e64, 3, seed=1, precision=prec).shape == (3, 0)
        assert rollout.rollout(pol, env.reset(5, device=dev), 0, seed=1, precision=prec).shape == (0, 5)
    # GameURM: rollout kernels and the update ops
    from g2048 import urm_ops
    from g2048.policy import GameURM, GameURMConfig
    um = GameURM(GameURMConfig(dropout=0.0)).to(dev)
    upol = rollout.pack_policy(um.eval())
    for prec in ("x3", "fp16"):
        assert rollout.rollout(upol, e64, 2, seed=1, precision=prec).shape == (2, 0)
    assert urm_ops.attention(torch.empty((0, 16, 192), device=dev)).shape == (0, 16, 64)
    assert urm_ops.residual_norm(torch.empty((0, 64), device=dev), torch.empty((0, 64), device=dev), 1e-5).shape == (0, 64)
    y = urm_ops.conv_swiglu(torch.empty((0, 16, 120), device=dev), torch.empty((0, 16, 120), device=dev), um.layers[0].mlp.dwconv.weight.detach(),
                            um.layers[0].mlp.dwconv.bias.detach())
    assert y.shape == (0, 16, 120)


@pytest.mark.parametrize("n", [(1 << 17) - 1, 1 << 17, (1 << 17) + 1, 148 * 1024 + 1, 148 * 2048 - 1])
def test_sizes_around_the_staging_threshold(n):
    """The staged (shared-memory table, persistent) and direct (L2 table) kernels agree with the oracle
    on both sides of the switch-over size and at grid-stride remainders."""
    from g2048 import env
    rng = np.random.default_rng(n)
    e = rng.integers(1, 14, (n, 16))
    e[rng.random((n, 16)) < 0.3] = 0
    boards = (e.astype(np.uint64) << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(axis=1).astype(np.uint64)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    d = torch.from_numpy(boards.view(np.int64)).cuda()
    r = env.step(d, torch.from_numpy(actions).cuda(), seed=1, env0=7, ctr=2)
    want_b, want = O.step_batch(boards, actions, seed=1, env0=7, ctr=2)
    np.testing.assert_array_equal(r["boards"].cpu().numpy().view(np.uint64), want_b)
    np.testing.assert_array_equal(r["points"].cpu().numpy(), want["points"])
    ex = env.expand4(d, want_max_tile=True)
    succ, points, mt, legal = O.expand4_batch(boards)
    np.testing.assert_array_equal(ex["succ"].cpu().numpy().view(np.uint64), succ)
    np.testing.assert_array_equal(ex["points"].cpu().numpy(), points)
    np.testing.assert_array_equal(ex["max_tile"].cpu().numpy(), mt)
    np.testing.assert_array_equal(ex["legal"].cpu().numpy(), legal)


def test_bad_arguments_are_reported_not_crashes():
    from g2048 import _lib, env, rollout
    from g2048.policy import GameMLP, MLPConfig
    env.init(0)
    lib = _lib.lib()
    rc = lib.g2048_step(None, None, None, None, None, None, None, 5, None, 0, 0, 0, None)
    assert rc == -1 and b"NULL" in lib.g2048_last_error()
    rc = lib.g2048_step(None, None, None, None, None, None, None, -1, None, 0, 0, 0, None)
    assert rc == -1 and b"n < 0" in lib.g2048_last_error()
    rc = lib.g2048_step4(None, None, None, None, None, None, 5, None, 0, 0, 0, None)
    assert rc == -1 and b"NULL" in lib.g2048_last_error()
    b = torch.zeros(8, dtype=torch.int64, device="cuda")
    odd = torch.zeros(4 * 8 + 1, dtype=torch.uint8, device="cuda")[1:]            # flags not 4-byte aligned
    with pytest.raises(_lib.G2048Error):
        _lib.call("g2048_step4", env._ptr(env.lut(torch.device("cuda:0"))), env._ptr(b), env._ptr(torch.zeros((8, 4), dtype=torch.int64, device="cuda")),
                  env._ptr(torch.zeros((8, 4), dtype=torch.int32, device="cuda")), env._ptr(odd), None, 8, None, 0, 0, 0, None)
    from g2048 import ppo  # noqa: F401  (registers g2048_masked_kl)
    st = torch.zeros(3, dtype=torch.float64, device="cuda")
    rc = lib.g2048_masked_kl(None, None, None, None, 5, None, env._ptr(st), None, None)
    assert rc == -1 and b"NULL" in lib.g2048_last_error()
    rc = lib.g2048_masked_kl(None, None, None, None, -2, None, env._ptr(st), None, None)
    assert rc == -1 and b"n < 0" in lib.g2048_last_error()
    rc = lib.g2048_masked_kl(None, None, None, None, 0, None, None, None, None)
    assert rc == -1 and b"stats_out" in lib.g2048_last_error()
    with pytest.raises(ValueError):
        rollout.pack_policy(GameMLP(MLPConfig(hidden_dim=512, num_layers=1)).cuda())
    with pytest.raises(_lib.G2048Error):
        _lib.call("g2048_init", 99)
    with pytest.raises(TypeError):
        env.step(torch.zeros(4, dtype=torch.int32, device="cuda"), torch.zeros(4, dtype=torch.uint8, device="cuda"))
