"""CPU: the policy mirror reproduces the reference network bit-for-bit on the shipped checkpoint,
and the torch restatement of the loss (the checker of the fused CUDA loss) reproduces the stats and
clipped gradients that the reference's own model_optimize_step produced (tests/golden/loss.npz)."""
import numpy as np
import torch

from helpers import ref_ppo_loss_torch, rollout_as_tb


def load_policy(golden, dropout=0.0):
    from g2048 import policy
    g = golden("model_best")
    m = policy.GameMLP(policy.MLPConfig(hidden_dim=int(g["hidden_dim"]), num_layers=int(g["num_layers"]),
                                        dropout=dropout))
    m.load_state_dict(policy.load_state_dict_from_npz(g))
    return m, g


def policy_b(golden):
    """The update-side policy of the loss fixture (oracle/make_golden.py gen_loss)."""
    m, _ = load_policy(golden)
    with torch.no_grad():
        m.action_head.weight.mul_(1.3)
        m.value_head.bias.add_(0.3)
    return m


def test_policy_mirror_matches_reference_forward(golden):
    m, g = load_policy(golden)
    m.eval()
    with torch.no_grad():
        logits, v = m(torch.from_numpy(g["inputs"]))
    np.testing.assert_array_equal(logits.numpy(), g["logits"])
    np.testing.assert_array_equal(v.numpy(), g["value"])
    assert [d.value for d in m.directions] == ["up", "down", "left", "right"]
    assert sum(p.numel() for p in m.parameters()) == 192 * 48 + 2 * 192 + 2 * (192 * 192 + 2 * 192) + 5 * 192 + 5


def test_policy_state_dict_keys_h196():
    from g2048 import policy
    m = policy.GameMLP(policy.MLPConfig(hidden_dim=196))
    keys = set(m.state_dict().keys())
    assert {"stem.0.weight", "stem.1.weight", "stem.1.bias", "backbone.0.mlp.0.weight", "backbone.1.mlp.1.bias",
            "action_head.weight", "action_head.bias", "value_head.weight", "value_head.bias"} <= keys
    assert sum(p.numel() for p in m.parameters()) == 88401
    groups = m.get_param_groups(1e-4, 1e-3)
    assert [len(g["params"]) for g in groups] == [4, 7, 1, 1]


def test_loss_restatement_matches_reference_optimize_step(golden):
    from oracle import oracle as O
    gr, ga, gl = golden("rollout"), golden("advantage"), golden("loss")
    m = policy_b(golden)
    m.train()
    x = torch.from_numpy(O.encode_batch(gr["board"]))
    adv = torch.from_numpy(ga["readme__adv"].astype(np.float32))
    gn = torch.from_numpy(ga["readme__g_norm"].astype(np.float32))
    for name in ("readme", "alt"):
        ent, crit = gl[name + "__coef"].tolist()
        m.zero_grad()
        logits, v = m(x)
        loss, parts = ref_ppo_loss_torch(logits, v, torch.from_numpy(gr["logp"]), torch.from_numpy(gr["action"]),
                                         torch.from_numpy(gr["legal"]), adv, gn, 0.2, crit, ent)
        loss.backward()
        gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        want = gl[name + "__stats"]
        np.testing.assert_allclose(float(loss), want[0], rtol=1e-5)
        np.testing.assert_allclose(-float(parts["ppo"]), want[1], rtol=1e-5)
        np.testing.assert_allclose(crit * float(parts["vl"]), want[2], rtol=1e-5)
        np.testing.assert_allclose(float(parts["ent"]), want[3], rtol=1e-5)
        np.testing.assert_allclose(float(gnorm), want[4], rtol=1e-4)
        for k, p in m.named_parameters():
            ref = gl[name + "__grad__" + k.replace(".", "__")]
            np.testing.assert_allclose(p.grad.numpy(), ref, rtol=1e-3, atol=1e-6 * max(1.0, np.abs(ref).max()), err_msg=k)


def test_rtg_moments_update_matches_reference(golden):
    from g2048.ppo import RtgMoments
    from oracle import oracle as O
    g, adv = golden("rollout"), golden("advantage")
    a, _ = rollout_as_tb(g)
    for name in ("readme", "warm"):
        gamma, wp, wm, we, beta, step, mu, m2 = adv[name + "__cfg"].tolist()
        mom = RtgMoments(mu=mu, m2=m2, step=int(step))
        mu_c, sd = mom.corrected(beta)
        graw = adv[name + "__g_raw"]
        mom.update(beta, float(graw.sum()), float((graw * graw).sum()), len(graw))
        np.testing.assert_allclose([mom.mu, mom.m2], adv[name + "__moments_out"], rtol=1e-10)
        gn = (graw - mu_c) / (sd + 1e-8)
        np.testing.assert_allclose(gn, adv[name + "__g_norm"], rtol=1e-12, atol=1e-12)


def test_urm_mirror_matches_reference_forward(golden):
    from g2048 import policy
    g = golden("model_urm")
    m = policy.GameURM(policy.GameURMConfig()).eval()
    m.load_state_dict(policy.load_state_dict_from_npz(g))
    with torch.no_grad():
        logits, v = m(torch.from_numpy(g["inputs"]))
    np.testing.assert_array_equal(logits.numpy(), g["logits"])
    np.testing.assert_array_equal(v.numpy(), g["value"])
    assert m.layers[0].mlp.inter == 120


def test_optimize_epoch_order_is_the_dataloader_shuffle():
    """g2048.optimize._epoch_order consumes torch's global RNG exactly like DataLoader(shuffle=True)
    (train.py:438-443): same permutation, same RNG state afterwards, epoch after epoch."""
    from torch.utils.data import DataLoader, Dataset

    from g2048 import optimize

    class D(Dataset):
        def __len__(self):
            return 1163

        def __getitem__(self, i):
            return i

    torch.manual_seed(777)
    want = [[int(i) for b in DataLoader(D(), batch_size=128, shuffle=True, collate_fn=lambda b: b) for i in b] for _ in range(3)]
    tail_want = float(torch.rand(1))
    torch.manual_seed(777)
    got = [optimize._epoch_order(1163).tolist() for _ in range(3)]
    assert got == want and float(torch.rand(1)) == tail_want


def test_optimize_collation_matches_reference_fields(golden):
    """episodes_to_batch = MyDataset + collate_fn (train.py:360-411) for the whole dataset: boards re-packed from
    game_state, legal bits = complement of action_mask, float32 advantage / future_reward / log-probs."""
    from g2048 import optimize
    gr, ga = golden("rollout"), golden("advantage")
    pos = np.arange(16)
    moves = []
    for k in range(64):
        exps = (gr["board"][k] >> (4 * pos).astype(np.uint64)) & np.uint64(15)
        x = np.stack([exps.astype(np.float32), (pos // 4 / 3).astype(np.float32), (pos % 4 / 3).astype(np.float32)], 1)
        moves.append({"game_state": torch.from_numpy(x.reshape(48)), "selected_direction": int(gr["action"][k]),
                      "action_mask": [not (int(gr["legal"][k]) >> d) & 1 for d in range(4)],
                      "advantage": float(ga["readme__adv"][k]), "future_reward": float(ga["readme__g_norm"][k]),
                      "policy_logprobs": [float(v) for v in gr["logp"][k]]})
    b = optimize.episodes_to_batch([{"moves": moves[:40]}, {"moves": moves[40:]}], torch.device("cpu"))
    np.testing.assert_array_equal(b["boards"].numpy().view(np.uint64), gr["board"][:64])
    np.testing.assert_array_equal(b["actions"].numpy(), gr["action"][:64])
    np.testing.assert_array_equal(b["legal"].numpy(), gr["legal"][:64])
    np.testing.assert_array_equal(b["adv"].numpy(), ga["readme__adv"][:64].astype(np.float32))
    np.testing.assert_array_equal(b["g_norm"].numpy(), ga["readme__g_norm"][:64].astype(np.float32))
    np.testing.assert_array_equal(b["logp"].numpy(), gr["logp"][:64])


def test_augmentation_sampling_policy():
    """ppo.sample_augmentation follows train.py:776-863: int(N * ratio) distinct recorded steps; each yields a
    mirror copy w.p. 1/2 (two axes, uniform) and a rotation copy w.p. 1/2 (three angles, uniform)."""
    from g2048 import ppo
    from g2048.env import MIRROR_H, MIRROR_V, ROT90, ROT180, ROT270
    n = 200000
    valid = torch.ones(n, dtype=torch.bool)
    valid[::7] = False
    gen = torch.Generator().manual_seed(5)
    src, ops = ppo.sample_augmentation(valid, 0.25, gen)
    k = int(int(valid.sum()) * 0.25)
    assert valid[src].all() and ops.dtype == torch.uint8 and src.numel() == ops.numel()
    is_m = ops <= MIRROR_V
    for part in (src[is_m], src[~is_m]):                       # at most one mirror and one rotation per drawn step
        assert part.unique().numel() == part.numel()
    assert torch.cat([src[is_m], src[~is_m]]).unique().numel() <= k
    frac = lambda c: float((ops == c).sum()) / k
    for c, want in ((MIRROR_H, 0.25), (MIRROR_V, 0.25), (ROT90, 1 / 6), (ROT180, 1 / 6), (ROT270, 1 / 6)):
        assert abs(frac(c) - want) < 0.01, (c, frac(c))
    assert ppo.sample_augmentation(valid, 0.0, gen)[0].numel() == 0
