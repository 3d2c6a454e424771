"""Pins the C oracle (oracle/oracle2048.c) against fixtures produced by the reference itself
(oracle/make_golden.py) and against SURVEY.md section 4's known answers.  CPU only."""
import hashlib

import numpy as np

from oracle import oracle as O


def cells_of(boards):
    b = np.asarray(boards, dtype=np.uint64)
    sh = (np.arange(16, dtype=np.uint64) * np.uint64(4))
    return ((b[:, None] >> sh[None, :]) & np.uint64(15)).astype(np.uint8)


def test_row_table_known_answers():
    out4, score, mt = O.row_table()
    idx = np.arange(65536)
    packed = out4[:, 0].astype(np.int64) | out4[:, 1].astype(np.int64) << 4 | \
        out4[:, 2].astype(np.int64) << 8 | out4[:, 3].astype(np.int64) << 12
    # rows that create exponent 16 cannot be compared through 4-bit packing
    assert int(((packed != idx) | (mt == 16)).sum()) >= 21210
    assert int(score.sum()) == 100660224
    assert int((mt == 16).sum()) == 767
    assert int(score.max()) == 131072
    small = ((idx & 15) < 15) & (((idx >> 4) & 15) < 15) & (((idx >> 8) & 15) < 15) & ((idx >> 12) < 15)
    assert int(score[small].sum()) == 44234100
    h = hashlib.sha256()
    for c0 in range(16):
        for c1 in range(16):
            for c2 in range(16):
                for c3 in range(16):
                    i = c0 | c1 << 4 | c2 << 8 | c3 << 12
                    h.update(bytes(out4[i].tolist()) + int(score[i]).to_bytes(4, "little") + bytes([int(mt[i])]))
    assert h.hexdigest() == "e0ace12e5d81807e91d24c724545abc556b9f49b9ad903f739be3f751e445688"


def test_row_examples():
    out4, score, mt = O.row_table()
    row = lambda a: a[0] | a[1] << 4 | a[2] << 8 | a[3] << 12
    assert out4[row([1, 1, 1, 1])].tolist() == [2, 2, 0, 0] and score[row([1, 1, 1, 1])] == 8
    assert out4[row([2, 1, 1, 2])].tolist() == [2, 2, 2, 0] and score[row([2, 1, 1, 2])] == 4
    assert out4[row([3, 3, 0, 3])].tolist() == [4, 3, 0, 0] and score[row([3, 3, 0, 3])] == 16
    assert out4[row([1, 0, 0, 1])].tolist() == [2, 0, 0, 0] and score[row([1, 0, 0, 1])] == 4


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox_raw([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox_raw([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox_raw([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_potentials_match_reference(golden):
    g = golden("potentials")
    got = O.potentials_batch(g["board"])
    np.testing.assert_array_equal(got, g["values"])


def test_survey_board_values():
    B = O.pack_grid([[1, 2, 3, 4], [8, 7, 6, 5], [9, 10, 11, 12], [0, 0, 0, 13]])
    Cb = O.pack_grid([[5, 5, 1, 0], [0, 5, 0, 0], [0, 0, 0, 0], [0, 0, 0, 5]])
    D = O.pack_grid([[0, 5, 1, 0], [0, 5, 0, 0], [0, 0, 0, 0], [0, 0, 0, 5]])
    E = 0
    p = O.potentials_batch(np.array([B, Cb, D, E], dtype=np.uint64))
    assert p[0, :4].tolist() == [30, 3, -42, 13]
    assert p[1, 0] == 6
    assert p[2, 0] == 1 and p[2, 3] == 5
    assert p[3, :4].tolist() == [0, 16, 0, 0]


def test_env_step_matches_reference(golden):
    g = golden("env_step")
    out, info = O.step_batch(g["board"], g["action"], replay=g["draw"])
    ovf = (g["out_cells"] > 15).any(axis=1)
    assert ovf.sum() > 0
    np.testing.assert_array_equal(info["overflow"].astype(bool), ovf)
    np.testing.assert_array_equal(cells_of(out)[~ovf], g["out_cells"][~ovf])
    np.testing.assert_array_equal(cells_of(out)[ovf], np.minimum(g["out_cells"][ovf], 15))
    for k in ("points", "done", "invalid", "mono_before", "mono_after", "empt_before", "empt_after",
              "max_tile_created", "max_exp_before", "max_exp_after", "smooth_before", "smooth_after",
              "corner_before", "corner_after", "legal_before", "legal_after"):
        np.testing.assert_array_equal(info[k], g[k], err_msg=k)
    np.testing.assert_array_equal(info["smooth_after"] - info["smooth_before"], g["smooth_delta"])
    np.testing.assert_array_equal(info["corner_after"] - info["corner_before"], g["corner_delta"])


def test_expand4_matches_reference(golden):
    g = golden("env_step")
    boards = g["board"][0::4]
    succ, points, max_tile, legal = O.expand4_batch(boards)
    pre = g["pre_spawn_cells"].reshape(-1, 4, 16)
    ok = (pre <= 15).all(axis=2)
    got = cells_of(succ.reshape(-1)).reshape(-1, 4, 16)
    np.testing.assert_array_equal(got[ok], pre[ok])
    np.testing.assert_array_equal(points.reshape(-1), g["points"])
    np.testing.assert_array_equal(max_tile.reshape(-1), g["max_tile_created"])
    np.testing.assert_array_equal(legal, g["legal_before"][0::4])
    # legal <=> the move changes the board
    changed = (succ != boards[:, None])
    np.testing.assert_array_equal(changed, ((legal[:, None] >> np.arange(4)) & 1).astype(bool))


def test_best_game_replay(golden):
    g = golden("best_game")
    before, after, action, points = g["before"], g["after"], g["action"], g["points"]
    succ, pts, _, legal = O.expand4_batch(before)
    idx = np.arange(len(action))
    moved = succ[idx, action]
    assert ((legal >> action) & 1).all()
    np.testing.assert_array_equal(pts[idx, action], points)
    assert int(points.sum()) == int(g["score"]) == 24792
    mc, ac = cells_of(moved).astype(int), cells_of(after).astype(int)
    diff = ac != mc
    assert (diff.sum(axis=1) == 1).all()                     # exactly one spawned tile
    assert (mc[diff] == 0).all() and np.isin(ac[diff], (1, 2)).all()
    assert int((ac[diff] == 2).sum()) == 133
    np.testing.assert_array_equal(before[1:], after[:-1])   # 1248 chained links
    assert np.bincount(action, minlength=4).tolist() == [210, 249, 5, 785]


def test_rollout_boards_follow_philox_draws(golden):
    g = golden("rollout")
    seed = int(g["seed"])
    for env in range(len(g["ep_len"])):
        sel = g["env"] == env
        boards, acts, res = g["board"][sel], g["action"][sel], g["result"][sel]
        assert boards[0] == O.reset_batch(1, seed=seed, env0=env, ctr=0)[0]
        for t in range(len(acts)):
            out, info = O.step_batch(boards[t:t + 1], acts[t:t + 1], seed=seed, env0=env, ctr=1 + t)
            assert out[0] == res[t]
            assert info["legal_before"][0] == g["legal"][sel][t]
            assert info["points"][0] == g["points"][sel][t]
            assert bool(info["done"][0]) == bool(g["done"][sel][t])
            if t + 1 < len(acts):
                assert boards[t + 1] == res[t]


def test_encode_matches_reference(golden):
    g = golden("model_best")
    np.testing.assert_array_equal(O.encode_batch(g["board"]), g["inputs"])


def _rollout_as_tb(g):
    """Lay the golden episodes out as time-major [T,B] columns, one game per column."""
    n_env = len(g["ep_len"])
    T = int(g["ep_len"].max())
    z = lambda dt: np.zeros((T, n_env), dtype=dt)
    a = dict(points=z(np.int32), mono_b=z(np.uint8), mono_a=z(np.uint8), empt_b=z(np.uint8),
             empt_a=z(np.uint8), done=z(np.uint8), valid=z(np.uint8), value=z(np.float32))
    order = []
    for env in range(n_env):
        sel = np.nonzero(g["env"] == env)[0]
        t = g["t"][sel]
        a["points"][t, env] = g["points"][sel]
        # undo the caller-side terminal fix-up; the oracle re-applies it from `done`
        a["mono_b"][t, env] = g["mono_before"][sel]
        a["mono_a"][t, env] = g["mono_after"][sel]
        a["empt_b"][t, env] = g["empt_before"][sel]
        a["empt_a"][t, env] = g["empt_after"][sel]
        a["done"][t, env] = g["done"][sel]
        a["valid"][t, env] = 1
        a["value"][t, env] = g["value"][sel]
        order.append((t, env, sel))
    return a, order


def test_rtg_advantage_matches_reference(golden):
    g = golden("rollout")
    adv = golden("advantage")
    a, order = _rollout_as_tb(g)
    for name in ("readme", "warm"):
        gamma, wp, wm, we, beta, step, mu, m2 = adv[name + "__cfg"].tolist()
        r = O.rtg_adv(a["points"], a["mono_b"], a["mono_a"], a["empt_b"], a["empt_a"], a["done"], a["valid"],
                      a["value"], gamma, wp, wm, we, beta, int(step), mu, m2)
        for key, ref in (("reward", "reward"), ("g_raw", "g_raw"), ("g_norm", "g_norm"), ("adv", "adv")):
            got = np.concatenate([r[key][t, env] for t, env, _ in order])
            want = np.concatenate([adv[name + "__" + ref][sel] for _, _, sel in order])
            np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-6, atol=1e-6, err_msg=name + key)
        np.testing.assert_allclose([r["rtg_mu"], r["rtg_m2"]], adv[name + "__moments_out"], rtol=1e-12)


def test_float_potentials_match_reference(golden):
    g = golden("potentials_ext")
    got = O.potentials_ext_batch(g["before"], g["after"])
    np.testing.assert_array_equal(got, g["values"])          # doubles, bit-identical
