"""GPU parity of the environment kernels, called through the C ABI (ctypes), against the
oracle, the reference-generated fixtures and size-independent properties.  Bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

SH_KEYS = ("mono_before", "mono_after", "empt_before", "empt_after", "max_tile_created", "max_exp_before",
           "max_exp_after", "smooth_before", "smooth_after", "corner_before", "corner_after")


@pytest.fixture(scope="module")
def env():
    from g2048 import env as e
    e.init(0)
    return e


def dev_boards(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()


def host_u64(t):
    return t.cpu().numpy().view(np.uint64)


def random_boards(n, seed, hi=12, p_empty=0.3):
    g = torch.Generator().manual_seed(seed)
    e = torch.randint(1, hi, (n, 16), generator=g, dtype=torch.int64)
    e[torch.rand((n, 16), generator=g) < p_empty] = 0
    sh = torch.arange(16, dtype=torch.int64) * 4
    return (e << sh).sum(dim=1).numpy().view(np.uint64)


def check_step(env, boards, actions, draws, n_expected=None):
    r = env.step(dev_boards(boards), torch.from_numpy(actions).cuda(),
                 replay=torch.from_numpy(draws.view(np.int32)).cuda())
    want_b, want = O.step_batch(boards, actions, replay=draws)
    fl = r["flags"].cpu().numpy()
    np.testing.assert_array_equal(r["points"].cpu().numpy(), want["points"])
    np.testing.assert_array_equal((fl >> 5) & 1, want["invalid"])
    np.testing.assert_array_equal((fl >> 6) & 1, want["overflow"])
    # A transition flagged OVERFLOW created exponent 16, which the reference keeps as an int
    # but a nibble cannot hold: only points / invalid / overflow are defined for it.
    ok = want["overflow"] == 0
    np.testing.assert_array_equal(host_u64(r["boards"])[ok], want_b[ok])
    np.testing.assert_array_equal((fl & 0x0F)[ok], want["legal_after"][ok])
    np.testing.assert_array_equal(((fl >> 4) & 1)[ok], want["done"][ok])
    sh = env.decode_shaping(r["shaping"].cpu().numpy())
    for k in SH_KEYS:
        np.testing.assert_array_equal(sh[k][ok], want[k][ok], err_msg=k)
    return r, want


def test_row_table_matches_oracle(env):
    both = env.lut(0).cpu().numpy().view(np.uint32)
    lut, mv_slots = both[:65536], both[65536:]
    rows = np.arange(65536)
    mv = mv_slots[rows ^ ((rows >> 8) & 31)]          # the move table is stored bank-hashed
    out4, score, mt = O.row_table()
    # move table of the 4-move expansion: exact for rows whose cells are all <= 11
    idx = np.arange(65536)
    small = ((idx & 15) <= 11) & (((idx >> 4) & 15) <= 11) & (((idx >> 8) & 15) <= 11) & ((idx >> 12) <= 11)
    assert small.sum() == 12 ** 4
    np.testing.assert_array_equal(np.stack([(mv >> (4 * k)) & 15 for k in range(4)], axis=1)[small], out4[small])
    np.testing.assert_array_equal((((mv >> 16) & 0xFFF) * 4)[small], score[small])
    np.testing.assert_array_equal((mv >> 28)[small], mt[small])
    res = np.stack([(lut >> (4 * k)) & 15 for k in range(4)], axis=1)
    np.testing.assert_array_equal(res, np.minimum(out4, 15))
    c1, c2 = (lut >> 16) & 15, (lut >> 20) & 15
    sc = np.where(c1 > 0, 2 << c1.astype(np.int64), 0) + np.where(c2 > 0, 2 << c2.astype(np.int64), 0)
    np.testing.assert_array_equal(sc, score)
    np.testing.assert_array_equal(np.maximum(np.where(c1 > 0, c1 + 1, 0), np.where(c2 > 0, c2 + 1, 0)), mt)
    assert int(sc.sum()) == 100660224


def test_step_matches_reference_fixture(env, golden):
    g = golden("env_step")
    r = env.step(dev_boards(g["board"]), torch.from_numpy(g["action"]).cuda(),
                 replay=torch.from_numpy(g["draw"].view(np.int32)).cuda())
    got = host_u64(r["boards"])
    sh16 = np.arange(16, dtype=np.uint64) * np.uint64(4)
    cells = ((got[:, None] >> sh16) & np.uint64(15)).astype(np.uint8)
    ovf = (g["out_cells"] > 15).any(axis=1)     # exponent 16 created: flagged, otherwise undefined
    ok = ~ovf
    assert ovf.sum() > 0
    np.testing.assert_array_equal(cells[ok], g["out_cells"][ok])
    np.testing.assert_array_equal(r["points"].cpu().numpy(), g["points"])
    fl = r["flags"].cpu().numpy()
    np.testing.assert_array_equal((fl & 0x0F)[ok], g["legal_after"][ok])
    np.testing.assert_array_equal(((fl >> 4) & 1)[ok], g["done"][ok])
    np.testing.assert_array_equal((fl >> 5) & 1, g["invalid"])
    np.testing.assert_array_equal(((fl >> 6) & 1).astype(bool), ovf)
    sh = env.decode_shaping(r["shaping"].cpu().numpy())
    for k in SH_KEYS:
        np.testing.assert_array_equal(sh[k][ok], g[k][ok], err_msg=k)
    np.testing.assert_array_equal((sh["smooth_after"] - sh["smooth_before"])[ok], g["smooth_delta"][ok])
    np.testing.assert_array_equal((sh["corner_after"] - sh["corner_before"])[ok], g["corner_delta"][ok])


def test_step_without_shaping_small_and_large(env):
    for n in (1000, 1 << 18):
        boards = random_boards(n, 5)
        actions = (np.arange(n) % 4).astype(np.uint8)
        r = env.step(dev_boards(boards), torch.from_numpy(actions).cuda(), seed=9, env0=17, ctr=3, shaping=False)
        want_b, want = O.step_batch(boards, actions, seed=9, env0=17, ctr=3)
        assert want["overflow"].sum() == 0
        np.testing.assert_array_equal(host_u64(r["boards"]), want_b)
        np.testing.assert_array_equal(r["points"].cpu().numpy(), want["points"])
        assert r["shaping"] is None


@pytest.mark.parametrize("n", [1, 31, 1000, (1 << 17) + 77])
def test_step_random_boards_replay(env, n):
    rng = np.random.default_rng(n)
    boards = random_boards(n, n)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    draws = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    check_step(env, boards, actions, draws)


def test_step_high_exponents_and_overflow(env):
    n = 1 << 17
    rng = np.random.default_rng(3)
    boards = random_boards(n, 3, hi=16, p_empty=0.2)
    actions = rng.integers(0, 4, n).astype(np.uint8)
    draws = rng.integers(0, 2**32, (n, 2), dtype=np.uint64).astype(np.uint32)
    r, want = check_step(env, boards, actions, draws)
    assert want["overflow"].sum() > 0


def test_step_all_rows_exhaustive(env):
    """Every row with exponents <= 14 tiled into the 4 rows of a board, and its transpose."""
    vals = np.arange(15, dtype=np.uint64)
    rows = (vals[:, None, None, None] | vals[None, :, None, None] << np.uint64(4) |
            vals[None, None, :, None] << np.uint64(8) | vals[None, None, None, :] << np.uint64(12)).reshape(-1)
    assert rows.size == 50625
    horiz = rows | rows << np.uint64(16) | rows << np.uint64(32) | rows << np.uint64(48)
    c = [(rows >> np.uint64(4 * k)) & np.uint64(15) for k in range(4)]
    col = lambda v: v | v << np.uint64(4) | v << np.uint64(8) | v << np.uint64(12)
    vert = col(c[0]) | col(c[1]) << np.uint64(16) | col(c[2]) << np.uint64(32) | col(c[3]) << np.uint64(48)
    boards = np.concatenate([horiz, vert])
    rng = np.random.default_rng(1)
    for a in range(4):
        actions = np.full(boards.size, a, dtype=np.uint8)
        draws = rng.integers(0, 2**32, (boards.size, 2), dtype=np.uint64).astype(np.uint32)
        check_step(env, boards, actions, draws)


def test_dense_step_tables_match_oracle(env):
    """The dense tables of the fused step kernel (M: u64 per row with cells <= 11, S: u16 per row with cells
    <= 12) against the oracle's row table and a direct count of the per-line potentials."""
    raw = env.lut(0).cpu().numpy()
    m = raw[2 * 65536 * 4:2 * 65536 * 4 + 167040].view(np.uint64)
    s16 = raw[2 * 65536 * 4 + 167040:].view(np.uint16)
    out4, score, mt = O.row_table()

    def stats(rows):
        c = np.stack([(rows >> (4 * k)) & 15 for k in range(4)], axis=1).astype(np.int64)
        both = (c[:, :-1] > 0) & (c[:, 1:] > 0)
        ge = (both & (c[:, :-1] >= c[:, 1:])).sum(1)
        le = (both & (c[:, :-1] <= c[:, 1:])).sum(1)
        sm = (both * np.abs(c[:, :-1] - c[:, 1:])).sum(1)
        return ge | le << 4 | sm << 8

    d = np.arange(12)
    c0, c1, c2, c3 = np.meshgrid(d, d, d, d, indexing="ij")
    rows = (c0 | c1 << 4 | c2 << 8 | c3 << 12).reshape(-1)
    slot = (c0 + 12 * c1 + 145 * (c2 + 12 * c3)).reshape(-1)
    e = m[slot]
    lo, hi = (e & np.uint64(0xFFFFFFFF)).astype(np.int64), (e >> np.uint64(32)).astype(np.int64)
    res = (out4[rows].astype(np.int64) << (4 * np.arange(4))).sum(1)
    np.testing.assert_array_equal(lo & 0xFFFF, res)
    np.testing.assert_array_equal(((lo >> 16) & 0xFFF) * 4, score[rows])
    np.testing.assert_array_equal(lo >> 28, mt[rows])
    np.testing.assert_array_equal(hi & 0xFFFF, stats(rows))
    np.testing.assert_array_equal(hi >> 16, stats(res))
    used = np.zeros(m.size, bool)
    used[slot] = True
    assert not m[~used].any()
    d = np.arange(13)
    c0, c1, c2, c3 = np.meshgrid(d, d, d, d, indexing="ij")
    rows = (c0 | c1 << 4 | c2 << 8 | c3 << 12).reshape(-1)
    slot = (c0 + 13 * c1 + 169 * c2 + 2197 * c3).reshape(-1)
    np.testing.assert_array_equal(s16[slot].astype(np.int64), stats(rows))


def test_step_all_small_rows_through_the_dense_kernel(env):
    """Every row with exponents <= 11 tiled into the 4 rows / 4 columns of a board, padded with random boards
    past the staging threshold so that the dense-table kernel (not the direct one) sees all of them."""
    vals = np.arange(12, dtype=np.uint64)
    rows = (vals[:, None, None, None] | vals[None, :, None, None] << np.uint64(4) |
            vals[None, None, :, None] << np.uint64(8) | vals[None, None, None, :] << np.uint64(12)).reshape(-1)
    horiz = rows | rows << np.uint64(16) | rows << np.uint64(32) | rows << np.uint64(48)
    c = [(rows >> np.uint64(4 * k)) & np.uint64(15) for k in range(4)]
    col = lambda v: v | v << np.uint64(4) | v << np.uint64(8) | v << np.uint64(12)
    vert = col(c[0]) | col(c[1]) << np.uint64(16) | col(c[2]) << np.uint64(32) | col(c[3]) << np.uint64(48)
    mixed = rows | np.roll(rows, 1) << np.uint64(16) | np.roll(rows, 7) << np.uint64(32) | np.roll(rows, 1234) << np.uint64(48)
    base = np.concatenate([horiz, vert, mixed, random_boards(1 << 16, 77, hi=14)])
    boards = np.concatenate([base] * 4)
    actions = np.repeat(np.arange(4, dtype=np.uint8), base.size)
    assert boards.size >= 1 << 17
    rng = np.random.default_rng(5)
    draws = rng.integers(0, 2**32, (boards.size, 2), dtype=np.uint64).astype(np.uint32)
    draws[::7, 1] = 3865470567 - (np.arange(draws[::7].shape[0]) % 2)         # both sides of the 4-tile threshold
    draws[::11, 0] = np.uint32(0xFFFFFFFF)
    draws[::13, 0] = 0
    check_step(env, boards, actions, draws)


def test_philox_path_matches_oracle(env):
    n = 1 << 17
    boards = random_boards(n, 11)
    actions = (np.arange(n) % 4).astype(np.uint8)
    r = env.step(dev_boards(boards), torch.from_numpy(actions).cuda(), seed=0xDEADBEEFCAFE, env0=(1 << 33) + 5, ctr=7)
    want_b, _ = O.step_batch(boards, actions, seed=0xDEADBEEFCAFE, env0=(1 << 33) + 5, ctr=7)
    np.testing.assert_array_equal(host_u64(r["boards"]), want_b)


def test_reset_matches_oracle(env):
    for n in (1, 1000, 100000):
        got = host_u64(env.reset(n, device=0, seed=2048, env0=3, ctr=0))
        np.testing.assert_array_equal(got, O.reset_batch(n, seed=2048, env0=3, ctr=0))
    rng = np.random.default_rng(0)
    draws = rng.integers(0, 2**32, (5000, 4), dtype=np.uint64).astype(np.uint32)
    got = host_u64(env.reset(5000, device=0, replay=torch.from_numpy(draws.view(np.int32)).cuda()))
    np.testing.assert_array_equal(got, O.reset_batch(5000, replay=draws))
    cells = ((got[:, None] >> (np.arange(16, dtype=np.uint64) * np.uint64(4))) & np.uint64(15))
    assert ((cells > 0).sum(axis=1) == 2).all()


@pytest.mark.parametrize("n", [7, 4097, 1 << 18])
def test_expand4_matches_oracle(env, n):
    boards = random_boards(n, n + 1, hi=15, p_empty=0.25)
    r = env.expand4(dev_boards(boards), want_max_tile=True)
    succ, points, mt, legal = O.expand4_batch(boards)
    np.testing.assert_array_equal(host_u64(r["succ"]), succ)
    np.testing.assert_array_equal(r["points"].cpu().numpy(), points)
    np.testing.assert_array_equal(r["max_tile"].cpu().numpy(), mt)
    np.testing.assert_array_equal(r["legal"].cpu().numpy(), legal)


def test_expand4_best_game_replay(env, golden):
    g = golden("best_game")
    r = env.expand4(dev_boards(g["before"]))
    idx = np.arange(len(g["action"]))
    moved = host_u64(r["succ"])[idx, g["action"]]
    np.testing.assert_array_equal(r["points"].cpu().numpy()[idx, g["action"]], g["points"])
    sh16 = np.arange(16, dtype=np.uint64) * np.uint64(4)
    mc = ((moved[:, None] >> sh16) & np.uint64(15)).astype(int)
    ac = ((g["after"][:, None] >> sh16) & np.uint64(15)).astype(int)
    diff = mc != ac
    assert (diff.sum(axis=1) == 1).all() and (mc[diff] == 0).all() and np.isin(ac[diff], (1, 2)).all()


def test_potentials_match_reference_fixture(env, golden):
    g = golden("potentials")
    got = env.potentials(dev_boards(g["board"])).cpu().numpy()
    np.testing.assert_array_equal(got, g["values"])


def test_encode_matches_reference_fixture(env, golden):
    g = golden("model_best")
    np.testing.assert_array_equal(env.encode(dev_boards(g["board"])).cpu().numpy(), g["inputs"])


def test_symmetry_properties_full_size(env):
    """Size-independent properties at the C2 size (2^20 boards): a move commutes with the
    board's transpose (UP<->LEFT, DOWN<->RIGHT) and legality <=> the board changes."""
    n = 1 << 20
    boards = random_boards(n, 2048)
    r = env.expand4(dev_boards(boards))
    succ, legal = host_u64(r["succ"]), r["legal"].cpu().numpy()
    changed = succ != boards[:, None]
    np.testing.assert_array_equal(changed, ((legal[:, None] >> np.arange(4)) & 1).astype(bool))

    def transpose(b):
        out = np.zeros_like(b)
        for rr in range(4):
            for cc in range(4):
                out |= ((b >> np.uint64(4 * (4 * rr + cc))) & np.uint64(15)) << np.uint64(4 * (4 * cc + rr))
        return out

    rt = env.expand4(dev_boards(transpose(boards)))
    st = host_u64(rt["succ"])
    np.testing.assert_array_equal(transpose(st[:, 2]), succ[:, 0])   # LEFT of transpose == UP
    np.testing.assert_array_equal(transpose(st[:, 3]), succ[:, 1])   # RIGHT of transpose == DOWN
    np.testing.assert_array_equal(rt["points"].cpu().numpy()[:, [2, 3, 0, 1]], r["points"].cpu().numpy())
    # tile-value sum is conserved by a move (merging 2^e + 2^e = 2^(e+1))
    val = lambda b: sum((np.uint64(1) << ((b >> np.uint64(4 * k)) & np.uint64(15))) * (((b >> np.uint64(4 * k)) & np.uint64(15)) > 0) for k in range(16))
    for d in range(4):
        np.testing.assert_array_equal(val(succ[:, d]), val(boards))


def test_facade_matches_reference_semantics(env):
    from g2048.env import Direction, Game2048
    g = Game2048([[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    assert set(g.current_valid_directions()) == {Direction.DOWN, Direction.RIGHT}
    grid, pts, done, info = g.step(Direction.UP)
    assert info["invalid_move"] and pts == 0 and not done and grid[0][0] == 1
    with pytest.raises(ValueError):
        g.move(Direction.UP)
    chk = Game2048([[1, 2, 1, 2], [2, 1, 2, 1], [1, 2, 1, 2], [2, 1, 2, 1]])
    assert not chk.has_next_step()
    grid, pts, done, info = chk.step(Direction.UP)
    assert done and info["invalid_move"]
    B = [[1, 2, 3, 4], [8, 7, 6, 5], [9, 10, 11, 12], [0, 0, 0, 13]]
    assert Game2048.monotonicity(B) == 30 and Game2048.smoothness_score(B) == -42.0
    assert Game2048.corner_bonus(B) == 13.0 and Game2048.emptiness(B) == 3
    assert Game2048.simulate_move([[1, 1, 1, 1]] + [[0] * 4] * 3, Direction.LEFT) == \
        ([[2, 2, 0, 0]] + [[0] * 4] * 3, 8, 2)
    q = [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 0]]
    assert Game2048.mirror_grid(q, "horizontal")[0] == [4, 3, 2, 1] and Game2048.mirror_grid(q, "vertical")[0] == [13, 14, 15, 0]
    assert Game2048.rotate_grid(q, 90)[0] == [13, 9, 5, 1] and Game2048.rotate_grid(q, "south")[3] == [4, 3, 2, 1]
    assert Game2048.rotate_grid(q, 0) == q and Game2048.calculate_grid_score([[1, 2, 0, 0]] + [[0] * 4] * 3) == 6
    with pytest.raises(ValueError):
        Game2048.mirror_grid(q, "diagonal")
    g2 = Game2048(seed=2048, env_id=0)
    grid = g2.reset()
    assert sum(1 for r in grid for c in r if c) == 2
    assert g2.to_model_format().shape == (48,)
    new, pts, done, info = g2.step(g2.current_valid_directions()[0])
    assert not info["invalid_move"] and "monotonicity_before" in info


def test_augment_matches_reference_and_commutes_with_the_env(env, golden):
    """Boards against the reference's mirror_grid / rotate_grid (fixture); the direction remap against the
    environment itself: moving the transformed board in the remapped direction gives the transformed successor."""
    g = golden("augment")
    boards = g["board"]
    n = len(boards)
    ex = env.expand4(dev_boards(boards))
    succ, legal = host_u64(ex["succ"]), ex["legal"]
    rng = np.random.default_rng(0)
    action = rng.integers(0, 4, n).astype(np.uint8)
    logp = torch.randn((n, 4), generator=torch.Generator().manual_seed(1)).cuda()
    after = succ[np.arange(n), action]
    for op in range(5):
        r = env.augment(dev_boards(boards), dev_boards(after), torch.from_numpy(action).cuda(), legal, logp,
                        torch.full((n,), op, dtype=torch.uint8, device="cuda"))
        np.testing.assert_array_equal(host_u64(r["before"]), g["transformed"][:, op])
        ex2 = env.expand4(r["before"])
        np.testing.assert_array_equal(r["legal"].cpu().numpy(), ex2["legal"].cpu().numpy())
        a2 = r["action"].cpu().numpy()
        np.testing.assert_array_equal(host_u64(ex2["succ"])[np.arange(n), a2], host_u64(r["after"]))
        got = torch.gather(r["logp"], 1, r["action"].long()[:, None]).squeeze(1)
        want = torch.gather(logp, 1, torch.from_numpy(action).cuda().long()[:, None]).squeeze(1)
        assert torch.equal(got, want)
        assert torch.equal(r["logp"].sort(dim=1).values, logp.sort(dim=1).values)


def test_host_stepper_matches_device_step(env):
    n = (1 << 20) + 12345
    boards = random_boards(n, 77)
    actions = (np.arange(n) % 4).astype(np.uint8)
    hb = torch.from_numpy(boards.view(np.int64).copy()).pin_memory()
    ha = torch.from_numpy(actions).pin_memory()
    h_out = dict(boards=torch.empty(n, dtype=torch.int64).pin_memory(), points=torch.empty(n, dtype=torch.int32).pin_memory(),
                 flags=torch.empty(n, dtype=torch.uint8).pin_memory(), shaping=torch.empty(n, dtype=torch.int64).pin_memory())
    st = env.HostStepper(n, device=0)
    st.step(hb, ha, h_out, seed=3, env0=100, ctr=9)
    torch.cuda.synchronize()
    r = env.step(hb.cuda(), ha.cuda(), seed=3, env0=100, ctr=9)
    for k in ("boards", "points", "flags", "shaping"):
        assert torch.equal(h_out[k], r[k].cpu()), k
    with pytest.raises(ValueError):
        st.step(hb.cuda(), ha, h_out)


def test_float_potentials_match_reference_fixture_and_oracle(env, golden):
    g = golden("potentials_ext")
    got = env.potentials_ext(dev_boards(g["before"]), dev_boards(g["after"])).cpu().numpy()
    np.testing.assert_array_equal(got, g["values"])          # float64, bit-identical to the Python floats
    boards = random_boards(200000, 5, hi=16, p_empty=0.2)
    succ = host_u64(env.expand4(dev_boards(boards))["succ"])
    after = succ[np.arange(len(boards)), np.arange(len(boards)) % 4]
    np.testing.assert_array_equal(env.potentials_ext(dev_boards(boards), dev_boards(after)).cpu().numpy(),
                                  O.potentials_ext_batch(boards, after))


def test_facade_info_dict_is_complete(env):
    from g2048.env import Direction, Game2048
    g = Game2048([[1, 2, 3, 4], [8, 7, 6, 5], [9, 10, 11, 12], [0, 0, 0, 13]], seed=1)
    _, _, _, info = g.step(Direction.DOWN)
    assert set(info) == {"invalid_move", "smoothness_delta", "max_tile_created", "max_exponent_before",
                         "max_exponent_after", "corner_delta", "adjacency_delta", "chain_delta",
                         "monotonicity_before", "monotonicity_after", "emptiness_before", "emptiness_after",
                         "topological_delta", "topological_anchor"}
    assert info["topological_anchor"] == (3, 3) and info["monotonicity_before"] == 30


def test_dense_step_chain_back_to_back_and_in_place(env):
    """Six consecutive g2048_step launches of the persistent dense kernel, each consuming the boards the previous
    one wrote, issued back to back without a host sync (programmatic dependent launch: a launch stages its table
    while its predecessor drains and must not read the boards before that one completed), out-of-place through
    two buffers and in place.  Every step against the oracle."""
    n = (1 << 17) + 5
    boards = random_boards(n, 21)
    rng = np.random.default_rng(21)
    acts = [rng.integers(0, 4, n).astype(np.uint8) for _ in range(6)]
    want = [boards]
    for t, a in enumerate(acts):
        want.append(O.step_batch(want[-1], a, seed=77, env0=3, ctr=10 + t)[0])
    d_acts = [torch.from_numpy(a).cuda() for a in acts]
    mk = lambda: dict(boards=torch.empty(n, dtype=torch.int64, device="cuda"), points=torch.empty(n, dtype=torch.int32, device="cuda"),
                      flags=torch.empty(n, dtype=torch.uint8, device="cuda"), shaping=torch.empty(n, dtype=torch.int64, device="cuda"))
    bufs = [mk(), mk()]
    cur = dev_boards(boards)
    snaps = []
    torch.cuda.synchronize()
    for t in range(6):
        out = bufs[t % 2]
        env.step(cur, d_acts[t], seed=77, env0=3, ctr=10 + t, out=out)
        snaps.append(out["boards"].clone())         # stream-ordered copy between the launches
        cur = out["boards"]
    torch.cuda.synchronize()
    for t in range(6):
        np.testing.assert_array_equal(host_u64(snaps[t]), want[t + 1], err_msg=f"step {t}")
    # in place: boards_out == boards_in
    inplace = mk()
    inplace["boards"].copy_(dev_boards(boards))
    for t in range(6):
        env.step(inplace["boards"], d_acts[t], seed=77, env0=3, ctr=10 + t, out=inplace)
    np.testing.assert_array_equal(host_u64(inplace["boards"]), want[6])


def test_c2_full_size_step_matches_oracle_and_the_direct_kernel(env):
    """BASELINE config C2 at its full size (2^20 boards x 4 moves = 4 194 304 transitions, Philox draws): the
    persistent dense-table kernel against the oracle on every output, and against the direct kernel (row table
    through L2, bit-parallel potentials: an independent implementation) run over the same transitions in chunks
    below the staging threshold."""
    nb = 1 << 20
    b = random_boards(nb, 2048)
    boards = np.tile(b, 4)
    actions = np.repeat(np.arange(4, dtype=np.uint8), nb)
    d_b, d_a = dev_boards(boards), torch.from_numpy(actions).cuda()
    r = env.step(d_b, d_a, seed=2048, env0=0, ctr=1)
    want_b, want = O.step_batch(boards, actions, seed=2048, env0=0, ctr=1)
    assert want["overflow"].sum() == 0
    np.testing.assert_array_equal(host_u64(r["boards"]), want_b)
    np.testing.assert_array_equal(r["points"].cpu().numpy(), want["points"])
    fl = r["flags"].cpu().numpy()
    np.testing.assert_array_equal(fl & 0x0F, want["legal_after"])
    np.testing.assert_array_equal((fl >> 4) & 1, want["done"])
    np.testing.assert_array_equal((fl >> 5) & 1, want["invalid"])
    sh = env.decode_shaping(r["shaping"].cpu().numpy())
    for k in SH_KEYS:
        np.testing.assert_array_equal(sh[k], want[k], err_msg=k)
    chunk = 1 << 16
    for lo in range(0, boards.size, chunk):
        q = env.step(d_b[lo:lo + chunk], d_a[lo:lo + chunk], seed=2048, env0=lo, ctr=1)
        for k in ("boards", "points", "flags", "shaping"):
            assert torch.equal(q[k], r[k][lo:lo + chunk]), (k, lo)
    assert int(r["points"].sum()) == int(want["points"].sum())


@pytest.mark.parametrize("n,shaping", [(1, True), (1000, True), (1000, False), (40000, True), (70001, True), (70001, False), (1 << 20, True)])
def test_step4_equals_step_on_every_board_and_move(n, shaping):
    """g2048_step4 (the C2 form: one thread plays the four moves of a board and shares their common work) against g2048_step on
    the 4 n (board, move) pairs with the same env ids -- every output bit for bit -- and, on a subset, against the oracle; both
    the direct (small n) and the dense-table kernel (4 n >= 2^17), boards with 4096+ tiles (the L2 path) included."""
    from g2048 import env
    rng = np.random.default_rng(n)
    e = rng.integers(1, 12, (n, 16))
    e[rng.random((n, 16)) < 0.3] = 0
    big = rng.random(n) < 0.05                                   # a few boards with a 4096..16384 tile
    e[big, rng.integers(0, 16, int(big.sum()))] = rng.integers(12, 15, int(big.sum()))
    b = (e.astype(np.uint64) << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(axis=1).astype(np.uint64)
    boards = torch.from_numpy(b.view(np.int64)).cuda()
    got = env.step4(boards, seed=77, env0=1000, ctr=3, shaping=shaping)
    want = env.step(boards.repeat_interleave(4), torch.arange(4, dtype=torch.uint8, device="cuda").repeat(n), seed=77, env0=1000, ctr=3,
                    shaping=shaping)
    torch.cuda.synchronize()
    for k in ("boards", "points", "flags") + (("shaping",) if shaping else ()):
        assert torch.equal(got[k].reshape(-1), want[k]), k
    m = min(n, 3000)
    ob, info = O.step_batch(np.repeat(b[:m], 4), np.tile(np.arange(4, dtype=np.uint8), m), seed=77, env0=1000, ctr=3)
    np.testing.assert_array_equal(got["boards"][:m].reshape(-1).cpu().numpy().view(np.uint64), ob)
    np.testing.assert_array_equal(got["points"][:m].reshape(-1).cpu().numpy(), info["points"])


def test_step4_replay_and_host_stepper():
    from g2048 import env
    n = 50000
    rng = np.random.default_rng(5)
    e = rng.integers(1, 12, (n, 16))
    e[rng.random((n, 16)) < 0.3] = 0
    b = (e.astype(np.uint64) << (np.arange(16, dtype=np.uint64) * np.uint64(4))).sum(axis=1).astype(np.uint64)
    boards = torch.from_numpy(b.view(np.int64)).cuda()
    draws = torch.from_numpy(rng.integers(0, 2**32, (n, 4, 2), dtype=np.uint64).astype(np.uint32).view(np.int32)).cuda()
    got = env.step4(boards, replay=draws)
    want = env.step(boards.repeat_interleave(4), torch.arange(4, dtype=torch.uint8, device="cuda").repeat(n), replay=draws.reshape(-1, 2))
    torch.cuda.synchronize()
    for k in ("boards", "points", "flags", "shaping"):
        assert torch.equal(got[k].reshape(-1), want[k]), k
    hs = env.HostStepper4(n, device=0, chunk=1 << 14)
    h_in = boards.cpu().pin_memory()
    h_out = dict(boards=torch.empty((n, 4), dtype=torch.int64).pin_memory(), points=torch.empty((n, 4), dtype=torch.int32).pin_memory(),
                 flags=torch.empty((n, 4), dtype=torch.uint8).pin_memory(), shaping=torch.empty((n, 4), dtype=torch.int64).pin_memory())
    hs.step(h_in, h_out, seed=9, env0=64, ctr=2)
    torch.cuda.synchronize()
    ref = env.step4(boards, seed=9, env0=64, ctr=2)
    for k in ("boards", "points", "flags", "shaping"):
        assert torch.equal(h_out[k], ref[k].cpu()), k


def test_c2_full_size_step4_on_replayed_draws_matches_oracle(env):
    """BASELINE config 2 as it is worded -- 2^20 random boards x 4 moves, bit-exact on REPLAYED spawn draws -- through
    g2048_step4 (the kernel `bench.py` times, replay form): every output of the 4 194 304 transitions against the oracle."""
    nb = 1 << 20
    b = random_boards(nb, 4096)
    rng = np.random.default_rng(31)
    draws = rng.integers(0, 2**32, (nb, 4, 2), dtype=np.uint64).astype(np.uint32)
    draws[::97, :, 1] = 3865470566                       # the threshold pair of the 2 / 4 decision (game.py:937)
    draws[1::97, :, 1] = 3865470567
    got = env.step4(dev_boards(b), replay=torch.from_numpy(draws.view(np.int32)).cuda())
    torch.cuda.synchronize()
    want_b, want = O.step_batch(np.repeat(b, 4), np.tile(np.arange(4, dtype=np.uint8), nb), replay=draws.reshape(-1, 2))
    assert want["overflow"].sum() == 0
    np.testing.assert_array_equal(host_u64(got["boards"].reshape(-1)), want_b)
    np.testing.assert_array_equal(got["points"].reshape(-1).cpu().numpy(), want["points"])
    fl = got["flags"].reshape(-1).cpu().numpy()
    np.testing.assert_array_equal(fl & 0x0F, want["legal_after"])
    np.testing.assert_array_equal((fl >> 4) & 1, want["done"])
    np.testing.assert_array_equal((fl >> 5) & 1, want["invalid"])
    sh = env.decode_shaping(got["shaping"].reshape(-1).cpu().numpy())
    for k in SH_KEYS:
        np.testing.assert_array_equal(sh[k], want[k], err_msg=k)


def test_host_stepper4_pipelined_steps_equal_joined_steps():
    """HostStepper4.step(join=False): consecutive steps overlap on the stepper's streams; after join() every step's outputs are
    what the joined call returns."""
    from g2048 import env
    n = 70001
    b = random_boards(n, 77)
    h_in = torch.from_numpy(b.view(np.int64)).pin_memory()
    mk = lambda: dict(boards=torch.empty((n, 4), dtype=torch.int64).pin_memory(), points=torch.empty((n, 4), dtype=torch.int32).pin_memory(),
                      flags=torch.empty((n, 4), dtype=torch.uint8).pin_memory(), shaping=torch.empty((n, 4), dtype=torch.int64).pin_memory())
    hs = env.HostStepper4(n, device=0, chunk=1 << 13)
    outs = [mk() for _ in range(3)]
    for k in range(3):
        hs.step(h_in, outs[k], seed=3, env0=10, ctr=k, join=False)
    hs.join()
    torch.cuda.synchronize()
    want = mk()
    for k in range(3):
        hs.step(h_in, want, seed=3, env0=10, ctr=k)
        torch.cuda.synchronize()
        for key in want:
            assert torch.equal(outs[k][key], want[key]), (k, key)
