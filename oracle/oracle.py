"""ctypes front end of oracle/liboracle2048.so (the C restatement of the reference).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle2048.so")

INFO_FIELDS = (
    "points", "done", "invalid", "overflow", "mono_before", "mono_after", "empt_before",
    "empt_after", "max_tile_created", "max_exp_before", "max_exp_after", "corner_before",
    "corner_after", "smooth_before", "smooth_after", "legal_before", "legal_after",
)
INFO_DTYPE = np.dtype([(f, np.int32) for f in INFO_FIELDS])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle2048.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


def pack_grid(grid) -> int:
    """list[list[int]] of exponents -> uint64 (cell (r,c) in nibble 4*(4r+c))."""
    b = 0
    for r in range(4):
        for c in range(4):
            e = int(grid[r][c])
            assert 0 <= e <= 15, "exponent does not fit a nibble"
            b |= e << (4 * (4 * r + c))
    return b


def unpack_board(b: int):
    b = int(b)
    return [[(b >> (4 * (4 * r + c))) & 0xF for c in range(4)] for r in range(4)]


def step_batch(boards, actions, replay=None, seed=0, env0=0, ctr=0):
    boards = np.ascontiguousarray(boards, dtype=np.uint64)
    actions = np.ascontiguousarray(actions, dtype=np.uint8)
    n = boards.shape[0]
    if replay is not None:
        replay = np.ascontiguousarray(replay, dtype=np.uint32).reshape(n, 2)
    out = np.empty(n, dtype=np.uint64)
    info = np.zeros(n, dtype=INFO_DTYPE)
    lib().orc_step_batch(_p(boards), _p(actions), _p(replay), C.c_uint64(seed), C.c_uint64(env0),
                         C.c_uint64(ctr), C.c_int64(n), _p(out), _p(info))
    return out, info


def expand4_batch(boards):
    boards = np.ascontiguousarray(boards, dtype=np.uint64)
    n = boards.shape[0]
    succ = np.empty((n, 4), dtype=np.uint64)
    points = np.empty((n, 4), dtype=np.int32)
    max_tile = np.empty((n, 4), dtype=np.uint8)
    legal = np.empty(n, dtype=np.uint8)
    lib().orc_expand4_batch(_p(boards), C.c_int64(n), _p(succ), _p(points), _p(max_tile), _p(legal))
    return succ, points, max_tile, legal


def reset_batch(n, replay=None, seed=0, env0=0, ctr=0):
    out = np.empty(n, dtype=np.uint64)
    if replay is not None:
        replay = np.ascontiguousarray(replay, dtype=np.uint32).reshape(n, 4)
    lib().orc_reset_batch(_p(out), C.c_int64(n), _p(replay), C.c_uint64(seed), C.c_uint64(env0),
                          C.c_uint64(ctr))
    return out


def potentials_batch(boards):
    """-> int32 [n,6]: mono, empt, smooth, corner, max_exp, legal mask."""
    boards = np.ascontiguousarray(boards, dtype=np.uint64)
    out = np.empty((boards.shape[0], 6), dtype=np.int32)
    lib().orc_potentials_batch(_p(boards), C.c_int64(boards.shape[0]), _p(out))
    return out


def potentials_ext_batch(before, after):
    """float64 [n,7]: adjacency b/a, chain b/a, topological b/a (anchor of `before`), anchor (4*row+col)."""
    before = np.ascontiguousarray(before, dtype=np.uint64)
    after = np.ascontiguousarray(after, dtype=np.uint64)
    out = np.empty((before.shape[0], 7), dtype=np.float64)
    lib().orc_potentials_ext_batch(_p(before), _p(after), C.c_int64(before.shape[0]), _p(out))
    return out


def row_table():
    out4 = np.empty((65536, 4), dtype=np.uint8)
    score = np.empty(65536, dtype=np.uint32)
    max_tile = np.empty(65536, dtype=np.uint8)
    lib().orc_row_table(_p(out4), _p(score), _p(max_tile))
    return out4, score, max_tile


def philox(seed: int, env_id: int, ctr: int):
    out = (C.c_uint32 * 4)()
    lib().orc_draws(C.c_uint64(seed), C.c_uint64(env_id), C.c_uint64(ctr), out)
    return [int(x) for x in out]


def philox_raw(ctr4, key2):
    c = (C.c_uint32 * 4)(*ctr4)
    k = (C.c_uint32 * 2)(*key2)
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, out)
    return [int(x) for x in out]


def encode_batch(boards):
    boards = np.ascontiguousarray(boards, dtype=np.uint64)
    out = np.empty((boards.shape[0], 48), dtype=np.float32)
    lib().orc_encode_batch(_p(boards), C.c_int64(boards.shape[0]), _p(out))
    return out


def rtg_adv(points, mono_b, mono_a, empt_b, empt_a, done, valid, value, gamma, w_points, w_mono,
            w_empt, rtg_beta, rtg_step, rtg_mu, rtg_m2):
    """All inputs time-major [T,B].  Returns dict with reward, g_raw, g_norm, adv (f32 [T,B]),
    stats {sum, sumsq, n, mean, var} and the updated (rtg_mu, rtg_m2)."""
    points = np.ascontiguousarray(points, dtype=np.int32)
    T, B = points.shape
    u8 = lambda a: np.ascontiguousarray(a, dtype=np.uint8)
    mono_b, mono_a, empt_b, empt_a, done, valid = map(u8, (mono_b, mono_a, empt_b, empt_a, done, valid))
    value = np.ascontiguousarray(value, dtype=np.float32)
    reward = np.zeros((T, B), dtype=np.float32)
    g_raw = np.zeros((T, B), dtype=np.float32)
    g_norm = np.zeros((T, B), dtype=np.float32)
    adv = np.zeros((T, B), dtype=np.float32)
    moments = np.array([rtg_mu, rtg_m2], dtype=np.float64)
    stats = np.zeros(5, dtype=np.float64)
    lib().orc_rtg_adv(_p(points), _p(mono_b), _p(mono_a), _p(empt_b), _p(empt_a), _p(done), _p(valid),
                      _p(value), C.c_int64(T), C.c_int64(B), C.c_double(gamma), C.c_double(w_points),
                      C.c_double(w_mono), C.c_double(w_empt), C.c_double(rtg_beta), C.c_int64(rtg_step),
                      _p(moments), _p(reward), _p(g_raw), _p(g_norm), _p(adv), _p(stats))
    return dict(reward=reward, g_raw=g_raw, g_norm=g_norm, adv=adv,
                stats=dict(sum=stats[0], sumsq=stats[1], n=int(stats[2]), mean=stats[3], var=stats[4]),
                rtg_mu=float(moments[0]), rtg_m2=float(moments[1]))
