"""Batched 2048 environment on the GPU + a drop-in facade for the reference's game.py API.

Reference interfaces mirrored here (file:line in RobotSail/2048-PPO):
  Direction                         game.py:14-18
  Game2048.reset/step/...           game.py:45-1030 (see the method docstrings)
Boards are int64 CUDA tensors holding the uint64 bit pattern described in include/g2048.h.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum

import numpy as np
import torch

from . import _lib

UP, DOWN, LEFT, RIGHT = 0, 1, 2, 3

FLAG_LEGAL_MASK, FLAG_DONE, FLAG_INVALID, FLAG_OVERFLOW = 0x0F, 0x10, 0x20, 0x40


class Direction(Enum):  # game.py:14-18
    UP = "up"
    DOWN = "down"
    LEFT = "left"
    RIGHT = "right"


DIRECTIONS = [Direction.UP, Direction.DOWN, Direction.LEFT, Direction.RIGHT]  # train.py:266 order
_DIR_INDEX = {d: i for i, d in enumerate(DIRECTIONS)}


def direction_index(d) -> int:
    if isinstance(d, int):
        return d
    if isinstance(d, Direction):
        return _DIR_INDEX[d]
    return _DIR_INDEX[Direction(getattr(d, "value", d))]  # the reference's own enum members


# ----------------------------------------------------------------------------- plumbing

def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


_LUTS: dict[int, torch.Tensor] = {}
_INITED: set[int] = set()


def init(device: torch.device | int | None = None) -> torch.device:
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise ValueError("g2048 runs on CUDA devices only")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _INITED:
        _lib.call("g2048_init", idx)
        _INITED.add(idx)
    return torch.device("cuda", idx)


def lut(device=None) -> torch.Tensor:
    """The table buffer of `device` (row tables + dense step tables, g2048_lut_bytes() bytes), built on first use."""
    dev = init(device)
    if dev.index not in _LUTS:
        with torch.cuda.device(dev):
            t = torch.empty(int(_lib.lib().g2048_lut_bytes()), dtype=torch.uint8, device=dev)
            _lib.call("g2048_build_lut", _ptr(t), _stream())
        _LUTS[dev.index] = t
    return _LUTS[dev.index]


# ----------------------------------------------------------------------------- batched API

def reset(n: int, *, device=None, seed: int = 0, env0: int = 0, ctr: int = 0,
          replay: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """game.py:942-950 for n boards.  replay: uint32-as-int32 [n,4] draws or None (Philox)."""
    dev = init(device if out is None else out.device)
    with torch.cuda.device(dev):
        boards = torch.empty(n, dtype=torch.int64, device=dev) if out is None else _req(out, torch.int64, "out")
        if replay is not None:
            replay = _req(replay, torch.int32, "replay")
            assert replay.numel() == 4 * n
        _lib.call("g2048_reset", _ptr(boards), n, _ptr(replay), seed, env0, ctr, _stream())
    return boards


def step(boards: torch.Tensor, actions: torch.Tensor, *, seed: int = 0, env0: int = 0, ctr: int = 0,
         replay: torch.Tensor | None = None, shaping: bool = True, out: dict | None = None) -> dict:
    """game.py:952-1030 for every (board, action) pair.

    Returns dict(boards int64[n], points int32[n], flags uint8[n], shaping int64[n] | None).
    """
    boards = _req(boards, torch.int64, "boards")
    actions = _req(actions, torch.uint8, "actions")
    n = boards.numel()
    assert actions.numel() == n
    dev = init(boards.device)
    with torch.cuda.device(dev):
        table = lut(dev)
        if out is None:
            out = dict(boards=torch.empty_like(boards), points=torch.empty(n, dtype=torch.int32, device=dev),
                       flags=torch.empty(n, dtype=torch.uint8, device=dev),
                       shaping=torch.empty(n, dtype=torch.int64, device=dev) if shaping else None)
        if replay is not None:
            replay = _req(replay, torch.int32, "replay")
            assert replay.numel() == 2 * n
        _lib.call("g2048_step", _ptr(table), _ptr(boards), _ptr(actions), _ptr(out["boards"]), _ptr(out["points"]),
                  _ptr(out["flags"]), _ptr(out.get("shaping")), n, _ptr(replay), seed, env0, ctr, _stream())
    return out


def step4(boards: torch.Tensor, *, seed: int = 0, env0: int = 0, ctr: int = 0, replay: torch.Tensor | None = None,
          shaping: bool = True, out: dict | None = None) -> dict:
    """game.py:952-1030 for all four moves of every board, each with its own spawn (BASELINE config 2: boards x 4 moves).

    Transition (b, m) = move m (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT) on board b with the draws of env id env0 + 4 b + m: the same results as
    `step(boards.repeat_interleave(4), tensor([0, 1, 2, 3]).repeat(n), env0=env0, ...)`, from one thread per board that shares the
    work the four moves have in common.  Returns dict(boards int64[n,4], points int32[n,4], flags uint8[n,4], shaping int64[n,4] | None)."""
    boards = _req(boards, torch.int64, "boards")
    n = boards.numel()
    dev = init(boards.device)
    with torch.cuda.device(dev):
        table = lut(dev)
        if out is None:
            out = dict(boards=torch.empty((n, 4), dtype=torch.int64, device=dev), points=torch.empty((n, 4), dtype=torch.int32, device=dev),
                       flags=torch.empty((n, 4), dtype=torch.uint8, device=dev),
                       shaping=torch.empty((n, 4), dtype=torch.int64, device=dev) if shaping else None)
        if replay is not None:
            replay = _req(replay, torch.int32, "replay")
            assert replay.numel() == 8 * n
        _lib.call("g2048_step4", _ptr(table), _ptr(boards), _ptr(out["boards"]), _ptr(out["points"]), _ptr(out["flags"]),
                  _ptr(out.get("shaping")), n, _ptr(replay), seed, env0, ctr, _stream())
    return out


def bind_host_thread_to_gpu(device=None):
    """Pin the calling thread to the CPUs NVML reports as local to `device` (its NUMA node / PCIe root), so that
    pinned host buffers allocated afterwards are first-touched next to the GPU and the copies do not cross the
    socket interconnect (several ranks per multi-socket host).  On this pool's boxes (one NUMA node, every GPU
    local to all 32 cpus) it changes nothing: the 8-rank end-to-end figure there (4.0e9 env-steps/s in total
    against 2.2e9 for one rank) is bounded by the host side of the PCIe fabric, not by placement.  Returns the
    previous affinity set (for os.sched_setaffinity) or None when NVML or the affinity call is unavailable;
    never raises."""
    import os
    try:
        import pynvml
        dev = init(device)
        want = str(torch.cuda.get_device_properties(dev).uuid)
        pynvml.nvmlInit()
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(h)
            u = u.decode() if isinstance(u, bytes) else u
            if u.replace("GPU-", "") == want.replace("GPU-", ""):
                old = os.sched_getaffinity(0)
                pynvml.nvmlDeviceSetCpuAffinity(h)
                return old
    except Exception:
        return None
    return None


class HostStepper:
    """Game2048.step for HOST (pinned) arrays: the batch is cut into chunks that flow through
    host->device copy, g2048_step and device->host copies on a ring of CUDA streams, so the two PCIe
    directions and the kernel overlap.  Same results as step(); this is the host-buffer entry point
    that bench.py's `e2e` figure measures."""

    def __init__(self, n: int, *, device=None, chunk: int = 1 << 20, streams: int = 2, shaping: bool = True):
        self.n, self.chunk, self.shaping = n, min(chunk, max(n, 1)), shaping
        self.dev = init(device)
        self.table = lut(self.dev)
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(streams)]
            c = self.chunk
            self.slots = [dict(boards=torch.empty(c, dtype=torch.int64, device=self.dev),
                               actions=torch.empty(c, dtype=torch.uint8, device=self.dev),
                               out_boards=torch.empty(c, dtype=torch.int64, device=self.dev),
                               points=torch.empty(c, dtype=torch.int32, device=self.dev),
                               flags=torch.empty(c, dtype=torch.uint8, device=self.dev),
                               shaping=torch.empty(c, dtype=torch.int64, device=self.dev) if shaping else None)
                          for _ in range(streams)]

    def step(self, h_boards, h_actions, h_out: dict, *, seed: int = 0, env0: int = 0, ctr: int = 0) -> dict:
        """h_boards int64[n], h_actions uint8[n], h_out = dict(boards, points, flags[, shaping]) of pinned
        host tensors.  Returns h_out after all copies completed."""
        for t in (h_boards, h_actions, *[v for v in h_out.values() if v is not None]):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("HostStepper works on pinned host tensors")
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            st.wait_stream(cur)
        k = 0
        for lo in range(0, self.n, self.chunk):
            hi = min(self.n, lo + self.chunk)
            m = hi - lo
            st, sl = self.streams[k % len(self.streams)], self.slots[k % len(self.slots)]
            k += 1
            with torch.cuda.stream(st):
                sl["boards"][:m].copy_(h_boards[lo:hi], non_blocking=True)
                sl["actions"][:m].copy_(h_actions[lo:hi], non_blocking=True)
                _lib.call("g2048_step", _ptr(self.table), _ptr(sl["boards"]), _ptr(sl["actions"]), _ptr(sl["out_boards"]),
                          _ptr(sl["points"]), _ptr(sl["flags"]), _ptr(sl["shaping"]), m, None, seed, env0 + lo, ctr,
                          C.c_void_p(st.cuda_stream))
                h_out["boards"][lo:hi].copy_(sl["out_boards"][:m], non_blocking=True)
                h_out["points"][lo:hi].copy_(sl["points"][:m], non_blocking=True)
                h_out["flags"][lo:hi].copy_(sl["flags"][:m], non_blocking=True)
                if self.shaping:
                    h_out["shaping"][lo:hi].copy_(sl["shaping"][:m], non_blocking=True)
        for st in self.streams:
            cur.wait_stream(st)
        return h_out


class HostStepper4:
    """`step4` for HOST (pinned) arrays, chunked over a ring of CUDA streams like HostStepper: h_boards int64[n] in,
    dict(boards [n,4], points [n,4], flags [n,4][, shaping [n,4]]) out."""

    def __init__(self, n: int, *, device=None, chunk: int = 1 << 18, streams: int = 2, shaping: bool = True):
        self.n, self.chunk, self.shaping = n, min(chunk, max(n, 1)), shaping
        self.dev = init(device)
        self.table = lut(self.dev)
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(streams)]
            c = self.chunk
            self.slots = [dict(boards=torch.empty(c, dtype=torch.int64, device=self.dev),
                               out_boards=torch.empty((c, 4), dtype=torch.int64, device=self.dev),
                               points=torch.empty((c, 4), dtype=torch.int32, device=self.dev),
                               flags=torch.empty((c, 4), dtype=torch.uint8, device=self.dev),
                               shaping=torch.empty((c, 4), dtype=torch.int64, device=self.dev) if shaping else None)
                          for _ in range(streams)]

    def step(self, h_boards, h_out: dict, *, seed: int = 0, env0: int = 0, ctr: int = 0, join: bool = True) -> dict:
        """join=False: the call only enqueues (uploads, kernels, downloads on the stepper's own streams) and the caller's stream
        does not wait for it -- consecutive steps then overlap (the upload and kernel of step k+1 run under the download of step
        k; a device slot is reused in stream order, so there is no race on the device) and `join()` makes the current stream
        wait for everything enqueued so far.  The caller keeps one set of host buffers per step in flight."""
        for t in (h_boards, *[v for v in h_out.values() if v is not None]):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("HostStepper4 works on pinned host tensors")
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            st.wait_stream(cur)
        k = 0
        for lo in range(0, self.n, self.chunk):
            hi = min(self.n, lo + self.chunk)
            m = hi - lo
            st, sl = self.streams[k % len(self.streams)], self.slots[k % len(self.slots)]
            k += 1
            with torch.cuda.stream(st):
                sl["boards"][:m].copy_(h_boards[lo:hi], non_blocking=True)
                _lib.call("g2048_step4", _ptr(self.table), _ptr(sl["boards"]), _ptr(sl["out_boards"]), _ptr(sl["points"]), _ptr(sl["flags"]),
                          _ptr(sl["shaping"]), m, None, seed, env0 + 4 * lo, ctr, C.c_void_p(st.cuda_stream))
                h_out["boards"][lo:hi].copy_(sl["out_boards"][:m], non_blocking=True)
                h_out["points"][lo:hi].copy_(sl["points"][:m], non_blocking=True)
                h_out["flags"][lo:hi].copy_(sl["flags"][:m], non_blocking=True)
                if self.shaping:
                    h_out["shaping"][lo:hi].copy_(sl["shaping"][:m], non_blocking=True)
        if join:
            self.join()
        return h_out

    def join(self) -> None:
        """The current stream waits for every step enqueued so far."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)


def expand4(boards: torch.Tensor, *, want_max_tile: bool = False, out: dict | None = None) -> dict:
    """All four pre-spawn successors (game.py:121-184, 295-299)."""
    boards = _req(boards, torch.int64, "boards")
    n = boards.numel()
    dev = init(boards.device)
    with torch.cuda.device(dev):
        table = lut(dev)
        if out is None:
            out = dict(succ=torch.empty((n, 4), dtype=torch.int64, device=dev),
                       points=torch.empty((n, 4), dtype=torch.int32, device=dev),
                       legal=torch.empty(n, dtype=torch.uint8, device=dev),
                       max_tile=torch.empty((n, 4), dtype=torch.uint8, device=dev) if want_max_tile else None)
        _lib.call("g2048_expand4", _ptr(table), _ptr(boards), _ptr(out["succ"]), _ptr(out["points"]),
                  _ptr(out["legal"]), _ptr(out.get("max_tile")), n, _stream())
    return out


def potentials(boards: torch.Tensor) -> torch.Tensor:
    """int32 [n,6]: monotonicity, emptiness, smoothness, corner bonus, max exponent, legal mask."""
    boards = _req(boards, torch.int64, "boards")
    dev = init(boards.device)
    with torch.cuda.device(dev):
        out = torch.empty((boards.numel(), 6), dtype=torch.int32, device=dev)
        _lib.call("g2048_potentials", _ptr(lut(dev)), _ptr(boards), _ptr(out), boards.numel(), _stream())
    return out


def potentials_ext(before: torch.Tensor, after: torch.Tensor) -> torch.Tensor:
    """float64 [n,7]: adjacency b/a, chain b/a, topological b/a (anchor of `before`), anchor (4*row+col);
    `after` is the pre-spawn successor (game.py:981-1001)."""
    before, after = _req(before, torch.int64, "before"), _req(after, torch.int64, "after")
    dev = init(before.device)
    with torch.cuda.device(dev):
        out = torch.empty((before.numel(), 7), dtype=torch.float64, device=dev)
        _lib.call("g2048_potentials_ext", _ptr(before), _ptr(after), _ptr(out), before.numel(), _stream())
    return out


def encode(boards: torch.Tensor) -> torch.Tensor:
    """game.py:92-101 to_model_format for a batch: float32 [n,48]."""
    boards = _req(boards, torch.int64, "boards")
    dev = init(boards.device)
    with torch.cuda.device(dev):
        out = torch.empty((boards.numel(), 48), dtype=torch.float32, device=dev)
        _lib.call("g2048_encode", _ptr(boards), _ptr(out), boards.numel(), _stream())
    return out


MIRROR_H, MIRROR_V, ROT90, ROT180, ROT270 = 0, 1, 2, 3, 4


def augment(before, after, action, legal, logp, op) -> dict:
    """Mirror / rotate recorded steps (train.py:774-881).  op: uint8[n] of MIRROR_H .. ROT270."""
    before, after = _req(before, torch.int64, "before"), _req(after, torch.int64, "after")
    action, legal, op = _req(action, torch.uint8, "action"), _req(legal, torch.uint8, "legal"), _req(op, torch.uint8, "op")
    logp = _req(logp, torch.float32, "logp")
    n = before.numel()
    dev = init(before.device)
    with torch.cuda.device(dev):
        out = dict(before=torch.empty_like(before), after=torch.empty_like(after), action=torch.empty_like(action),
                   legal=torch.empty_like(legal), logp=torch.empty_like(logp))
        _lib.call("g2048_augment", _ptr(before), _ptr(after), _ptr(action), _ptr(legal), _ptr(logp), _ptr(op),
                  _ptr(out["before"]), _ptr(out["after"]), _ptr(out["action"]), _ptr(out["legal"]), _ptr(out["logp"]),
                  n, _stream())
    return out


# ----------------------------------------------------------------------------- host helpers

def pack_grid(grid) -> int:
    """list[list[int]] exponents -> signed int64 bit pattern of the packed board."""
    b = 0
    for r in range(4):
        for c in range(4):
            e = int(grid[r][c])
            if not 0 <= e <= 15:
                raise ValueError(f"exponent {e} at ({r},{c}) does not fit the 4-bit board packing (valid: 0..15)")
            b |= e << (4 * (4 * r + c))
    return b - (1 << 64) if b >= (1 << 63) else b


def unpack_board(b: int) -> list[list[int]]:
    b = int(b) & ((1 << 64) - 1)
    return [[(b >> (4 * (4 * r + c))) & 0xF for c in range(4)] for r in range(4)]


def decode_shaping(w) -> dict:
    """Unpack the u64 shaping records (include/g2048.h G2048_SH_*) into int arrays."""
    w = np.asarray(w).astype(np.int64).view(np.uint64)
    f = lambda sh, m: ((w >> np.uint64(sh)) & np.uint64(m)).astype(np.int32)
    mb, ma = f(27, 15), f(32, 15)
    cb = np.where(f(31, 1) == 1, mb, -mb)
    ca = np.where(f(36, 1) == 1, ma, -ma)
    return dict(mono_before=f(0, 63), mono_after=f(6, 63), empt_before=f(12, 31), empt_after=f(17, 31),
                max_tile_created=f(22, 31), max_exp_before=mb, max_exp_after=ma, corner_before=cb,
                corner_after=ca, smooth_before=-f(37, 511), smooth_after=-f(46, 511))


# ----------------------------------------------------------------------------- facade

class Game2048:
    """Drop-in for the reference's Game2048 (game.py:45-1030), one board, state on the GPU.

    Every query is a kernel call through the C ABI; this class only converts between the
    reference's Grid (list[list[int]]) and the packed board.  Spawns come from the
    counter-based Philox stream (seed, env_id, move counter) instead of Python's `random`.
    """

    def __init__(self, state=None, *, device=None, seed: int = 0, env_id: int = 0):
        self.device = init(device)
        self.seed, self.env_id, self._ctr = seed, env_id, 0
        if state:
            if not all(int(s) in range(0, 16) for row in state for s in row):   # game.py:58-60 (nibble: <= 15)
                raise AssertionError("exponents must be in 0..15")
            self._b = torch.tensor([pack_grid(state)], dtype=torch.int64, device=self.device)
        else:
            self._b = torch.zeros(1, dtype=torch.int64, device=self.device)

    # -- state
    @property
    def grid(self):
        return unpack_board(self._b.item())

    @grid.setter
    def grid(self, g):
        self._b = torch.tensor([pack_grid(g)], dtype=torch.int64, device=self.device)

    def score(self) -> int:  # game.py:63-64
        return sum(2 ** k for row in self.grid for k in row if k > 0)

    get_score = score

    def _pot(self):
        return potentials(self._b)[0].tolist()

    # -- legality (game.py:103-119, 295-299)
    def _legal(self) -> int:
        return self._pot()[5]

    def has_next_step(self) -> bool:
        return self._legal() != 0

    def direction_has_step(self, direction) -> bool:
        return bool((self._legal() >> direction_index(direction)) & 1)

    def current_valid_directions(self):
        m = self._legal()
        return [d for i, d in enumerate(DIRECTIONS) if (m >> i) & 1]

    # -- moves
    @staticmethod
    def simulate_move(grid, direction, *, device=None):  # game.py:121-160
        dev = init(device)
        b = torch.tensor([pack_grid(grid)], dtype=torch.int64, device=dev)
        r = expand4(b, want_max_tile=True)
        d = direction_index(direction)
        return unpack_board(r["succ"][0, d].item()), int(r["points"][0, d]), int(r["max_tile"][0, d])

    def preview_move_rewards(self):  # game.py:167-184
        r = expand4(self._b)
        pts = r["points"][0].tolist()
        return {d: int(pts[i]) for i, d in enumerate(DIRECTIONS)}

    def move(self, direction):  # game.py:186-216
        if not self.direction_has_step(direction):
            raise ValueError(f"Cannot move in direction {getattr(direction, 'value', direction)}")
        r = expand4(self._b)
        self._b = r["succ"][0, direction_index(direction)].reshape(1).clone()
        return self.grid

    def to_model_format(self) -> torch.Tensor:  # game.py:92-101
        return encode(self._b)[0]

    def reset(self):  # game.py:942-950
        self._ctr = 0
        self._b = reset(1, device=self.device, seed=self.seed, env0=self.env_id, ctr=0)
        self._ctr = 1
        return self.grid

    def step(self, direction):  # game.py:952-1030
        a = torch.tensor([direction_index(direction)], dtype=torch.uint8, device=self.device)
        before = self._b
        r = step(self._b, a, seed=self.seed, env0=self.env_id, ctr=self._ctr, shaping=True)
        flags = int(r["flags"].item())
        if flags & FLAG_OVERFLOW:
            raise OverflowError("a merge produced exponent 16, which the 4-bit board packing cannot represent")
        self._b = r["boards"]
        done = bool(flags & FLAG_DONE)
        if flags & FLAG_INVALID:  # game.py:959-978
            info = {"invalid_move": True, "smoothness_delta": 0.0, "max_tile_created": 0, "corner_delta": 0.0,
                    "adjacency_delta": 0.0, "chain_delta": 0.0, "monotonicity_before": 0.0,
                    "monotonicity_after": 0.0, "topological_delta": 0.0, "emptiness_before": 0.0,
                    "emptiness_after": 0.0}
            return self.grid, 0, done, info
        self._ctr += 1
        s = {k: int(v[0]) for k, v in decode_shaping(r["shaping"].cpu().numpy()).items()}
        pre_spawn = expand4(before)["succ"][0, direction_index(direction)].reshape(1)
        ext = potentials_ext(before, pre_spawn)[0].tolist()
        anchor = int(ext[6])
        info = {  # game.py:1012-1029
            "invalid_move": False,
            "adjacency_delta": ext[1] - ext[0],
            "chain_delta": ext[3] - ext[2],
            "topological_delta": ext[5] - ext[4],
            "topological_anchor": (anchor // 4, anchor % 4),
            "smoothness_delta": float(s["smooth_after"] - s["smooth_before"]),
            "max_tile_created": s["max_tile_created"],
            "max_exponent_before": s["max_exp_before"],
            "max_exponent_after": s["max_exp_after"],
            "corner_delta": float(s["corner_after"] - s["corner_before"]),
            "monotonicity_before": s["mono_before"],
            "monotonicity_after": s["mono_after"],
            "emptiness_before": s["empt_before"],
            "emptiness_after": s["empt_after"],
        }
        return self.grid, int(r["points"].item()), done, info

    # -- small helpers of the reference API
    @staticmethod
    def calculate_grid_score(grid) -> int:  # game.py:162-165
        return sum(2 ** k for row in grid for k in row if k > 0)

    @staticmethod
    def create_random_board(generator: torch.Generator | None = None):  # game.py:76-90
        board = torch.zeros((4, 4))
        for idx in torch.randperm(16, generator=generator)[:2]:
            board[idx // 4, idx % 4] = 1
        return board.tolist()

    @staticmethod
    def _symmetry(grid, op: int, device=None):
        dev = init(device)
        b = torch.tensor([pack_grid(grid)], dtype=torch.int64, device=dev)
        z8 = torch.zeros(1, dtype=torch.uint8, device=dev)
        r = augment(b, b, z8, z8, torch.zeros((1, 4), dtype=torch.float32, device=dev),
                    torch.tensor([op], dtype=torch.uint8, device=dev))
        return unpack_board(r["before"].item())

    @staticmethod
    def mirror_grid(grid, direction: str):  # game.py:508-535
        if direction not in ("horizontal", "vertical"):
            raise ValueError(f"Invalid direction: {direction}. Must be 'horizontal' or 'vertical'")
        return Game2048._symmetry(grid, MIRROR_H if direction == "horizontal" else MIRROR_V)

    @staticmethod
    def rotate_grid(grid, rotation):  # game.py:537-590
        degrees = {"north": 0, "up": 0, 0: 0, "east": 90, "right": 90, 90: 90, "south": 180, "down": 180, 180: 180,
                   "west": 270, "left": 270, 270: 270}.get(rotation)
        if degrees is None:
            raise ValueError(f"Invalid rotation: {rotation}")
        if degrees == 0:
            return [row[:] for row in grid]
        return Game2048._symmetry(grid, {90: ROT90, 180: ROT180, 270: ROT270}[degrees])

    # -- potentials (static in the reference: game.py:339-399, 671-800)
    @staticmethod
    def _pot_of(grid, idx, device=None):
        dev = init(device)
        return int(potentials(torch.tensor([pack_grid(grid)], dtype=torch.int64, device=dev))[0, idx])

    @staticmethod
    def monotonicity(grid, require_corner_max: bool = False) -> int:
        return Game2048._pot_of(grid, 0)

    @staticmethod
    def emptiness(grid) -> int:
        return Game2048._pot_of(grid, 1)

    @staticmethod
    def smoothness_score(grid) -> float:
        return float(Game2048._pot_of(grid, 2))

    @staticmethod
    def corner_bonus(grid) -> float:
        return float(Game2048._pot_of(grid, 3))

    @staticmethod
    def state_has_next_step(state) -> bool:
        return Game2048._pot_of(state, 5) != 0
