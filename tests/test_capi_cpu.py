"""CPU-only: the C-ABI library loads and exports every symbol include/g2048.h declares; host
helpers (packing, shaping decode) behave.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g2048.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g2048_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from g2048 import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"libg2048.so does not export {n}"
    assert lib.g2048_lut_bytes() == 2 * 65536 * 4 + 167040 + 57136   # row tables + dense step tables (M | S)
    assert b"sm_100a" in lib.g2048_version()


def test_error_convention_without_gpu():
    import torch
    from g2048 import _lib
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    lib = _lib.lib()
    rc = lib.g2048_init(0)
    assert rc < 0
    assert len(lib.g2048_last_error()) > 0
    with pytest.raises(_lib.G2048Error):
        _lib.check(rc)


def test_no_cpu_fallback_in_host_api():
    import torch
    from g2048 import env
    with pytest.raises((ValueError, RuntimeError)):
        env.step(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8))


def test_pack_unpack_roundtrip():
    from g2048 import env
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    for _ in range(200):
        g = rng.integers(0, 16, size=(4, 4)).tolist()
        b = env.pack_grid(g)
        assert env.unpack_board(b) == g
        assert (b & ((1 << 64) - 1)) == O.pack_grid(g)
    with pytest.raises(ValueError):
        env.pack_grid([[16, 0, 0, 0]] + [[0] * 4] * 3)


def test_shaping_decode_matches_header_macros():
    from g2048 import env
    w = (30 | 28 << 6 | 3 << 12 | 4 << 17 | 5 << 22 | 13 << 27 | 1 << 31 | 12 << 32 | 0 << 36 | 42 << 37 | 40 << 46)
    d = env.decode_shaping(np.array([w], dtype=np.uint64).view(np.int64))
    assert {k: int(v[0]) for k, v in d.items()} == dict(
        mono_before=30, mono_after=28, empt_before=3, empt_after=4, max_tile_created=5, max_exp_before=13,
        max_exp_after=12, corner_before=13, corner_after=-12, smooth_before=-42, smooth_after=-40)


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, load or name it."""
    pkg = os.path.join(ROOT, "2048-ppo_b200")
    hits = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                if re.search(r"\boracle\b", text):
                    hits.append(os.path.join(d, f))
    assert hits == [], hits
    shim = open(os.path.join(ROOT, "batched_rollout.py")).read()
    assert "oracle" not in shim
