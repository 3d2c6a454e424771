"""ctypes binding of libg2048.so.  No fallback: a missing library or a failing call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G2048_LIB", os.path.join(HERE, "libg2048.so"))   # override: A/B builds only


class G2048Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libg2048 error {code}: {message}")
        self.code = code


_lib = None

vp, u64, i64, i32, f32, f64 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.c_float, C.c_double

_SIGNATURES = {
    "g2048_init": [C.c_int],
    "g2048_build_lut": [vp, vp],
    "g2048_reset": [vp, i64, vp, u64, u64, u64, vp],
    "g2048_step": [vp, vp, vp, vp, vp, vp, vp, i64, vp, u64, u64, u64, vp],
    "g2048_step4": [vp, vp, vp, vp, vp, vp, i64, vp, u64, u64, u64, vp],
    "g2048_expand4": [vp, vp, vp, vp, vp, vp, i64, vp],
    "g2048_potentials": [vp, vp, vp, i64, vp],
    "g2048_encode": [vp, vp, i64, vp],
    "g2048_augment": [vp] * 11 + [i64, vp],
    "g2048_potentials_ext": [vp, vp, vp, i64, vp],
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the CUDA path)")
        l = C.CDLL(LIB_PATH)
        l.g2048_last_error.restype = C.c_char_p
        l.g2048_version.restype = C.c_char_p
        l.g2048_lut_bytes.restype = C.c_int64
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = l
    return _lib


def register(name: str, argtypes: list) -> None:
    """Used by the other host modules to declare the entry points they call."""
    fn = getattr(lib(), name)
    fn.argtypes = argtypes
    fn.restype = C.c_int


def check(rc: int) -> None:
    if rc != 0:
        raise G2048Error(rc, lib().g2048_last_error().decode())


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))
