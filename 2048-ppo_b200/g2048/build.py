"""Builds libg2048.so (hand-written sm_100a CUDA behind the C ABI of include/g2048.h) in-tree.

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it
travels with a repo snapshot.  There is no JIT and no fallback: if the library is missing,
importing g2048._lib fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.environ.get("G2048_LIB", os.path.join(HERE, "libg2048.so"))
SOURCES = ["g2048_host.cu", "g2048_env.cu", "g2048_train.cu", "g2048_rollout.cu", "g2048_rollout_tc.cu", "g2048_rollout_x3.cu", "g2048_rollout_urm.cu", "g2048_rollout_urm_x3.cu", "g2048_urm_train.cu", "g2048_tc.cu",
           "g2048_update.cu", "g2048_linear.cu", "g2048_update_x3.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libg2048.so")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "g2048.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("G2048_NVCC_DEFS", "").split()      # e.g. -DG2048_SHIFT_ON_FMA=0 for A/B runs
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB + ".tmp", *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
