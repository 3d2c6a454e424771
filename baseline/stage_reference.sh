#!/bin/bash
# Stages the UNMODIFIED reference under baseline/_ref/ (git-ignored; it travels to the GPU box with the gpurun snapshot).
# `python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>` succeeds but
# installs only metadata: the reference's pyproject.toml declares no packages or py-modules (it is a directory of scripts),
# so the scripts the hot path and its callers live in are copied verbatim instead.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
mkdir -p "$ROOT/baseline/_ref/docs/data"
cp "$REF/game.py" "$REF/train.py" "$REF/logger.py" "$REF/README.md" "$REF/pyproject.toml" "$ROOT/baseline/_ref/"
cp "$REF/docs/data/best_model.pt" "$REF/docs/data/model_config.json" "$ROOT/baseline/_ref/docs/data/"
echo "staged $(ls "$ROOT/baseline/_ref" | wc -l) entries under baseline/_ref"
