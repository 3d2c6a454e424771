"""g2048 -- host side of the B200-native 2048-PPO hot path.

Mirrors the reference's interfaces for this path:
  g2048.env      Game2048 / Direction facade (game.py) over batched CUDA kernels
  g2048.policy   GameMLP / MLPConfig (game.py:24-28,1033-1220), same state_dict keys
  g2048.rollout  play_games_batched (the slot train.py:30 imports) and the [T,B] rollout engine
  g2048.ppo      calculate_advantage / PPO-clip loss (train.py:414-772) on fused kernels
All compute goes through libg2048.so (include/g2048.h); nothing here falls back to the CPU.
"""
from ._lib import G2048Error, lib  # noqa: F401

__all__ = ["G2048Error", "lib"]
