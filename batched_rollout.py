"""Drop-in for the module the reference imports but does not ship (train.py:30):

    from batched_rollout import play_games_batched

Put this repository's root on PYTHONPATH next to the reference checkout and
`python train.py train --episodes N --gpu ...` runs its rollouts on the fused B200 kernel.
"""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "2048-ppo_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from g2048.rollout import play_games_batched  # noqa: E402,F401

__all__ = ["play_games_batched"]
