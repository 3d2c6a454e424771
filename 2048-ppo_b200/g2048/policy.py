"""Policy / value network interface of the reference (game.py:24-28, 1033-1220), same module
tree and state_dict keys, so `best_model.pt` loads unchanged:

    stem.0.weight (h,48)  stem.1.{weight,bias} (h)
    backbone.{i}.mlp.0.weight (h,h)  backbone.{i}.mlp.1.{weight,bias} (h)
    action_head.{weight (4,h), bias (4)}  value_head.{weight (1,h), bias (1)}

forward(x:(B,48)) -> (logits (B,4), value (B,1)).  This torch module is the update-path model
(autograd + cuBLAS, per SURVEY section 8: trunk backward stays in torch); the rollout never calls
it -- it reads the same weights through pack_weights() into the fused CUDA rollout kernel.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass

import torch
import torch.nn as nn

from .env import DIRECTIONS


@dataclass
class MLPConfig:  # game.py:24-28
    hidden_dim: int = 64
    num_layers: int = 2
    dropout: float = 0.1
    decouple_critic: bool = False

    def model_dump(self) -> dict:
        return asdict(self)


class ResidualBlock(nn.Module):  # game.py:1033-1046
    def __init__(self, hidden_dim: int, dropout: float = 0.1):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim, bias=False),
            nn.LayerNorm(hidden_dim),
            nn.ReLU(),
            nn.Dropout(dropout),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.mlp(x)


class GameMLP(nn.Module):  # game.py:1049-1220
    N = 16
    NUM_ACTIONS = 4

    @staticmethod
    def init_kaiming(m: nn.Module) -> None:  # game.py:1054-1059
        if isinstance(m, nn.Linear):
            nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def __init__(self, config: MLPConfig) -> None:
        super().__init__()
        self.config = config
        self.decouple_critic = config.decouple_critic
        h = config.hidden_dim
        self.stem = nn.Sequential(nn.Linear(self.N * 3, h, bias=False), nn.LayerNorm(h), nn.ReLU())
        self.backbone = nn.ModuleList([ResidualBlock(h, config.dropout) for _ in range(config.num_layers)])
        self.action_head = nn.Linear(h, self.NUM_ACTIONS, bias=True)
        self.value_head = nn.Linear(h, 1, bias=True)
        self.apply(GameMLP.init_kaiming)

    @property
    def directions(self):  # game.py:1087-1092
        return list(DIRECTIONS)

    def get_param_groups(self, value_lr: float, other_lr: float) -> list[dict]:  # game.py:1093-1127
        v1, v2, o1, o2 = [], [], [], []
        for name, module in self.named_children():
            for p in module.parameters():
                tgt = (v1 if p.ndim == 1 else v2) if name == "value_head" else (o1 if p.ndim == 1 else o2)
                tgt.append(p)
        return [{"params": o2, "lr": other_lr}, {"params": o1, "lr": other_lr},
                {"params": v2, "lr": value_lr}, {"params": v1, "lr": value_lr}]

    def get_1d_and_2d_params(self):  # game.py:1129-1143
        p1 = [p for p in self.parameters() if p.ndim == 1]
        p2 = [p for p in self.parameters() if p.ndim >= 2]
        return p1, p2

    def forward(self, inputs: torch.Tensor):  # game.py:1145-1220
        if inputs.ndim <= 1:
            raise ValueError(f"input must consist of shape (batch, channel), got: {inputs.shape}")
        assert inputs.shape[-1] == self.N * 3, f"{inputs.shape[-1]} does not equal {self.N * 3}"
        x = self.stem(inputs.to(dtype=torch.float32))
        for layer in self.backbone:
            x = layer(x)
        logits = self.action_head(x)
        value = self.value_head(x.detach() if self.decouple_critic else x)
        return logits, value


def load_state_dict_from_npz(npz, prefix: str = "sd__") -> dict:
    """Rebuild a state_dict from tests/golden/model_best.npz (keys stored with '__' for '.')."""
    return {k[len(prefix):].replace("__", "."): torch.from_numpy(npz[k].copy())
            for k in npz.files if k.startswith(prefix)}
