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


# ----------------------------------------------------------------------------- GameURM (config #5)

@dataclass
class GameURMConfig:  # game.py:31-42
    hidden_dim: int = 64
    num_layers: int = 2
    num_heads: int = 4
    expansion: float = 2.67
    dropout: float = 0.1
    num_loops: int = 4
    num_truncated_loops: int = 1
    conv_kernel: int = 2
    rms_norm_eps: float = 1e-5


def rms_norm(x: torch.Tensor, eps: float) -> torch.Tensor:  # game.py:1223-1229
    x32 = x.to(torch.float32)
    return (x32 * torch.rsqrt(x32.square().mean(-1, keepdim=True) + eps)).to(x.dtype)


class GameConvSwiGLU(nn.Module):  # game.py:1232-1276
    def __init__(self, hidden_size: int, expansion: float, conv_kernel: int = 2):
        super().__init__()
        inter = ((round(expansion * hidden_size * 2 / 3) + 7) // 8) * 8
        self.inter = inter
        self.gate_up_proj = nn.Linear(hidden_size, inter * 2, bias=False)
        self.dwconv = nn.Conv1d(inter, inter, kernel_size=conv_kernel, padding=conv_kernel // 2, groups=inter, bias=True)
        self.down_proj = nn.Linear(inter, hidden_size, bias=False)

    def forward(self, x):
        gate, up = self.gate_up_proj(x).chunk(2, dim=-1)
        h = torch.nn.functional.silu(gate) * up
        c = self.dwconv(h.transpose(1, 2))[..., : h.size(1)]
        return self.down_proj(torch.nn.functional.silu(c).transpose(1, 2).contiguous())


class GameURMAttention(nn.Module):  # game.py:1279-1317
    def __init__(self, hidden_size: int, num_heads: int, dropout: float = 0.0):
        super().__init__()
        self.hidden_size, self.num_heads, self.head_dim, self.dropout = hidden_size, num_heads, hidden_size // num_heads, dropout
        self.qkv_proj = nn.Linear(hidden_size, hidden_size * 3, bias=False)
        self.o_proj = nn.Linear(hidden_size, hidden_size, bias=False)

    def forward(self, x):
        b, s, _ = x.shape
        qkv = self.qkv_proj(x).view(b, s, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        o = torch.nn.functional.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2],
                                                             dropout_p=self.dropout if self.training else 0.0, is_causal=False)
        return self.o_proj(o.transpose(1, 2).contiguous().view(b, s, self.hidden_size))


class GameURMBlock(nn.Module):  # game.py:1320-1352
    def __init__(self, config: GameURMConfig):
        super().__init__()
        self.attn = GameURMAttention(config.hidden_dim, config.num_heads, config.dropout)
        self.mlp = GameConvSwiGLU(config.hidden_dim, config.expansion, config.conv_kernel)
        self.norm_eps = config.rms_norm_eps

    def forward(self, h):
        h = rms_norm(h + self.attn(h), self.norm_eps)
        return rms_norm(h + self.mlp(h), self.norm_eps)


class GameURM(nn.Module):  # game.py:1355-1458
    N = 16
    NUM_ACTIONS = 4

    def __init__(self, config: GameURMConfig):
        super().__init__()
        self.config = config
        self.stem = nn.Sequential(nn.Linear(3, config.hidden_dim, bias=False), nn.LayerNorm(config.hidden_dim), nn.SiLU())
        self.layers = nn.ModuleList([GameURMBlock(config) for _ in range(config.num_layers)])
        self.init_hidden = nn.Parameter(torch.zeros(1, self.N, config.hidden_dim))
        nn.init.trunc_normal_(self.init_hidden, std=0.02)
        self.action_head = nn.Linear(config.hidden_dim, self.NUM_ACTIONS, bias=True)
        self.value_head = nn.Linear(config.hidden_dim, 1, bias=True)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    @property
    def directions(self):
        return list(DIRECTIONS)

    def get_param_groups(self, value_lr: float, other_lr: float) -> list[dict]:
        """Same grouping as GameMLP.get_param_groups (game.py:1093-1127), which the reference's optimiser set-up
        (train.py:1587-1597) needs and its GameURM lacks (its CLI refuses --model-type urm, train.py:1523-1532):
        matrices -> Muon, everything else (LayerNorm, biases, init_hidden [1,16,h], depthwise conv [c,1,k]) -> AdamW,
        the value head on its own learning rate."""
        v1, v2, o1, o2 = [], [], [], []
        for name, p in self.named_parameters():
            value = name.startswith("value_head")
            (v2 if value else o2).append(p) if p.ndim == 2 else (v1 if value else o1).append(p)
        return [{"params": o2, "lr": other_lr}, {"params": o1, "lr": other_lr},
                {"params": v2, "lr": value_lr}, {"params": v1, "lr": value_lr}]

    def forward(self, inputs: torch.Tensor):
        if inputs.ndim == 1:
            inputs = inputs.unsqueeze(0)
        b = inputs.shape[0]
        emb = self.stem(inputs.view(b, self.N, 3))
        h = self.init_hidden.expand(b, -1, -1).clone()
        trunc = self.config.num_truncated_loops
        if trunc > 0:
            with torch.no_grad():
                for _ in range(trunc):
                    h = h + emb
                    for layer in self.layers:
                        h = layer(h)
        for _ in range(self.config.num_loops - trunc):
            h = h + emb
            for layer in self.layers:
                h = layer(h)
        pooled = h.mean(dim=1)
        return self.action_head(pooled), self.value_head(pooled)
