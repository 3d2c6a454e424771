"""ONNX export of the policy (reference train.py:33-78) -- CPU only.  The file this library writes is read back with
tests/onnx_mini.py (pinned on the reference's own docs/data/model.onnx by oracle/make_golden.py --onnx-only) and compared
with the reference's shipped file: same structure (nodes, names, attributes, initializer shapes, inputs / outputs) and,
for the same weights, the same outputs."""
import json
import os

import numpy as np
import torch

import onnx_mini
from g2048 import export
from g2048.policy import GameMLP, MLPConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _best_model():
    d = np.load(os.path.join(GOLD, "model_best.npz"))
    m = GameMLP(MLPConfig(hidden_dim=int(d["hidden_dim"]), num_layers=int(d["num_layers"])))
    m.load_state_dict({k[4:].replace("__", "."): torch.from_numpy(d[k]) for k in d.files if k.startswith("sd__")})
    return m, d


def test_export_has_the_structure_of_the_reference_file(tmp_path):
    m, _ = _best_model()
    m.train()                                                    # the export is eval-mode whatever the caller left
    export.export_model_to_onnx(m, tmp_path / "web" / "model.onnx", m.config)
    assert m.training
    got = onnx_mini.structure(onnx_mini.load(tmp_path / "web" / "model.onnx"))
    want = json.load(open(os.path.join(GOLD, "onnx_structure.json")))
    assert got == want
    cfg = json.load(open(tmp_path / "web" / "model_config.json"))
    assert cfg == {"hidden_dim": 192, "num_layers": 2, "dropout": 0.1, "decouple_critic": False}
    assert cfg == json.loads(str(np.load(os.path.join(GOLD, "onnx_reference.npz"))["config"]))


def test_export_reproduces_the_reference_file_outputs(tmp_path):
    m, d = _best_model()
    export.export_model_to_onnx(m, tmp_path / "model.onnx", m.config)
    out = onnx_mini.run(onnx_mini.load(tmp_path / "model.onnx"), {"board_state": d["inputs"]})
    ref = np.load(os.path.join(GOLD, "onnx_reference.npz"))
    # same weights, same graph, same evaluator: the reference's shipped file gives these outputs bit for bit
    assert np.array_equal(out["action_logits"], ref["action_logits"]) and np.array_equal(out["value"], ref["value"])
    # and they are the reference model's own forward outputs (fixture of the imported reference)
    np.testing.assert_allclose(out["action_logits"], d["logits"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(out["value"], d["value"], rtol=1e-5, atol=2e-5)


def test_export_follows_the_model_shape(tmp_path):
    torch.manual_seed(3)
    for h, L in ((196, 2), (64, 1), (96, 4)):
        m = GameMLP(MLPConfig(hidden_dim=h, num_layers=L, dropout=0.0)).eval()
        for p in m.parameters():                                  # non-trivial LayerNorm parameters and biases
            p.data.add_(0.05 * torch.randn_like(p))
        mo = onnx_mini.load(export.onnx_bytes(m))
        assert [n["op"] for n in mo["nodes"]].count("Add") == L and len(mo["initializers"]) == 3 * (L + 1) + 4
        x = torch.randn(37, 48)
        out = onnx_mini.run(mo, {"board_state": x.numpy()})
        with torch.no_grad():
            lg, v = m(x)
        np.testing.assert_allclose(out["action_logits"], lg.numpy(), rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(out["value"], v.numpy(), rtol=1e-5, atol=2e-5)


def test_export_refuses_other_models():
    import pytest

    with pytest.raises(TypeError):
        export.onnx_bytes(torch.nn.Linear(4, 4))
