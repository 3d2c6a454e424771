"""`export_model_to_onnx` of the reference (train.py:33-78): the GameMLP policy as `model.onnx` + `model_config.json` for
the browser client (docs/js), same graph the reference's export produces -- input `board_state` (batch, 48), outputs
`action_logits` (batch, 4) and `value` (batch, 1), weights embedded, node / initializer names of the torch legacy exporter
(the names `docs/data/model.onnx` of the reference carries; pinned in tests/golden/onnx_structure.json).

The reference goes through `torch.onnx.export` + the `onnx` package, and neither `onnx` nor `onnxscript` is in this image
(torch's exporter refuses to finish without them).  An ONNX file is a protobuf message of a dozen field kinds, so the file
is written here directly: `_Msg` is the wire encoder (varint / length-delimited / fixed32 fields), the graph is the fixed
chain MatMul -> LayerNormalization -> ReLU (+ residual Add) per block, Gemm heads.  Host-side, once per checkpoint: not
part of the hot path."""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np
import torch

IR_VERSION = 9          # what torch 2.9's exporter wrote into the reference's docs/data/model.onnx
OPSET = 20
_FLOAT, _INT = 1, 2     # AttributeProto.AttributeType
_DT_FLOAT = 1           # TensorProto.DataType


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1                                      # negative int64 -> ten-byte two's complement, as protobuf does
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


class _Msg:
    """A protobuf message under construction (fields are appended in the order given)."""

    def __init__(self) -> None:
        self.buf = bytearray()

    def int(self, field: int, v: int) -> "_Msg":
        self.buf += _varint(field << 3) + _varint(int(v))
        return self

    def bytes(self, field: int, v) -> "_Msg":
        b = v.buf if isinstance(v, _Msg) else (v.encode() if isinstance(v, str) else bytes(v))
        self.buf += _varint((field << 3) | 2) + _varint(len(b)) + b
        return self

    def float(self, field: int, v: float) -> "_Msg":
        self.buf += _varint((field << 3) | 5) + struct.pack("<f", v)
        return self


def _attr(name: str, value) -> _Msg:        # AttributeProto: name = 1, f = 2, i = 3, type = 20
    a = _Msg().bytes(1, name)
    return a.float(2, value).int(20, _FLOAT) if isinstance(value, float) else a.int(3, value).int(20, _INT)


def _node(op: str, name: str, inputs, outputs, **attrs) -> _Msg:   # NodeProto: input 1, output 2, name 3, op_type 4, attribute 5
    n = _Msg()
    for i in inputs:
        n.bytes(1, i)
    for o in outputs:
        n.bytes(2, o)
    n.bytes(3, name).bytes(4, op)
    for k, v in attrs.items():
        n.bytes(5, _attr(k, v))
    return n


def _tensor(name: str, t: torch.Tensor) -> _Msg:                   # TensorProto: dims 1, data_type 2, name 8, raw_data 9
    a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
    m = _Msg()
    for d in a.shape:
        m.int(1, d)
    return m.int(2, _DT_FLOAT).bytes(8, name).bytes(9, a.astype("<f4").tobytes())


def _value_info(name: str, dims) -> _Msg:                          # ValueInfoProto{name, type{tensor_type{elem_type, shape{dim}}}}
    shape = _Msg()
    for d in dims:
        shape.bytes(1, _Msg().int(1, d))
    return _Msg().bytes(1, name).bytes(2, _Msg().bytes(1, _Msg().int(1, _DT_FLOAT).bytes(2, shape)))


def onnx_bytes(model) -> bytes:
    """The serialized ModelProto of a GameMLP (game.py:1049-1220), eval-mode semantics (Dropout is the identity)."""
    if not (hasattr(model, "stem") and hasattr(model, "backbone") and hasattr(model, "action_head")):
        raise TypeError("export_model_to_onnx takes a GameMLP (the reference exports nothing else, train.py:33)")
    sd = model.state_dict()
    nodes, inits = [], []
    for k in ("stem.1.weight", "stem.1.bias"):
        inits.append(_tensor(k, sd[k]))
    L = len(model.backbone)
    for i in range(L):
        for k in (f"backbone.{i}.mlp.1.weight", f"backbone.{i}.mlp.1.bias"):
            inits.append(_tensor(k, sd[k]))
    for k in ("action_head.weight", "action_head.bias", "value_head.weight", "value_head.bias"):
        inits.append(_tensor(k, sd[k]))
    eps = float(model.stem[1].eps)
    wname = lambda j: f"onnx::MatMul_{31 + j}"     # the names in the reference's shipped file (the counter depends on the torch version)

    def block(prefix: str, x: str, j: int, wkey: str, lnkey: str) -> str:
        inits.append(_tensor(wname(j), sd[wkey].t()))               # Linear(bias=False) is exported as MatMul by W^T
        mm, ln, relu = (f"/{prefix}.0/MatMul", f"/{prefix}.1/LayerNormalization", f"/{prefix}.2/Relu")
        nodes.append(_node("MatMul", mm, [x, wname(j)], [mm + "_output_0"]))
        nodes.append(_node("LayerNormalization", ln, [mm + "_output_0", lnkey + ".weight", lnkey + ".bias"], [ln + "_output_0"],
                           axis=-1, epsilon=eps))
        nodes.append(_node("Relu", relu, [ln + "_output_0"], [relu + "_output_0"]))
        return relu + "_output_0"

    nodes.append(_node("Cast", "/Cast", ["board_state"], ["/Cast_output_0"], to=_DT_FLOAT))
    x = block("stem/stem", "/Cast_output_0", 0, "stem.0.weight", "stem.1")
    for i in range(L):
        y = block(f"backbone.{i}/mlp/mlp", x, i + 1, f"backbone.{i}.mlp.0.weight", f"backbone.{i}.mlp.1")
        nodes.append(_node("Add", f"/backbone.{i}/Add", [x, y], [f"/backbone.{i}/Add_output_0"]))
        x = f"/backbone.{i}/Add_output_0"
    for head, out in (("action_head", "action_logits"), ("value_head", "value")):
        nodes.append(_node("Gemm", f"/{head}/Gemm", [x, head + ".weight", head + ".bias"], [out], alpha=1.0, beta=1.0, transB=1))

    g = _Msg()                                                      # GraphProto: node 1, name 2, initializer 5, input 11, output 12
    for n in nodes:
        g.bytes(1, n)
    g.bytes(2, "main_graph")
    for t in inits:
        g.bytes(5, t)
    g.bytes(11, _value_info("board_state", (1, model.stem[0].in_features)))
    g.bytes(12, _value_info("action_logits", (1, model.action_head.out_features)))
    g.bytes(12, _value_info("value", (1, model.value_head.out_features)))
    # ModelProto: ir_version 1, producer_name 2, producer_version 3, graph 7, opset_import 8 {version 2}
    m = _Msg().int(1, IR_VERSION).bytes(2, "g2048").bytes(3, "0.1").bytes(7, g).bytes(8, _Msg().int(2, OPSET))
    return bytes(m.buf)


def export_model_to_onnx(model, output_path, config) -> None:
    """train.py:33-78, same arguments: writes `output_path` (weights embedded) and `model_config.json` beside it."""
    was_training = model.training
    model.eval()
    out = Path(output_path)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_bytes(onnx_bytes(model))
    cfg = config.model_dump() if hasattr(config, "model_dump") else dict(config)
    with open(out.parent / "model_config.json", "w") as f:
        json.dump(cfg, f, indent=2)
    model.train(was_training)
