"""Test infrastructure: a reader for ONNX files (protobuf wire format, the fields a feed-forward graph uses) and a numpy
evaluator for the six operators of the reference's exported policy (docs/data/model.onnx: Cast, MatMul,
LayerNormalization, Relu, Add, Gemm).  `onnx` / `onnxruntime` are not in this image; `oracle/make_golden.py --onnx-only`
runs this reader on the reference's own shipped file to produce tests/golden/onnx_structure.json."""
from __future__ import annotations

import struct

import numpy as np


def _varint(b: bytes, i: int):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, i


def fields(b: bytes):
    """[(field number, wire type, value)] of one message; length-delimited values stay bytes."""
    i, out = 0, []
    while i < len(b):
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"wire type {w}")
        out.append((f, w, v))
    return out


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _attr(b: bytes):
    name, val = None, None
    for f, w, v in fields(b):
        if f == 1:
            name = v.decode()
        elif f == 2:
            val = struct.unpack("<f", v)[0]
        elif f == 3:
            val = _signed(v)
    return name, val


def _value_info(b: bytes):
    name, dims, elem = None, [], None
    for f, _, v in fields(b):
        if f == 1:
            name = v.decode()
        elif f == 2:
            for f2, _, v2 in fields(v):
                if f2 == 1:                                        # tensor_type
                    for f3, _, v3 in fields(v2):
                        if f3 == 1:
                            elem = v3
                        elif f3 == 2:
                            for f4, _, v4 in fields(v3):
                                if f4 == 1:
                                    d = [x for ff, _, x in fields(v4) if ff == 1]
                                    dims.append(d[0] if d else None)
    return {"name": name, "elem_type": elem, "dims": dims}


def load(path_or_bytes) -> dict:
    b = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    model = {"ir_version": None, "opset": [], "nodes": [], "initializers": {}, "inputs": [], "outputs": [], "graph_name": None}
    for f, w, v in fields(bytes(b)):
        if f == 1:
            model["ir_version"] = v
        elif f == 8:
            model["opset"].append(dict((("domain", x.decode()) if ff == 1 else ("version", x)) for ff, _, x in fields(v)))
        elif f == 7:
            for gf, _, gv in fields(v):
                if gf == 1:
                    n = {"inputs": [], "outputs": [], "name": None, "op": None, "attrs": {}}
                    for nf, _, nv in fields(gv):
                        if nf == 1:
                            n["inputs"].append(nv.decode())
                        elif nf == 2:
                            n["outputs"].append(nv.decode())
                        elif nf == 3:
                            n["name"] = nv.decode()
                        elif nf == 4:
                            n["op"] = nv.decode()
                        elif nf == 5:
                            k, val = _attr(nv)
                            n["attrs"][k] = val
                    model["nodes"].append(n)
                elif gf == 2:
                    model["graph_name"] = gv.decode()
                elif gf == 5:
                    dims, dt, name, raw, fl = [], None, None, None, []
                    for tf, tw, tv in fields(gv):
                        if tf == 1:
                            dims.append(tv)
                        elif tf == 2:
                            dt = tv
                        elif tf == 8:
                            name = tv.decode()
                        elif tf == 9:
                            raw = tv
                        elif tf == 4:                              # float_data (packed or not)
                            fl += list(struct.unpack(f"<{len(tv) // 4}f", tv))
                    assert dt == 1, f"initializer {name}: data type {dt} (only float32 is read)"
                    a = np.frombuffer(raw, dtype="<f4") if raw is not None else np.asarray(fl, dtype=np.float32)
                    model["initializers"][name] = a.reshape(dims).astype(np.float32)
                elif gf == 11:
                    model["inputs"].append(_value_info(gv))
                elif gf == 12:
                    model["outputs"].append(_value_info(gv))
    return model


def structure(model: dict) -> dict:
    """What a consumer of the file depends on, without the weights (JSON-serialisable)."""
    return {"ir_version": model["ir_version"], "opset": [o.get("version") for o in model["opset"]], "graph_name": model["graph_name"],
            "nodes": [{"op": n["op"], "name": n["name"], "inputs": n["inputs"], "outputs": n["outputs"],
                       "attrs": {k: (round(v, 12) if isinstance(v, float) else v) for k, v in sorted(n["attrs"].items())}}
                      for n in model["nodes"]],
            "initializers": {k: list(v.shape) for k, v in model["initializers"].items()},
            "inputs": model["inputs"], "outputs": model["outputs"]}


def run(model: dict, feeds: dict) -> dict:
    """Evaluate the graph in float32 numpy (nodes are in topological order in an ONNX file)."""
    env = dict(model["initializers"])
    env.update({k: np.asarray(v) for k, v in feeds.items()})
    for n in model["nodes"]:
        x = [env[i] for i in n["inputs"]]
        a = n["attrs"]
        if n["op"] == "Cast":
            assert a["to"] == 1
            y = x[0].astype(np.float32)
        elif n["op"] == "MatMul":
            y = x[0] @ x[1]
        elif n["op"] == "Relu":
            y = np.maximum(x[0], 0)
        elif n["op"] == "Add":
            y = x[0] + x[1]
        elif n["op"] == "Gemm":
            A = x[0].T if a.get("transA", 0) else x[0]
            B = x[1].T if a.get("transB", 0) else x[1]
            y = np.float32(a.get("alpha", 1.0)) * (A @ B) + np.float32(a.get("beta", 1.0)) * x[2]
        elif n["op"] == "LayerNormalization":
            assert a.get("axis", -1) in (-1, x[0].ndim - 1)
            mu = x[0].mean(-1, keepdims=True, dtype=np.float32)
            var = ((x[0] - mu) ** 2).mean(-1, keepdims=True, dtype=np.float32)
            y = (x[0] - mu) / np.sqrt(var + np.float32(a.get("epsilon", 1e-5))) * x[1] + x[2]
        else:
            raise NotImplementedError(n["op"])
        env[n["outputs"][0]] = y.astype(np.float32)
    return {o["name"]: env[o["name"]] for o in model["outputs"]}
