"""Minimal ONNX (protobuf wire format) reader -- no ``onnx`` / ``protobuf`` dependency.

Enough of ModelProto/GraphProto/NodeProto/AttributeProto/TensorProto to import the reference's
in-tree known-answer model ``models/onnx/mnist-v1.3/model.onnx`` and its ``test_data_set_*/*.pb``
TensorProto files (reference ``examples/30_PyTensorRT/server.py:19-31`` is the test that pins them).
Field numbers follow onnx.proto3 (IR version 3, opset 8).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message; value is int or bytes."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"onnx: unsupported wire type {wt}")
        yield fno, wt, v


def _packed_varints(v) -> List[int]:
    if isinstance(v, int):
        return [v]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64, 2: np.uint8, 3: np.int8}


def parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    int64s: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += [_signed(x) for x in _packed_varints(v)]
        elif fno == 2:
            dtype = v
        elif fno == 4:  # float_data
            if wt == 2:
                floats += list(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 7:  # int64_data
            int64s += [_signed(x) for x in _packed_varints(v)]
        elif fno == 8:
            name = v.decode()
        elif fno == 9:
            raw = v
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dtype).copy()
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    else:
        arr = np.asarray(int64s, dtype=np_dtype)
    return name, arr.reshape(dims)


def _parse_attr(buf: bytes):
    name = ""
    val = None
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = v.decode(errors="replace")
        elif fno == 5:
            val = parse_tensor(v)[1]
        elif fno == 7:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            ints += [_signed(x) for x in _packed_varints(v)]
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


def _parse_node(buf: bytes) -> dict:
    node = {"inputs": [], "outputs": [], "name": "", "op": "", "attrs": {}}
    for fno, _, v in _fields(buf):
        if fno == 1:
            node["inputs"].append(v.decode())
        elif fno == 2:
            node["outputs"].append(v.decode())
        elif fno == 3:
            node["name"] = v.decode()
        elif fno == 4:
            node["op"] = v.decode()
        elif fno == 5:
            k, a = _parse_attr(v)
            node["attrs"][k] = a
    return node


def _value_info_name(buf: bytes) -> str:
    for fno, _, v in _fields(buf):
        if fno == 1:
            return v.decode()
    return ""


def _value_info_shape(buf: bytes) -> List[int]:
    """Static dims of a ValueInfoProto (type.tensor_type.shape.dim[*].dim_value); symbolic dims read as 0."""
    dims: List[int] = []
    for fno, _, v in _fields(buf):
        if fno != 2:
            continue
        for f2, _, tt in _fields(v):           # TypeProto: tensor_type = 1
            if f2 != 1:
                continue
            for f3, _, shp in _fields(tt):     # Tensor: elem_type = 1, shape = 2
                if f3 != 2:
                    continue
                for f4, _, dim in _fields(shp):  # TensorShapeProto: dim = 1
                    if f4 != 1:
                        continue
                    val = 0
                    for f5, _, dv in _fields(dim):
                        if f5 == 1:
                            val = _signed(dv)
                    dims.append(val)
    return dims


def parse_model(buf: bytes) -> dict:
    """-> {nodes: [...], initializers: {name: ndarray}, inputs: [names], outputs: [names], input_shapes: {name: dims}}
    A truncated or corrupt file is a ValueError, whatever the low-level symptom."""
    import struct
    try:
        return _parse_model(buf)
    except ValueError:
        raise
    except (IndexError, struct.error, OverflowError, MemoryError, KeyError, TypeError) as ex:
        raise ValueError(f"onnx: truncated or corrupt ModelProto ({type(ex).__name__}: {ex})") from ex


def _parse_model(buf: bytes) -> dict:
    graph = None
    for fno, _, v in _fields(buf):
        if fno == 7:
            graph = v
    if graph is None:
        raise ValueError("onnx: no graph in model")
    nodes, inits, inputs, outputs = [], {}, [], []
    in_shapes: Dict[str, List[int]] = {}
    for fno, _, v in _fields(graph):
        if fno == 1:
            nodes.append(_parse_node(v))
        elif fno == 5:
            n, a = parse_tensor(v)
            inits[n] = a
        elif fno == 11:
            inputs.append(_value_info_name(v))
            in_shapes[inputs[-1]] = _value_info_shape(v)
        elif fno == 12:
            outputs.append(_value_info_name(v))
    return {"nodes": nodes, "initializers": inits,
            "inputs": [i for i in inputs if i not in inits], "outputs": outputs,
            "input_shapes": {k: d for k, d in in_shapes.items() if k not in inits}}


def load_model(path: str) -> dict:
    with open(path, "rb") as f:
        return parse_model(f.read())


def load_tensor(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        return parse_tensor(f.read())[1]


def same_upper_pads(size: int, k: int, stride: int = 1) -> Tuple[int, int]:
    """ONNX ``auto_pad=SAME_UPPER``: output = ceil(in/stride); the extra pad goes at the END."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def mnist_to_caffe_like(model: dict) -> Tuple[dict, Dict[str, dict]]:
    """Translate the MNIST CNTK-exported graph into (raw layer list, raw weights).

    Graph (reference ``models/onnx/mnist-v1.3/model.onnx``): Conv5x5(SAME_UPPER)+Add(bias) -> Relu ->
    MaxPool 2/2 -> Conv5x5+Add -> Relu -> MaxPool 3/3 -> Reshape -> MatMul -> Add.  SAME_UPPER with
    k=5,s=1 is symmetric pad 2, so the Caffe-style layer list expresses it exactly.
    """
    inits = model["initializers"]
    produced = {}  # tensor name -> ("const", array) for Reshape-of-initializer
    layers: List[dict] = []
    weights: Dict[str, dict] = {}
    alias: Dict[str, str] = {}

    def res(n):
        while n in alias:
            n = alias[n]
        return n

    inp = model["inputs"][0]
    cur_shape = None
    pending_conv = None
    for node in model["nodes"]:
        op = node["op"]
        ins = [res(x) for x in node["inputs"]]
        out = node["outputs"][0]
        if op == "Conv":
            W = inits[ins[1]]
            k = int(W.shape[2])
            pl, pr = same_upper_pads(28, k) if node["attrs"].get("auto_pad") == "SAME_UPPER" else (0, 0)
            if pl != pr:
                raise ValueError("asymmetric SAME_UPPER pad not expressible")
            name = node["name"] or out
            layers.append(dict(name=name, type="Convolution", bottoms=[ins[0]], tops=[out],
                               num_output=int(W.shape[0]), kernel_size=k, pad=pl, stride=1, bias_term=True))
            weights[name] = {"W": W.astype(np.float32), "b": np.zeros(W.shape[0], np.float32)}
            pending_conv = (out, name)
        elif op == "Add":
            a, b = ins
            const = produced.get(b, inits.get(b))
            if const is None:
                const, a = produced.get(a, inits.get(a)), b
            target = [L for L in layers if L["tops"][0] == a][-1]
            weights[target["name"]]["b"] = np.asarray(const, np.float32).reshape(-1)
            alias[out] = a
        elif op == "Relu":
            layers.append(dict(name=node["name"] or out, type="ReLU", bottoms=[ins[0]], tops=[ins[0]]))
            alias[out] = ins[0]
        elif op == "MaxPool":
            k = int(node["attrs"]["kernel_shape"][0])
            s = int(node["attrs"]["strides"][0])
            pads = node["attrs"].get("pads", [0, 0, 0, 0])
            if any(pads):
                raise ValueError("padded MaxPool not supported")
            layers.append(dict(name=node["name"] or out, type="Pooling", bottoms=[ins[0]], tops=[out],
                               pool="MAX", kernel_size=k, stride=s, pad=0, ceil_mode=False))
        elif op == "Reshape":
            if ins[0] in inits:
                shape = [int(x) for x in (inits[ins[1]] if len(ins) > 1 else node["attrs"]["shape"])]
                produced[out] = inits[ins[0]].reshape(shape)
            else:
                alias[out] = ins[0]  # flatten of the activation: InnerProduct flattens C,H,W itself
        elif op == "MatMul":
            Wm = produced.get(ins[1], inits.get(ins[1]))  # [K, N]
            name = node["name"] or out
            layers.append(dict(name=name, type="InnerProduct", bottoms=[ins[0]], tops=[out],
                               num_output=int(Wm.shape[1]), bias_term=True))
            weights[name] = {"W": np.ascontiguousarray(Wm.T).astype(np.float32),
                             "b": np.zeros(Wm.shape[1], np.float32)}
        else:
            raise ValueError(f"onnx: unsupported op {op}")
    net = {"name": "mnist-v1.3", "input": inp, "input_dims": [1, 1, 28, 28], "layers": layers}
    return net, weights
