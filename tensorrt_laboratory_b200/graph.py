"""Layer-graph IR for the inference hot path.

Two levels:

* **raw Caffe layers** -- a list of plain dicts ``{name, type, bottoms, tops, ...params}`` exactly as a
  deploy prototxt states them (Convolution / BatchNorm / Scale / ReLU / Pooling / Eltwise /
  InnerProduct / Softmax).  Produced either by :func:`parse_prototxt` (front-end for the files the
  reference feeds to ``trtexec --deploy``, reference ``models/setup.py:53-55``) or by
  :func:`resnet_caffe` (a programmatic generator of the same layer list, so the GPU box does not need
  the prototxt).  The CPU oracle executes THIS level, unfused.

* **lowered ops** -- what the engine executes: every Convolution absorbs its BatchNorm+Scale (folded to
  per-channel scale/bias at build time), its ReLU, and -- for the last conv of a bottleneck -- the
  Eltwise SUM + ReLU.  See :func:`lower`.

Caffe semantics that matter (reference ``models/ResNet-50-deploy.prototxt``):
  * Pooling output size uses CEIL:  out = ceil((in + 2p - k)/s) + 1        (``:48-58`` pool1 -> 56)
  * BatchNorm is ``use_global_stats`` followed by a separate Scale layer with ``bias_term``
  * stride 2 sits on the 1x1 ``branch2a`` / ``branch1`` convs (Caffe-v1 ResNet)
  * ``conv1`` has a bias, every other conv says ``bias_term: false``
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Tuple

# --------------------------------------------------------------------------------------------------
# prototxt front-end
# --------------------------------------------------------------------------------------------------

_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([{}:])|"((?:[^"\\]|\\.)*)"|([^\s{}:"#]+))')


def _tokenize(text: str):
    pos = 0
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError(f"prototxt: cannot tokenize at offset {pos}: {text[pos:pos+40]!r}")
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            yield ("sym", m.group(2))
        elif m.group(3) is not None:
            yield ("str", m.group(3))
        else:
            yield ("atom", m.group(4))


def _parse_message(tokens, i, top=False):
    """Parse protobuf text format into {field: [values...]} (every field repeated)."""
    out: Dict[str, list] = {}
    while i < len(tokens):
        kind, val = tokens[i]
        if kind == "sym" and val == "}":
            if top:
                raise ValueError("prototxt: unbalanced '}'")
            return out, i + 1
        if kind != "atom":
            raise ValueError(f"prototxt: expected field name, got {val!r}")
        name = val
        i += 1
        kind, val = tokens[i]
        if kind == "sym" and val == ":":
            i += 1
            kind, val = tokens[i]
            if kind == "sym" and val == "{":
                sub, i = _parse_message(tokens, i + 1)
                out.setdefault(name, []).append(sub)
            else:
                out.setdefault(name, []).append(_scalar(kind, val))
                i += 1
        elif kind == "sym" and val == "{":
            sub, i = _parse_message(tokens, i + 1)
            out.setdefault(name, []).append(sub)
        else:
            raise ValueError(f"prototxt: expected ':' or '{{' after {name}")
    if not top:
        raise ValueError("prototxt: missing '}'")
    return out, i


def _scalar(kind, val):
    if kind == "str":
        return val
    if val in ("true", "false"):
        return val == "true"
    try:
        return int(val)
    except ValueError:
        pass
    try:
        return float(val)
    except ValueError:
        return val  # enum such as MAX / AVE


def _one(msg, key, default=None):
    v = msg.get(key)
    return v[0] if v else default


def parse_prototxt(text: str) -> dict:
    """Parse a Caffe deploy prototxt into ``{name, input, input_dims, layers}`` (raw layer dicts)."""
    tokens = list(_tokenize(text))
    msg, _ = _parse_message(tokens, 0, top=True)
    layers = []
    for L in msg.get("layer", []):
        ltype = _one(L, "type")
        rec = {
            "name": _one(L, "name"),
            "type": ltype,
            "bottoms": list(L.get("bottom", [])),
            "tops": list(L.get("top", [])),
        }
        if ltype == "Convolution":
            p = _one(L, "convolution_param", {})
            rec.update(
                num_output=_one(p, "num_output"),
                kernel_size=_one(p, "kernel_size"),
                pad=_one(p, "pad", 0),
                stride=_one(p, "stride", 1),
                bias_term=_one(p, "bias_term", True),
            )
        elif ltype == "BatchNorm":
            p = _one(L, "batch_norm_param", {})
            rec.update(use_global_stats=_one(p, "use_global_stats", True), eps=_one(p, "eps", 1e-5))
        elif ltype == "Scale":
            p = _one(L, "scale_param", {})
            rec.update(bias_term=_one(p, "bias_term", False))
        elif ltype == "Pooling":
            p = _one(L, "pooling_param", {})
            rec.update(
                pool=_one(p, "pool", "MAX"),
                kernel_size=_one(p, "kernel_size"),
                stride=_one(p, "stride", 1),
                pad=_one(p, "pad", 0),
            )
        elif ltype == "InnerProduct":
            p = _one(L, "inner_product_param", {})
            rec.update(num_output=_one(p, "num_output"), bias_term=_one(p, "bias_term", True))
        elif ltype == "Eltwise":
            p = _one(L, "eltwise_param", {})
            rec.update(operation=_one(p, "operation", "SUM"))
        elif ltype in ("ReLU", "Softmax"):
            pass
        else:
            raise ValueError(f"prototxt: unsupported layer type {ltype!r} ({rec['name']})")
        layers.append(rec)
    dims = [int(d) for d in msg.get("input_dim", [])]
    return {
        "name": _one(msg, "name", "net"),
        "input": _one(msg, "input", "data"),
        "input_dims": dims,  # [1, C, H, W]
        "layers": layers,
    }


# --------------------------------------------------------------------------------------------------
# programmatic generator of the Caffe-v1 ResNet deploy nets (same layer list as the prototxt)
# --------------------------------------------------------------------------------------------------

_RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def _block_names(count: int, stage: int, depth: int) -> List[str]:
    # ResNet-50 names blocks a,b,c,...; the deeper nets name them a, b1, b2, ... in stages 3 and 4.
    if depth == 50 or count <= 3:
        return [chr(ord("a") + i) for i in range(count)]
    return ["a"] + [f"b{i}" for i in range(1, count)]


def resnet_caffe(depth: int = 50) -> dict:
    """Generate the raw layer list of ``models/ResNet-{50,152}-deploy.prototxt`` (reference)."""
    if depth not in _RESNET_BLOCKS:
        raise ValueError(f"unsupported ResNet depth {depth}")
    L: List[dict] = []

    def conv(name, bottom, nout, k, pad, stride, bias):
        L.append(dict(name=name, type="Convolution", bottoms=[bottom], tops=[name], num_output=nout,
                      kernel_size=k, pad=pad, stride=stride, bias_term=bias))

    def bn_scale(suffix, blob):
        L.append(dict(name="bn" + suffix, type="BatchNorm", bottoms=[blob], tops=[blob],
                      use_global_stats=True, eps=1e-5))
        L.append(dict(name="scale" + suffix, type="Scale", bottoms=[blob], tops=[blob], bias_term=True))

    def relu(name, blob):
        L.append(dict(name=name, type="ReLU", bottoms=[blob], tops=[blob]))

    conv("conv1", "data", 64, 7, 3, 2, depth == 50)  # the 152 prototxt says bias_term: false
    bn_scale("_conv1", "conv1")
    relu("conv1_relu", "conv1")
    L.append(dict(name="pool1", type="Pooling", bottoms=["conv1"], tops=["pool1"], pool="MAX",
                  kernel_size=3, stride=2, pad=0))
    prev = "pool1"
    for si, count in enumerate(_RESNET_BLOCKS[depth]):
        stage = si + 2
        mid = 64 << si
        out = mid * 4
        for bi, bname in enumerate(_block_names(count, stage, depth)):
            tag = f"{stage}{bname}"
            stride = 2 if (bi == 0 and stage > 2) else 1
            if bi == 0:
                conv(f"res{tag}_branch1", prev, out, 1, 0, stride, False)
                bn_scale(f"{tag}_branch1", f"res{tag}_branch1")
                shortcut = f"res{tag}_branch1"
            else:
                shortcut = prev
            conv(f"res{tag}_branch2a", prev, mid, 1, 0, stride, False)
            bn_scale(f"{tag}_branch2a", f"res{tag}_branch2a")
            relu(f"res{tag}_branch2a_relu", f"res{tag}_branch2a")
            conv(f"res{tag}_branch2b", f"res{tag}_branch2a", mid, 3, 1, 1, False)
            bn_scale(f"{tag}_branch2b", f"res{tag}_branch2b")
            relu(f"res{tag}_branch2b_relu", f"res{tag}_branch2b")
            conv(f"res{tag}_branch2c", f"res{tag}_branch2b", out, 1, 0, 1, False)
            bn_scale(f"{tag}_branch2c", f"res{tag}_branch2c")
            L.append(dict(name=f"res{tag}", type="Eltwise", bottoms=[shortcut, f"res{tag}_branch2c"],
                          tops=[f"res{tag}"], operation="SUM"))
            relu(f"res{tag}_relu", f"res{tag}")
            prev = f"res{tag}"
    L.append(dict(name="pool5", type="Pooling", bottoms=[prev], tops=["pool5"], pool="AVE",
                  kernel_size=7, stride=1, pad=0))
    L.append(dict(name="fc1000", type="InnerProduct", bottoms=["pool5"], tops=["fc1000"],
                  num_output=1000, bias_term=True))
    L.append(dict(name="prob", type="Softmax", bottoms=["fc1000"], tops=["prob"]))
    return {"name": f"ResNet-{depth}", "input": "data", "input_dims": [1, 3, 224, 224], "layers": L}


# --------------------------------------------------------------------------------------------------
# shape inference on raw layers
# --------------------------------------------------------------------------------------------------

def conv_out(size: int, k: int, pad: int, stride: int) -> int:
    return (size + 2 * pad - k) // stride + 1


def pool_out_ceil(size: int, k: int, pad: int, stride: int, ceil_mode: bool = True) -> int:
    """Caffe pooling: ceil mode, and the last window must start inside the (padded) image.
    ``ceil_mode=False`` is the ONNX/floor convention (used by the MNIST import)."""
    if not ceil_mode:
        return (size + 2 * pad - k) // stride + 1
    out = int(math.ceil((size + 2 * pad - k) / stride)) + 1
    if pad > 0 and (out - 1) * stride >= size + pad:
        out -= 1
    return out


def infer_shapes(net: dict) -> Dict[str, Tuple[int, int, int]]:
    """blob name -> (C, H, W) after the LAST layer that writes it (in-place layers keep the shape)."""
    _, c, h, w = net["input_dims"]
    shapes = {net["input"]: (c, h, w)}
    for L in net["layers"]:
        t = L["type"]
        c, h, w = shapes[L["bottoms"][0]]
        if t == "Convolution":
            k, p, s = L["kernel_size"], L["pad"], L["stride"]
            shapes[L["tops"][0]] = (L["num_output"], conv_out(h, k, p, s), conv_out(w, k, p, s))
        elif t == "Pooling":
            k, p, s = L["kernel_size"], L["pad"], L["stride"]
            cm = L.get("ceil_mode", True)
            shapes[L["tops"][0]] = (c, pool_out_ceil(h, k, p, s, cm), pool_out_ceil(w, k, p, s, cm))
        elif t == "InnerProduct":
            shapes[L["tops"][0]] = (L["num_output"], 1, 1)
        elif t == "Eltwise":
            for b in L["bottoms"][1:]:
                if shapes[b] != (c, h, w):
                    raise ValueError(f"Eltwise {L['name']}: shape mismatch {shapes[b]} vs {(c, h, w)}")
            shapes[L["tops"][0]] = (c, h, w)
        else:  # BatchNorm / Scale / ReLU / Softmax
            shapes[L["tops"][0]] = (c, h, w)
    return shapes


# --------------------------------------------------------------------------------------------------
# lowering: raw layers (+ raw weights) -> fused ops (+ folded weights)
# --------------------------------------------------------------------------------------------------

OP_CONV, OP_MAXPOOL, OP_AVGPOOL, OP_FC, OP_SOFTMAX = "conv", "maxpool", "avgpool", "fc", "softmax"


def lower(net: dict, weights: Optional[dict] = None) -> dict:
    """Fuse Conv+BatchNorm+Scale(+ReLU)(+Eltwise SUM+ReLU) and version in-place blobs.

    Returns ``{name, input, input_shape (C,H,W), tensors {name: (C,H,W)}, ops [...], output}``.
    When ``weights`` (raw, keyed by layer name; see :mod:`weights`) is given each conv/fc op carries
    folded fp32 parameters: ``W`` in OHWI order ``[Cout, kh, kw, Cin]`` already multiplied by the
    per-channel BN*Scale factor, and ``bias`` ``[Cout]``.
    """
    import numpy as np

    layers = net["layers"]
    shapes = infer_shapes(net)
    # last writer of each blob decides the SSA tensor that consumers see
    ops: List[dict] = []
    producer: Dict[str, dict] = {}  # blob -> op dict that currently defines it
    tensors: Dict[str, Tuple[int, int, int]] = {net["input"]: shapes[net["input"]]}

    def fold_state(op):
        return op.setdefault("_fold", {"scale": None, "shift": None})

    consumed_by_eltwise = set()
    i = 0
    while i < len(layers):
        L = layers[i]
        t = L["type"]
        name = L["name"]
        if t == "Convolution":
            cin = tensors[L["bottoms"][0]][0]
            op = dict(type=OP_CONV, name=name, input=L["bottoms"][0], output=L["tops"][0], residual=None,
                      cin=cin, cout=L["num_output"], k=L["kernel_size"], stride=L["stride"], pad=L["pad"],
                      relu=False)
            if weights is not None:
                w = np.asarray(weights[name]["W"], dtype=np.float64)  # [Cout, Cin, k, k]
                b = (np.asarray(weights[name]["b"], dtype=np.float64) if L["bias_term"]
                     else np.zeros(L["num_output"], dtype=np.float64))
                op["_w"], op["_b"] = w, b
            ops.append(op)
            producer[L["tops"][0]] = op
            tensors[L["tops"][0]] = shapes[L["tops"][0]]
        elif t == "BatchNorm":
            op = producer[L["bottoms"][0]]
            if op["type"] != OP_CONV or op.get("_sealed"):
                raise ValueError(f"BatchNorm {name} does not follow a foldable Convolution")
            if weights is not None:
                mean = np.asarray(weights[name]["mean"], dtype=np.float64)
                var = np.asarray(weights[name]["var"], dtype=np.float64)
                inv = 1.0 / np.sqrt(var + L.get("eps", 1e-5))
                op["_w"] = op["_w"] * inv[:, None, None, None]
                op["_b"] = (op["_b"] - mean) * inv
        elif t == "Scale":
            op = producer[L["bottoms"][0]]
            if op["type"] != OP_CONV or op.get("_sealed"):
                raise ValueError(f"Scale {name} does not follow a foldable Convolution")
            if weights is not None:
                g = np.asarray(weights[name]["gamma"], dtype=np.float64)
                op["_w"] = op["_w"] * g[:, None, None, None]
                op["_b"] = op["_b"] * g
                if L.get("bias_term"):
                    op["_b"] = op["_b"] + np.asarray(weights[name]["beta"], dtype=np.float64)
        elif t == "ReLU":
            op = producer[L["bottoms"][0]]
            if op["type"] != OP_CONV:
                raise ValueError(f"ReLU {name}: only conv-fused ReLU is supported")
            op["relu"] = True
            op["_sealed"] = True
        elif t == "Eltwise":
            if L.get("operation", "SUM") != "SUM" or len(L["bottoms"]) != 2:
                raise ValueError(f"Eltwise {name}: only 2-input SUM is supported")
            a, b = L["bottoms"]
            # fuse into whichever input was produced LAST by a conv that nobody else has read yet
            cand = [x for x in (a, b) if producer.get(x, {}).get("type") == OP_CONV
                    and not producer[x].get("_sealed") and x not in consumed_by_eltwise]
            if not cand:
                raise ValueError(f"Eltwise {name}: no fusable conv input")
            fuse = max(cand, key=lambda x: ops.index(producer[x]))
            other = b if fuse == a else a
            op = producer[fuse]
            op["residual"] = other
            op["output"] = L["tops"][0]
            op["_sealed"] = True
            del tensors[fuse]
            tensors[L["tops"][0]] = shapes[L["tops"][0]]
            producer[L["tops"][0]] = op
            consumed_by_eltwise.add(fuse)
        elif t == "Pooling":
            c, h, w = tensors[L["bottoms"][0]]
            if L["pool"] == "MAX":
                op = dict(type=OP_MAXPOOL, name=name, input=L["bottoms"][0], output=L["tops"][0],
                          k=L["kernel_size"], stride=L["stride"], pad=L["pad"],
                          ceil_mode=L.get("ceil_mode", True))
            else:
                if not (L["kernel_size"] == h == w and L["pad"] == 0):
                    raise ValueError(f"Pooling {name}: only global AVE pooling is supported")
                op = dict(type=OP_AVGPOOL, name=name, input=L["bottoms"][0], output=L["tops"][0],
                          k=L["kernel_size"], stride=L["stride"], pad=0, ceil_mode=True)
            ops.append(op)
            producer[L["tops"][0]] = op
            tensors[L["tops"][0]] = shapes[L["tops"][0]]
        elif t == "InnerProduct":
            c, h, w = tensors[L["bottoms"][0]]
            op = dict(type=OP_FC, name=name, input=L["bottoms"][0], output=L["tops"][0],
                      cin=c * h * w, cout=L["num_output"], in_chw=(c, h, w))
            if weights is not None:
                W = np.asarray(weights[name]["W"], dtype=np.float64).reshape(L["num_output"], c, h, w)
                # Caffe flattens C,H,W; the engine keeps activations NHWC -> permute K to (h, w, c)
                op["W"] = np.ascontiguousarray(W.transpose(0, 2, 3, 1).reshape(L["num_output"], -1)).astype(np.float32)
                op["bias"] = (np.asarray(weights[name]["b"], dtype=np.float32) if L["bias_term"]
                              else np.zeros(L["num_output"], np.float32))
            ops.append(op)
            producer[L["tops"][0]] = op
            tensors[L["tops"][0]] = shapes[L["tops"][0]]
        elif t == "Softmax":
            op = dict(type=OP_SOFTMAX, name=name, input=L["bottoms"][0], output=L["tops"][0])
            ops.append(op)
            producer[L["tops"][0]] = op
            tensors[L["tops"][0]] = shapes[L["tops"][0]]
        else:
            raise ValueError(f"unsupported layer type {t}")
        i += 1

    for op in ops:
        if op["type"] == OP_CONV and "_w" in op:
            op["W"] = np.ascontiguousarray(op.pop("_w").transpose(0, 2, 3, 1)).astype(np.float32)  # OHWI
            op["bias"] = op.pop("_b").astype(np.float32)
        op.pop("_sealed", None)
        op.pop("_fold", None)
    return {
        "name": net["name"],
        "input": net["input"],
        "input_shape": tuple(shapes[net["input"]]),
        "tensors": tensors,
        "ops": ops,
        "output": ops[-1]["output"],
    }


def conv_flops(lowered: dict) -> int:
    """2*MAC over conv + fc ops, per image (the algorithmic FLOP count used by the roofline)."""
    total = 0
    for op in lowered["ops"]:
        if op["type"] == OP_CONV:
            c, h, w = lowered["tensors"][op["output"]]
            total += 2 * h * w * op["cout"] * op["cin"] * op["k"] * op["k"]
        elif op["type"] == OP_FC:
            total += 2 * op["cin"] * op["cout"]
    return total
