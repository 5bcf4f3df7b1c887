"""INT8 post-training quantization of a lowered graph (the builder-side role of the reference's
``examples/ONNX/resnet50/int8.py:5-22`` + ``calibrator.py:61`` + ``build.py:63-65``: calibrate on a batch set, hand the
scales to the engine builder).

Scheme (what the INT8 kernels of this repository implement, bit for bit):
  * activations: symmetric per-tensor scale ``s = max|x| / 127`` from max-abs calibration over the calibration inputs;
  * weights: symmetric per-OUTPUT-CHANNEL scale ``s_w[c] = max|W[c]| / 127``, ``Wq = clip(rint(W / s_w), -127, 127)``;
  * convolution: ``acc = sum(q_in * Wq)`` exact in int32; epilogue in fp32, two fused multiply-adds (one rounding each)
        t = fma(float(acc), m[c], b[c])               m[c] = fl(s_in * s_w[c] / s_out),  b[c] = fl(bias[c] / s_out)
        t = fma(float(q_res), r, t)                   r = fl(s_res / s_out)                      (fused residual)
        t = max(t, 0)                                                                           (fused ReLU)
        q_out = clip(rint(t), -127, 127)              round-half-even
  * the thin-input stem convolution, the max pool behind it, the classifier (FC) and the softmax stay fp16: a
    ``quantize`` op ``q = clip(rint(fl(float(h) * fl(1/s))), -127, 127)`` sits between the fp16 part and the first INT8
    convolution; the global average pool reads INT8 and writes fp16: ``h = fp16(fl(float(sum q) * fl(s / HW)))``.

A convolution runs in INT8 when both its channel counts are multiples of 64 (all bottleneck convolutions of the ResNets).

The calibration forward pass runs on the host (torch CPU, fp32) -- build-time work like the rest of this module, never
part of the request path.
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

import numpy as np

from . import graph as G

QMAX = 127.0


def _is_int8_conv(op: dict) -> bool:
    return op["type"] == G.OP_CONV and op["cin"] % 64 == 0 and op["cout"] % 64 == 0


def calibrate(lowered: dict, calib_inputs: np.ndarray) -> Dict[str, float]:
    """max |x| of every tensor of the lowered (fused, folded) graph over ``calib_inputs`` [N, C, H, W], fp32 arithmetic."""
    import torch
    import torch.nn.functional as F

    from .graph import pool_out_ceil

    amax: Dict[str, float] = {}
    with torch.no_grad():
        blobs = {lowered["input"]: torch.from_numpy(np.ascontiguousarray(calib_inputs, dtype=np.float32))}
        amax[lowered["input"]] = float(blobs[lowered["input"]].abs().max())
        for op in lowered["ops"]:
            a = blobs[op["input"]]
            t = op["type"]
            if t == G.OP_CONV:
                w = torch.from_numpy(np.ascontiguousarray(op["W"], dtype=np.float32)).permute(0, 3, 1, 2).contiguous()
                y = F.conv2d(a, w, torch.from_numpy(np.asarray(op["bias"], dtype=np.float32)), stride=op["stride"], padding=op["pad"])
                if op["residual"] is not None:
                    y = y + blobs[op["residual"]]
                if op["relu"]:
                    y = torch.relu(y)
            elif t == G.OP_MAXPOOL:
                k, s, p = op["k"], op["stride"], op["pad"]
                ho = pool_out_ceil(a.shape[2], k, p, s) if op["ceil_mode"] else (a.shape[2] + 2 * p - k) // s + 1
                wo = pool_out_ceil(a.shape[3], k, p, s) if op["ceil_mode"] else (a.shape[3] + 2 * p - k) // s + 1
                need_h = (ho - 1) * s + k - a.shape[2] - p
                need_w = (wo - 1) * s + k - a.shape[3] - p
                y = F.max_pool2d(F.pad(a, (p, max(need_w, 0), p, max(need_h, 0)), value=float("-inf")), k, s)
            elif t == G.OP_AVGPOOL:
                y = a.mean(dim=(2, 3), keepdim=True)
            elif t == G.OP_FC:
                flat = a.permute(0, 2, 3, 1).reshape(a.shape[0], -1)
                y = (flat @ torch.from_numpy(np.asarray(op["W"], dtype=np.float32)).t() + torch.from_numpy(np.asarray(op["bias"], dtype=np.float32)))
                y = y.view(a.shape[0], -1, 1, 1)
            elif t == G.OP_SOFTMAX:
                y = torch.softmax(a, dim=1)
            else:
                raise ValueError(f"calibrate: unsupported op {t}")
            blobs[op["output"]] = y
            amax[op["output"]] = float(y.abs().max())
    return amax


def quantize_lowered(lowered: dict, calib_inputs: np.ndarray, amax: Optional[Dict[str, float]] = None) -> dict:
    """-> a lowered graph whose eligible convolutions carry INT8 parameters (``Wq`` int8 OHWI, ``m`` / ``b`` fp32 per
    output channel, ``r`` fp32 or None, ``in_scale`` / ``out_scale``), with ``quantize`` ops inserted where an INT8
    convolution reads an fp16 tensor, ``in_scale`` on an average pool that reads INT8, and ``tensor_scales`` {name: s}
    for every INT8 tensor.  ``lowered`` itself is not modified."""
    amax = amax or calibrate(lowered, calib_inputs)
    q = copy.copy(lowered)
    q["tensors"] = dict(lowered["tensors"])
    q["int8"] = True
    scales: Dict[str, float] = {}   # INT8 tensors only
    ops = []
    alias: Dict[str, str] = {}      # fp16 tensor -> its quantized copy

    def scale_of(name: str) -> float:
        # scales are fp32 numbers (that is what the plan stores); everything derived from them starts from the rounded value
        return float(np.float32(max(amax[name], 1e-12) / QMAX))

    for op in lowered["ops"]:
        op = dict(op)
        if _is_int8_conv(op):
            src = op["input"]
            if src not in scales:  # produced by the fp16 part: quantize it once
                if src not in alias:
                    qn = src + "_q"
                    alias[src] = qn
                    q["tensors"][qn] = q["tensors"][src]
                    scales[qn] = scale_of(src)
                    ops.append(dict(type="quantize", name="quantize:" + src, input=src, output=qn, scale=scales[qn],
                                    inv_scale=np.float32(1.0 / scales[qn])))
                op["input"] = alias[src]
            if op["residual"] is not None:
                if op["residual"] not in scales:
                    raise ValueError(f"conv {op['name']}: the residual input of an INT8 convolution must be an INT8 tensor")
            s_in = scales[op["input"]]
            s_out = scale_of(op["output"])
            W = np.asarray(op["W"], dtype=np.float64)                      # [O, kh, kw, I]
            s_w = np.maximum(np.abs(W).reshape(W.shape[0], -1).max(axis=1), 1e-12) / QMAX
            op["Wq"] = np.clip(np.rint(W / s_w[:, None, None, None]), -QMAX, QMAX).astype(np.int8)
            op["m"] = (s_in * s_w / s_out).astype(np.float32)
            op["b"] = (np.asarray(op["bias"], dtype=np.float64) / s_out).astype(np.float32)
            op["r"] = np.float32(scales[op["residual"]] / s_out) if op["residual"] is not None else None
            op["in_scale"], op["out_scale"], op["w_scale"] = s_in, s_out, s_w
            op["int8"] = True
            scales[op["output"]] = s_out
        else:
            for key in ("input", "residual"):
                if op.get(key) in scales and op["type"] != G.OP_AVGPOOL:
                    raise ValueError(f"{op['name']}: an fp16 operator reads the INT8 tensor {op[key]}")
            if op["type"] == G.OP_AVGPOOL and op["input"] in scales:
                c, h, w = q["tensors"][op["input"]]
                op["in_scale"] = scales[op["input"]]
                op["k_scale"] = np.float32(scales[op["input"]] / float(h * w))
        ops.append(op)
    q["ops"] = ops
    q["tensor_scales"] = scales
    q["calib_amax"] = amax
    return q
