"""INT8 CPU ORACLE -- test infrastructure only.  Never imported by the product package.

Executes a quantized lowered graph (``tensorrt_laboratory_b200.quantize.quantize_lowered``) with EXACT integer
accumulation and the explicit fp32 rounding steps the scheme prescribes, so the GPU INT8 path (``tcgen05.mma.kind::i8``
convolutions with the requantising epilogue) can be compared BIT FOR BIT.

What it restates.  The reference reaches INT8 through TensorRT's builder (``examples/ONNX/resnet50/build.py:63-65``
``--int8`` + the entropy calibrator of ``calibrator.py:61``; ``int8.py:5-22`` feeds it batches); TensorRT is closed
source and absent here, so -- as for the fp16 path -- the oracle restates PUBLISHED semantics: symmetric linear
quantization, int8 x int8 -> int32 accumulation, per-output-channel weight scales, per-tensor activation scales, fp32
requantisation with round-half-to-even (the arithmetic of TensorRT's / ONNX's QLinearConv).  **Parity unpinned**: the
reference holds no INT8 golden vector in-tree; what this oracle pins is the repository's own scheme (quantize.py), and
the accuracy claim is oracle-int8 vs oracle-fp32 on the synthetic inputs (tests/test_int8.py).

Rounding contract (every ``fl`` / ``fma`` is one IEEE fp32 operation, round-to-nearest-even):
    quantize   q = clip(rint(fl(h * inv_s)), -127, 127)                          h: the fp16 value as fp32
    conv       t = fma(float(acc), m[c], b[c]);  t = fma(float(q_res), r, t);  t = max(t, 0);
               q = clip(rint(t), -127, 127)            (fused multiply-adds: ``fma32`` below restates them exactly)
    avg pool   h = fp16(fl(float(sum q) * k))
    output     y = fl(float(q) * fl(s))                                          (INT8 tensor exposed as fp32 binding)
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle.caffe_forward import _pool_out

f32 = np.float32


def fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """IEEE fp32 fused multiply-add, round(a*b + c) with ONE rounding, for float32 arrays -- exact, not approximately:
    the product of two 24-bit significands is exact in float64; the float64 sum p + c may round, so its error is recovered
    with TwoSum, and the only case in which rounding the float64 sum to float32 differs from rounding the true value -- the
    float64 sum lands exactly on a float32 tie while the true value is off the tie -- is resolved by the sign of that error."""
    shape = np.broadcast_shapes(np.shape(a), np.shape(b), np.shape(c))
    ta, tb, tc = (torch.broadcast_to(torch.from_numpy(np.array(v, dtype=np.float32)).double(), shape).reshape(-1)
                  for v in (a, b, c))                # torch: the same IEEE float64 operations, on all host cores
    p = ta * tb                                    # exact
    s = p + tc                                     # rounded to float64
    r = s.float()                                  # round-to-nearest-even of s
    # a float64 is a float32 tie iff its low 29 significand bits are 1000...0 (float32-normal magnitudes); everything
    # else -- subnormal float32 results, overflow -- goes through the generic neighbour comparison below
    mag = s.abs()
    normal = (mag >= 2.0 ** -126) & (mag < 2.0 ** 127)
    suspect = (((s.view(torch.int64) & 0x1FFFFFFF) == 0x10000000) & normal) | (~normal & torch.isfinite(s) & (mag > 0))
    out = r.numpy()
    cand = torch.nonzero(suspect).reshape(-1).numpy()
    if cand.size:
        sc, pc, cc = s.numpy()[cand], p.numpy()[cand], tc.numpy()[cand]
        bb = sc - pc
        err = (pc - (sc - bb)) + (cc - bb)         # TwoSum: the true value is sc + err, exactly
        rc = out[cand]
        rc64 = rc.astype(np.float64)
        other = np.where(rc64 > sc, np.nextafter(rc, np.float32(-np.inf)), np.nextafter(rc, np.float32(np.inf)))
        o64 = other.astype(np.float64)
        tie = (rc64 != sc) & (np.abs(rc64 - sc) == np.abs(o64 - sc)) & (err != 0)
        flip = tie & (np.sign(err) == np.sign(o64 - sc))  # float() went to the even neighbour; the truth lies on err's side
        out[cand[flip]] = other[flip]
    return out.reshape(shape)


def _requant(acc: np.ndarray, op: dict, res_q: Optional[np.ndarray]) -> np.ndarray:
    """acc [N, C, H, W] int64 (exact) -> int8 values as int32 array."""
    m = np.broadcast_to(op["m"].astype(f32).reshape(1, -1, 1, 1), acc.shape)
    b = np.broadcast_to(op["b"].astype(f32).reshape(1, -1, 1, 1), acc.shape)
    t = fma32(acc.astype(f32), m, b)   # int -> fp32 rounds to nearest even (|acc| can exceed 2^24); fma(float(acc), m[c], b[c])
    if res_q is not None:
        t = fma32(res_q.astype(f32), np.broadcast_to(f32(op["r"]), acc.shape), t)
    if op["relu"]:
        t = np.maximum(t, f32(0))
    return np.clip(np.rint(t), -127, 127).astype(np.int32)


def conv_int8(qx: np.ndarray, op: dict) -> np.ndarray:
    """Exact int32 accumulation: the products are integers below 2^14 and K <= 2^15 taps*channels, so float64 holds the
    sums exactly (< 2^53)."""
    w = torch.from_numpy(op["Wq"].astype(np.float64)).permute(0, 3, 1, 2).contiguous()  # OHWI -> OIHW
    y = F.conv2d(torch.from_numpy(qx.astype(np.float64)), w, None, stride=op["stride"], padding=op["pad"])
    acc = y.numpy()
    assert np.abs(acc).max() < 2 ** 31
    return np.rint(acc).astype(np.int64)


def int8_forward(lowered_q: dict, x: np.ndarray, keep: Optional[list] = None, start_from: Optional[Dict[str, np.ndarray]] = None):
    """Run the quantized graph on ``x`` [N, C, H, W].  fp16 parts follow the fp16 engine's numerics plan (as
    ``caffe_forward.lowered_forward_f16emu``).  ``start_from`` {tensor: values}: take these tensors as given (e.g. the
    GPU's fp16 ``pool1``) and skip the ops that produce them, so the INT8 part can be compared bit for bit downstream of a
    floating-point stem.  Returns (out [N, -1] float64, {tensor: ndarray}) -- INT8 tensors are returned as int32 arrays."""
    def r16(t):
        return t.to(torch.float16).to(torch.float64)

    scales = lowered_q["tensor_scales"]
    blobs: Dict[str, object] = {lowered_q["input"]: r16(torch.from_numpy(np.ascontiguousarray(x)).double())}
    given = dict(start_from or {})
    for k, v in given.items():
        blobs[k] = np.asarray(v).astype(np.int32) if k in scales else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
    snap = {}
    with torch.no_grad():
        for op in lowered_q["ops"]:
            if op["output"] in given:
                continue
            t = op["type"]
            a = blobs[op["input"]]
            if t == "quantize":
                h = a.numpy().astype(f32)                      # exact: fp16 values
                y = np.clip(np.rint(h * f32(op["inv_scale"])), -127, 127).astype(np.int32)
            elif t == "conv" and op.get("int8"):
                res = blobs[op["residual"]] if op["residual"] is not None else None
                y = _requant(conv_int8(a, op), op, res)
            elif t == "conv":
                w = r16(torch.from_numpy(op["W"]).double()).permute(0, 3, 1, 2).contiguous()
                y = F.conv2d(a, w, None, stride=op["stride"], padding=op["pad"])
                y = y + torch.from_numpy(op["bias"]).double().view(1, -1, 1, 1)
                if op["residual"] is not None:
                    y = y + blobs[op["residual"]]
                if op["relu"]:
                    y = torch.relu(y)
                y = r16(y)
            elif t == "maxpool":
                k, s, p = op["k"], op["stride"], op["pad"]
                ho = _pool_out(a.shape[2], k, p, s, op["ceil_mode"])
                wo = _pool_out(a.shape[3], k, p, s, op["ceil_mode"])
                need_h = (ho - 1) * s + k - a.shape[2] - p
                need_w = (wo - 1) * s + k - a.shape[3] - p
                y = F.max_pool2d(F.pad(a, (p, max(need_w, 0), p, max(need_h, 0)), value=float("-inf")), k, s)
            elif t == "avgpool":
                if "k_scale" in op:  # INT8 in, fp16 out
                    ssum = a.astype(np.int64).sum(axis=(2, 3), keepdims=True)
                    pf = ssum.astype(f32) * f32(op["k_scale"])
                    y = torch.from_numpy(pf.astype(np.float16).astype(np.float64))
                else:
                    y = r16(a.mean(dim=(2, 3), keepdim=True).float().double())
            elif t == "fc":
                W = r16(torch.from_numpy(op["W"]).double())
                flat = a.permute(0, 2, 3, 1).reshape(a.shape[0], -1)
                y = (flat @ W.t() + torch.from_numpy(op["bias"]).double()).float().double().view(a.shape[0], -1, 1, 1)
            elif t == "softmax":
                y = torch.softmax(a.float(), dim=1).double()
            else:
                raise ValueError(f"int8 oracle: unsupported op {t}")
            blobs[op["output"]] = y
            if keep and op["output"] in keep:
                snap[op["output"]] = y.copy() if isinstance(y, np.ndarray) else y.numpy().copy()
    out = blobs[lowered_q["output"]]
    if isinstance(out, np.ndarray):  # INT8 graph output: dequantised the way the output cast does
        out = (out.astype(f32) * f32(scales[lowered_q["output"]])).astype(np.float64)
    else:
        out = out.numpy()
    out = out.reshape(out.shape[0], -1)
    return (out, snap) if keep is not None else out
