"""CPU ORACLE -- test infrastructure only.  Never imported by the product package.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; it is the checker, never the thing measured as the product or shipped.

What it restates.  The reference (NVIDIA/tensorrt-laboratory) contains no arithmetic of its own for the
forward pass: ``ExecutionContext::Infer`` hands the bindings to closed-source TensorRT
(``trtlab/tensorrt/src/workspace.cc:47,52`` ``enqueueV2``; legacy contract
``examples/10_Internals/README.md:50-52``).  TensorRT 7.1 (``Dockerfile:6``
``nvcr.io/nvidia/tensorrt:20.06-py3``) is absent from /root/reference and cannot be run here, so the
oracle restates the PUBLISHED operator semantics of the model files the reference feeds it:

* Caffe deploy nets ``models/ResNet-{50,152}-deploy.prototxt`` (built by ``models/setup.py:53-55``):
  Convolution, BatchNorm(use_global_stats)+Scale, ReLU, Pooling (MAX/AVE, **ceil** output size),
  Eltwise SUM, InnerProduct, Softmax -- executed UNFUSED, layer by layer, in fp32 (or fp64).
* ONNX opset-8 MNIST ``models/onnx/mnist-v1.3/model.onnx`` (Conv SAME_UPPER, Add, Relu, MaxPool floor,
  Reshape, MatMul) -- expressed with the same layer vocabulary (``ceil_mode=False`` on its pools).

Pinning: the MNIST path is pinned against the reference's in-tree golden vectors
(``models/onnx/mnist-v1.3/test_data_set_{0,1,2}``, tolerance ``decimal=3`` as in
``examples/30_PyTensorRT/server.py:31``) by ``tests/test_oracle.py``.  For ResNet-50/152 the reference
holds NO golden vector in-tree (its ResNet fixtures are a network download,
``examples/ONNX/resnet50/fetch.sh:3-9``): **parity unpinned** for those graphs beyond the operator
family that MNIST exercises (conv+bias, relu, maxpool, matmul+bias).

Inputs are fp32 NCHW, the reference's binding contract (``trtlab/tensorrt/src/bindings.cc:128-175``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _pool_out(size, k, pad, stride, ceil_mode):
    if not ceil_mode:
        return (size + 2 * pad - k) // stride + 1
    out = int(math.ceil((size + 2 * pad - k) / stride)) + 1
    if pad > 0 and (out - 1) * stride >= size + pad:
        out -= 1
    return out


def caffe_forward(net: dict, weights: Dict[str, dict], x: np.ndarray, dtype=torch.float32,
                  keep: Optional[list] = None, threads: Optional[int] = None):
    """Run the raw layer list on ``x`` (N,C,H,W fp32).  Returns the last top as float64 ndarray;
    with ``keep=[blob names]`` returns (out, {blob: ndarray NCHW}) snapshot right after each blob's
    last in-place writer."""
    if threads:
        torch.set_num_threads(threads)
    blobs = {net["input"]: torch.from_numpy(np.ascontiguousarray(x)).to(dtype)}
    snap = {}
    layers = net["layers"]
    last_writer = {}
    for i, L in enumerate(layers):
        last_writer[L["tops"][0]] = i
    with torch.no_grad():
        for i, L in enumerate(layers):
            t = L["type"]
            name = L["name"]
            a = blobs[L["bottoms"][0]]
            if t == "Convolution":
                w = torch.from_numpy(weights[name]["W"]).to(dtype)
                b = torch.from_numpy(weights[name]["b"]).to(dtype) if L["bias_term"] else None
                y = F.conv2d(a, w, b, stride=L["stride"], padding=L["pad"])
            elif t == "BatchNorm":
                mean = torch.from_numpy(weights[name]["mean"]).to(dtype).view(1, -1, 1, 1)
                var = torch.from_numpy(weights[name]["var"]).to(dtype).view(1, -1, 1, 1)
                y = (a - mean) / torch.sqrt(var + L.get("eps", 1e-5))
            elif t == "Scale":
                y = a * torch.from_numpy(weights[name]["gamma"]).to(dtype).view(1, -1, 1, 1)
                if L.get("bias_term"):
                    y = y + torch.from_numpy(weights[name]["beta"]).to(dtype).view(1, -1, 1, 1)
            elif t == "ReLU":
                y = torch.relu(a)
            elif t == "Pooling":
                k, s, p = L["kernel_size"], L["stride"], L["pad"]
                cm = L.get("ceil_mode", True)
                ho = _pool_out(a.shape[2], k, p, s, cm)
                wo = _pool_out(a.shape[3], k, p, s, cm)
                # explicit right/bottom padding reproduces Caffe's clipped windows
                need_h = (ho - 1) * s + k - a.shape[2] - p
                need_w = (wo - 1) * s + k - a.shape[3] - p
                if L["pool"] == "MAX":
                    ap = F.pad(a, (p, max(need_w, 0), p, max(need_h, 0)), value=float("-inf"))
                    y = F.max_pool2d(ap, k, s)
                else:
                    if p or need_h > 0 or need_w > 0:
                        raise ValueError("oracle: only unpadded, exact AVE pooling is restated")
                    y = F.avg_pool2d(a, k, s)
                assert y.shape[2] == ho and y.shape[3] == wo
            elif t == "Eltwise":
                y = a
                for bname in L["bottoms"][1:]:
                    y = y + blobs[bname]
            elif t == "InnerProduct":
                w = torch.from_numpy(weights[name]["W"]).to(dtype)
                b = torch.from_numpy(weights[name]["b"]).to(dtype) if L["bias_term"] else None
                y = F.linear(a.reshape(a.shape[0], -1), w, b).view(a.shape[0], -1, 1, 1)
            elif t == "Softmax":
                y = torch.softmax(a, dim=1)
            else:
                raise ValueError(f"oracle: unsupported layer {t}")
            blobs[L["tops"][0]] = y
            if keep and L["tops"][0] in keep and last_writer[L["tops"][0]] == i:
                snap[L["tops"][0]] = y.double().numpy().copy()
    out = blobs[layers[-1]["tops"][0]]
    out = out.reshape(out.shape[0], -1).double().numpy()
    return (out, snap) if keep else out


# ------------------------------------------------------------------------------------------------
# fp16-rounding emulation of the ENGINE's numerics plan (separates kernel bugs from rounding)
# ------------------------------------------------------------------------------------------------

def lowered_forward_f16emu(lowered: dict, x: np.ndarray, keep: Optional[list] = None, round16: bool = True):
    """Execute *lowered* ops (folded fp32 W in OHWI, bias) with the rounding points of the fp16 engine:

    input -> fp16; weights -> fp16; conv accumulates in fp32 (here fp64, i.e. exact) then
    ``+bias (+residual) -> relu -> fp16``; max-pool exact in fp16; global avg-pool fp32 sum / HW -> fp16;
    FC fp16 weights, fp32 accumulate + fp32 bias -> fp32 logits; softmax fp32.
    ``round16=False`` evaluates the same fused graph exactly (fp64): used to check BN/Scale folding.
    """
    def r16(t):
        return t.to(torch.float16).to(torch.float64) if round16 else t.to(torch.float64)

    blobs = {lowered["input"]: r16(torch.from_numpy(np.ascontiguousarray(x)).double())}
    snap = {}
    with torch.no_grad():
        for op in lowered["ops"]:
            a = blobs[op["input"]]
            t = op["type"]
            if t == "conv":
                w = r16(torch.from_numpy(op["W"]).double()).permute(0, 3, 1, 2).contiguous()  # OHWI->OIHW
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
                y = r16(a.mean(dim=(2, 3), keepdim=True).float().double())
            elif t == "fc":
                c, h, w_ = op["in_chw"]
                W = r16(torch.from_numpy(op["W"]).double())  # [out, (h,w,c)]
                flat = a.permute(0, 2, 3, 1).reshape(a.shape[0], -1)
                y = (flat @ W.t() + torch.from_numpy(op["bias"]).double()).float().double()
                y = y.view(a.shape[0], -1, 1, 1)
            elif t == "softmax":
                y = torch.softmax(a.float(), dim=1).double()
            else:
                raise ValueError(t)
            blobs[op["output"]] = y
            if keep and op["output"] in keep:
                snap[op["output"]] = y.numpy().copy()
    out = blobs[lowered["output"]]
    out = out.reshape(out.shape[0], -1).numpy()
    return (out, snap) if keep else out
