#!/usr/bin/env python
"""Extract the reference's in-tree MNIST known-answer fixture into tests/golden/mnist_v1_3.npz.

Source (read-only, only available in the build container): /root/reference/models/onnx/mnist-v1.3/
{model.onnx, test_data_set_{0,1,2}/{input_0.pb,output_0.pb}} -- the vectors the reference's own
functional test asserts at decimal=3 (reference examples/30_PyTensorRT/server.py:19-31).
The npz carries: raw layer list (json), raw weights, the three inputs [1,1,28,28] and expected logits.
Run:  python tools/make_mnist_golden.py [/root/reference]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tensorrt_laboratory_b200 import onnx_lite  # noqa: E402


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    base = os.path.join(ref, "models/onnx/mnist-v1.3")
    model = onnx_lite.load_model(os.path.join(base, "model.onnx"))
    net, weights = onnx_lite.mnist_to_caffe_like(model)
    arrays = {"net_json": np.frombuffer(json.dumps(net).encode(), dtype=np.uint8)}
    for lname, rec in weights.items():
        for k, v in rec.items():
            arrays[f"w/{lname}/{k}"] = v
    for i in range(3):
        arrays[f"input_{i}"] = onnx_lite.load_tensor(os.path.join(base, f"test_data_set_{i}/input_0.pb"))
        arrays[f"output_{i}"] = onnx_lite.load_tensor(os.path.join(base, f"test_data_set_{i}/output_0.pb"))
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_v1_3.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", os.path.normpath(out), {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
