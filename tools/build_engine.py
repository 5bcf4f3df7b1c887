#!/usr/bin/env python
"""Build a B2ENGINE plan file -- the step the reference performs with `trtexec` (reference models/setup.py:32-56,
examples/ONNX/resnet50/build.py:35-67).

  python tools/build_engine.py --model resnet50 --precision fp16 --batch 8 -o rn50_b8_fp16.plan
  python tools/build_engine.py --prototxt /path/ResNet-152-deploy.prototxt --precision fp16 --batch 32 -o rn152.plan
  python tools/build_engine.py --model mnist --precision fp32 --batch 1 -o mnist.plan
Weights: deterministic synthetic weights (the reference's benchmark engines are weightless too, models/README.md:6-7),
except MNIST which carries its real ONNX weights.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, graph, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", choices=["resnet50", "resnet152", "mnist"])
    ap.add_argument("--prototxt")
    ap.add_argument("--onnx", help="ONNX CNN classifier (Conv / BatchNormalization / Relu / Add / MaxPool / AveragePool / "
                                   "GlobalAveragePool / Flatten / Reshape / Gemm / MatMul / Softmax), e.g. an ONNX-zoo ResNet")
    ap.add_argument("--precision", choices=["fp16", "fp32"], default="fp16")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("-o", "--output", required=True)
    a = ap.parse_args()
    prec = builder.PREC_FP16 if a.precision == "fp16" else builder.PREC_FP32
    if a.prototxt:
        with open(a.prototxt) as f:
            net = graph.parse_prototxt(f.read())
        wts = weights.random_weights(net, a.seed)
    elif a.onnx:
        from tensorrt_laboratory_b200 import onnx_import, onnx_lite
        net, wts = onnx_import.import_onnx(onnx_lite.load_model(a.onnx), name=os.path.splitext(os.path.basename(a.onnx))[0])
    elif a.model == "mnist":
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from tests import helpers
        net, wts, _, _ = helpers.load_mnist_golden()
    else:
        net = graph.resnet_caffe(int(a.model[6:]))
        wts = weights.random_weights(net, a.seed)
    blob = builder.build_plan(graph.lower(net, wts), prec, a.batch)
    with open(a.output, "wb") as f:
        f.write(blob)
    print(f"wrote {a.output}: {len(blob) / 1e6:.1f} MB, {net['name']}, {a.precision}, max batch {a.batch}")


if __name__ == "__main__":
    main()
