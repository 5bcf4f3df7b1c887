#!/usr/bin/env python
"""Device-resident ResNet-50 throughput (4 contexts) with and without an unrelated pinned->device copy loop running:
how much does PCIe input traffic alone slow the forward passes down?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
ring = weights.synthetic_input(8, ring=8)
os.environ["B2_PROBE_TINY_H2D"] = "0"
for bg in ("0", "1", "64", "0", "64"):
    os.environ["B2_PROBE_BG_H2D"] = bg
    ms, _ = capi.device_throughput(blob, 4, 8, 1600, 20, ring)
    print(json.dumps({"background_h2d_burst": bg, "contexts": 4, "img_s": round(1600 * 8 / (ms * 1e-3))}), flush=True)
