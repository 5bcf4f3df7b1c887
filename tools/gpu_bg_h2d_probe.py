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
for bg, st in (("0", "0"), ("0", "1"), ("1", "1"), ("0", "0"), ("0", "1")):
    os.environ["B2_PROBE_BG_H2D"] = bg
    os.environ["B2_PROBE_STAGGER"] = st
    for n in (2, 3, 4):
        ms, _ = capi.device_throughput(blob, n, 8, 800, 20, ring)
        print(json.dumps({"bg_h2d": bg, "stagger": st, "contexts": n, "img_s": round(800 * 8 / (ms * 1e-3))}), flush=True)
