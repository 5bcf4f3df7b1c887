#!/usr/bin/env python
"""ResNet-152 fp16 batch 32 (the network and batch of BASELINE.json configs[2], in fp16 -- the INT8 kernels do not exist
yet) device-resident throughput at 1/2/4/8 contexts."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

t0 = time.time()
blob = builder.build_resnet_plan(152, builder.PREC_FP16, 32)
ring = weights.synthetic_input(32, ring=2)
rec = {"model": "ResNet-152 fp16", "batch": 32, "plan_MB": round(len(blob) / 1e6, 1)}
for n in (1, 2, 4, 8):
    steps = 40 * max(n, 2)
    ms, launches = capi.device_throughput(blob, n, 32, steps, 6, ring)
    rec[f"ctx{n}_img_s"] = round(steps * 32 / (ms * 1e-3))
    rec["launches_per_forward"] = launches
rec["wall_s"] = round(time.time() - t0, 1)
print(json.dumps(rec), flush=True)
