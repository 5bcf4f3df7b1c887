#!/usr/bin/env python
"""Device-resident ResNet-50 fp16 throughput as a function of the batch size of one forward pass and the number of
concurrent contexts: what dynamic batching in front of the engine (BatchedInferRunner) can buy."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

for batch in (8, 16, 32, 64):
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, batch)
    ring = weights.synthetic_input(batch, ring=2)
    rec = {"batch": batch}
    for n in (1, 2, 4):
        steps = max(60, 1600 // batch)
        ms, _ = capi.device_throughput(blob, n, batch, steps, 10, ring)
        rec[f"ctx{n}_img_s"] = round(steps * batch / (ms * 1e-3))
    print(json.dumps(rec), flush=True)
