#!/usr/bin/env python
"""Per-launch event timing of one forward pass (b2_context_profile: direct launches, one event pair each, single stream),
grouped by kernel kind -- where the non-convolution time of a plan goes.
  PROFILE_DEPTH=152 PROFILE_BATCH=32 PROFILE_PREC=int8 python tools/gpu_launch_profile.py"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

depth = int(os.environ.get("PROFILE_DEPTH", "50"))
batch = int(os.environ.get("PROFILE_BATCH", "8"))
prec = builder.PREC_INT8 if os.environ.get("PROFILE_PREC", "fp16") == "int8" else builder.PREC_FP16
eng = capi.Engine(builder.build_resnet_plan(depth, prec, batch))
eng.tune(streams=int(os.environ.get("PROFILE_TUNE_STREAMS", "4")))
sess = capi.Session(eng, {"graph": 0})
sess.infer(weights.synthetic_input(batch))
runs = [sess.profile(batch) for _ in range(5)]
best = [min(r[i]["ms"] for r in runs) for i in range(len(runs[0]))]
kinds = collections.OrderedDict()
for rec, ms in zip(runs[0], best):
    kind = rec["name"].split(":")[0]
    kinds.setdefault(kind, [0, 0.0])
    kinds[kind][0] += 1
    kinds[kind][1] += ms
total = sum(best)
print(f"ResNet-{depth} {os.environ.get('PROFILE_PREC', 'fp16')} batch {batch}: {len(best)} launches, {total * 1e3:.1f} us serial")
for k, (n, ms) in kinds.items():
    print(f"  {k:18s} x{n:3d}  {ms * 1e3:8.1f} us  {100 * ms / total:5.1f} %")
for rec, ms in zip(runs[0], best):
    if not rec["name"].startswith("conv"):
        print(f"    {rec['name'][:70]:70s} {ms * 1e3:7.1f} us")
sess.close()
eng.destroy()
