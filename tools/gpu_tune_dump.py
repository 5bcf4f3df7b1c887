#!/usr/bin/env python
"""Dump the tuner's view of every convolution of a ResNet at one batch size: time per launch of every candidate tactic
with `streams` concurrent copies of the layer (B2_TUNE_VERBOSE=2).  The sum over layers of the winners is the forward
pass's cost if nothing but the per-layer saturated throughput mattered.
  TUNE_DEPTH=50 TUNE_BATCH=8 TUNE_STREAMS=4 python tools/gpu_tune_dump.py > gpurun_out/tune_dump.log 2>&1"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2_TUNE_VERBOSE", "2")
from tensorrt_laboratory_b200 import builder, capi  # noqa: E402

depth = int(os.environ.get("TUNE_DEPTH", "50"))
batch = int(os.environ.get("TUNE_BATCH", "8"))
streams = int(os.environ.get("TUNE_STREAMS", "4"))
prec = {"fp16": builder.PREC_FP16, "int8": builder.PREC_INT8}[os.environ.get("TUNE_PREC", "fp16")]
eng = capi.Engine(builder.build_resnet_plan(depth, prec, batch))
n = eng.tune(streams=streams)
print(f"tuned {n} tactics, depth {depth}, batch {batch}, streams {streams}", file=sys.stderr)
