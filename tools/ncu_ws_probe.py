#!/usr/bin/env python
"""One wide 1x1 convolution with a residual (res2 branch2c shape: 64 -> 256 channels, 56x56, batch 8) through the C ABI with
a forced tactic, for an ncu capture of exactly that kernel.
  PROBE_OPTS="ws=148,bn=64,stages=2" python tools/ncu_ws_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi  # noqa: E402
from tests import helpers  # noqa: E402

opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("PROBE_OPTS", "ws=148,bn=64,stages=2").split(","))}
cin, h, cout, batch = 64, 56, 256, 8
net, wts, low = helpers.conv_case(cin, h, h, cout, 1, 1, 0, relu=True, residual=True, seed=0)
x = np.random.default_rng(1).standard_normal((batch, cin, h, h), dtype=np.float32)
blob = builder.build_plan(low, builder.PREC_FP16, batch)
eng = capi.Engine(blob)
sess = capi.Session(eng, dict(opts, graph=0, autotune=0))
for _ in range(int(os.environ.get("PROBE_ITERS", "6"))):
    out = sess.infer(x)
n = sess.nb_launches(batch)
print([capi.load().b2_context_launch_name(sess.ctx, batch, i).decode() for i in range(n)])
sess.close()
eng.destroy()
