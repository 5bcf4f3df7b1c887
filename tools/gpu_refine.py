"""Network-level tactic refinement (b2_engine_refine_tactics) of ResNet-50 fp16 batch 8 for 4 contexts: before / after
device-resident throughput (both from plans that CARRY their tactic table, so nothing is tuned in the timed runs)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

CTX = int(os.environ.get("REFINE_CTX", "4"))
PASSES = int(os.environ.get("REFINE_PASSES", "1"))
blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8, seed=0)
ring = weights.synthetic_input(8, seed=1234, ring=32)
eng = capi.Engine(blob)
t0 = time.perf_counter()
eng.tune(streams=CTX)
t_tune = time.perf_counter() - t0
base_tactics = eng.tactics().copy()
t0 = time.perf_counter()
gain = eng.refine_tactics(streams=CTX, passes=PASSES)
t_refine = time.perf_counter() - t0
refined = eng.tactics().copy()
eng.destroy()
changed = int((base_tactics != refined).any(axis=1).sum())
np.save("gpurun_out/tactics_rn50_b8_base.npy", base_tactics)
np.save("gpurun_out/tactics_rn50_b8_refined.npy", refined)
out = {"contexts": CTX, "tune_s": t_tune, "refine_s": t_refine, "refine_reported_gain": gain, "layers_changed": changed}
for tag, tac in (("base", base_tactics), ("refined", refined), ("base2", base_tactics), ("refined2", refined)):
    b = builder.attach_tactics(blob, tac)
    ms, _ = capi.device_throughput(b, CTX, 8, 1500, 40, ring)
    out[tag + "_img_per_s"] = 1500 * 8 / (ms * 1e-3)
print(json.dumps(out))
for a, b in zip(base_tactics, refined):
    if (a != b).any():
        print("op", a[0], "bn/st/sps/halo", a[2], a[3], a[5], a[8], "->", b[2], b[3], b[5], b[8])
