#!/usr/bin/env python
"""Sweep the caller-side knobs of the reference pipeline (contexts, buffers, cuda/post threads) on the e2e path."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402


def main():
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
    ring = weights.synthetic_input(8, ring=16)
    for arg in sys.argv[1:]:
        ctx, buf, cuda_t, post_t = [int(v) for v in arg.split(",")]
        mgr = capi.InferenceManager(ctx, buf, 1, cuda_t, post_t)
        mgr.register_model("rn50", blob)
        mgr.update_resources()
        mgr.prefill_inputs("rn50", ring[:buf])
        mgr.bench("rn50", 8, 600.0, int(os.environ.get("PROBE_WARMUP", "80")), False)
        res, lats = mgr.bench("rn50", 8, 600.0, 1500, True)
        print(json.dumps(dict(contexts=ctx, buffers=buf, cuda_threads=cuda_t, post_threads=post_t,
                              img_s=res["kInferencesPerSecond"], gpu_ms_per_batch=res["kGpuComputeTimePerBatch"] * 1e3,
                              depth=os.environ.get("TRTLAB_ENQUEUE_DEPTH", "2"), p50_ms=float(np.percentile(lats, 50) * 1e3),
                              p99_ms=float(np.percentile(lats, 99) * 1e3))), flush=True)
        mgr.close()


if __name__ == "__main__":
    main()
