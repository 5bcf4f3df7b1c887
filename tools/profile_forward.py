#!/usr/bin/env python
"""Run ResNet-50 fp16 batch-8 forward passes with direct launches (no CUDA graph) so that ncu sees every kernel.
Plan building + tactic autotuning + warm-up happen BEFORE cudaProfilerStart, so profile with
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/profile_forward.py 1
Launch order of one pass: input_cast, conv1, pool1, 52 convs, pool5, fc1000, prob  (58 kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    opts = {"graph": 0}
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        opts[k] = int(v)
    depth = int(os.environ.get("PROFILE_DEPTH", "50"))
    batch = int(os.environ.get("PROFILE_BATCH", "8"))
    prec = builder.PREC_INT8 if os.environ.get("PROFILE_PREC", "fp16") == "int8" else builder.PREC_FP16
    blob = builder.build_resnet_plan(depth, prec, batch)
    eng = capi.Engine(blob)
    eng.tune(streams=int(os.environ.get("PROFILE_TUNE_STREAMS", "4")))  # tactics are timed at load (B2_TUNE_CACHE pins them)
    sess = capi.Session(eng, opts)
    x = weights.synthetic_input(batch)
    sess.host_array(0, batch)[...] = x
    sess.h2d(batch)
    for _ in range(3):  # builds the plan, warms caches -- not profiled
        sess.enqueue(batch)
        sess.stream.sync()
    capi.check(capi.load().b2_profiler_start())
    for _ in range(passes):
        sess.enqueue(batch)
        sess.stream.sync()
    capi.check(capi.load().b2_profiler_stop())
    n = sess.nb_launches(batch)
    names = [capi.load().b2_context_launch_name(sess.ctx, batch, i).decode() for i in range(n)]
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "launch_names.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    sess.close()
    eng.destroy()


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main()
