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
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
    eng = capi.Engine(blob)
    sess = capi.Session(eng, opts)
    x = weights.synthetic_input(8)
    sess.host_array(0, 8)[...] = x
    sess.h2d(8)
    for _ in range(3):  # builds the plan, runs the autotuner, warms caches -- not profiled
        sess.enqueue(8)
        sess.stream.sync()
    capi.check(capi.load().b2_profiler_start())
    for _ in range(passes):
        sess.enqueue(8)
        sess.stream.sync()
    capi.check(capi.load().b2_profiler_stop())
    n = sess.nb_launches(8)
    names = [capi.load().b2_context_launch_name(sess.ctx, 8, i).decode() for i in range(n)]
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "launch_names.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    sess.close()
    eng.destroy()


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main()
