#!/usr/bin/env python
"""Throughput of ResNet-50 fp16 b=8 (inputs resident) as a function of (a) the number of streams the on-device
tactic autotuner loads the GPU with and (b) the number of concurrent ExecutionContexts.  Each autotune setting is
tuned once into its own tactic-cache file (B2_TUNE_CACHE), then replayed with every context count.
usage: python tools/gpu_ctx_sweep.py [autotune values, default 4,8,16] [context counts, default 1,2,4,8,16] [ENV=V,ENV=V ...]
Every further argument is a set of extra environment settings (e.g. B2_FORCE_CN=2) swept as its own variant."""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, graph, weights  # noqa: E402


def main():
    tunes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,8,16").split(",")]
    ctxs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8,16").split(",")]
    out_dir = os.environ.get("SWEEP_OUT", "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    net = graph.resnet_caffe(50)
    low = graph.lower(net, weights.random_weights(net, 0))
    blob = builder.build_plan(low, builder.PREC_FP16, 8)
    ring = weights.synthetic_input(8, ring=8)
    variants = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[3:]] or [{}]
    for var in variants:
      for t in tunes:
        for k in list(os.environ):
            if k.startswith("B2_FORCE_") or k.startswith("B2_PDL"):
                del os.environ[k]
        os.environ["B2_PDL"] = "1"  # process-wide switch: always restate it
        os.environ.update(var)
        tag = "".join(f"_{k[3:].lower()}{v}" for k, v in var.items())
        cache = os.path.join(out_dir, f"tactics_autotune{t}{tag}.txt")
        if os.path.exists(cache):
            os.remove(cache)
        os.environ["B2_TUNE_CACHE"] = cache
        os.environ["B2_AUTOTUNE"] = str(t)
        rec = {"autotune_streams": t, "env": var}
        for n in ctxs:
            steps = 100 * max(n, 2)
            ms, _ = capi.device_throughput(blob, n, 8, steps, 20, ring)
            rec[f"ctx{n}_img_s"] = round(steps * 8 / (ms * 1e-3))
        fields = [l.split() for l in open(cache) if not l.startswith("#")]
        rec["bn"] = dict(collections.Counter(f[3] for f in fields))
        rec["stages"] = dict(collections.Counter(f[4] + "x" + f[6] for f in fields))
        rec["splits"] = dict(collections.Counter(f[5] for f in fields))
        rec["cn"] = dict(collections.Counter(f[8] for f in fields))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
