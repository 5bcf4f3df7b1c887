"""Invariant check: a TUNED engine, an untuned (cost-model) engine and the InferenceManager pipeline (tuned at registration,
two lanes, dynamic batcher) of the same plan must agree bit for bit, whatever tactics the load-time tuner picked in this
process.  Loops `n_tuned` engines and `n_managers` managers and prints the unusual tactics (split-K / persistent / halo) of
each.  Written to chase a once-seen 4e-3 difference that turned out to be split-K being chosen by the timing.
  python tools/gpu_tactic_invariance.py [n_tuned=12] [n_managers=6]"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from tensorrt_laboratory_b200 import builder, capi, weights
blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8, seed=0)
x = weights.synthetic_input(8)
e0 = capi.Engine(blob); s0 = capi.Session(e0); d0 = s0.infer(x)["prob"].copy()
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    e1 = capi.Engine(blob); e1.tune(streams=4); s1 = capi.Session(e1); d1 = s1.infer(x)["prob"].copy()
    t = e1.tactics()
    eq = np.array_equal(d0, d1)
    odd = t[(t[:, 4] > 1) | (t[:, 6] > 0) | (t[:, 8] > 0)]
    print(it, "tuned==untuned", eq, "odd tactics (splits/ws/halo):", odd.tolist(), flush=True)
    if not eq:
        bad += 1
        for key, opts in (("tail0", {"fuse_tail": 0}),):
            s = capi.Session(e1, opts); print("   ", key, np.array_equal(s.infer(x)["prob"], d0)); s.close()
    s1.close(); e1.destroy()
mg_bad = 0
xx = np.concatenate([x, x[:5]])
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    mgr = capi.InferenceManager(max_exec_concurrency=2, max_copy_concurrency=4)
    mgr.register_model("rn50", blob); mgr.update_resources()
    got, nb = mgr.infer_batched("rn50", xx, window_us=20000)
    ok = np.array_equal(got[:8], d0) and np.array_equal(got[8:], d0[:5])
    one = np.array_equal(mgr.infer("rn50", x), d0)
    print("manager", it, "batched ok", ok, "single ok", one, nb, flush=True)
    mg_bad += (not ok)
    mgr.close()
print("bad tuned:", bad, "bad manager:", mg_bad)
