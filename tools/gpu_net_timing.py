"""Per-role wait/busy cycle counters of the persistent network kernel (instrumented instantiation), RN50 b=8."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402

ROLES = ["A-producer", "MMA", "B-producer", "scheduler", "epilogue(w4)"]
FIELDS = {
    "A-producer": ["sched", "dep", "stage-empty", "res-empty", "tiles", "-"],
    "MMA": ["sched", "full", "acc-empty", "-", "-", "-"],
    "B-producer": ["sched", "stage-empty", "-", "-", "-", "-"],
    "scheduler": ["-", "atomic", "ring-empty", "-", "-", "-"],
    "epilogue(w4)": ["sched", "acc-full", "tmem->global", "fence+publish", "-", "-"],
}


def main():
    ctas_list = [int(x) for x in (sys.argv[1:] or ["296", "74"])]
    lib = capi.load()
    lib.b2_context_debug_net_timing.restype = C.c_int
    lib.b2_context_debug_net_timing.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8, seed=0)
    x = weights.synthetic_input(8)
    eng = capi.Engine(blob)
    for ctas in ctas_list:
        sess = capi.Session(eng, {"net": 1, "net_ctas": ctas, "graph": 0})
        sess.infer(x)
        sess.infer(x)
        buf = np.zeros((296, 8, 8), np.int64)
        n = C.c_int(0)
        capi.check(lib.b2_context_debug_net_timing(sess.ctx, 8, sess._ptrs, sess.stream.handle, buf.ctypes.data, 296, C.byref(n)))
        d = buf[: n.value]
        print(f"== ctas={n.value}: kernel cycles (mean over CTAs of the role's lifetime) ==")
        for r, role in enumerate(ROLES):
            tot = d[:, r, 6].mean()
            parts = ", ".join(f"{name} {d[:, r, k].mean():.0f} ({100 * d[:, r, k].mean() / max(tot, 1):.0f}%)"
                              for k, name in enumerate(FIELDS[role]) if name != "-")
            print(f"  {role:13s} total {tot:9.0f} cyc : {parts}")
        print(f"  tiles per CTA: mean {d[:, 0, 4].mean():.1f} min {d[:, 0, 4].min()} max {d[:, 0, 4].max()}")
        sess.close()
    eng.destroy()


if __name__ == "__main__":
    main()
