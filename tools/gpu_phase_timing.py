#!/usr/bin/env python
"""Per-CTA phase timing of selected conv layers (debug instrumentation in conv_f16_tcgen05).
Phases (SM clock cycles, median over CTAs): prologue, wait-for-dependency, first TMA landed, main loop,
accumulator ready, epilogue, teardown; plus the kernel's wall span from globaltimer."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi, weights  # noqa: E402


def main():
    lib = capi.load()
    lib.b2_context_debug_conv_timing.restype = C.c_int
    lib.b2_context_debug_conv_timing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
    eng = capi.Engine(blob)
    opts = {}
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        opts[k] = int(v)
    sess = capi.Session(eng, opts)
    sess.infer(weights.synthetic_input(8))
    n = sess.nb_launches(8)
    cap = 4096
    buf = np.zeros((cap, 16), dtype=np.int64)
    for i in range(n):
        name = lib.b2_context_launch_name(sess.ctx, 8, i).decode()
        if not name.startswith("conv_tcgen05"):
            continue
        nct = C.c_int()
        for reps in (1, 4):
            rc = lib.b2_context_debug_conv_timing(sess.ctx, 8, i, reps, sess.stream.handle, buf.ctypes.data, cap, C.byref(nct))
            if rc:
                print(name, "ERR", lib.b2_last_error().decode())
                break
            t = buf[:nct.value]
            d = lambda a, b: float(np.median(t[:, b] - t[:, a]))
            span_us = (t[:, 9].max() - t[:, 8].min()) / 1e3
            cta_us = float(np.median(t[:, 9] - t[:, 8])) / 1e3
            start_spread = (t[:, 8].max() - t[:, 8].min()) / 1e3
            sms = len(np.unique(t[:, 10]))
            print(f"{name[13:60]:48s} reps={reps} ctas={nct.value:4d} sms={sms:3d} span={span_us:6.2f}us cta={cta_us:5.2f}us startspread={start_spread:5.2f}us | "
                  f"prolog={d(0,1):6.0f} depwait={d(1,2):6.0f} firstTMA={d(2,3):6.0f} mainloop={d(3,4):7.0f} accum={d(4,5):6.0f} epi={d(5,6):6.0f} tear={d(6,7):6.0f} cyc"
                  f" | prod: wait_empty={np.median(t[:,11]):7.0f} expect={np.median(t[:,12]):6.0f} tma={np.median(t[:,13]):6.0f}  mma: wait_full={np.median(t[:,14]):7.0f} issue+commit={np.median(t[:,15]):6.0f}")
    sess.close()


if __name__ == "__main__":
    main()
