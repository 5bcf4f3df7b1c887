#!/usr/bin/env python
"""Phase accounting of the persistent warp-specialised conv tactic (conv_f16_tcgen05_ws) on one layer shape.
  python tools/gpu_ws_timing.py cin=64 h=56 cout=256 res=1 ws=148 bn=64 stages=2"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import builder, capi  # noqa: E402
from tests import helpers  # noqa: E402

kv = dict(a.split("=") for a in sys.argv[1:])
cin, h, cout, k = int(kv.pop("cin", 64)), int(kv.pop("h", 56)), int(kv.pop("cout", 256)), int(kv.pop("k", 1))
res, batch = int(kv.pop("res", 1)), int(kv.pop("batch", 8))
opts = {a: int(b) for a, b in kv.items()}
lib = capi.load()
lib.b2_context_debug_conv_timing.restype = C.c_int
lib.b2_context_debug_conv_timing.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
net, wts, low = helpers.conv_case(cin, h, h, cout, k, 1, k // 2, relu=True, residual=bool(res), seed=0)
x = np.random.default_rng(1).standard_normal((batch, cin, h, h), dtype=np.float32)
eng = capi.Engine(builder.build_plan(low, builder.PREC_FP16, batch))
sess = capi.Session(eng, dict(opts, graph=0, autotune=0))
sess.infer(x)
n = sess.nb_launches(batch)
buf = np.zeros((4096, 16), dtype=np.int64)
for i in range(n):
    name = lib.b2_context_launch_name(sess.ctx, batch, i).decode()
    if "ws=" not in name:
        continue
    for reps in (1, 4, 16):
        nct = C.c_int()
        capi.check(lib.b2_context_debug_conv_timing(sess.ctx, batch, i, reps, sess.stream.handle, buf.ctypes.data, 4096, C.byref(nct)))
        t = buf[:nct.value].astype(np.float64)
        tiles = np.maximum(t[:, 3], 1)
        med = lambda v: float(np.median(v))
        print(f"{name[13:70]:58s} reps={reps:2d} ctas={nct.value} | CTA life {med(t[:,2]-t[:,0]):7.0f} cyc  prologue {med(t[:,1]-t[:,0]):5.0f}  "
              f"tiles/grp0 {med(t[:,3]):.0f} | per tile (grp 0): wait_acc {med(t[:,4]/tiles):6.0f} wait_res {med(t[:,5]/tiles):6.0f} "
              f"store+barA {med(t[:,6]/tiles):6.0f} math {med(t[:,7]/tiles):6.0f} barB+issue {med(t[:,8]/tiles):6.0f} | "
              f"mma: wait_acc_free {med(t[:,9]):7.0f} wait_operands {med(t[:,10]):7.0f}  producer: wait_res_buf {med(t[:,11]):7.0f} wait_stage {med(t[:,12]):7.0f}")
sess.close()
eng.destroy()
