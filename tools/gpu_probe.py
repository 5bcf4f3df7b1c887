#!/usr/bin/env python
"""GPU bring-up probe: runs each experiment in its own subprocess (a trapped kernel poisons the CUDA
context) with a timeout, and writes a JSON line per experiment to gpurun_out/probe.jsonl."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

CASES = {
    # name: (cin,h,w,cout,k,stride,pad,relu,residual,batch, options)
    "c1x1_tiled": (64, 56, 56, 64, 1, 1, 0, True, False, 2, {}),
    "c1x1_im2col": (64, 56, 56, 64, 1, 1, 0, True, False, 2, {"im2col": 1}),
    "c1x1_k256_tiled": (256, 28, 28, 128, 1, 1, 0, True, False, 2, {}),
    "c1x1_s2": (256, 56, 56, 128, 1, 2, 0, True, False, 2, {}),
    "c3x3": (64, 56, 56, 64, 3, 1, 1, True, False, 2, {}),
    "c3x3_res": (128, 28, 28, 128, 3, 1, 1, True, True, 2, {}),
    "c3x3_tail": (512, 7, 7, 512, 3, 1, 1, True, False, 1, {}),
    "c7x7_stem": (3, 224, 224, 64, 7, 2, 3, True, False, 2, {}),
    "c1x1_bn32": (64, 56, 56, 64, 1, 1, 0, True, False, 2, {"bn": 32}),
    "c1x1_bn64": (64, 56, 56, 64, 1, 1, 0, True, False, 2, {"bn": 64}),
    "c1x1_bn128": (64, 56, 56, 256, 1, 1, 0, False, False, 2, {"bn": 128}),
    "c1x1_bn256": (64, 56, 56, 256, 1, 1, 0, False, False, 2, {"bn": 256}),
    "c1x1_simt": (64, 56, 56, 64, 1, 1, 0, True, False, 2, {"simt": 1}),
    "c3x3_nograph": (64, 56, 56, 64, 3, 1, 1, True, False, 2, {"graph": 0}),
}


def run_case(name):
    import numpy as np
    from oracle.caffe_forward import lowered_forward_f16emu
    from tensorrt_laboratory_b200 import builder
    from tests import helpers
    cin, h, w, cout, k, s, p, relu, res, batch, opts = CASES[name]
    net, wts, low = helpers.conv_case(cin, h, w, cout, k, s, p, relu, res)
    x = np.random.default_rng(1).standard_normal((batch, cin, h, w), dtype=np.float32)
    ref = lowered_forward_f16emu(low, x)
    t0 = time.time()
    out = helpers.run_engine(low, x, builder.PREC_FP16, opts)
    got = list(out.values())[0].reshape(batch, -1)
    err = np.abs(got - ref)
    bad = int((err > 2e-2 * max(np.abs(ref).max(), 1)).sum())
    rec = dict(case=name, max_abs=float(err.max()), ref_max=float(np.abs(ref).max()), rel=helpers.rel_err(got, ref),
               bad=bad, n=int(ref.size), secs=time.time() - t0)
    if bad:
        idx = np.argwhere(err > 2e-2 * max(np.abs(ref).max(), 1))[:8]
        rec["first_bad"] = [[int(i) for i in ix] + [float(got[tuple(ix)]), float(ref[tuple(ix)])] for ix in idx]
        # summarise by output channel / pixel to expose layout bugs
        co = low["ops"][-1]["cout"]
        e3 = err.reshape(batch, co, -1)
        rec["bad_by_channel"] = [int(v) for v in (e3 > 2e-2 * np.abs(ref).max()).sum(axis=(0, 2))[:64]]
        px = (e3 > 2e-2 * np.abs(ref).max()).sum(axis=1).reshape(batch, -1)
        rec["bad_pixels_first"] = [int(v) for v in np.argwhere(px[0] > 0)[:32, 0]]
    print(json.dumps(rec))


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        return
    names = sys.argv[1:] or list(CASES)
    with open(os.path.join(OUT, "probe.jsonl"), "a") as log:
        for name in names:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=180)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                rec = json.loads(line[-1]) if line else dict(case=name, error=(r.stderr or r.stdout)[-1500:], rc=r.returncode)
            except subprocess.TimeoutExpired:
                rec = dict(case=name, error="timeout")
            rec["wall"] = time.time() - t0
            log.write(json.dumps(rec) + "\n")
            log.flush()
            print(json.dumps(rec)[:600])


if __name__ == "__main__":
    main()
