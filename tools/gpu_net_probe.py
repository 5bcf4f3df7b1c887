#!/usr/bin/env python
"""GPU bring-up probe for whole networks (each experiment in its own subprocess, JSON lines to
gpurun_out/net_probe.jsonl): MNIST known-answer (fp32 engine), ResNet-50 fp32 SIMT engine vs fp32 oracle,
ResNet-50 fp16 tcgen05 engine vs fp16-emulating oracle with intermediate tensors, per-layer timing."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
TAPS = ["conv1", "pool1", "res2a", "res2c", "res3a", "res3d", "res4a", "res4f", "res5a", "res5c", "pool5", "fc1000", "prob"]


def exp_mnist():
    import numpy as np
    from tensorrt_laboratory_b200 import builder, graph
    from tests import helpers
    net, w, xs, ys = helpers.load_mnist_golden()
    low = graph.lower(net, w)
    res = []
    for x, y in zip(xs, ys):
        out = helpers.run_engine(low, x, builder.PREC_FP32)
        got = list(out.values())[0]
        res.append(dict(max_abs=float(np.abs(got - y).max()), argmax=int(got.argmax()), expect=int(y.argmax())))
    # batch of 3 in one call
    out = helpers.run_engine(low, np.concatenate(xs, 0), builder.PREC_FP32)
    got = list(out.values())[0]
    res.append(dict(batched_max_abs=float(np.abs(got - np.concatenate(ys, 0)).max())))
    return dict(exp="mnist_fp32", results=res)


def _resnet(precision, batch, options, emu):
    import numpy as np
    from oracle.caffe_forward import caffe_forward, lowered_forward_f16emu
    from tensorrt_laboratory_b200 import builder, graph, weights
    from tests import helpers
    net = graph.resnet_caffe(50)
    wts = weights.random_weights(net, 0)
    low = graph.lower(net, wts)
    x = weights.synthetic_input(batch)
    t0 = time.time()
    if emu:
        ref, snaps = lowered_forward_f16emu(low, x, keep=TAPS)
    else:
        ref, snaps = caffe_forward(net, wts, x, keep=TAPS)
    t_oracle = time.time() - t0
    out = helpers.run_engine(low, x, precision, options, outputs=TAPS)
    rows = {}
    for name in TAPS:
        got = out[name].reshape(batch, -1)
        want = snaps[name].reshape(batch, -1)
        rows[name] = dict(rel=helpers.rel_err(got, want), max_abs=float(np.abs(got - want).max()), ref_max=float(np.abs(want).max()))
    prob = out["prob"]
    return dict(oracle_s=t_oracle, taps=rows, argmax_gpu=[int(v) for v in prob.argmax(1)],
                argmax_ref=[int(v) for v in ref.argmax(1)], pmax_gpu=[float(v) for v in prob.max(1)],
                pmax_ref=[float(v) for v in ref.max(1)])


def exp_rn50_fp32():
    return dict(exp="rn50_fp32_simt", **_resnet(0, 1, {}, emu=False))


def exp_rn50_fp16_simt():
    return dict(exp="rn50_fp16_simt", **_resnet(1, 1, {"simt": 1}, emu=True))


def exp_rn50_fp16():
    return dict(exp="rn50_fp16_tcgen05", **_resnet(1, 2, {}, emu=True))


def exp_rn50_fp16_nograph():
    return dict(exp="rn50_fp16_tcgen05_nograph", **_resnet(1, 2, {"graph": 0}, emu=True))


def exp_profile(options=None, tag="profile_b8"):
    import numpy as np
    from tensorrt_laboratory_b200 import builder, capi, weights
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
    eng = capi.Engine(blob)
    sess = capi.Session(eng, options)
    x = weights.synthetic_input(8)
    sess.infer(x)
    sess.infer(x)
    prof = sess.profile(8)
    prof = sess.profile(8)
    # graph-replay timing, single stream
    ev0, ev1 = capi.Event(), capi.Event()
    for _ in range(5):
        sess.enqueue(8)
    sess.stream.sync()
    ev0.record(sess.stream)
    iters = 50
    for _ in range(iters):
        sess.enqueue(8)
    ev1.record(sess.stream)
    sess.stream.sync()
    ms = ev0.elapsed_ms(ev1) / iters
    total = sum(p["ms"] for p in prof)
    conv = sum(p["ms"] for p in prof if p["name"].startswith("conv"))
    sess.close()
    return dict(exp=tag, options=options, graph_ms_per_batch=ms, img_per_s=8 / ms * 1e3, serial_sum_ms=total, conv_ms=conv,
                tflops_graph=eng.flops(8) / ms / 1e9, layers=prof)


def _variant(name, opts):
    def f():
        return exp_profile(opts, name)
    f.__name__ = "exp_" + name
    return f


VARIANTS = {
    "p_nopdl": {"pdl": 0},
    "p_trig0": {"pdl": 1, "pdl_trigger": 0},
    "p_nosplit": {"splits": 1},
    "p_bn64": {"bn": 64},
    "p_bn64_nosplit": {"bn": 64, "splits": 1},
    "p_bn128": {"bn": 128},
    "p_st4": {"stages": 4},
    "p_st4_nosplit": {"stages": 4, "splits": 1},
}
EXPS = {f.__name__[4:]: f for f in (exp_mnist, exp_rn50_fp32, exp_rn50_fp16_simt, exp_rn50_fp16, exp_rn50_fp16_nograph, exp_profile)}
EXPS.update({k: _variant(k, v) for k, v in VARIANTS.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--exp":
        print(json.dumps(EXPS[sys.argv[2]]()))
        return
    names = sys.argv[1:] or list(EXPS)
    with open(os.path.join(OUT, "net_probe.jsonl"), "a") as log:
        for name in names:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "--exp", name], capture_output=True, text=True, timeout=600)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                rec = json.loads(line[-1]) if line else dict(exp=name, error=(r.stderr or r.stdout)[-2000:], rc=r.returncode)
            except subprocess.TimeoutExpired:
                rec = dict(exp=name, error="timeout")
            rec["wall"] = time.time() - t0
            log.write(json.dumps(rec) + "\n")
            log.flush()
            s = json.dumps({k: v for k, v in rec.items() if k != "layers"})
            print(s[:1500])


if __name__ == "__main__":
    main()
