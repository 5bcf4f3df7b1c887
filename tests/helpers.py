"""Shared test helpers: build small engines through the public builder and run them through the C ABI."""
from __future__ import annotations

import json
import os

import numpy as np

from tensorrt_laboratory_b200 import builder, capi, graph, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_mnist_golden():
    z = np.load(os.path.join(GOLDEN, "mnist_v1_3.npz"))
    net = json.loads(bytes(z["net_json"]).decode())
    w = {}
    for k in z.files:
        if k.startswith("w/"):
            _, lname, field = k.split("/")
            w.setdefault(lname, {})[field] = z[k]
    inputs = [z[f"input_{i}"] for i in range(3)]
    outputs = [z[f"output_{i}"] for i in range(3)]
    return net, w, inputs, outputs


def conv_case(cin, h, w, cout, k, stride, pad, relu=True, residual=False, seed=0):
    """-> (raw net, raw weights, lowered graph with folded weights)"""
    net = builder.single_conv_net(cin, h, w, cout, k, stride, pad, relu=relu, residual=residual)
    wts = weights.random_weights(net, seed)
    return net, wts, graph.lower(net, wts)


LAST_LAUNCH_NAMES = []  # launch names (kernel + tactic) of the most recent run_engine call


def run_engine(lowered, x, precision, options=None, outputs=None, max_batch=None):
    blob = builder.build_plan(lowered, precision, max_batch or x.shape[0], outputs=outputs)
    eng = capi.Engine(blob)
    sess = capi.Session(eng, options)
    try:
        out = sess.infer(x)
        n = sess.nb_launches(x.shape[0])
        LAST_LAUNCH_NAMES[:] = [capi.load().b2_context_launch_name(sess.ctx, x.shape[0], i).decode() for i in range(n)]
    finally:
        sess.close()
        eng.destroy()
    return out


def rel_err(a, b):
    """max |a-b| / max|b|  (scale-relative error)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
