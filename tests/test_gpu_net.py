"""The persistent whole-network kernel (net_f16_tcgen05) against the per-layer kernels and the oracle.

The persistent kernel runs the same products in the same fp32 order with the same epilogue arithmetic as
conv_f16_tcgen05, so its results must be BIT-IDENTICAL to the per-layer path -- whatever the number of CTAs, however
the tickets interleave, across repeated launches (the arrival counters are re-armed by the last CTA), for partial
batches, and when several contexts share the GPU (one persistent kernel each, as under InferenceManager).
"""
import threading

import numpy as np
import pytest

from oracle.caffe_forward import lowered_forward_f16emu
from tensorrt_laboratory_b200 import builder, capi, graph, weights
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rn50():
    net = graph.resnet_caffe(50)
    wts = weights.random_weights(net, 0)
    low = graph.lower(net, wts)
    x = weights.synthetic_input(8)
    blob = builder.build_plan(low, builder.PREC_FP16, 8)
    return dict(low=low, x=x, blob=blob)


def _names(sess, batch):
    return [capi.load().b2_context_launch_name(sess.ctx, batch, i).decode() for i in range(sess.nb_launches(batch))]


def test_net_kernel_is_selected_and_bit_identical_to_per_layer_kernels(gpu, rn50):
    eng = capi.Engine(rn50["blob"])
    ref_s = capi.Session(eng, {"net": 0, "autotune": 0})
    net_s = capi.Session(eng, {"net": 1})
    try:
        names = _names(net_s, 8)
        assert sum(n.startswith("net_tcgen05") for n in names) == 1, names
        assert sum(n.startswith("conv_tcgen05") for n in names) == 1, names  # the 7x7 stem keeps its own kernel
        assert not any(n.startswith("net_tcgen05") for n in _names(ref_s, 8))
        ref = ref_s.infer(rn50["x"])["prob"]
        got = net_s.infer(rn50["x"])["prob"]
        assert np.array_equal(ref, got)
        for _ in range(4):  # counters re-armed by the last CTA of every launch
            assert np.array_equal(net_s.infer(rn50["x"])["prob"], ref)
        # partial batches through the same context (their own plan and tables)
        for b in (1, 3, 5):
            assert np.array_equal(net_s.infer(rn50["x"][:b])["prob"], ref_s.infer(rn50["x"][:b])["prob"])
    finally:
        ref_s.close()
        net_s.close()
        eng.destroy()


@pytest.mark.parametrize("ctas", [1, 7, 37, 74, 148])
def test_net_kernel_result_independent_of_cta_count(gpu, rn50, ctas):
    eng = capi.Engine(rn50["blob"])
    a = capi.Session(eng, {"net": 1, "net_ctas": 148})
    b = capi.Session(eng, {"net": 1, "net_ctas": ctas})
    try:
        assert np.array_equal(a.infer(rn50["x"])["prob"], b.infer(rn50["x"])["prob"])
        assert f"ctas={ctas}" in " ".join(_names(b, 8))
    finally:
        a.close()
        b.close()
        eng.destroy()


def test_net_kernel_intermediate_tensors_match_oracle(gpu, rn50):
    taps = ["res2a", "res2c", "res3d", "res4f", "res5c"]
    x = rn50["x"][:2]
    _, snaps = lowered_forward_f16emu(rn50["low"], x, keep=taps)
    out = helpers.run_engine(rn50["low"], x, builder.PREC_FP16, options={"net": 1, "net_ctas": 37}, outputs=taps)
    assert any(n.startswith("net_tcgen05") for n in helpers.LAST_LAUNCH_NAMES)
    for name in taps:
        assert helpers.rel_err(out[name], snaps[name]) <= 2.0 ** -8, name


def test_net_kernel_narrow_tiles_bit_identical(gpu, rn50):
    eng = capi.Engine(rn50["blob"])
    a = capi.Session(eng, {"net": 1})
    b = capi.Session(eng, {"net": 1, "net_bn": 64})
    try:
        assert np.array_equal(a.infer(rn50["x"])["prob"], b.infer(rn50["x"])["prob"])
    finally:
        a.close()
        b.close()
        eng.destroy()


def test_four_contexts_share_the_gpu(gpu, rn50):
    """BASELINE configs[1]: four ExecutionContexts on four streams, one persistent kernel each."""
    eng = capi.Engine(rn50["blob"])
    sessions = [capi.Session(eng, {"net": 1, "net_ctas": 37}) for _ in range(4)]
    xs = [weights.synthetic_input(8, seed=100 + i) for i in range(4)]
    try:
        ref = [sessions[0].infer(x)["prob"] for x in xs]
        results = [[None] * 6 for _ in range(4)]

        def work(i):
            for k in range(6):
                results[i][k] = sessions[i].infer(xs[i])["prob"]

        ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for i in range(4):
            for k in range(6):
                assert np.array_equal(results[i][k], ref[i]), (i, k)
    finally:
        for s in sessions:
            s.close()
        eng.destroy()
