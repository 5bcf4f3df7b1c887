"""Round-2 GPU tests: the parity holes the round-1 review listed, and the ahead-of-time engine work
(tactics timed at load or carried by the plan, binding-independent pre-instantiated graphs, fused classifier tail)."""
import time

import numpy as np
import pytest

from oracle.caffe_forward import caffe_forward, lowered_forward_f16emu
from tensorrt_laboratory_b200 import builder, capi, graph, onnx_import, onnx_lite, weights
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rn50(gpu):
    net = graph.resnet_caffe(50)
    wts = weights.random_weights(net, 0)
    low = graph.lower(net, wts)
    x = weights.synthetic_input(8)
    blob = builder.build_plan(low, builder.PREC_FP16, 8)
    eng = capi.Engine(blob)
    sess = capi.Session(eng)
    direct = sess.infer(x)["prob"]
    emu = lowered_forward_f16emu(low, x)
    yield dict(net=net, wts=wts, low=low, x=x, blob=blob, eng=eng, sess=sess, direct=direct, emu=emu)
    sess.close()
    eng.destroy()


# ------------------------------------------------------------------------------------------------------------------
# parity holes ("Next round" item 1)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("managed", [False, True])
def test_v2_workspace_output_matches_oracle_and_direct_path(rn50, managed):
    """StaticSingleModelGraphWorkspace / BenchmarkWorkspace: the CALLER captures b2_context_enqueue into its own graph
    (reference workspace.cc:51-56,75); `prob` read back through async_d2h must be the direct path's, bit for bit, and
    within the north-star tolerance of the oracle.  managed=True loads the weights through ManagedRuntime."""
    prob = capi.workspace_infer(rn50["blob"], rn50["x"], managed_runtime=managed, iters=3)
    np.testing.assert_array_equal(prob, rn50["direct"])
    assert (prob.argmax(1) == rn50["emu"].argmax(1)).all()
    assert (np.abs(prob - rn50["emu"]) / rn50["emu"].max(1, keepdims=True)).max() <= 1e-4


@pytest.mark.parametrize("managed", [False, True])
def test_cyclic_buffers_hot_path_on_the_gpu(rn50, managed):
    """CyclicBuffers<CudaPinnedHostMemory, CudaDeviceMemory> (buffers.h:122-154) with CUDA memory: 7 requests cut from a
    3-segment ring (it wraps twice) through CreateBindings / CopyToDevice / Infer / CopyFromDevice / Synchronize."""
    out, sec = capi.cyclic_infer(rn50["blob"], rn50["x"], managed_runtime=managed, rounds=7)
    np.testing.assert_array_equal(out, rn50["direct"])
    assert 0 < sec < 0.05
    out3, _ = capi.cyclic_infer(rn50["blob"], rn50["x"][:3], managed_runtime=managed, rounds=4)
    np.testing.assert_array_equal(out3, rn50["direct"][:3])


def test_onnx_imported_resnet50_on_the_gpu(rn50):
    """Caffe ResNet-50 -> ONNX bytes -> generic importer -> plan -> GPU: same `prob` as the prototxt path, and within
    tolerance of the oracle run on the IMPORTED graph."""
    model = onnx_lite.parse_model(onnx_import.export_onnx(rn50["net"], rn50["wts"]))
    net2, w2 = onnx_import.import_onnx(model, name="ResNet-50-onnx")
    low2 = graph.lower(net2, w2)
    x = rn50["x"][:4]
    got = list(helpers.run_engine(low2, x, builder.PREC_FP16, max_batch=8).values())[0]  # (the importer names the output after the ONNX value)
    ref32 = caffe_forward(net2, w2, x)
    assert (got.argmax(1) == ref32.argmax(1)).all()
    assert (np.abs(got - ref32) / ref32.max(1, keepdims=True)).max() <= 1e-3
    np.testing.assert_allclose(got, rn50["direct"][:4], rtol=0, atol=2e-6)  # same lowering up to the importer's fp32 fold


def test_resnet152_fp16_batch32_tuned_tactics_match_oracle(gpu):
    """BASELINE configs[2]'s network and batch (fp16): ResNet-152 at batch 32 with the tactics the tuner picks AT 32."""
    net = graph.resnet_caffe(152)
    wts = weights.random_weights(net, 0)
    low = graph.lower(net, wts)
    x = weights.synthetic_input(32, seed=11)
    blob = builder.build_plan(low, builder.PREC_FP16, 32, outputs=["fc1000", "prob"])
    eng = capi.Engine(blob)
    assert eng.tune(streams=2) > 100  # one tactic per tcgen05 conv, timed at batch 32
    sess = capi.Session(eng)
    try:
        out = sess.infer(x)
        names = [capi.load().b2_context_launch_name(sess.ctx, 32, i).decode() for i in range(sess.nb_launches(32))]
    finally:
        sess.close()
        eng.destroy()
    assert sum(n.startswith("conv_tcgen05") for n in names) == 155
    emu, emu_t = lowered_forward_f16emu(low, x, keep=["fc1000"])
    ref32, ref_t = caffe_forward(net, wts, x, keep=["fc1000"])
    prob, logits = out["prob"], out["fc1000"].reshape(32, -1)
    assert (prob.argmax(1) == ref32.argmax(1)).all() and (prob.argmax(1) == emu.argmax(1)).all()
    assert helpers.rel_err(logits, ref_t["fc1000"].reshape(32, -1)) <= 1e-3
    assert helpers.rel_err(logits, emu_t["fc1000"].reshape(32, -1)) <= 1e-3
    np.testing.assert_allclose(prob.sum(1), 1.0, atol=1e-5)


# ------------------------------------------------------------------------------------------------------------------
# ahead-of-time tactics and graphs
# ------------------------------------------------------------------------------------------------------------------
def test_tactics_are_timed_at_load_and_travel_in_the_plan(gpu, rn50):
    eng = capi.Engine(rn50["blob"])
    try:
        assert capi.load().b2_engine_nb_tactics(eng.handle) == 0
        n = eng.tune(streams=4)
        assert n == 53  # every tcgen05 convolution of ResNet-50, at max batch
        tactics = eng.tactics()
        assert tactics.shape == (53, 10) and (tactics[:, 1] == 8).all() and (tactics[:, 2] >= 32).all()
        assert eng.tune(streams=4) == 53  # idempotent: nothing re-timed
        sess = capi.Session(eng)
        try:
            np.testing.assert_array_equal(sess.infer(rn50["x"])["prob"], rn50["direct"])  # every tactic computes the same bits
        finally:
            sess.close()
    finally:
        eng.destroy()
    # the table rides in the blob (the role of a TensorRT plan's tactics): an engine loaded from it never tunes
    blob2 = builder.attach_tactics(rn50["blob"], tactics)
    eng2 = capi.Engine(blob2)
    try:
        assert capi.load().b2_engine_nb_tactics(eng2.handle) == 53
        np.testing.assert_array_equal(eng2.tactics(), tactics)
        t0 = time.perf_counter()
        assert eng2.tune(streams=4) == 53
        assert time.perf_counter() - t0 < 0.05  # no-op
        sess = capi.Session(eng2)
        try:
            np.testing.assert_array_equal(sess.infer(rn50["x"])["prob"], rn50["direct"])
        finally:
            sess.close()
    finally:
        eng2.destroy()


def test_first_request_at_an_unseen_batch_size_never_tunes(gpu, rn50):
    """Nothing is timed, captured or instantiated on the request path once the context is prepared: the first request at
    every batch size costs less than twice a steady-state one (VERDICT r1 item 4's bar)."""
    eng = capi.Engine(rn50["blob"])
    eng.tune(streams=2)
    sess = capi.Session(eng)
    try:
        for b in range(1, 9):
            sess.prepare(b)

        def timed(b):
            t0 = time.perf_counter()
            out = sess.infer(rn50["x"][:b])["prob"]
            return time.perf_counter() - t0, out

        firsts = {b: timed(b) for b in (3, 5, 7, 2)}
        steady = min(timed(5)[0] for _ in range(10))
        slow = []
        for b, (dt, out) in firsts.items():
            np.testing.assert_array_equal(out, rn50["direct"][:b])
            if not dt < max(2.0 * steady, steady + 1.5e-3):
                slow.append((b, dt, steady))
        # tuning or capturing on the request path would make EVERY first request slow (tens of ms to seconds); one slow
        # request out of four is a scheduling hiccup of the host (observed: ~2 ms once in a few hundred requests)
        assert len(slow) <= 1 and all(dt < 0.05 for _, dt, _ in slow), slow
    finally:
        sess.close()
        eng.destroy()
    # and an UNPREPARED context still never tunes: building a plan is tensor maps + one graph, tens of milliseconds at most
    eng = capi.Engine(rn50["blob"])
    sess = capi.Session(eng)
    try:
        t0 = time.perf_counter()
        sess.infer(rn50["x"][:6])
        assert time.perf_counter() - t0 < 1.0   # timing the tactics of ResNet-50 takes 5-8 s
    finally:
        sess.close()
        eng.destroy()


def test_graph_is_independent_of_the_binding_pointers(gpu, rn50):
    """One captured graph per (context, batch): the same context serves different device buffers (what pooled Buffers
    hand it) without capturing again, and the results follow the data."""
    lib = capi.load()
    sess = rn50["sess"]
    xs = [weights.synthetic_input(8, seed=50 + i) for i in range(3)]
    want = [sess.infer(x)["prob"] for x in xs]
    import ctypes as C
    alt_in = capi.DeviceBuffer(sess.engine.bindings[0]["item_bytes"] * 8)
    alt_out = capi.DeviceBuffer(sess.engine.bindings[1]["item_bytes"] * 8)
    ptrs = (C.c_void_p * 2)(alt_in.ptr, alt_out.ptr)
    for x, w in zip(xs, want):
        sess.host_array(0, 8)[...] = x
        capi.check(lib.b2_memcpy_h2d(alt_in.ptr, sess.host[0].ptr, x.nbytes, sess.stream.handle))
        capi.check(lib.b2_context_enqueue(sess.ctx, 8, ptrs, sess.stream.handle, None))
        capi.check(lib.b2_memcpy_d2h(sess.host[1].ptr, alt_out.ptr, w.nbytes, sess.stream.handle))
        sess.stream.sync()
        np.testing.assert_array_equal(sess.host_array(1, 8), w)
    alt_in.free()
    alt_out.free()


def test_manager_prepares_everything_before_the_first_request(gpu, rn50):
    """InferenceManager: tactics at RegisterModel, lane-pinned contexts with every plan and graph built in
    AllocateResources -> the first requests (any batch size, any pooled Buffers) are as fast as the later ones."""
    mgr = capi.InferenceManager(max_exec_concurrency=2, max_copy_concurrency=4)
    try:
        mgr.register_model("rn50", rn50["blob"])
        mgr.update_resources()
        first = []
        for b in (8, 3, 8, 5, 8, 1, 8, 8):
            t0 = time.perf_counter()
            out = mgr.infer("rn50", rn50["x"][:b])
            first.append(time.perf_counter() - t0)
            np.testing.assert_array_equal(out, rn50["direct"][:b])
        steady = min(first[-3:])
        # one outlier among the eight is host jitter; lazy work on the request path would hit every new batch size
        assert sorted(first)[-2] < max(3.0 * steady, steady + 2e-3) and max(first) < 0.05, first
        res, lats = mgr.bench("rn50", 8, seconds=30.0, max_batches=40)
        assert res["kBatchesComputed"] == 40
        assert np.sort(lats)[-2] < 2.0 * np.percentile(lats, 50) + 1e-4, np.sort(lats)[-4:]   # same: tolerate ONE hiccup in 40
    finally:
        mgr.close()


# ------------------------------------------------------------------------------------------------------------------
# fused classifier tail
# ------------------------------------------------------------------------------------------------------------------
def test_fused_tail_is_one_launch_and_bit_identical(gpu, rn50):
    eng = capi.Engine(rn50["blob"])
    fused = capi.Session(eng)
    plain = capi.Session(eng, {"fuse_tail": 0})
    try:
        nf = [capi.load().b2_context_launch_name(fused.ctx, 8, i).decode() for i in range(fused.nb_launches(8))]
        npl = [capi.load().b2_context_launch_name(plain.ctx, 8, i).decode() for i in range(plain.nb_launches(8))]
        assert len(npl) == 58 and len(nf) == 56
        assert nf[-1].startswith("tail_pool_fc_softmax:pool5+fc1000+prob")
        assert sum(not n.startswith("conv_tcgen05") for n in nf) == 3  # input cast, max pool, tail
        for x in (rn50["x"], rn50["x"][:3], rn50["x"][:1]):
            a, b = fused.infer(x)["prob"], plain.infer(x)["prob"]
            np.testing.assert_array_equal(a, b)
        for _ in range(5):  # the tail's counters are re-armed by its last CTA
            np.testing.assert_array_equal(fused.infer(rn50["x"])["prob"], rn50["direct"])
    finally:
        fused.close()
        plain.close()
        eng.destroy()
    # tapping the intermediate tensors still works: the pooled tensor and the logits are the fused kernel's scratch
    taps = ["pool5", "prob"]
    _, snaps = lowered_forward_f16emu(rn50["low"], rn50["x"][:2], keep=["pool5"])
    out = helpers.run_engine(rn50["low"], rn50["x"][:2], builder.PREC_FP16, outputs=taps)
    assert helpers.rel_err(out["pool5"].reshape(2, -1), snaps["pool5"].reshape(2, -1)) <= 4e-3
    np.testing.assert_allclose(out["prob"], rn50["direct"][:2], rtol=0, atol=0)


# ------------------------------------------------------------------------------------------------------------------
# process-level modes of the host pipeline: cuda_sync<userspace_threads> in the post stage, zero-copy input
# ------------------------------------------------------------------------------------------------------------------
_MODE_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from tensorrt_laboratory_b200 import builder, capi, weights
blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8, seed=0)
x = weights.synthetic_input(8)
mgr = capi.InferenceManager(max_exec_concurrency=2, max_copy_concurrency=4)
mgr.register_model("rn50", blob)
mgr.update_resources()
outs = [mgr.infer("rn50", x[:b]) for b in (8, 3, 8, 1)]
res, lats = mgr.bench("rn50", 8, seconds=30.0, max_batches=24)
assert res["kBatchesComputed"] == 24
np.savez(sys.argv[2], *outs)
mgr.close()
"""


@pytest.mark.parametrize("mode", [{"TRTLAB_SYNC": "yield"}, {"TRTLAB_ZERO_COPY_INPUT": "1"},
                                  {"TRTLAB_ZERO_COPY_INPUT": "1", "TRTLAB_ZERO_COPY_CTAS": "8", "TRTLAB_SYNC": "yield"}],
                         ids=["yielding_sync", "zero_copy_input", "both_small_grid"])
def test_pipeline_modes_change_no_bit(gpu, rn50, tmp_path, mode):
    """TRTLAB_SYNC=yield (post stage polls through cuda_sync<userspace_threads>, reference trtlab/cuda/sync.h:16-48) and
    TRTLAB_ZERO_COPY_INPUT=1 (the input cast reads the mapped pinned buffer over PCIe; CopyToDevice stages nothing) are
    read once per process: run the manager pipeline in a child under each mode and compare with this process's direct path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "out.npz"
    env = dict(os.environ, **mode)
    r = subprocess.run([sys.executable, "-c", _MODE_SCRIPT, root, str(out)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    for arr, b in zip([got[k] for k in got.files], (8, 3, 8, 1)):
        np.testing.assert_array_equal(arr, rn50["direct"][:b])
