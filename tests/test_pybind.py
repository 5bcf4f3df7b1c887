"""Python module `trtlab` (pybind11): the reference's trtlab/pybind surface (infer.cc:683-720) on this runtime.
The GPU case replays the reference's results-pinning script examples/30_PyTensorRT/server.py:19-31 on the golden vectors."""
import os
import sys

import numpy as np
import pytest

from tests import helpers

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tensorrt_laboratory_b200")


def _module():
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import trtlab  # built in-tree by __graft_entry__.build()
    return trtlab


def test_module_surface_matches_the_reference():
    trtlab = _module()
    assert {"InferenceManager", "InferRunner", "InferFuture", "RemoteInferenceManager"} <= set(dir(trtlab))
    mgr = trtlab.InferenceManager(max_exec_concurrency=2)  # keywords of infer.cc:686-688
    for name in ("register_tensorrt_engine", "update_resources", "infer_runner", "get_models", "serve"):
        assert hasattr(mgr, name)
    for name in ("infer", "input_bindings", "output_bindings", "max_batch_size"):
        assert hasattr(trtlab.InferRunner, name)
    assert hasattr(trtlab.InferFuture, "get") and hasattr(trtlab.InferFuture, "wait")
    assert mgr.get_models() == {}
    with pytest.raises(RuntimeError):  # missing engine file: std::runtime_error, as runtime.cc:83-86
        mgr.register_tensorrt_engine("nope", "/nonexistent/engine.plan")
    srv = mgr.serve(port=0, block=False)     # TRTIS GRPCService in front of the manager (grpcio restatement, trtis.py)
    try:
        remote = trtlab.RemoteInferenceManager(hostname=f"127.0.0.1:{srv.port}")
        assert remote.is_healthy() and remote.get_models() == []   # nothing registered yet
        remote.close()
    finally:
        srv.shutdown()


@pytest.mark.gpu
def test_mnist_known_answer_through_the_python_surface(gpu, tmp_path):
    from tensorrt_laboratory_b200 import builder, graph
    trtlab = _module()
    net, w, xs, ys = helpers.load_mnist_golden()
    plan = tmp_path / "mnist-v1.3.plan"
    plan.write_bytes(builder.build_plan(graph.lower(net, w), builder.PREC_FP32, 1))
    models = trtlab.InferenceManager(max_exec_concurrency=2)
    mnist = models.register_tensorrt_engine("mnist", str(plan))
    models.update_resources()
    ins, outs = mnist.input_bindings(), mnist.output_bindings()
    assert list(ins) == ["Input3"] and ins["Input3"]["shape"] == [1, 28, 28] and ins["Input3"]["dtype"] == np.float32
    assert len(outs) == 1 and mnist.max_batch_size() == 1
    results = [mnist.infer(Input3=x) for x in xs]          # futures, all in flight
    results = [r.get() for r in results]
    for r, e, want in zip(results, ys, (2, 0, 9)):
        for key, val in r.items():
            np.testing.assert_almost_equal(val.reshape((1, 10)), e.reshape((1, 10)), decimal=3)  # server.py:31
            assert int(val.argmax()) == want
    assert "mnist" in models.get_models() and 'model="mnist"' not in models.metrics_text()  # InferBench feeds metrics, infer() does not
    with pytest.raises(ValueError):
        mnist.infer(Input3=np.zeros((2, 1, 28, 28), np.float32))  # batch > max_batch_size
    with pytest.raises(KeyError):
        mnist.infer(Nope=np.zeros((1, 1, 28, 28), np.float32))
    np.testing.assert_almost_equal(mnist.infer(Input3=xs[0]).get()[list(outs)[0]].reshape(1, 10), ys[0], decimal=3)  # still serving
    # more requests in flight than pooled Buffers (2 x max_exec_concurrency = 4): infer() must not wait for a Buffers with
    # the GIL held (the post stage of the requests in flight needs it to release theirs); dropped futures are harmless
    futs = [mnist.infer(Input3=xs[i % 3]) for i in range(24)]
    del futs[::2]
    for i, f in zip(range(1, 24, 2), futs):
        assert int(f.get()[list(outs)[0]].argmax()) == (2, 0, 9)[i % 3]


@pytest.mark.gpu
def test_serve_and_remote_inference_manager_replay_the_golden_vectors(gpu, tmp_path):
    """examples/30_PyTensorRT: server.py serves the manager (`models.serve()`), client.py reaches it through
    `trtlab.RemoteInferenceManager(hostname=...)`, `get_models()`, `infer_runner("mnist").infer(Input3=x).get()` -- here over
    the TRTIS GRPCService protocol on a loopback port, against the reference's golden vectors."""
    from tensorrt_laboratory_b200 import builder, graph
    trtlab = _module()
    net, w, xs, ys = helpers.load_mnist_golden()
    plan = tmp_path / "mnist-v1.3.plan"
    plan.write_bytes(builder.build_plan(graph.lower(net, w), builder.PREC_FP32, 1))
    models = trtlab.InferenceManager(max_exec_concurrency=2)
    local = models.register_tensorrt_engine("mnist", str(plan))
    models.update_resources()
    srv = models.serve(port=0, block=False)
    try:
        remote = trtlab.RemoteInferenceManager(hostname=f"127.0.0.1:{srv.port}")
        assert remote.get_models() == ["mnist"]
        mnist = remote.infer_runner("mnist")
        assert mnist.input_bindings() == {"Input3": {"shape": [1, 28, 28], "dtype": np.dtype(np.float32)}}
        assert mnist.max_batch_size() == 1
        futures = [mnist.infer(Input3=x) for x in xs]
        for f, x, e, want in zip(futures, xs, ys, (2, 0, 9)):
            (val,) = f.get().values()
            np.testing.assert_almost_equal(val.reshape((1, 10)), e.reshape((1, 10)), decimal=3)
            assert int(val.argmax()) == want
            (direct,) = local.infer(Input3=x).get().values()
            np.testing.assert_array_equal(val.reshape(-1), np.asarray(direct).reshape(-1))   # the wire changes no bit
        remote.close()
    finally:
        srv.shutdown()
