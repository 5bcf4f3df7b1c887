"""nvrpc-role unary service (BASELINE.json configs[0]: unary echo + host pinned-pool round trip; runs without a GPU).
Cases follow trtlab/nvrpc/tests/test_pingpong.cc:182-228 (UnaryTest)."""
import threading

import grpc
import numpy as np
import pytest

from tensorrt_laboratory_b200 import rpc


def test_messages_use_the_reference_field_numbers():
    Input, Output = rpc.message("simple.Input"), rpc.message("simple.Output")
    m = Input(batch_id=300, raw_bytes=b"\x01\x02")
    assert m.SerializeToString() == b"\x08\xac\x02\x12\x02\x01\x02"  # field 1 varint 300, field 2 bytes
    assert Output.FromString(b"\x08\x07").batch_id == 7
    bi = rpc.message("ssd.BatchInput")(batch_id=5, batch_size=8, data=b"ab")
    assert bi.SerializeToString() == b"\x10\x05\x18\x08\x32\x02ab"   # fields 2, 3, 6 of demo/inference.proto
    bp = rpc.message("ssd.BatchPredictions")()
    p = bp.elements.add().predictions.add()
    p.class_id, p.score = 3, 0.5
    assert rpc.message("ssd.BatchPredictions").FromString(bp.SerializeToString()).elements[0].predictions[0].class_id == 3


def test_unary_echo_pingpong():
    server = rpc.build_echo_server(contexts=10, executor_threads=4).async_start()
    assert server.running() and server.port > 0
    client = rpc.ClientUnary(f"127.0.0.1:{server.port}", "/simple.Inference/Compute", rpc.message("simple.Input"),
                             rpc.message("simple.Output"))
    lock = threading.Lock()
    state = {"count": 0, "recv": 0}
    send = 100
    futs = []
    for i in range(1, send + 1):
        with lock:
            state["count"] += 1

        def on_complete(inp, out, status, i=i):
            assert status == grpc.StatusCode.OK and out.batch_id == i
            with lock:
                state["count"] -= 1
                state["recv"] += 1

        futs.append(client.enqueue(rpc.message("simple.Input")(batch_id=i), on_complete, {"x-content-model": "flowers-152"}))
    for f in futs:
        f.result(timeout=30)
    assert state == {"count": 0, "recv": send}
    assert server.running()
    client.close()
    server.shutdown()
    assert not server.running()


def test_echo_payload_round_trips_through_the_host_buffer_pool():
    res = rpc.EchoResources(threads=3, buffers=2, buffer_bytes=1 << 20, pinned=False)
    server = rpc.build_echo_server(resources=res, contexts=4).async_start()
    client = rpc.ClientUnary(f"127.0.0.1:{server.port}", "/simple.Inference/Compute", rpc.message("simple.Input"),
                             rpc.message("simple.Output"))
    rng = np.random.default_rng(0)
    payloads = {i: rng.integers(0, 256, size=200_000 + i, dtype=np.uint8).tobytes() for i in range(1, 9)}
    futs = [client.enqueue(rpc.message("simple.Input")(batch_id=i, raw_bytes=p)) for i, p in payloads.items()]
    assert sorted(f.result(timeout=30).batch_id for f in futs) == list(payloads)
    assert res.pool.bytes_staged == sum(len(p) for p in payloads.values())
    for i, p in payloads.items():  # what came back out of the pool buffer is what went in
        assert res.checksums[i] == int(np.frombuffer(p, dtype=np.uint8).sum(dtype=np.uint64))
    # a payload larger than a pool buffer is an RPC error, not a crash
    status = client.enqueue(rpc.message("simple.Input")(batch_id=99, raw_bytes=bytes(2 << 20)), lambda i, o, s: s).result(timeout=30)
    assert status != grpc.StatusCode.OK
    client.close()
    server.shutdown()


def test_unknown_method_and_unregistered_contexts():
    server = rpc.Server()
    svc = server.register_async_service("simple.Inference")
    svc.register_rpc("Compute", rpc.message("simple.Input"), rpc.message("simple.Output"), rpc.EchoContext)  # no contexts
    server.register_executor(rpc.Executor(1))
    server.async_start()
    ok = rpc.ClientUnary(f"127.0.0.1:{server.port}", "/simple.Inference/Compute", rpc.message("simple.Input"), rpc.message("simple.Output"))
    assert ok.enqueue(rpc.message("simple.Input")(batch_id=1), lambda i, o, s: s).result(timeout=30) == grpc.StatusCode.UNAVAILABLE
    bad = rpc.ClientUnary(f"127.0.0.1:{server.port}", "/simple.Inference/Nope", rpc.message("simple.Input"), rpc.message("simple.Output"))
    assert bad.enqueue(rpc.message("simple.Input")(batch_id=1), lambda i, o, s: s).result(timeout=30) == grpc.StatusCode.UNIMPLEMENTED
    ok.close(), bad.close()
    server.shutdown()


def test_router_spreads_requests_over_replicas_and_routes_by_model_header():
    """Replica front end: round-robin over N backends like the reference's Envoy cluster, model-aware
    routing by the `custom-metadata-model-name` header, least-outstanding as the alternative policy."""
    servers = [rpc.build_echo_server(contexts=4).async_start() for _ in range(3)]
    targets = [f"127.0.0.1:{s.port}" for s in servers]
    router = rpc.Router(targets, routes={"flowers-152": targets[2:]}).async_start()
    Input, Output = rpc.message("simple.Input"), rpc.message("simple.Output")
    client = rpc.ClientUnary(f"127.0.0.1:{router.port}", "/simple.Inference/Compute", Input, Output)
    outs = [client.enqueue(Input(batch_id=i)).result(timeout=30) for i in range(1, 91)]   # sequential: exact round robin
    assert [o.batch_id for o in outs] == list(range(1, 91))
    assert list(router.served().values()) == [30, 30, 30]
    futs = [client.enqueue(Input(batch_id=i), headers={rpc.Router.HEADER: "flowers-152"}) for i in range(100, 120)]
    assert sorted(f.result(timeout=30).batch_id for f in futs) == list(range(100, 120))
    assert list(router.served().values()) == [30, 30, 50]                                  # the routed model has one replica
    # errors of the backend travel through (unknown method), the router stays up
    bad = rpc.ClientUnary(f"127.0.0.1:{router.port}", "/simple.Inference/Nope", Input, Output)
    assert bad.enqueue(Input(batch_id=1), lambda i, o, s: s).result(timeout=30) == grpc.StatusCode.UNIMPLEMENTED
    assert router.running()
    lo = rpc.Router(targets, policy="least_outstanding").async_start()
    c2 = rpc.ClientUnary(f"127.0.0.1:{lo.port}", "/simple.Inference/Compute", Input, Output)
    futs = [c2.enqueue(Input(batch_id=i)) for i in range(60)]
    assert sorted(f.result(timeout=30).batch_id for f in futs) == list(range(60))
    assert sum(lo.served().values()) == 60 and min(lo.served().values()) > 0
    for c in (client, bad, c2):
        c.close()
    router.shutdown(), lo.shutdown()
    for s in servers:
        s.shutdown()
    with pytest.raises(ValueError):
        rpc.Router([])


@pytest.mark.gpu
def test_inference_service_matches_direct_path(gpu):
    from tensorrt_laboratory_b200 import builder, capi, weights
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8)
    x = weights.synthetic_input(8)
    mgr = capi.InferenceManager(max_exec_concurrency=2, max_copy_concurrency=4)
    try:
        mgr.register_model("rn50", blob)
        mgr.update_resources()
        direct = mgr.infer("rn50", x)
        server = rpc.build_inference_server(mgr, "rn50").async_start()
        client = rpc.ClientUnary(f"127.0.0.1:{server.port}", "/ssd.Inference/Compute", rpc.message("ssd.BatchInput"),
                                 rpc.message("ssd.BatchPredictions"))
        futs = [client.enqueue(rpc.message("ssd.BatchInput")(batch_id=i, batch_size=n, data=x[:n].tobytes()))
                for i, n in enumerate((8, 3, 8, 1), start=1)]
        for (i, n), f in zip(enumerate((8, 3, 8, 1), start=1), futs):
            out = f.result(timeout=120)
            assert out.batch_id == i and len(out.elements) == n and out.total_time > 0
            assert 0 < out.compute_time < out.total_time  # device time of the forward pass (server.cc:169), not the wall time
            assert [e.predictions[0].class_id for e in out.elements] == list(direct[:n].argmax(axis=1))
            np.testing.assert_allclose([e.predictions[0].score for e in out.elements], direct[:n].max(axis=1), rtol=1e-6)
        client.close()
        server.shutdown()
    finally:
        mgr.close()
