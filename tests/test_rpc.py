"""nvrpc-role unary service (BASELINE.json configs[0]: unary echo + host pinned-pool round trip; runs without a GPU).
Cases follow trtlab/nvrpc/tests/test_pingpong.cc:182-228 (UnaryTest)."""
import threading
import time

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
        # the batching life cycle on the same path: 11 single-image requests and a 3-image one on ONE stream come back as
        # their own BatchPredictions, computed in ceil(14 / 8) = 2 merged forward passes
        got = []
        stream = rpc.ClientStreaming(f"127.0.0.1:{server.port}", "/ssd.Inference/BatchedCompute", rpc.message("ssd.BatchInput"),
                                     rpc.message("ssd.BatchPredictions"), on_response=got.append)
        picks = [(i, [i % 8]) for i in range(5)] + [(5, [1, 2, 3])] + [(i, [i % 8]) for i in range(6, 12)]
        for bid, rows in picks:
            stream.write(rpc.message("ssd.BatchInput")(batch_id=bid, batch_size=len(rows), data=x[rows].tobytes()))
        assert stream.done().result(120) == grpc.StatusCode.OK
        assert [o.batch_id for o in got] == [bid for bid, _ in picks]
        for o, (bid, rows) in zip(got, picks):
            assert [e.predictions[0].class_id for e in o.elements] == list(direct[rows].argmax(axis=1))
            np.testing.assert_allclose([e.predictions[0].score for e in o.elements], direct[rows].max(axis=1), rtol=1e-6)
        stream.close()
        server.shutdown()
    finally:
        mgr.close()


# ------------------------------------------------------------------------------------------------------------------
# streaming and batching life cycles (trtlab/nvrpc/tests/test_pingpong.{h,cc}: PingPongStreaming, ...EarlyFinish, ...EarlyCancel)
# ------------------------------------------------------------------------------------------------------------------
def _testing_service(contexts):
    In, Out = rpc.message("nvrpc.testing.Input"), rpc.message("nvrpc.testing.Output")
    server = rpc.Server()
    svc = server.register_async_service("nvrpc.testing.TestService")
    ex = server.register_executor(rpc.Executor(4))
    for method, ctx in contexts.items():
        ex.register_contexts(svc.register_rpc(method, In, Out, ctx), None, 2)
    return server.async_start(), In, Out


def test_streaming_life_cycle_pingpong_early_finish_and_early_cancel():
    import grpc

    class PingPong(rpc.StreamingContext):          # one response per request, the context never ends the stream itself
        def stream_initialized(self, stream):
            self.count = 0

        def request_received(self, request, stream):
            self.count += 1
            assert stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=request.batch_id))

    class EarlyFinish(rpc.StreamingContext):       # finishes after the 3rd request: later requests are dropped, status OK
        def stream_initialized(self, stream):
            self.count = 0

        def request_received(self, request, stream):
            self.count += 1
            stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=request.batch_id))
            if self.count == 3:
                assert stream.finish_stream() and not stream.is_connected()
                assert not stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=99))

    class EarlyCancel(EarlyFinish):                # cancels instead: the client sees CANCELLED
        def request_received(self, request, stream):
            self.count += 1
            stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=request.batch_id))
            if self.count == 3:
                stream.cancel_stream()

    class Fanout(rpc.StreamingContext):            # zero-to-many responses per request, some written after the client is done
        def stream_initialized(self, stream):
            self.seen = []

        def request_received(self, request, stream):
            self.seen.append(request.batch_id)
            for _ in range(request.batch_id % 3):
                stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=request.batch_id))

        def requests_finished(self, stream):
            stream.write_response(rpc.message("nvrpc.testing.Output")(batch_id=1000 + len(self.seen)))

    class Broken(rpc.StreamingContext):
        def request_received(self, request, stream):
            raise RuntimeError("callback failed")

    server, In, Out = _testing_service({"Streaming": PingPong, "EarlyFinish": EarlyFinish, "EarlyCancel": EarlyCancel,
                                        "Fanout": Fanout, "Broken": Broken})
    try:
        def run(method, n):
            c = rpc.ClientStreaming(f"127.0.0.1:{server.port}", f"/nvrpc.testing.TestService/{method}", In, Out)
            for i in range(n):
                c.write(In(batch_id=i))
            status = c.done().result(20)
            ids = [r.batch_id for r in c.responses]
            c.close()
            return status, ids

        assert run("Streaming", 10) == (grpc.StatusCode.OK, list(range(10)))
        assert run("Streaming", 0) == (grpc.StatusCode.OK, [])
        assert run("EarlyFinish", 10) == (grpc.StatusCode.OK, [0, 1, 2])
        status, ids = run("EarlyCancel", 10)
        assert status == grpc.StatusCode.CANCELLED and ids[:3] == [0, 1, 2][:len(ids)]
        assert run("Fanout", 6) == (grpc.StatusCode.OK, [1, 2, 2, 4, 5, 5, 1006])
        assert run("Broken", 2)[0] == grpc.StatusCode.INTERNAL
        assert run("Streaming", 3) == (grpc.StatusCode.OK, [0, 1, 2])      # the contexts went back to their pools
        # responses can be consumed while requests are still being written (true bidirectional streaming)
        got = []
        c = rpc.ClientStreaming(f"127.0.0.1:{server.port}", "/nvrpc.testing.TestService/Streaming", In, Out, on_response=got.append)
        c.write(In(batch_id=7))
        deadline = time.time() + 10
        while not got and time.time() < deadline:
            time.sleep(0.01)
        assert [r.batch_id for r in got] == [7]
        c.write(In(batch_id=8))
        assert c.done().result(20) == grpc.StatusCode.OK and [r.batch_id for r in got] == [7, 8]
        c.close()
    finally:
        server.shutdown()


def test_batching_life_cycle_all_in_then_all_out():
    import grpc
    received = []

    class Batch(rpc.BatchingContext):
        def on_request_received(self, request):
            received.append(request.batch_id)

        def execute_rpc(self, requests, responses):
            assert [r.batch_id for r in requests] == received[-len(requests):] if requests else True
            total = sum(r.batch_id for r in requests)                   # something only the WHOLE batch can know
            for r in requests:
                responses.append(rpc.message("nvrpc.testing.Output")(batch_id=total * 100 + r.batch_id))

    class Short(rpc.BatchingContext):
        def execute_rpc(self, requests, responses):
            responses.extend(rpc.message("nvrpc.testing.Output")(batch_id=r.batch_id) for r in requests[:-1])

    server, In, Out = _testing_service({"Batching": Batch, "Short": Short})
    try:
        c = rpc.ClientStreaming(f"127.0.0.1:{server.port}", "/nvrpc.testing.TestService/Batching", In, Out)
        for i in (1, 2, 3, 4):
            c.write(In(batch_id=i))
        time.sleep(0.2)
        assert c.responses == []                                         # nothing comes back before the client is done
        assert c.done().result(20) == grpc.StatusCode.OK
        assert [r.batch_id for r in c.responses] == [1001, 1002, 1003, 1004] and received == [1, 2, 3, 4]
        c.close()
        c = rpc.ClientStreaming(f"127.0.0.1:{server.port}", "/nvrpc.testing.TestService/Batching", In, Out)
        assert c.done().result(20) == grpc.StatusCode.OK and c.responses == []   # an empty batch is a valid batch
        c.close()
        c = rpc.ClientStreaming(f"127.0.0.1:{server.port}", "/nvrpc.testing.TestService/Short", In, Out)
        c.write(In(batch_id=1)), c.write(In(batch_id=2))
        assert c.done().result(20) == grpc.StatusCode.INTERNAL           # one response per request is the contract
        c.close()
    finally:
        server.shutdown()
