"""TRTIS GRPCService protocol front end and remote client (tensorrt_laboratory_b200/trtis.py): wire format against
hand-encoded protobuf bytes with the reference's field numbers, Status / Health / Infer against a device-free backend, the
request failures TRTIS reports in `request_status`, and -- on the GPU -- the capi InferenceManager behind the service."""
import numpy as np
import pytest

from tensorrt_laboratory_b200 import trtis


class ToyBackend:
    """y0 = (sum, max) per item, y1 = 2*x; two outputs so the request's output selection and order are exercised."""

    def __init__(self):
        self.calls = 0

    def models(self):
        return {"toy": dict(max_batch=4, inputs={"x": ((3,), np.dtype(np.float32))},
                            outputs={"y0": ((2,), np.dtype(np.float32)), "y1": ((3,), np.dtype(np.float32))}),
                "other": dict(max_batch=1, inputs={"in": ((2, 2), np.dtype(np.float16))}, outputs={"out": ((1,), np.dtype(np.float32))})}

    def infer(self, model, inputs):
        self.calls += 1
        if model == "other":
            raise RuntimeError("backend exploded")
        x = inputs["x"]
        return {"y0": np.stack([x.sum(1), x.max(1)], 1).astype(np.float32), "y1": (2 * x).astype(np.float32)}, 0.25e-3


@pytest.fixture()
def served():
    backend = ToyBackend()
    srv = trtis.build_trtis_server(backend).async_start()
    mgr = trtis.RemoteInferenceManager(f"127.0.0.1:{srv.port}")
    yield backend, srv, mgr
    mgr.close()
    srv.shutdown()


def test_wire_format_uses_the_reference_field_numbers():
    # examples/11_Protos/inference/nvidia_inference.proto: InferRequest{model_name=1, meta_data=3, raw_input=4, batch_id=100};
    # api.proto: InferRequestHeader{batch_size=1, input=2{name=1, byte_size=2}, output=3{name=1}}
    r = trtis.message("InferRequest")(model_name="m", batch_id=7)
    r.meta_data.batch_size = 2
    i = r.meta_data.input.add()
    i.name, i.byte_size = "x", 24
    r.meta_data.output.add().name = "y"
    r.raw_input.append(b"\x01\x02")
    want = (b"\x0a\x01m" + b"\x1a\x0e" + b"\x08\x02" + b"\x12\x05\x0a\x01x\x10\x18" + b"\x1a\x03\x0a\x01y" + b"\x22\x02\x01\x02" + b"\xa0\x06\x07")
    assert r.SerializeToString() == want
    # InferResponse{request_status=1{code=1}, meta_data=2{model_name=1, batch_size=3, output=4{name=1, raw=2{byte_size=1}}},
    #               raw_output=3, compute_time=101}
    resp = trtis.message("InferResponse").FromString(
        b"\x0a\x02\x08\x01" + b"\x12\x0e\x0a\x01m\x18\x02\x22\x07\x0a\x01y\x12\x02\x08\x08" + b"\x1a\x08" + bytes(8) + b"\xad\x06\x00\x00\x80\x3f")
    assert resp.request_status.code == trtis.SUCCESS and resp.meta_data.batch_size == 2
    assert resp.meta_data.output[0].name == "y" and resp.meta_data.output[0].raw.byte_size == 8
    assert resp.raw_output[0] == bytes(8) and resp.compute_time == 1.0
    # server_status.proto: ServerStatus{id=1, model_status=4 map<string, ModelStatus{config=1{name=1, max_batch_size=4}}>, ready_state=7}
    ss = trtis.message("ServerStatus").FromString(b"\x0a\x02id" + b"\x22\x0b\x0a\x01m\x12\x06\x0a\x04\x0a\x00\x20\x08" + b"\x38\x02")
    assert ss.id == "id" and ss.ready_state == trtis.SERVER_READY and ss.model_status["m"].config.max_batch_size == 8


def test_status_health_and_model_discovery(served):
    backend, srv, mgr = served
    assert mgr.is_healthy("live") and mgr.is_healthy("ready") and not mgr.is_healthy("bogus")
    assert mgr.get_models() == ["other", "toy"]
    status = mgr.server_status("toy")
    assert list(status.model_status) == ["toy"] and status.ready_state == trtis.SERVER_READY
    cfg = status.model_status["toy"].config
    assert cfg.max_batch_size == 4 and [(i.name, list(i.dims), i.data_type) for i in cfg.input] == [("x", [3], trtis.TYPE_FP32)]
    assert [(o.name, list(o.dims)) for o in cfg.output] == [("y0", [2]), ("y1", [3])]
    assert status.model_status["toy"].version_status[1].ready_state == trtis.MODEL_READY
    with pytest.raises(trtis.TrtisError) as e:
        mgr.server_status("missing")
    assert e.value.code == trtis.NOT_FOUND
    run = mgr.infer_runner("other")
    assert run.input_bindings() == {"in": {"shape": [2, 2], "dtype": np.dtype(np.float16)}}
    with pytest.raises(KeyError):
        mgr.infer_runner("missing")


def test_infer_round_trip_many_in_flight(served):
    backend, srv, mgr = served
    run = mgr.infer_runner("toy")
    rng = np.random.default_rng(0)
    xs = [rng.standard_normal((b, 3), dtype=np.float32) for b in (1, 4, 2, 3, 4, 1, 2, 4)]
    futs = [run.infer(x=x) for x in xs]
    for x, f in zip(xs, futs):
        out = f.get(30)
        np.testing.assert_array_equal(out["y0"], np.stack([x.sum(1), x.max(1)], 1))
        np.testing.assert_array_equal(out["y1"], 2 * x)
    assert backend.calls == len(xs)
    one = run.infer(x=np.ones(3, np.float32)).get(30)      # a single item without its batch dimension
    assert one["y0"].shape == (1, 2) and one["y0"][0, 0] == 3.0
    with pytest.raises(ValueError):
        run.infer(x=np.ones((5, 3), np.float32))
    with pytest.raises(ValueError):
        run.infer(z=np.ones((1, 3), np.float32))


def test_request_failures_come_back_in_request_status(served):
    backend, srv, mgr = served
    Req = trtis.message("InferRequest")
    call = mgr._infer

    def status_of(req):
        return call.enqueue(req).result(30).request_status

    def good():
        r = Req(model_name="toy")
        r.meta_data.batch_size = 2
        i = r.meta_data.input.add()
        i.name, i.byte_size = "x", 24
        r.meta_data.output.add().name = "y1"
        r.raw_input.append(np.ones((2, 3), np.float32).tobytes())
        return r

    resp = call.enqueue(good()).result(30)
    assert resp.request_status.code == trtis.SUCCESS and [o.name for o in resp.meta_data.output] == ["y1"] and len(resp.raw_output) == 1
    assert abs(resp.compute_time - 0.25e-3) < 1e-9 and resp.request_time > 0 and resp.request_status.request_id > 0
    r = good(); r.model_name = "nope"
    assert status_of(r).code == trtis.NOT_FOUND
    r = good(); r.meta_data.batch_size = 9
    assert status_of(r).code == trtis.INVALID_ARG
    r = good(); r.raw_input[0] = b"123"
    assert status_of(r).code == trtis.INVALID_ARG
    r = good(); r.meta_data.input[0].name = "w"
    assert status_of(r).code == trtis.NOT_FOUND
    r = good(); r.meta_data.output[0].name = "y9"
    assert status_of(r).code == trtis.NOT_FOUND
    r = good(); r.meta_data.output[0].cls.count = 3
    assert status_of(r).code == 7            # UNSUPPORTED: classification post-processing is not offered
    calls = backend.calls
    o = Req(model_name="other")
    o.meta_data.batch_size = 1
    i = o.meta_data.input.add(); i.name = "in"
    o.raw_input.append(np.zeros((1, 2, 2), np.float16).tobytes())
    st = status_of(o)
    assert st.code == trtis.INTERNAL and "backend exploded" in st.msg and backend.calls == calls + 1
    assert call.enqueue(good()).result(30).request_status.code == trtis.SUCCESS   # the service survives a failed request


@pytest.mark.gpu
def test_capi_manager_behind_the_trtis_service(gpu):
    from tensorrt_laboratory_b200 import builder, capi, weights
    blob = builder.build_resnet_plan(50, builder.PREC_FP16, 8, seed=0)
    x = weights.synthetic_input(8)
    mgr = capi.InferenceManager(max_exec_concurrency=2, max_copy_concurrency=4)
    srv = None
    try:
        mgr.register_model("rn50", blob)
        mgr.update_resources()
        direct = mgr.infer("rn50", x)
        srv = trtis.build_trtis_server(trtis.CapiBackend(mgr)).async_start()
        remote = trtis.RemoteInferenceManager(f"127.0.0.1:{srv.port}")
        assert remote.get_models() == ["rn50"]
        run = remote.infer_runner("rn50")
        (in_name,) = run.input_bindings()
        assert run.input_bindings()[in_name]["shape"] == [3, 224, 224] and run.max_batch_size() == 8
        futs = [run.infer(**{in_name: x[:b]}) for b in (8, 3, 8, 1)]
        for b, f in zip((8, 3, 8, 1), futs):
            (y,) = f.get(60).values()
            np.testing.assert_array_equal(y.reshape(b, -1), direct[:b].reshape(b, -1))
        remote.close()
    finally:
        if srv is not None:
            srv.shutdown()
        mgr.close()
