"""Unary RPC service around the runtime -- the caller side of the hot path (BASELINE.json configs[0]).

The reference's transport is nvrpc, a C++ wrapper over gRPC's async API (trtlab/nvrpc/include/nvrpc/{server,service,rpc,
executor,context}.h; examples/01_Basic_GRPC/src/server.cpp:89-181; examples/02_TensorRT_GRPC/src/server.cc:148-185).
gRPC C++ and protoc are not in this image; grpcio (Python) and the protobuf runtime are, so the same roles are
restated here in Python over grpcio:

  Server / AsyncService / register_rpc / Executor / Context.execute_rpc / Resources      <- nvrpc (unary life cycle)
  BatchingContext.execute_rpc(requests, responses)   all-in / all-out on one stream      <- nvrpc/life_cycle_batching.h
  StreamingContext.request_received(request, stream) + ServerStream                      <- nvrpc/life_cycle_streaming.h
  ClientUnary.enqueue(request, on_complete, headers) -> future                          <- nvrpc/client/client_unary.h
  ClientStreaming.write(request) / done() -> future of the status                       <- nvrpc/client/client_streaming.h
  simple.Inference/Compute  (Input{batch_id, raw_bytes|sysv} -> Output{batch_id})        <- examples/11_Protos/echo/echo.proto
  ssd.Inference/Compute     (BatchInput -> BatchPredictions)                             <- examples/11_Protos/demo/inference.proto

Messages are built at run time from descriptors with the reference's package names, message names and FIELD NUMBERS, so
the wire format is what the reference's own clients and servers exchange.  Nothing here touches the GPU by itself: the
echo service needs no device (it stages payloads through a host buffer pool -- pinned when a CUDA device is present),
the inference service drives `capi.InferenceManager`.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent import futures
from typing import Callable, Dict, Optional

import grpc
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None, oneof=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _build_pool():
    pool = descriptor_pool.DescriptorPool()
    # ---- package simple (echo.proto) ----
    fd = descriptor_pb2.FileDescriptorProto(name="b2/simple_echo.proto", package="simple", syntax="proto3")
    m = fd.message_type.add(name="SystemV")
    _field(m, "shm_id", 1, _F.TYPE_UINT64), _field(m, "offset", 2, _F.TYPE_UINT64), _field(m, "size", 3, _F.TYPE_UINT64)
    m = fd.message_type.add(name="Input")
    m.oneof_decl.add(name="data")
    _field(m, "batch_id", 1, _F.TYPE_UINT64)
    _field(m, "raw_bytes", 2, _F.TYPE_BYTES, oneof=0)
    _field(m, "sysv", 3, _F.TYPE_MESSAGE, type_name=".simple.SystemV", oneof=0)
    m = fd.message_type.add(name="Output")
    _field(m, "batch_id", 1, _F.TYPE_UINT64)
    pool.Add(fd)
    # ---- package nvrpc.testing (trtlab/nvrpc/tests/testing.proto: TestService{Unary, Streaming}) ----
    fd = descriptor_pb2.FileDescriptorProto(name="b2/nvrpc_testing.proto", package="nvrpc.testing", syntax="proto3")
    m = fd.message_type.add(name="SystemV")
    _field(m, "shm_id", 1, _F.TYPE_UINT64), _field(m, "offset", 2, _F.TYPE_UINT64), _field(m, "size", 3, _F.TYPE_UINT64)
    m = fd.message_type.add(name="Input")
    m.oneof_decl.add(name="data")
    _field(m, "batch_id", 1, _F.TYPE_UINT64)
    _field(m, "raw_bytes", 2, _F.TYPE_BYTES, oneof=0)
    _field(m, "sysv", 3, _F.TYPE_MESSAGE, type_name=".nvrpc.testing.SystemV", oneof=0)
    m = fd.message_type.add(name="Output")
    _field(m, "batch_id", 1, _F.TYPE_UINT64)
    pool.Add(fd)
    # ---- package ssd (demo/inference.proto; the fields the classification path uses) ----
    fd = descriptor_pb2.FileDescriptorProto(name="b2/ssd_inference.proto", package="ssd", syntax="proto3")
    m = fd.message_type.add(name="BatchInput")
    _field(m, "engine_id", 1, _F.TYPE_UINT32), _field(m, "batch_id", 2, _F.TYPE_UINT64)
    _field(m, "batch_size", 3, _F.TYPE_UINT32), _field(m, "int_offset", 4, _F.TYPE_UINT32)
    _field(m, "sysv_offset", 5, _F.TYPE_UINT64), _field(m, "data", 6, _F.TYPE_BYTES)
    m = fd.message_type.add(name="Prediction")
    _field(m, "class_id", 1, _F.TYPE_UINT32), _field(m, "class_str", 2, _F.TYPE_STRING), _field(m, "score", 3, _F.TYPE_FLOAT)
    m = fd.message_type.add(name="Element")
    _field(m, "predictions", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".ssd.Prediction")
    m = fd.message_type.add(name="BatchPredictions")
    _field(m, "elements", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".ssd.Element")
    _field(m, "batch_id", 2, _F.TYPE_UINT64), _field(m, "compute_time", 3, _F.TYPE_FLOAT), _field(m, "total_time", 4, _F.TYPE_FLOAT)
    pool.Add(fd)
    return pool


_POOL = _build_pool()


def message(full_name: str):
    """Message class by full name, e.g. ``message("simple.Input")``."""
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(full_name))


# ------------------------------------------------------------------------------------------------
# nvrpc roles
# ------------------------------------------------------------------------------------------------
class Resources:
    """Shared state of the contexts of one RPC (nvrpc Resources: thread pools, models, buffer pools)."""


class Context:
    """One RPC invocation.  Subclasses implement ``execute_rpc(request, response)`` (nvrpc Context::ExecuteRPC);
    the response is sent when the method returns (the reference's FinishResponse())."""

    def __init__(self, resources: Optional[Resources] = None):
        self._resources = resources
        self.metadata: Dict[str, str] = {}

    def get_resources(self):
        return self._resources

    def execute_rpc(self, request, response) -> None:  # pragma: no cover - interface
        raise NotImplementedError


class BatchingContext(Context):
    """All-in, then all-out batching on one bidirectional stream (nvrpc LifeCycleBatching, life_cycle_batching.h:64-255): the
    client writes requests until it half-closes, ``execute_rpc(requests, responses)`` then runs ONCE on all of them and must
    append exactly one response per request; they go back on the stream in request order.  ``on_request_received`` is the
    per-message hook (called as each request arrives, before the batch runs)."""

    life_cycle = "batching"

    def on_request_received(self, request) -> None:
        pass

    def execute_rpc(self, requests, responses) -> None:  # pragma: no cover - interface
        raise NotImplementedError


class ServerStream:
    """Write side of a streaming RPC handed to the context's callbacks (LifeCycleStreaming::ServerStream,
    life_cycle_streaming.h:80-170).  May be kept and written from any thread until finished or cancelled."""

    _FINISH, _CANCEL = object(), object()

    def __init__(self):
        self._out: "queue.Queue" = queue.Queue()
        self._lock = threading.Lock()
        self._open = True

    def is_connected(self) -> bool:
        return self._open

    def write_response(self, response) -> bool:
        with self._lock:
            if not self._open:
                return False
            self._out.put(response)
            return True

    def finish_stream(self) -> bool:
        with self._lock:
            if not self._open:
                return False
            self._open = False
            self._out.put(self._FINISH)
            return True

    def cancel_stream(self) -> bool:
        with self._lock:
            if not self._open:
                return False
            self._open = False
            self._out.put(self._CANCEL)
            return True


class StreamingContext(Context):
    """Bidirectional streaming (nvrpc LifeCycleStreaming, life_cycle_streaming.h:40-420): ``request_received(request, stream)``
    for every message of the client, with ``stream_initialized`` before the first and ``requests_finished`` after the client
    half-closes; responses are written through the ``ServerStream`` at any time -- zero, one or many per request -- and the
    RPC ends when the context calls ``finish_stream()`` (OK) or ``cancel_stream()`` (CANCELLED).  A context that has not
    ended the stream by the time ``requests_finished`` returns is finished for it (the reference warns and cancels)."""

    life_cycle = "streaming"

    def stream_initialized(self, stream: ServerStream) -> None:
        pass

    def request_received(self, request, stream: ServerStream) -> None:  # pragma: no cover - interface
        raise NotImplementedError

    def requests_finished(self, stream: ServerStream) -> None:
        pass


class RPC:
    def __init__(self, service: "AsyncService", method: str, request_cls, response_cls, context_cls):
        self.service, self.method = service, method
        self.request_cls, self.response_cls, self.context_cls = request_cls, response_cls, context_cls
        self.resources: Optional[Resources] = None
        self.contexts: Optional[queue.Queue] = None  # registered execution contexts (bounds concurrency per RPC)

    def _handle(self, request, grpc_ctx):
        if self.contexts is None:
            grpc_ctx.abort(grpc.StatusCode.UNAVAILABLE, "no execution contexts registered for " + self.method)
        ctx = self.contexts.get()  # blocks when every context of this RPC is busy, like the reference's context pool
        try:
            ctx.metadata = {k: v for k, v in grpc_ctx.invocation_metadata()}
            response = self.response_cls()
            ctx.execute_rpc(request, response)
            return response
        finally:
            self.contexts.put(ctx)

    def _handle_batching(self, request_iterator, grpc_ctx):
        if self.contexts is None:
            grpc_ctx.abort(grpc.StatusCode.UNAVAILABLE, "no execution contexts registered for " + self.method)
        ctx = self.contexts.get()
        try:
            ctx.metadata = {k: v for k, v in grpc_ctx.invocation_metadata()}
            requests = []
            for request in request_iterator:  # ends when the client half-closes (WritesDone)
                ctx.on_request_received(request)
                requests.append(request)
            responses = []
            ctx.execute_rpc(requests, responses)
            if len(responses) != len(requests):
                grpc_ctx.abort(grpc.StatusCode.INTERNAL, f"{self.method}: {len(responses)} responses for {len(requests)} requests")
            for response in responses:
                yield response
        finally:
            self.contexts.put(ctx)

    def _handle_streaming(self, request_iterator, grpc_ctx):
        if self.contexts is None:
            grpc_ctx.abort(grpc.StatusCode.UNAVAILABLE, "no execution contexts registered for " + self.method)
        ctx = self.contexts.get()
        stream = ServerStream()
        failure = []

        def reader():  # the client's messages arrive on their own thread so that responses can flow meanwhile
            try:
                ctx.stream_initialized(stream)
                for request in request_iterator:
                    if not stream.is_connected():
                        break  # finished or cancelled early: further requests are dropped, as the reference does
                    ctx.request_received(request, stream)
                ctx.requests_finished(stream)
                stream.finish_stream()  # no-op when the context already ended the stream
            except Exception as exc:  # a failing callback ends the RPC with an error instead of hanging it
                failure.append(exc)
                stream.cancel_stream()

        try:
            ctx.metadata = {k: v for k, v in grpc_ctx.invocation_metadata()}
            t = threading.Thread(target=reader, daemon=True)
            t.start()
            while True:
                item = stream._out.get()
                if item is ServerStream._FINISH:
                    break
                if item is ServerStream._CANCEL:
                    if failure:
                        grpc_ctx.abort(grpc.StatusCode.INTERNAL, f"{type(failure[0]).__name__}: {failure[0]}")
                    grpc_ctx.abort(grpc.StatusCode.CANCELLED, "stream cancelled by the service")
                yield item
        finally:
            stream._open = False
            self.contexts.put(ctx)


class AsyncService:
    def __init__(self, name: str):
        self.name = name
        self.rpcs: Dict[str, RPC] = {}

    def register_rpc(self, method: str, request_cls, response_cls, context_cls) -> RPC:
        rpc = RPC(self, method, request_cls, response_cls, context_cls)
        self.rpcs[method] = rpc
        return rpc

    def _generic_handler(self):
        handlers = {}
        for name, rpc in self.rpcs.items():
            kind = getattr(rpc.context_cls, "life_cycle", "unary")  # the context class names its life cycle, as in nvrpc
            kw = dict(request_deserializer=rpc.request_cls.FromString, response_serializer=lambda m: m.SerializeToString())
            if kind == "unary":
                handlers[name] = grpc.unary_unary_rpc_method_handler(rpc._handle, **kw)
            elif kind == "batching":
                handlers[name] = grpc.stream_stream_rpc_method_handler(rpc._handle_batching, **kw)
            elif kind == "streaming":
                handlers[name] = grpc.stream_stream_rpc_method_handler(rpc._handle_streaming, **kw)
            else:
                raise ValueError(f"unknown life cycle '{kind}' on {rpc.context_cls.__name__}")
        return grpc.method_handlers_generic_handler(self.name, handlers)


class Executor:
    """Message-processing threads (nvrpc Executor).  ``register_contexts(rpc, resources, n)`` creates the n execution
    contexts that may run concurrently for that RPC."""

    def __init__(self, num_threads: int = 1):
        self.num_threads = num_threads

    def register_contexts(self, rpc: RPC, resources: Optional[Resources], count: int):
        rpc.resources = resources
        rpc.contexts = queue.Queue()
        for _ in range(count):
            rpc.contexts.put(rpc.context_cls(resources))


class Server:
    def __init__(self, address: str = "127.0.0.1:0"):
        self.address = address
        self.services = []
        self.executors = []
        self._server = None
        self.port = None

    def register_async_service(self, name: str) -> AsyncService:
        svc = AsyncService(name)
        self.services.append(svc)
        return svc

    def register_executor(self, executor: Executor) -> Executor:
        self.executors.append(executor)
        return executor

    def async_start(self):
        threads = max(1, sum(e.num_threads for e in self.executors))
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=threads),
                                   options=[("grpc.max_receive_message_length", 64 << 20), ("grpc.max_send_message_length", 64 << 20)])
        for svc in self.services:
            self._server.add_generic_rpc_handlers((svc._generic_handler(),))
        self.port = self._server.add_insecure_port(self.address)
        self._server.start()
        return self

    def running(self) -> bool:
        return self._server is not None

    def shutdown(self):
        if self._server is not None:
            self._server.stop(grace=1.0).wait()
            self._server = None


class ClientUnary:
    """Async unary client (nvrpc/client/client_unary.h): ``enqueue(request, on_complete, headers)`` -> future of
    ``on_complete(request, response, status)``'s return value."""

    def __init__(self, target: str, method: str, request_cls, response_cls):
        self._channel = grpc.insecure_channel(target, options=[("grpc.max_receive_message_length", 64 << 20),
                                                                ("grpc.max_send_message_length", 64 << 20)])
        self._call = self._channel.unary_unary(method, request_serializer=lambda m: m.SerializeToString(),
                                               response_deserializer=response_cls.FromString)

    def enqueue(self, request, on_complete: Optional[Callable] = None, headers: Optional[Dict[str, str]] = None):
        result: "futures.Future" = futures.Future()
        call = self._call.future(request, metadata=tuple((headers or {}).items()))

        def done(c):
            try:
                response, status = c.result(), grpc.StatusCode.OK
            except grpc.RpcError as e:
                response, status = None, e.code()
            try:
                result.set_result(on_complete(request, response, status) if on_complete else response)
            except Exception as exc:  # the callback's failure belongs to the caller
                result.set_exception(exc)

        call.add_done_callback(done)
        return result

    def close(self):
        self._channel.close()


class ClientStreaming:
    """Bidirectional streaming client (nvrpc/client/client_streaming.h): ``write(request)`` any number of times, ``done()``
    half-closes and returns a future of the final ``grpc.StatusCode``; every response is handed to ``on_response`` on the
    reader thread (and collected in ``responses`` when no callback is given)."""

    def __init__(self, target: str, method: str, request_cls, response_cls, on_response: Optional[Callable] = None,
                 headers: Optional[Dict[str, str]] = None):
        self._channel = grpc.insecure_channel(target, options=[("grpc.max_receive_message_length", 64 << 20),
                                                                ("grpc.max_send_message_length", 64 << 20)])
        self._requests: "queue.Queue" = queue.Queue()
        self.responses = []
        self._status: "futures.Future" = futures.Future()
        call = self._channel.stream_stream(method, request_serializer=lambda m: m.SerializeToString(),
                                           response_deserializer=response_cls.FromString)

        def feed():
            while True:
                item = self._requests.get()
                if item is None:
                    return
                yield item

        self._call = call(feed(), metadata=tuple((headers or {}).items()))

        def reader():
            try:
                for response in self._call:
                    (on_response or self.responses.append)(response)
                self._status.set_result(grpc.StatusCode.OK)
            except grpc.RpcError as e:
                self._status.set_result(e.code())

        self._reader = threading.Thread(target=reader, daemon=True)
        self._reader.start()
        self._closed = False

    def write(self, request) -> bool:
        if self._closed:
            return False
        self._requests.put(request)
        return True

    def done(self) -> "futures.Future":
        if not self._closed:
            self._closed = True
            self._requests.put(None)
        return self._status

    def close(self):
        self.done()
        self._reader.join(timeout=5)
        self._channel.close()


# ------------------------------------------------------------------------------------------------
# services
# ------------------------------------------------------------------------------------------------
class HostBufferPool:
    """Blocking pool of equally sized host staging buffers: pinned (cuda_malloc_host) when a CUDA device is usable,
    pageable numpy otherwise -- the plumbing of configs[0] runs without a GPU."""

    def __init__(self, count: int, nbytes: int, pinned: Optional[bool] = None):
        self.nbytes = nbytes
        self.pinned = False
        self._keep = []
        self._free: "queue.Queue[np.ndarray]" = queue.Queue()
        self.bytes_staged = 0
        self._lock = threading.Lock()
        if pinned is None or pinned:
            try:
                from . import capi
                if capi.device_count() > 0:
                    for _ in range(count):
                        pb = capi.PinnedBuffer(nbytes)
                        self._keep.append(pb)
                        self._free.put(pb.array(np.uint8, (nbytes,)))
                    self.pinned = True
            except Exception:
                if pinned:
                    raise
        if not self.pinned:
            for _ in range(count):
                self._free.put(np.empty(nbytes, dtype=np.uint8))

    def round_trip(self, payload: bytes) -> bytes:
        buf = self._free.get()
        try:
            n = len(payload)
            if n > self.nbytes:
                raise ValueError("payload larger than a pool buffer")
            buf[:n] = np.frombuffer(payload, dtype=np.uint8)
            out = buf[:n].tobytes()
            with self._lock:
                self.bytes_staged += n
            return out
        finally:
            self._free.put(buf)


class EchoResources(Resources):
    def __init__(self, threads: int = 3, buffers: int = 4, buffer_bytes: int = 8 << 20, pinned: Optional[bool] = None):
        self.workers = futures.ThreadPoolExecutor(max_workers=threads)  # examples/01_Basic_GRPC SimpleResources
        self.pool = HostBufferPool(buffers, buffer_bytes, pinned)
        self.checksums: Dict[int, int] = {}


class EchoContext(Context):
    """simple.Inference/Compute: the work is pushed to the resources' thread pool (the TPS thread stays free,
    server.cpp:118-133); a payload, when present, makes a round trip through a pooled host buffer."""

    def execute_rpc(self, request, response):
        res: EchoResources = self.get_resources()

        def work():
            if request.WhichOneof("data") == "raw_bytes":
                back = res.pool.round_trip(request.raw_bytes)
                res.checksums[request.batch_id] = int(np.frombuffer(back, dtype=np.uint8).sum(dtype=np.uint64))
            response.batch_id = request.batch_id

        res.workers.submit(work).result()


class InferenceResources(Resources):
    def __init__(self, manager, model_name: str):
        self.manager, self.model_name = manager, model_name
        meta = manager.models[model_name]
        self.in_binding = [b for b in meta.bindings if b["is_input"]][0]
        self.max_batch = meta.max_batch


class InferenceContext(Context):
    """ssd.Inference/Compute on the new runtime (examples/02_TensorRT_GRPC/src/server.cc:148-185): the request's tensor
    goes through InferRunner (pinned Buffers -> H2D -> forward -> D2H), the response carries the top-1 class and score
    per image plus compute / total time."""

    def execute_rpc(self, request, response):
        t0 = time.perf_counter()
        res: InferenceResources = self.get_resources()
        b = res.in_binding
        n = int(request.batch_size)
        if n < 1 or n > res.max_batch:
            raise ValueError(f"batch_size {n} outside [1, {res.max_batch}]")
        x = np.frombuffer(request.data, dtype=b["np_dtype"])
        if x.size != n * int(np.prod(b["shape"])):
            raise ValueError("tensor size does not match batch_size x input binding")
        prob, compute_s = res.manager.infer_timed(res.model_name, x.reshape((n,) + b["shape"]))
        prob = prob.reshape(n, -1)
        for row in prob:
            el = response.elements.add()
            p = el.predictions.add()
            p.class_id = int(row.argmax())
            p.score = float(row.max())
        response.batch_id = request.batch_id
        response.total_time = float(time.perf_counter() - t0)
        response.compute_time = float(compute_s)  # device time of the forward pass, as server.cc:169 (ctx->Synchronize())


class BatchedInferenceContext(BatchingContext):
    """ssd.Inference/BatchedCompute -- the batching life cycle on the inference path: every message of the stream is one
    request (its own ``batch_size`` images), the whole stream is merged into as few forward passes as the model's max batch
    allows (requests are never split), and each request gets its own BatchPredictions back, in order."""

    def execute_rpc(self, requests, responses):
        t0 = time.perf_counter()
        res: InferenceResources = self.get_resources()
        b = res.in_binding
        item = int(np.prod(b["shape"]))
        tensors = []
        for r in requests:
            n = int(r.batch_size)
            x = np.frombuffer(r.data, dtype=b["np_dtype"])
            if n < 1 or n > res.max_batch or x.size != n * item:
                raise ValueError(f"request {r.batch_id}: batch_size {n} / {x.size} elements do not fit the input binding")
            tensors.append(x.reshape((n,) + b["shape"]))
        groups, cur, cur_n = [], [], 0          # greedy, order-preserving packing into forward passes
        for i, x in enumerate(tensors):
            if cur and cur_n + x.shape[0] > res.max_batch:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(i)
            cur_n += x.shape[0]
        if cur:
            groups.append(cur)
        probs, compute = [None] * len(tensors), [0.0] * len(tensors)
        for g in groups:
            y, sec = res.manager.infer_timed(res.model_name, np.concatenate([tensors[i] for i in g], 0))
            at = 0
            for i in g:
                n = tensors[i].shape[0]
                probs[i], compute[i] = y[at:at + n].reshape(n, -1), sec
                at += n
        self.forward_passes = len(groups)
        for r, prob, sec in zip(requests, probs, compute):
            out = message("ssd.BatchPredictions")()
            for row in prob:
                p = out.elements.add().predictions.add()
                p.class_id, p.score = int(row.argmax()), float(row.max())
            out.batch_id, out.compute_time, out.total_time = r.batch_id, float(sec), float(time.perf_counter() - t0)
            responses.append(out)


def build_echo_server(address: str = "127.0.0.1:0", contexts: int = 10, executor_threads: int = 4,
                      resources: Optional[EchoResources] = None) -> Server:
    """The reference's simpleServer (examples/01_Basic_GRPC/src/server.cpp:137-181)."""
    server = Server(address)
    svc = server.register_async_service("simple.Inference")
    rpc = svc.register_rpc("Compute", message("simple.Input"), message("simple.Output"), EchoContext)
    executor = server.register_executor(Executor(executor_threads))
    executor.register_contexts(rpc, resources or EchoResources(), contexts)
    return server


def build_inference_server(manager, model_name: str, address: str = "127.0.0.1:0", contexts: int = 8,
                           executor_threads: int = 8) -> Server:
    server = Server(address)
    svc = server.register_async_service("ssd.Inference")
    rpc = svc.register_rpc("Compute", message("ssd.BatchInput"), message("ssd.BatchPredictions"), InferenceContext)
    batched = svc.register_rpc("BatchedCompute", message("ssd.BatchInput"), message("ssd.BatchPredictions"), BatchedInferenceContext)
    executor = server.register_executor(Executor(executor_threads))
    resources = InferenceResources(manager, model_name)
    executor.register_contexts(rpc, resources, contexts)
    executor.register_contexts(batched, resources, max(1, contexts // 2))
    return server


# ------------------------------------------------------------------------------------------------
# replica router: one front port, N replica servers behind it
# ------------------------------------------------------------------------------------------------
class _PassThrough(grpc.GenericRpcHandler):
    """Method-agnostic unary proxy: request and response travel as raw bytes, so any service can sit behind it."""

    def __init__(self, router: "Router"):
        self._router = router

    def service(self, handler_call_details):
        method = handler_call_details.method

        def forward(request_bytes, grpc_ctx):
            meta = tuple(grpc_ctx.invocation_metadata())
            backend = self._router._pick(dict(meta))
            try:
                call = backend.call_for(method)  # no (de)serializers: bytes in, bytes out
                return call(request_bytes, metadata=tuple((k, v) for k, v in meta if not k.startswith(":") and k != "user-agent"))
            except grpc.RpcError as e:
                grpc_ctx.abort(e.code(), e.details())
            finally:
                backend.done()

        return grpc.unary_unary_rpc_method_handler(forward)


class _Backend:
    def __init__(self, target: str):
        self.target = target
        self.channel = grpc.insecure_channel(target, options=[("grpc.max_receive_message_length", 64 << 20),
                                                              ("grpc.max_send_message_length", 64 << 20)])
        self.outstanding = 0
        self.served = 0
        self._lock = threading.Lock()
        self._calls: Dict[str, Callable] = {}

    def call_for(self, method: str):
        call = self._calls.get(method)
        if call is None:
            call = self._calls.setdefault(method, self.channel.unary_unary(method))
        return call

    def take(self):
        with self._lock:
            self.outstanding += 1
            self.served += 1

    def done(self):
        with self._lock:
            self.outstanding -= 1


class Router:
    """Front end of a replica set: the role Envoy plays in the reference (examples/99_LoadBalancer/lb-envoy.j2
    ``lb_policy: round_robin``; model-aware routing by the ``custom-metadata-model-name`` header,
    examples/Deployment/RouteRequests).  One replica = one process per GPU serving on its own port; requests are
    independent, so the router never touches payloads (no collective, no NCCL).

    ``policy``: "round_robin" or "least_outstanding".  ``routes``: optional {model name: [backend targets]} consulted
    when the request carries the routing header; unknown models fall back to the whole set."""

    HEADER = "custom-metadata-model-name"

    def __init__(self, backends, address: str = "127.0.0.1:0", policy: str = "round_robin", routes=None, threads: int = 16):
        if policy not in ("round_robin", "least_outstanding"):
            raise ValueError("policy must be round_robin or least_outstanding")
        if not backends:
            raise ValueError("a router needs at least one backend")
        self.backends = [_Backend(t) for t in backends]
        self._by_target = {b.target: b for b in self.backends}
        self.policy = policy
        self.routes = {m: [self._by_target[t] for t in ts] for m, ts in (routes or {}).items()}
        self.address, self.threads = address, threads
        self._next: Dict[int, int] = {}
        self._lock = threading.Lock()
        self._server = None
        self.port = None

    def _pick(self, metadata: Dict[str, str]) -> _Backend:
        pool = self.routes.get(metadata.get(self.HEADER, ""), self.backends)
        with self._lock:
            if self.policy == "round_robin":
                i = self._next.get(id(pool), 0)
                self._next[id(pool)] = (i + 1) % len(pool)
                b = pool[i]
            else:
                b = min(pool, key=lambda x: x.outstanding)
            b.take()
        return b

    def async_start(self):
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=self.threads),
                                   options=[("grpc.max_receive_message_length", 64 << 20), ("grpc.max_send_message_length", 64 << 20)])
        self._server.add_generic_rpc_handlers((_PassThrough(self),))
        self.port = self._server.add_insecure_port(self.address)
        self._server.start()
        return self

    def served(self) -> Dict[str, int]:
        return {b.target: b.served for b in self.backends}

    def running(self) -> bool:
        return self._server is not None

    def shutdown(self):
        if self._server is not None:
            self._server.stop(grace=1.0).wait()
            self._server = None
        for b in self.backends:
            b.channel.close()
