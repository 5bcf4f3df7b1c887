"""Unary RPC service around the runtime -- the caller side of the hot path (BASELINE.json configs[0]).

The reference's transport is nvrpc, a C++ wrapper over gRPC's async API (trtlab/nvrpc/include/nvrpc/{server,service,rpc,
executor,context}.h; examples/01_Basic_GRPC/src/server.cpp:89-181; examples/02_TensorRT_GRPC/src/server.cc:148-185).
gRPC C++ and protoc are not in this image; grpcio (Python) and the protobuf runtime are, so the same roles are
restated here in Python over grpcio:

  Server / AsyncService / register_rpc / Executor / Context.execute_rpc / Resources      <- nvrpc
  ClientUnary.enqueue(request, on_complete, headers) -> future                          <- nvrpc/client/client_unary.h
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


class AsyncService:
    def __init__(self, name: str):
        self.name = name
        self.rpcs: Dict[str, RPC] = {}

    def register_rpc(self, method: str, request_cls, response_cls, context_cls) -> RPC:
        rpc = RPC(self, method, request_cls, response_cls, context_cls)
        self.rpcs[method] = rpc
        return rpc

    def _generic_handler(self):
        handlers = {
            name: grpc.unary_unary_rpc_method_handler(rpc._handle, request_deserializer=rpc.request_cls.FromString,
                                                      response_serializer=lambda m: m.SerializeToString())
            for name, rpc in self.rpcs.items()}
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
    executor = server.register_executor(Executor(executor_threads))
    executor.register_contexts(rpc, InferenceResources(manager, model_name), contexts)
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
