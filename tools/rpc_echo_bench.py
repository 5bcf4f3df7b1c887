#!/usr/bin/env python
"""BASELINE.json configs[0] (unary echo + host-pool round trip; CPU only): closed-loop RPC rate and latency of the
grpcio restatement of the nvrpc roles, with and without a payload staged through the host buffer pool, direct and
through the replica router."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import rpc  # noqa: E402


def closed_loop(target, n, inflight, payload=None):
    Input, Output = rpc.message("simple.Input"), rpc.message("simple.Output")
    client = rpc.ClientUnary(target, "/simple.Inference/Compute", Input, Output)
    lat = []
    t0 = time.perf_counter()
    pending = []
    sent = 0

    def send(i):
        ts = time.perf_counter()
        msg = Input(batch_id=i, raw_bytes=payload) if payload is not None else Input(batch_id=i)
        return client.enqueue(msg, lambda a, b, c, ts=ts: lat.append(time.perf_counter() - ts))

    while sent < n or pending:
        while sent < n and len(pending) < inflight:
            pending.append(send(sent))
            sent += 1
        pending.pop(0).result(timeout=60)
    wall = time.perf_counter() - t0
    client.close()
    lat = np.asarray(lat) * 1e3
    return {"rpc_per_s": round(n / wall), "p50_ms": round(float(np.percentile(lat, 50)), 3), "p99_ms": round(float(np.percentile(lat, 99)), 3)}


def main():
    servers = [rpc.build_echo_server(contexts=16, executor_threads=8).async_start() for _ in range(2)]
    targets = [f"127.0.0.1:{s.port}" for s in servers]
    router = rpc.Router(targets).async_start()
    payload = bytes(np.random.default_rng(0).integers(0, 256, 3 * 224 * 224 * 4, dtype=np.uint8))  # one fp32 image
    out = {"cores": len(os.sched_getaffinity(0)), "pool_pinned": servers[0].services[0].rpcs["Compute"].resources.pool.pinned}
    closed_loop(targets[0], 200, 8)
    out["echo_direct"] = closed_loop(targets[0], 3000, 16)
    out["echo_via_router"] = closed_loop(f"127.0.0.1:{router.port}", 3000, 16)
    out["echo_602KB_payload_direct"] = closed_loop(targets[0], 600, 8, payload)
    out["echo_602KB_payload_via_router"] = closed_loop(f"127.0.0.1:{router.port}", 600, 8, payload)
    print(json.dumps(out))
    router.shutdown()
    for s in servers:
        s.shutdown()


if __name__ == "__main__":
    main()
