#!/usr/bin/env python
"""Raw pinned-host -> device copy bandwidth for the batch-8 input size (4.8 MB), 1..4 streams, with and without a
concurrent forward pass load: tells whether the e2e path is PCIe-bound."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorrt_laboratory_b200 import capi  # noqa: E402


def main():
    nbytes = 8 * 3 * 224 * 224 * 4
    for nstreams in (1, 2, 4, 8):
        streams = [capi.Stream() for _ in range(nstreams)]
        host = [capi.PinnedBuffer(nbytes) for _ in range(nstreams)]
        dev = [capi.DeviceBuffer(nbytes) for _ in range(nstreams)]
        lib = capi.load()
        iters = 200
        for rep in range(2):
            t0 = time.perf_counter()
            for i in range(iters):
                k = i % nstreams
                capi.check(lib.b2_memcpy_h2d(dev[k].ptr, host[k].ptr, nbytes, streams[k].handle))
            for s in streams:
                s.sync()
            dt = time.perf_counter() - t0
        print(json.dumps({"streams": nstreams, "GBps": iters * nbytes / dt / 1e9, "us_per_copy": dt / iters * 1e6}), flush=True)


if __name__ == "__main__":
    main()
