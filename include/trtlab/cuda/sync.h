// trtlab::cuda_sync<ThreadType> -- how a host task waits for device work (reference
// trtlab/cuda/include/trtlab/cuda/sync.h:13-62): `standard_threads` blocks the OS thread in the driver,
// `userspace_threads` polls the event / stream and yields to its scheduler between polls, so ONE OS thread can carry
// many in-flight requests.
//
// The reference's userspace scheduler is boost::fibers; Boost is not part of this build, so `userspace_threads::yield`
// is a hook: it defaults to std::this_thread::yield() and a fiber runtime installs its own
// (`userspace_threads::set_yield(&boost::this_fiber::yield)`).  The polling protocol -- query, anything but
// "not ready" is an error, yield, query again -- is the reference's.
#pragma once

#include <atomic>
#include <stdexcept>
#include <string>
#include <thread>

#include "b200cuda.h"
#include "b200infer.h"
#include "trtlab/core/batcher.h"  // standard_threads

namespace trtlab {

struct userspace_threads {
    using yield_fn = void (*)();
    static void yield() { hook().load(std::memory_order_relaxed)(); }
    static void set_yield(yield_fn f) { hook().store(f ? f : &os_yield, std::memory_order_relaxed); }

  private:
    static void os_yield() { std::this_thread::yield(); }
    static std::atomic<yield_fn>& hook() {
        static std::atomic<yield_fn> h{&os_yield};
        return h;
    }
};

template <typename ThreadType>
struct cuda_sync;

template <>
struct cuda_sync<standard_threads> {
    static void event_sync(b2_event_t event) {
        if (b2_event_sync(event) != B2_OK) throw std::runtime_error(std::string("cuda event sync failed: ") + b2_last_error());
    }
    static void stream_sync(b2_stream_t stream) {
        if (b2_stream_sync(stream) != B2_OK) throw std::runtime_error(std::string("cuda stream sync failed: ") + b2_last_error());
    }
};

template <>
struct cuda_sync<userspace_threads> {
    static void event_sync(b2_event_t event) {
        for (int rc = b2_event_query(event); rc != 0; rc = b2_event_query(event)) {
            if (rc < 0) throw std::runtime_error(std::string("cuda event query failed: ") + b2_last_error());
            userspace_threads::yield();
        }
    }
    static void stream_sync(b2_stream_t stream) {
        for (int rc = b2_stream_query(stream); rc != 0; rc = b2_stream_query(stream)) {
            if (rc < 0) throw std::runtime_error(std::string("cuda stream query failed: ") + b2_last_error());
            userspace_threads::yield();
        }
    }
};

}  // namespace trtlab
