// cuda_sync<userspace_threads> polling protocol against a fake device (no CUDA): the hook is called once per "not ready",
// an error code throws, standard_threads blocks through the sync entry points.  Build: see tests/test_cpp_host.py.
#include <cassert>
#include <cstdio>
#include <stdexcept>

#include "trtlab/cuda/sync.h"

static int g_pending = 0, g_yields = 0, g_fail = 0, g_syncs = 0;
extern "C" {
int b2_event_query(b2_event_t) { return g_fail ? -1 : (g_pending-- > 0 ? 1 : 0); }
int b2_stream_query(b2_stream_t) { return g_fail ? -1 : (g_pending-- > 0 ? 1 : 0); }
int b2_event_sync(b2_event_t) { return ++g_syncs, g_fail ? -2 : 0; }
int b2_stream_sync(b2_stream_t) { return ++g_syncs, g_fail ? -2 : 0; }
const char* b2_last_error(void) { return "fake failure"; }
}
static void count_yield() { ++g_yields; }

int main() {
    using namespace trtlab;
    userspace_threads::set_yield(&count_yield);
    g_pending = 5;
    cuda_sync<userspace_threads>::event_sync(nullptr);
    assert(g_yields == 5);
    g_pending = 3;
    cuda_sync<userspace_threads>::stream_sync(nullptr);
    assert(g_yields == 8);
    g_fail = 1;
    bool threw = false;
    try { cuda_sync<userspace_threads>::event_sync(nullptr); } catch (const std::runtime_error&) { threw = true; }
    assert(threw && g_yields == 8);
    threw = false;
    try { cuda_sync<standard_threads>::stream_sync(nullptr); } catch (const std::runtime_error&) { threw = true; }
    assert(threw && g_syncs == 1);
    g_fail = 0;
    cuda_sync<standard_threads>::event_sync(nullptr);
    assert(g_syncs == 2);
    userspace_threads::set_yield(nullptr);  // back to the OS yield
    g_pending = 2;
    cuda_sync<userspace_threads>::event_sync(nullptr);
    std::printf("test_sync OK\n");
    return 0;
}
