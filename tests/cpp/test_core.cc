// CPU-only unit tests of the host-side building blocks (no CUDA): Pool, ThreadPool, AsyncCompute,
// MemoryStack arithmetic, byte-string helpers.  Modelled on the reference's gtest suites
// (trtlab/core/tests/test_pool.cc, test_thread_pool.cc, test_async_compute.cc) but dependency-free.
#include <atomic>
#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "trtlab/core/batcher.h"
#include "trtlab/tensorrt/metrics.h"
#include "trtlab/core/hotpath_core.h"

using namespace trtlab;

#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) {                                                        \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                     \
        }                                                                     \
    } while (0)

struct Obj {
    explicit Obj(int v) : value(v) {}
    int value;
    int resets = 0;
};

static void test_pool_returns_on_release() {
    auto pool = Pool<Obj>::Create();
    pool->EmplacePush(new Obj(1));
    pool->EmplacePush(2);
    EXPECT(pool->Size() == 2);
    {
        auto a = pool->Pop([](Obj* o) { o->resets++; });
        auto b = pool->Pop();
        EXPECT(pool->Size() == 0);
        EXPECT(a->value + b->value == 3);
    }
    EXPECT(pool->Size() == 2);  // both came back through the shared_ptr deleters
    int resets = 0;
    for (int i = 0; i < 2; ++i) resets += pool->Pop()->resets;
    EXPECT(resets == 1);  // onReturn ran exactly once, on the object that asked for it
}

static void test_pool_blocks_until_available() {
    auto pool = Pool<Obj>::Create();
    pool->EmplacePush(7);
    auto held = pool->Pop();
    std::atomic<bool> got{false};
    std::thread t([&] {
        auto x = pool->Pop();  // must block while `held` is alive
        got = true;
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    EXPECT(!got.load());
    held.reset();
    t.join();
    EXPECT(got.load());
    EXPECT(pool->Size() == 1);
}

static void test_pool_outlives_handle() {
    // a resource handed out keeps the pool alive (the deleter holds a shared_ptr to it)
    std::shared_ptr<Obj> x;
    {
        auto pool = Pool<Obj>::Create();
        pool->EmplacePush(3);
        x = pool->Pop();
    }
    EXPECT(x->value == 3);
    x.reset();
}

static void test_thread_pool() {
    ThreadPool tp(3);
    EXPECT(tp.Size() == 3);
    std::atomic<int> sum{0};
    std::vector<std::future<int>> futs;
    for (int i = 1; i <= 100; ++i) futs.push_back(tp.enqueue([i, &sum] { sum += i; return i * 2; }));
    int doubled = 0;
    for (auto& f : futs) doubled += f.get();
    EXPECT(sum == 5050 && doubled == 10100);
    std::atomic<int> n{0};
    {
        ThreadPool drain(1);
        for (int i = 0; i < 10; ++i) drain.enqueue(std::function<void()>([&n] { std::this_thread::sleep_for(std::chrono::milliseconds(1)); n++; }));
    }  // destructor drains the queue before joining
    EXPECT(n == 10);
}

// The last owner of a pool's owner may be one of the pool's own tasks (InferRunner's post stage holds the manager): the pool
// is then destroyed ON its own thread.  Joining oneself throws EDEADLK and terminates the process; the worker must be let go.
static void test_thread_pool_destroyed_from_its_own_task() {
    struct Owner {
        explicit Owner(std::atomic<int>* d) : pool(2), destroyed(d) {}
        ~Owner() { destroyed->fetch_add(1); }
        ThreadPool pool;
        std::atomic<int>* destroyed;
    };
    for (int round = 0; round < 50; ++round) {
        std::atomic<int> destroyed{0};
        std::promise<void> go;
        std::shared_future<void> gate = go.get_future().share();
        auto owner = std::make_shared<Owner>(&destroyed);
        owner->pool.enqueue(std::function<void()>([owner, gate] { gate.wait(); }));  // this task keeps the owner alive ...
        owner.reset();                                                               // ... and becomes its LAST owner
        go.set_value();
        for (int i = 0; i < 2000 && destroyed.load() == 0; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        EXPECT(destroyed.load() == 1);
    }
}

static void test_async_compute() {
    using Wrapper = AsyncComputeWrapper<void(std::shared_ptr<int>&)>;
    auto compute = Wrapper::Wrap([](std::shared_ptr<int>& p) { return *p + 1; });
    auto fut = compute->Future();
    auto arg = std::make_shared<int>(41);
    (*compute)(arg);
    EXPECT(fut.get() == 42);
    auto vcompute = Wrapper::Wrap([](std::shared_ptr<int>& p) { p.reset(); });
    auto vfut = vcompute->Future();
    (*vcompute)(arg);
    vfut.get();
    EXPECT(arg == nullptr);
    auto thrower = Wrapper::Wrap([](std::shared_ptr<int>&) -> int { throw std::runtime_error("boom"); });
    auto tf = thrower->Future();
    (*thrower)(arg);
    bool threw = false;
    try { tf.get(); } catch (const std::runtime_error&) { threw = true; }
    EXPECT(threw);
}

// CyclicAllocator: segment accounting and recycling (cases of trtlab/core/tests/test_cyclic_allocator.cc)
static void test_cyclic_allocator() {
    const size_t mb = 1024 * 1024;
    CyclicAllocator<Malloc> ring(5, mb);
    EXPECT(ring.AvailableSegments() == 5);
    EXPECT(ring.AvailableBytes() == 5 * mb);
    ring.AddSegment();
    EXPECT(ring.AvailableSegments() == 6);
    ring.DropSegment();
    EXPECT(ring.AvailableSegments() == 5);
    {
        auto a = ring.Allocate(1);
        EXPECT(reinterpret_cast<uintptr_t>(a.get()) % ring.Alignment() == 0);
        EXPECT(ring.AvailableBytes() == 5 * mb - ring.Alignment());
    }
    // released, but the current segment is only recycled once the allocator has moved past it
    EXPECT(ring.AvailableBytes() == 5 * mb - ring.Alignment());
    {
        auto big = ring.Allocate(mb);  // does not fit behind `a` -> segment 0 detached (and recycled: no handles)
        EXPECT(ring.AvailableSegments() == 5);
        EXPECT(ring.AvailableBytes() == 4 * mb);
        auto next = ring.Allocate(mb / 2);  // segment 1 full -> detached but still referenced by `big`
        EXPECT(ring.AvailableSegments() == 4);
        EXPECT(ring.AvailableBytes() == 3 * mb + mb / 2);
    }
    EXPECT(ring.AvailableSegments() == 5);  // `big` released its segment
    bool threw = false;
    try {
        ring.Allocate(mb + 1);
    } catch (const std::length_error&) {
        threw = true;
    }
    EXPECT(threw);

    // back-pressure: with every segment referenced, Allocate() blocks until a handle is dropped
    CyclicAllocator<Malloc> two(2, 4096);
    auto h0 = two.Allocate(4096);
    auto h1 = two.Allocate(4096);
    std::atomic<bool> got{false};
    std::thread t([&] {
        auto h2 = two.Allocate(4096);
        got = true;
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    EXPECT(!got.load());
    h0.reset();
    t.join();
    EXPECT(got.load());
    // handles may outlive the allocator
    std::shared_ptr<void> survivor;
    {
        CyclicAllocator<Malloc> tmp(1, 4096);
        survivor = tmp.Allocate(128);
    }
    std::memset(survivor.get(), 0xab, 128);
}

// StandardBatcher / Dispatcher / DeferredShortTaskPool (cases of trtlab/core/tests/test_batcher.cc)
static void test_standard_batcher() {
    StandardBatcher<int, standard_threads> batcher(5);
    for (int i = 0; i < 9; i++) {
        auto f = batcher.enqueue(i);
        auto batch = batcher.update();
        if (i == 4) {
            EXPECT(batch.has_value());
            EXPECT(batch->items.size() == 5 && batch->batch_id == 0);
            EXPECT(f.wait_for(std::chrono::seconds(0)) == std::future_status::timeout);
            batch->promise.set_value();
            EXPECT(f.wait_for(std::chrono::seconds(0)) == std::future_status::ready);
        } else {
            EXPECT(!batch.has_value());
        }
    }
    EXPECT(!batcher.update().has_value());
    auto rest = batcher.close_batch();
    EXPECT(rest.has_value() && rest->items.size() == 4 && rest->batch_id == 1);
    EXPECT(batcher.empty() && !batcher.close_batch().has_value());
}

static void test_deferred_task_pool() {
    using namespace std::chrono_literals;
    using clock = std::chrono::high_resolution_clock;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> order;
    DeferredShortTaskPool pool;
    auto start = clock::now();
    auto push = [&](int ms) {
        pool.enqueue_deferred(start + std::chrono::milliseconds(ms), [&, ms] {
            {
                std::lock_guard<std::mutex> l(mu);
                order.push_back(ms);
            }
            cv.notify_one();
        });
    };
    push(25), push(5), push(10);
    {
        std::unique_lock<std::mutex> l(mu);
        EXPECT(order.empty());
        cv.wait(l, [&] { return order.size() == 3; });
    }
    EXPECT(order[0] == 5 && order[1] == 10 && order[2] == 25);  // by deadline, not by submission
    const auto wall = std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - start).count();
    EXPECT(wall >= 25 && wall < 60);
    pool.shutdown();
    bool threw = false;
    try {
        pool.enqueue_deferred(clock::now() + 3ms, [] {});
    } catch (const std::runtime_error&) {
        threw = true;
    }
    EXPECT(threw);
}

static void test_dispatcher_full_batch_and_window() {
    using namespace std::chrono_literals;
    std::mutex mu;
    std::vector<size_t> sizes;
    auto execute = [&](const std::vector<int>& batch, std::function<void()> release) {
        {
            std::lock_guard<std::mutex> l(mu);
            sizes.push_back(batch.size());
        }
        std::this_thread::sleep_for(2ms);
        release();
    };
    auto workers = std::make_shared<ThreadPool>(1);
    auto timers = std::make_shared<DeferredShortTaskPool>();
    StandardBatcher<int, standard_threads> batcher(5);
    Dispatcher<decltype(batcher)> dispatcher(std::move(batcher), 15ms, workers, timers, execute);
    std::vector<std::shared_future<void>> futures;
    for (int i = 0; i < 9; i++) futures.push_back(dispatcher.enqueue(i));
    for (int i = 0; i < 5; i++) futures[i].wait();                                        // the full batch ran at once
    EXPECT(futures[5].wait_for(0s) == std::future_status::timeout);                       // the rest waits for the window
    EXPECT(futures[5].wait_for(200ms) == std::future_status::ready);                      // ... which closes it
    for (auto& f : futures) f.wait();
    {
        std::lock_guard<std::mutex> l(mu);
        EXPECT(sizes.size() == 2 && sizes[0] == 5 && sizes[1] == 4);
    }
    // shutdown flushes an open batch without waiting for its window and refuses new items
    auto late = dispatcher.enqueue(42);
    dispatcher.shutdown();
    EXPECT(late.wait_for(0s) == std::future_status::ready);
    bool threw = false;
    try {
        dispatcher.enqueue(43);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    EXPECT(threw);
    // an exception in the user function reaches the callers of that batch
    StandardBatcher<int, standard_threads> b2(2);
    Dispatcher<decltype(b2)> failing(std::move(b2), 5ms, workers, timers,
                                      [](const std::vector<int>&, std::function<void()>) { throw std::runtime_error("boom"); });
    auto f0 = failing.enqueue(1);
    auto f1 = failing.enqueue(2);
    bool got = false;
    try {
        f1.get();
    } catch (const std::runtime_error&) {
        got = true;
    }
    EXPECT(got);
    (void)f0;
}

// Metrics: the reference service's series in Prometheus text format (examples/02_TensorRT_GRPC/src/server.cc:82-107)
static void test_metrics_exposition() {
    Metrics m;
    for (int i = 1; i <= 100; i++) m.ObserveRequest("rn50", i * 1e-3, i * 1.3e-3);  // load ratio 1.3 each
    m.ObserveRequest("mnist", 1e-3, 0.2);                                          // load ratio 200
    m.SetPower(0, 512.5);
    const std::string t = m.Expose();
    auto has = [&](const char* s) { return t.find(s) != std::string::npos; };
    EXPECT(has("# TYPE yais_inference_compute_duration_ms summary"));
    EXPECT(has("yais_inference_compute_duration_ms{model=\"rn50\",quantile=\"0.5\"} 51"));
    EXPECT(has("yais_inference_compute_duration_ms_count{model=\"rn50\"} 100"));
    EXPECT(has("yais_inference_request_duration_ms_sum{model=\"mnist\"} 200"));
    EXPECT(has("yais_inference_load_ratio_bucket{le=\"1.25\"} 0"));
    EXPECT(has("yais_inference_load_ratio_bucket{le=\"1.5\"} 100"));
    EXPECT(has("yais_inference_load_ratio_bucket{le=\"100\"} 100"));
    EXPECT(has("yais_inference_load_ratio_bucket{le=\"+Inf\"} 101"));
    EXPECT(has("yais_gpus_power_usage{gpu=\"0\"} 512.5"));
}

static void test_bytes() {
    EXPECT(BytesToString(512) == "512 B");
    EXPECT(BytesToString(1536) == "1.5 KiB");
    EXPECT(BytesToString(size_t(30.7 * 1024 * 1024)) == "30.7 MiB");
    EXPECT(StringToBytes("10b") == 10);
    EXPECT(StringToBytes("1KiB") == 1024);
    EXPECT(StringToBytes("10MiB") == 10ull * 1024 * 1024);
    EXPECT(StringToBytes("10MB") == 10000000ull);
    EXPECT(StringToBytes("2.5gb") == 2500000000ull);
    bool threw = false;
    try { StringToBytes("ten bytes"); } catch (const std::invalid_argument&) { threw = true; }
    EXPECT(threw);
    EXPECT(Align(1, 256) == 256 && Align(256, 256) == 256 && Align(257, 256) == 512 && Align(0, 64) == 0);
}

int main() {
    test_pool_returns_on_release();
    test_pool_blocks_until_available();
    test_pool_outlives_handle();
    test_thread_pool();
    test_thread_pool_destroyed_from_its_own_task();
    test_async_compute();
    test_bytes();
    test_cyclic_allocator();
    test_metrics_exposition();
    test_standard_batcher();
    test_deferred_task_pool();
    test_dispatcher_full_batch_and_window();
    std::printf("ALL OK\n");
    return 0;
}
