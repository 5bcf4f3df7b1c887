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
    test_async_compute();
    test_bytes();
    test_cyclic_allocator();
    std::printf("ALL OK\n");
    return 0;
}
