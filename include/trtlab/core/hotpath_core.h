// Minimal stand-ins for the parts of trtlab/core that the inference hot path consumes.
//
// trtlab/core is a "kept as-is" collaborator of the path but cannot be
// built in this image (it needs Boost.Fiber 1.72, cpuaff, glog).  The path only uses four of its
// facilities; they are re-stated here, std-thread only, with the reference's names and call shapes so
// the trtlab/tensorrt classes read the same:
//   ThreadPool::enqueue          trtlab/core/include/trtlab/core/thread_pool.h:142-145
//   Pool<T>::Create/Push/EmplacePush/Pop(onReturn)   trtlab/core/include/trtlab/core/pool.h:145-245
//   AsyncComputeWrapper / async_compute<void(Args...)>::wrap   .../async_compute.h:38-118
//   Resources                    trtlab/core/include/trtlab/core/resources.h:33-42
//   BytesToString / StringToBytes trtlab/core/src/utils.cc:44-76
// When integrating into a real trtlab tree define B2_USE_TRTLAB_CORE and these are replaced by the
// reference's own headers.
#pragma once

#ifdef B2_USE_TRTLAB_CORE
#include "trtlab/core/async_compute.h"
#include "trtlab/core/pool.h"
#include "trtlab/core/resources.h"
#include "trtlab/core/thread_pool.h"
#include "trtlab/core/utils.h"
#else

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <future>
#include <iostream>
#include <memory>
#include <mutex>
#include <new>
#include <queue>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace trtlab {

// ---- logging / checks (glog-shaped, stderr) ---------------------------------------------------
namespace detail {
struct LogLine {
    std::ostringstream os;
    bool fatal;
    bool enabled;
    LogLine(const char* sev, const char* file, int line, bool fatal_, bool enabled_) : fatal(fatal_), enabled(enabled_) {
        os << sev << " " << file << ":" << line << "] ";
    }
    ~LogLine() noexcept(false) {
        if (enabled || fatal) std::cerr << os.str() << std::endl;
        if (fatal) std::abort();
    }
    template <typename T>
    LogLine& operator<<(const T& v) {
        os << v;
        return *this;
    }
};
inline int verbosity() {
    static int v = [] {
        const char* e = std::getenv("TRTLAB_VERBOSE");
        return e ? std::atoi(e) : 0;
    }();
    return v;
}
}  // namespace detail

#define TRTLAB_LOG_INFO ::trtlab::detail::LogLine("I", __FILE__, __LINE__, false, ::trtlab::detail::verbosity() >= 1)
#define TRTLAB_LOG_WARNING ::trtlab::detail::LogLine("W", __FILE__, __LINE__, false, true)
#define TRTLAB_LOG_ERROR ::trtlab::detail::LogLine("E", __FILE__, __LINE__, false, true)
#define TRTLAB_LOG_FATAL ::trtlab::detail::LogLine("F", __FILE__, __LINE__, true, true)
#define TRTLAB_CHECK(cond) \
    if (!(cond)) ::trtlab::detail::LogLine("F", __FILE__, __LINE__, true, true) << "Check failed: " #cond " "
#define TRTLAB_CHECK_OP(a, op, b) \
    if (!((a)op(b)))              \
    ::trtlab::detail::LogLine("F", __FILE__, __LINE__, true, true) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs " << (b) << ") "

// ---- utils -------------------------------------------------------------------------------------
inline std::string BytesToString(size_t bytes) {
    char buf[64];
    const char prefixes[] = "KMGTPE";
    if (bytes < 1024) {
        snprintf(buf, sizeof buf, "%zu B", bytes);
        return buf;
    }
    int e = 0;
    double v = double(bytes);
    while (v >= 1024.0 && e < 6) {
        v /= 1024.0;
        ++e;
    }
    snprintf(buf, sizeof buf, "%.1f %ciB", v, prefixes[e - 1]);
    return buf;
}

// "10b", "1024B", "1KiB", "10MB", "2.4gb": an 'i' selects powers of 1024, otherwise powers of 1000
inline std::uint64_t StringToBytes(const std::string& s) {
    size_t pos = 0;
    double val = 0;
    try {
        val = std::stod(s, &pos);
    } catch (...) {
        throw std::invalid_argument("StringToBytes: cannot parse \"" + s + "\"");
    }
    std::string unit = s.substr(pos);
    while (!unit.empty() && unit.front() == ' ') unit.erase(unit.begin());
    if (unit.empty() || (unit.back() != 'b' && unit.back() != 'B'))
        throw std::invalid_argument("StringToBytes: expected a unit ending in b/B in \"" + s + "\"");
    unit.pop_back();
    bool binary = false;
    if (!unit.empty() && unit.back() == 'i') {
        binary = true;
        unit.pop_back();
    }
    int exp = 0;
    if (!unit.empty()) {
        switch (unit[0]) {
            case 'k': case 'K': exp = 1; break;
            case 'm': case 'M': exp = 2; break;
            case 'g': case 'G': exp = 3; break;
            case 't': case 'T': exp = 4; break;
            default: throw std::invalid_argument("StringToBytes: unknown prefix in \"" + s + "\"");
        }
        if (unit.size() != 1) throw std::invalid_argument("StringToBytes: bad unit in \"" + s + "\"");
    }
    return static_cast<std::uint64_t>(val * std::pow(binary ? 1024.0 : 1000.0, exp));
}

inline size_t Align(size_t size, size_t alignment) { return (size + alignment - 1) / alignment * alignment; }

// ---- Resources ---------------------------------------------------------------------------------
struct Resources : public std::enable_shared_from_this<Resources> {
    virtual ~Resources() {}
    template <class Target>
    std::shared_ptr<Target> casted_shared_from_this() {
        return std::dynamic_pointer_cast<Target>(Resources::shared_from_this());
    }
};

// ---- ThreadPool --------------------------------------------------------------------------------
class ThreadPool {
    // queue, lock and stop flag are shared with the workers, so a worker outlives the ThreadPool object safely -- which
    // happens when the LAST owner of the pool's owner is a task: InferRunner's stages capture shared_ptr<InferenceManager>,
    // and if the caller drops its reference while the post stage is still unwinding, the manager (and this pool) is
    // destroyed ON a pool thread.  Joining oneself is EDEADLK (std::system_error -> terminate): that worker is detached
    // instead, finds the stop flag in the shared state when its task returns, and leaves.
    struct State {
        std::deque<std::function<void()>> tasks;
        std::mutex mutex;
        std::condition_variable cv;
        bool stop = false;
    };

  public:
    explicit ThreadPool(size_t nthreads) : m_State(std::make_shared<State>()) {
        if (nthreads == 0) nthreads = 1;
        for (size_t i = 0; i < nthreads; ++i) m_Workers.emplace_back([st = m_State] { Loop(*st); });
    }
    ThreadPool(const ThreadPool&) = delete;
    ThreadPool& operator=(const ThreadPool&) = delete;
    ~ThreadPool() {
        {
            std::lock_guard<std::mutex> l(m_State->mutex);
            m_State->stop = true;
        }
        m_State->cv.notify_all();
        for (auto& t : m_Workers) {
            if (t.get_id() == std::this_thread::get_id())
                t.detach();  // destroyed from one of our own tasks: see above
            else
                t.join();
        }
    }

    template <class F, class... Args>
    auto enqueue(F&& f, Args&&... args) -> std::future<typename std::invoke_result<F, Args...>::type> {
        using R = typename std::invoke_result<F, Args...>::type;
        auto task = std::make_shared<std::packaged_task<R()>>(std::bind(std::forward<F>(f), std::forward<Args>(args)...));
        std::future<R> fut = task->get_future();
        enqueue(std::function<void()>([task] { (*task)(); }));
        return fut;
    }

    void enqueue(std::function<void()> task) {
        {
            std::lock_guard<std::mutex> l(m_State->mutex);
            if (m_State->stop) throw std::runtime_error("enqueue on stopped ThreadPool");
            m_State->tasks.push_back(std::move(task));
        }
        m_State->cv.notify_one();
    }

    int Size() const { return int(m_Workers.size()); }

  private:
    static void Loop(State& st) {
        for (;;) {
            std::function<void()> task;
            {
                std::unique_lock<std::mutex> l(st.mutex);
                st.cv.wait(l, [&st] { return st.stop || !st.tasks.empty(); });
                if (st.tasks.empty()) return;  // stop requested and drained
                task = std::move(st.tasks.front());
                st.tasks.pop_front();
            }
            task();
            task = nullptr;  // the task's captures die HERE, not at the top of the next iteration under the lock
        }
    }
    std::shared_ptr<State> m_State;
    std::vector<std::thread> m_Workers;
};

// ---- Pool<T>: blocking pool of shared resources; Pop() hands out a shared_ptr whose deleter returns
//      the resource (after onReturn) instead of destroying it -------------------------------------
template <typename T>
class Pool : public std::enable_shared_from_this<Pool<T>> {
  public:
    static std::shared_ptr<Pool<T>> Create() { return std::shared_ptr<Pool<T>>(new Pool<T>()); }

    void Push(std::shared_ptr<T> obj) {
        {
            std::lock_guard<std::mutex> l(m_Mutex);
            m_Items.push(std::move(obj));
        }
        m_Cv.notify_one();
    }
    void EmplacePush(T* raw) { Push(std::shared_ptr<T>(raw)); }
    template <typename... Args>
    void EmplacePush(Args&&... args) {
        EmplacePush(new T(std::forward<Args>(args)...));
    }

    std::shared_ptr<T> Pop() {
        return Pop([](T*) {});
    }
    // blocks until a resource is available
    std::shared_ptr<T> Pop(std::function<void(T*)> onReturn) {
        std::shared_ptr<T> held = PopWithoutReturn();
        T* raw = held.get();
        auto self = this->shared_from_this();
        return std::shared_ptr<T>(raw, [held, self, onReturn](T* p) mutable {
            onReturn(p);
            self->Push(std::move(held));
            self.reset();
        });
    }
    std::shared_ptr<T> PopWithoutReturn() {
        std::unique_lock<std::mutex> l(m_Mutex);
        m_Cv.wait(l, [this] { return !m_Items.empty(); });
        std::shared_ptr<T> v = std::move(m_Items.front());
        m_Items.pop();
        return v;
    }
    size_t Size() {
        std::lock_guard<std::mutex> l(m_Mutex);
        return m_Items.size();
    }

  private:
    Pool() = default;
    std::queue<std::shared_ptr<T>> m_Items;
    std::mutex m_Mutex;
    std::condition_variable m_Cv;
};

// ---- CyclicAllocator<MemoryType>: a ring of equally sized segments, each a bump stack.  Allocate() hands out
//      a shared handle; a segment is recycled only after the allocator has moved past it AND every handle cut
//      from it is gone (behaviour pinned by trtlab/core/tests/test_cyclic_allocator.cc:57-125).  Allocate()
//      blocks while every segment is still referenced -- natural back-pressure for a request pipeline.
//      MemoryType needs: static void* Allocate(size_t), static void Free(void*), static size_t DefaultAlignment().
struct Malloc {
    static const char* TypeName() { return "Malloc"; }
    static constexpr size_t DefaultAlignment() { return 64; }
    static void* Allocate(size_t bytes) { return std::aligned_alloc(64, (bytes + 63) / 64 * 64); }
    static void Free(void* p) { std::free(p); }
};

template <typename MemoryType>
class CyclicAllocator {
    struct Segment {
        explicit Segment(size_t bytes) : base(static_cast<char*>(MemoryType::Allocate(bytes))) {
            if (!base) throw std::bad_alloc();
        }
        ~Segment() { MemoryType::Free(base); }
        char* base;
        size_t used = 0;
    };
    struct State {  // outlives the allocator while handles are still out
        std::mutex mutex;
        std::condition_variable cv;
        std::queue<std::unique_ptr<Segment>> idle;
        size_t to_drop = 0;  // DropSegment() requests not yet honoured
    };

  public:
    using Descriptor = std::shared_ptr<void>;

    CyclicAllocator(size_t segments, size_t bytes_per_segment)
        : m_State(std::make_shared<State>()), m_SegmentSize(Align(bytes_per_segment, Alignment())) {
        if (segments == 0) throw std::invalid_argument("CyclicAllocator needs at least one segment");
        for (size_t i = 0; i < segments; i++) AddSegment();
    }
    CyclicAllocator(const CyclicAllocator&) = delete;
    CyclicAllocator& operator=(const CyclicAllocator&) = delete;

    static constexpr size_t Alignment() { return MemoryType::DefaultAlignment(); }
    size_t MaxAllocationSize() const { return m_SegmentSize; }

    void AddSegment() {
        auto seg = std::make_unique<Segment>(m_SegmentSize);
        {
            std::lock_guard<std::mutex> l(m_State->mutex);
            m_State->idle.push(std::move(seg));
        }
        m_State->cv.notify_one();
    }
    // removes one idle segment now, or the next one that comes back
    void DropSegment() {
        std::lock_guard<std::mutex> l(m_State->mutex);
        if (!m_State->idle.empty()) m_State->idle.pop();
        else m_State->to_drop++;
    }

    // blocks until a segment with room is available; throws std::length_error if size can never fit
    Descriptor Allocate(size_t size) {
        const size_t need = Align(size ? size : 1, Alignment());
        if (need > m_SegmentSize) throw std::length_error("allocation larger than a CyclicAllocator segment");
        std::lock_guard<std::mutex> serial(m_AllocMutex);
        if (m_Current && m_Current->used + need > m_SegmentSize) m_Current.reset();  // detach; recycles at refcount 0
        if (!m_Current) m_Current = NextSegment();
        char* p = m_Current->base + m_Current->used;
        m_Current->used += need;
        return Descriptor(m_Current, p);  // aliasing: shares the segment's lifetime, points at the slice
    }

    size_t AvailableSegments() {
        std::lock_guard<std::mutex> serial(m_AllocMutex);
        std::lock_guard<std::mutex> l(m_State->mutex);
        return m_State->idle.size() + (m_Current ? 1 : 0);
    }
    size_t AvailableBytes() {
        std::lock_guard<std::mutex> serial(m_AllocMutex);
        std::lock_guard<std::mutex> l(m_State->mutex);
        return m_State->idle.size() * m_SegmentSize + (m_Current ? m_SegmentSize - m_Current->used : 0);
    }

  private:
    std::shared_ptr<Segment> NextSegment() {
        std::unique_lock<std::mutex> l(m_State->mutex);
        m_State->cv.wait(l, [this] { return !m_State->idle.empty(); });
        std::unique_ptr<Segment> seg = std::move(m_State->idle.front());
        m_State->idle.pop();
        auto state = m_State;
        return std::shared_ptr<Segment>(seg.release(), [state](Segment* s) {
            std::unique_ptr<Segment> back(s);
            back->used = 0;
            {
                std::lock_guard<std::mutex> g(state->mutex);
                if (state->to_drop) {
                    state->to_drop--;
                    return;  // `back` frees the memory
                }
                state->idle.push(std::move(back));
            }
            state->cv.notify_one();
        });
    }

    std::shared_ptr<State> m_State;
    const size_t m_SegmentSize;
    std::mutex m_AllocMutex;
    std::shared_ptr<Segment> m_Current;
};

// ---- AsyncCompute: user completion function + promise of its result ----------------------------
template <typename Signature>
class AsyncCompute;

template <typename R, typename... Args>
class AsyncCompute<R(Args...)> {
  public:
    using Fn = std::function<R(Args...)>;
    explicit AsyncCompute(Fn fn) : m_Fn(std::move(fn)) {}
    std::future<R> Future() { return m_Promise.get_future(); }
    std::future<R> get_future() { return m_Promise.get_future(); }
    void operator()(Args... args) {
        try {
            if constexpr (std::is_void<R>::value) {
                m_Fn(args...);
                m_Promise.set_value();
            } else {
                m_Promise.set_value(m_Fn(args...));
            }
        } catch (...) {
            m_Promise.set_exception(std::current_exception());
        }
    }

  private:
    Fn m_Fn;
    std::promise<R> m_Promise;
};

template <typename Signature>
struct AsyncComputeWrapper;

template <typename... Args>
struct AsyncComputeWrapper<void(Args...)> {
    template <typename F>
    static auto Wrap(F&& f) {
        using R = typename std::invoke_result<F, Args...>::type;
        return std::make_shared<AsyncCompute<R(Args...)>>(std::forward<F>(f));
    }
};

template <typename Signature>
struct async_compute;
template <typename... Args>
struct async_compute<void(Args...)> {
    template <typename F>
    static auto wrap(F&& f) {
        return AsyncComputeWrapper<void(Args...)>::Wrap(std::forward<F>(f));
    }
};

}  // namespace trtlab

#ifndef DELETE_COPYABILITY
#define DELETE_COPYABILITY(T) \
    T(const T&) = delete;     \
    T& operator=(const T&) = delete;
#define DELETE_MOVEABILITY(T) \
    T(T&&) = delete;          \
    T& operator=(T&&) = delete;
#endif

#endif  // B2_USE_TRTLAB_CORE
