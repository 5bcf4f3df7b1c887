// Dynamic batching in front of the hot path (host part): StandardBatcher collects items into a batch
// and maps every item to the batch's shared future; Dispatcher adds the time window (a batch is closed when it is
// full OR when `window` has elapsed since its first item) and runs the user's function on a worker pool.
// Surface and behaviour follow the reference (trtlab/core/include/trtlab/core/batcher.h:23-154,
// dispatcher.h:30-182, task_pool.h; cases of trtlab/core/tests/test_batcher.cc:41-200), std::thread flavour only
// (`standard_threads`); the boost::fiber flavour (`userspace_threads`) needs Boost, which this build does not have.
// Inside a trtlab tree (B2_USE_TRTLAB_CORE) the tree's own trtlab/core headers come first on the include path and this
// file is never reached; the guard below keeps it inert if it is.
#pragma once
#ifndef B2_USE_TRTLAB_CORE
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <thread>
#include <vector>

#include "trtlab/core/hotpath_core.h"

namespace trtlab {

struct standard_threads {
    template <typename R>
    using promise = std::promise<R>;
    template <typename R>
    using shared_future = std::shared_future<R>;
    using mutex = std::mutex;
    using cv = std::condition_variable;
    template <typename TimePoint>
    static void sleep_until(TimePoint t) { std::this_thread::sleep_until(t); }
};

// One timer thread running short callbacks at (not before) their deadlines, earliest first.
class DeferredShortTaskPool {
  public:
    using clock_type = std::chrono::high_resolution_clock;
    DeferredShortTaskPool() : m_Thread([this] { Loop(); }) {}
    ~DeferredShortTaskPool() { shutdown(); }
    DeferredShortTaskPool(const DeferredShortTaskPool&) = delete;
    DeferredShortTaskPool& operator=(const DeferredShortTaskPool&) = delete;

    void enqueue_deferred(clock_type::time_point deadline, std::function<void()> task) {
        {
            std::lock_guard<std::mutex> l(m_Mutex);
            if (m_Stop) throw std::runtime_error("DeferredShortTaskPool is shut down");
            m_Tasks.emplace(deadline, std::move(task));
        }
        m_Cv.notify_one();
    }
    // pending tasks are dropped; the thread is joined
    void shutdown() {
        {
            std::lock_guard<std::mutex> l(m_Mutex);
            if (m_Stop) return;
            m_Stop = true;
        }
        m_Cv.notify_one();
        if (m_Thread.joinable()) m_Thread.join();
    }

  private:
    void Loop() {
        std::unique_lock<std::mutex> l(m_Mutex);
        while (!m_Stop) {
            if (m_Tasks.empty()) {
                m_Cv.wait(l);
                continue;
            }
            const auto next = m_Tasks.begin()->first;
            if (clock_type::now() < next) {
                m_Cv.wait_until(l, next);
                continue;
            }
            auto task = std::move(m_Tasks.begin()->second);
            m_Tasks.erase(m_Tasks.begin());
            l.unlock();
            task();
            l.lock();
        }
    }
    std::mutex m_Mutex;
    std::condition_variable m_Cv;
    std::multimap<clock_type::time_point, std::function<void()>> m_Tasks;
    bool m_Stop = false;
    std::thread m_Thread;  // last member: starts after everything above exists
};

// Batching state machine without threads or locks of its own (the Dispatcher serialises access).
template <typename T, typename ThreadType = standard_threads>
class StandardBatcher {
  public:
    using thread_type = ThreadType;
    using clock_type = std::chrono::high_resolution_clock;
    using future_type = typename ThreadType::template shared_future<void>;
    using batch_item = T;
    struct Batch {
        std::vector<T> items;
        mutable typename ThreadType::template promise<void> promise;  // fulfilled when the items may be reused
        std::size_t batch_id;
    };
    using batch_type = std::optional<Batch>;

    explicit StandardBatcher(std::size_t max_batch_size) : m_Max(max_batch_size) {
        if (max_batch_size == 0) throw std::invalid_argument("max_batch_size must be positive");
    }
    StandardBatcher(StandardBatcher&&) = default;
    virtual ~StandardBatcher() = default;

    // adds the item to the open batch (opening one -- and its time window -- if needed); the future is the batch's
    future_type enqueue(T item) {
        if (!m_Open) {
            m_Open.emplace();
            m_Open->batch.items.reserve(m_Max);
            m_Open->batch.batch_id = m_Counter++;
            m_Open->future = m_Open->batch.promise.get_future().share();
            m_Open->start = clock_type::now();
        }
        m_Open->batch.items.push_back(std::move(item));
        return m_Open->future;
    }
    // the batch, if it just became full
    batch_type update() { return (m_Open && m_Open->batch.items.size() >= m_Max) ? close_batch() : std::nullopt; }
    // whatever has been collected so far (the window expired)
    batch_type close_batch() {
        if (!m_Open) return std::nullopt;
        batch_type out(std::move(m_Open->batch));
        m_Open.reset();
        return out;
    }
    bool empty() const { return !m_Open.has_value(); }
    clock_type::time_point start_time() const { return m_Open ? m_Open->start : clock_type::time_point{}; }
    std::size_t max_batch_size() const { return m_Max; }

  private:
    struct Open {
        Batch batch;
        future_type future;
        clock_type::time_point start;
    };
    std::size_t m_Max;
    std::optional<Open> m_Open;
    std::size_t m_Counter = 0;
};

template <typename BatcherType>
class Dispatcher;

template <template <class, class> class BatcherT, typename T>
class Dispatcher<BatcherT<T, standard_threads>> : private BatcherT<T, standard_threads> {
    using batcher_type = BatcherT<T, standard_threads>;
    using clock_type = typename batcher_type::clock_type;

  public:
    using batch_t = std::vector<T>;
    using future_type = typename batcher_type::future_type;
    using release_fn = std::function<void()>;
    // execute_fn(items, release): `release()` fulfils the batch's promise -- call it once the items' memory is no
    // longer needed (it may be called before the function returns, e.g. right after the H2D copy)
    using execute_fn = std::function<void(const batch_t&, release_fn)>;

    Dispatcher(batcher_type&& batcher, std::chrono::nanoseconds batching_window, std::shared_ptr<ThreadPool> workers,
               std::shared_ptr<DeferredShortTaskPool> timers, execute_fn fn)
        : batcher_type(std::move(batcher)), m_Fn(std::move(fn)), m_Workers(std::move(workers)), m_Timers(std::move(timers)),
          m_Window(batching_window), m_State(std::make_shared<Shared>()) {}
    virtual ~Dispatcher() { shutdown(); }
    Dispatcher(const Dispatcher&) = delete;
    Dispatcher& operator=(const Dispatcher&) = delete;

    future_type enqueue(T item) {
        std::lock_guard<std::mutex> lock(m_State->mutex);
        if (m_State->shutdown) throw std::runtime_error("dispatcher shutting down; no new enqueues can be accepted");
        const bool opens_window = batcher_type::empty();
        auto future = batcher_type::enqueue(std::move(item));
        if (auto batch = batcher_type::update()) {
            Queue(std::move(*batch));  // full: runs now, the pending timer (if any) finds a newer dispatch id
        } else if (opens_window) {
            ArmTimer();
        }
        return future;
    }

    // no new items; the open batch (if any) is flushed immediately and every queued batch finishes
    void shutdown() {
        std::unique_lock<std::mutex> lock(m_State->mutex);
        if (m_State->shutdown) return;
        m_State->shutdown = true;
        if (auto batch = batcher_type::close_batch()) Queue(std::move(*batch));
        m_State->cv.wait(lock, [this] { return m_State->in_flight == 0; });
    }

  private:
    struct Shared {  // outlives the dispatcher inside timer callbacks
        std::mutex mutex;
        std::condition_variable cv;
        bool shutdown = false;
        std::size_t dispatch_id = 0;  // batches queued so far == id of the batch being collected
        std::size_t in_flight = 0;
    };

    void Queue(typename batcher_type::Batch&& batch) {  // mutex held
        auto state = m_State;
        state->dispatch_id++;
        state->in_flight++;
        auto shared_batch = std::make_shared<typename batcher_type::Batch>(std::move(batch));
        auto fn = m_Fn;
        m_Workers->enqueue([state, shared_batch, fn] {
            // shared and captured BY VALUE: the execute function may keep `release` and call it later, from another thread
            auto released = std::make_shared<std::atomic<bool>>(false);
            auto release = [released, shared_batch] {
                if (!released->exchange(true)) shared_batch->promise.set_value();
            };
            try {
                fn(shared_batch->items, release);
                release();  // a function that forgot to release must not strand its callers
            } catch (...) {
                if (!released->exchange(true)) shared_batch->promise.set_exception(std::current_exception());
            }
            {
                std::lock_guard<std::mutex> l(state->mutex);
                state->in_flight--;
            }
            state->cv.notify_all();
        });
    }
    void ArmTimer() {  // mutex held
        auto state = m_State;
        const std::size_t id = state->dispatch_id;
        m_Timers->enqueue_deferred(batcher_type::start_time() + m_Window, [this, state, id] {
            std::lock_guard<std::mutex> l(state->mutex);
            if (state->shutdown || state->dispatch_id != id) return;  // that batch already left (full or flushed)
            if (auto batch = batcher_type::close_batch()) Queue(std::move(*batch));
        });
    }

    execute_fn m_Fn;
    std::shared_ptr<ThreadPool> m_Workers;
    std::shared_ptr<DeferredShortTaskPool> m_Timers;
    std::chrono::nanoseconds m_Window;
    std::shared_ptr<Shared> m_State;
};

}  // namespace trtlab
#endif  // B2_USE_TRTLAB_CORE
