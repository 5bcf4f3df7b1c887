// Observability of the hot path: the four series the reference's inference service exports
// (examples/02_TensorRT_GRPC/src/server.cc:82-107,176-178; metrics.cc:32-62), without prometheus-cpp:
//   yais_inference_compute_duration_ms{model=...}   summary (quantiles 0.5 / 0.9 / 0.99, _sum, _count)
//   yais_inference_request_duration_ms{model=...}   summary
//   yais_inference_load_ratio                       histogram, buckets 1.25 1.5 2 10 100 (request / compute time)
//   yais_gpus_power_usage{gpu="N"}                  gauge (watts, NVML; absent when NVML is not loadable)
// `Expose()` renders the Prometheus text format 0.0.4; `MetricsExposer` serves it over HTTP (GET /metrics), the role of
// the prometheus::Exposer the reference service starts (examples/02_TensorRT_GRPC/src/metrics.cc:34-60).
#pragma once
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <functional>
#include <thread>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace trtlab {

class Metrics {
  public:
    // quantiles over a sliding window of the most recent observations (prometheus-cpp uses a decaying estimator; the
    // window keeps this dependency-free and exact over what it holds)
    class Summary {
      public:
        explicit Summary(size_t window = 4096) : m_Window(window) {}
        void Observe(double v) {
            std::lock_guard<std::mutex> l(m_Mutex);
            m_Sum += v;
            m_Count++;
            if (m_Samples.size() < m_Window) m_Samples.push_back(v);
            else m_Samples[m_Next] = v;
            m_Next = (m_Next + 1) % m_Window;
        }
        struct Snapshot {
            uint64_t count;
            double sum, q50, q90, q99;
        };
        Snapshot Read() const {
            std::lock_guard<std::mutex> l(m_Mutex);
            std::vector<double> s(m_Samples);
            std::sort(s.begin(), s.end());
            auto q = [&](double p) { return s.empty() ? 0.0 : s[std::min(s.size() - 1, size_t(p * double(s.size() - 1) + 0.5))]; };
            return Snapshot{m_Count, m_Sum, q(0.5), q(0.9), q(0.99)};
        }

      private:
        mutable std::mutex m_Mutex;
        size_t m_Window, m_Next = 0;
        std::vector<double> m_Samples;
        double m_Sum = 0;
        uint64_t m_Count = 0;
    };

    class Histogram {
      public:
        explicit Histogram(std::vector<double> upper_bounds) : m_Bounds(std::move(upper_bounds)), m_Counts(m_Bounds.size() + 1, 0) {}
        void Observe(double v) {
            std::lock_guard<std::mutex> l(m_Mutex);
            size_t i = 0;
            while (i < m_Bounds.size() && v > m_Bounds[i]) ++i;
            m_Counts[i]++;
            m_Sum += v;
        }
        void Read(std::vector<double>* bounds, std::vector<uint64_t>* cumulative, double* sum) const {
            std::lock_guard<std::mutex> l(m_Mutex);
            *bounds = m_Bounds;
            cumulative->assign(m_Counts.size(), 0);
            uint64_t run = 0;
            for (size_t i = 0; i < m_Counts.size(); ++i) (*cumulative)[i] = (run += m_Counts[i]);
            *sum = m_Sum;
        }

      private:
        mutable std::mutex m_Mutex;
        std::vector<double> m_Bounds;
        std::vector<uint64_t> m_Counts;  // last = +Inf
        double m_Sum = 0;
    };

    Metrics() : m_LoadRatio({1.25, 1.50, 2.0, 10.0, 100.0}) {}

    // one finished request (seconds, as ExecutionContext::Synchronize() and a wall clock report them)
    void ObserveRequest(const std::string& model, double compute_seconds, double request_seconds) {
        SummaryFor(m_Compute, model).Observe(compute_seconds * 1e3);
        SummaryFor(m_Request, model).Observe(request_seconds * 1e3);
        if (compute_seconds > 0) m_LoadRatio.Observe(request_seconds / compute_seconds);
    }
    void SetPower(int gpu, double watts) {
        std::lock_guard<std::mutex> l(m_Mutex);
        m_Power[gpu] = watts;
    }
    // reads the board power through NVML (loaded lazily with dlopen); false when NVML is not available
    bool SamplePower(int gpu) {
        typedef int (*init_t)();
        typedef int (*handle_t)(unsigned, void**);
        typedef int (*power_t)(void*, unsigned*);
        static void* lib = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        static init_t init = reinterpret_cast<init_t>(dlsym(lib, "nvmlInit_v2"));
        static handle_t handle = reinterpret_cast<handle_t>(dlsym(lib, "nvmlDeviceGetHandleByIndex_v2"));
        static power_t power = reinterpret_cast<power_t>(dlsym(lib, "nvmlDeviceGetPowerUsage"));
        static const bool ok = init && handle && power && init() == 0;
        if (!ok) return false;
        void* dev = nullptr;
        unsigned mw = 0;
        if (handle(unsigned(gpu), &dev) != 0 || power(dev, &mw) != 0) return false;
        SetPower(gpu, double(mw) * 1e-3);
        return true;
    }

    std::string Expose() const {
        std::ostringstream os;
        os.precision(10);
        auto summaries = [&](const char* name, const std::map<std::string, std::unique_ptr<Summary>>& fam) {
            os << "# TYPE " << name << " summary\n";
            for (const auto& kv : fam) {
                const auto s = kv.second->Read();
                const std::string l = "model=\"" + kv.first + "\"";
                os << name << "{" << l << ",quantile=\"0.5\"} " << s.q50 << "\n";
                os << name << "{" << l << ",quantile=\"0.9\"} " << s.q90 << "\n";
                os << name << "{" << l << ",quantile=\"0.99\"} " << s.q99 << "\n";
                os << name << "_sum{" << l << "} " << s.sum << "\n";
                os << name << "_count{" << l << "} " << s.count << "\n";
            }
        };
        std::lock_guard<std::mutex> l(m_Mutex);
        summaries("yais_inference_compute_duration_ms", m_Compute);
        summaries("yais_inference_request_duration_ms", m_Request);
        std::vector<double> bounds;
        std::vector<uint64_t> cum;
        double sum = 0;
        m_LoadRatio.Read(&bounds, &cum, &sum);
        os << "# TYPE yais_inference_load_ratio histogram\n";
        for (size_t i = 0; i < bounds.size(); ++i) os << "yais_inference_load_ratio_bucket{le=\"" << bounds[i] << "\"} " << cum[i] << "\n";
        os << "yais_inference_load_ratio_bucket{le=\"+Inf\"} " << cum.back() << "\n";
        os << "yais_inference_load_ratio_sum " << sum << "\n";
        os << "yais_inference_load_ratio_count " << cum.back() << "\n";
        if (!m_Power.empty()) {
            os << "# TYPE yais_gpus_power_usage gauge\n";
            for (const auto& kv : m_Power) os << "yais_gpus_power_usage{gpu=\"" << kv.first << "\"} " << kv.second << "\n";
        }
        return os.str();
    }

  private:
    Summary& SummaryFor(std::map<std::string, std::unique_ptr<Summary>>& fam, const std::string& model) {
        std::lock_guard<std::mutex> l(m_Mutex);
        auto& slot = fam[model];
        if (!slot) slot.reset(new Summary());
        return *slot;
    }
    mutable std::mutex m_Mutex;
    std::map<std::string, std::unique_ptr<Summary>> m_Compute, m_Request;
    Histogram m_LoadRatio;
    std::map<int, double> m_Power;
};


// Minimal HTTP/1.0 endpoint for the Prometheus scraper: one accept thread, one response per connection.
// GET /metrics (or /) -> 200 text/plain; version=0.0.4 with `render()`; anything else -> 404.
class MetricsExposer {
  public:
    // port 0 = let the kernel pick (see Port()); binds 0.0.0.0 like prometheus::Exposer("0.0.0.0:port")
    MetricsExposer(int port, std::function<std::string()> render) : m_Render(std::move(render)) {
        m_Fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (m_Fd < 0) throw std::runtime_error("MetricsExposer: socket() failed");
        int one = 1;
        ::setsockopt(m_Fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in addr{};
        addr.sin_family = AF_INET;
        addr.sin_addr.s_addr = htonl(INADDR_ANY);
        addr.sin_port = htons(static_cast<uint16_t>(port));
        socklen_t len = sizeof addr;
        if (::bind(m_Fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 || ::listen(m_Fd, 16) != 0 ||
            ::getsockname(m_Fd, reinterpret_cast<sockaddr*>(&addr), &len) != 0) {
            ::close(m_Fd);
            throw std::runtime_error("MetricsExposer: cannot listen on port " + std::to_string(port));
        }
        m_Port = ntohs(addr.sin_port);
        m_Thread = std::thread([this] { Serve(); });
    }
    ~MetricsExposer() {
        m_Stop = true;
        ::shutdown(m_Fd, SHUT_RDWR);
        ::close(m_Fd);
        if (m_Thread.joinable()) m_Thread.join();
    }
    MetricsExposer(const MetricsExposer&) = delete;
    MetricsExposer& operator=(const MetricsExposer&) = delete;
    int Port() const { return m_Port; }

  private:
    void Serve() {
        while (!m_Stop) {
            const int c = ::accept(m_Fd, nullptr, nullptr);
            if (c < 0) {
                if (m_Stop) break;
                continue;
            }
            char req[1024];
            const ssize_t n = ::recv(c, req, sizeof req - 1, 0);
            std::string head(req, n > 0 ? size_t(n) : 0);
            const bool ok = head.rfind("GET /metrics", 0) == 0 || head.rfind("GET / ", 0) == 0;
            const std::string body = ok ? m_Render() : std::string("not found\n");
            std::string resp = std::string(ok ? "HTTP/1.0 200 OK\r\n" : "HTTP/1.0 404 Not Found\r\n") +
                               "Content-Type: text/plain; version=0.0.4\r\nContent-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
            size_t off = 0;
            while (off < resp.size()) {
                const ssize_t w = ::send(c, resp.data() + off, resp.size() - off, MSG_NOSIGNAL);
                if (w <= 0) break;
                off += size_t(w);
            }
            ::close(c);
        }
    }
    std::function<std::string()> m_Render;
    int m_Fd = -1, m_Port = 0;
    std::atomic<bool> m_Stop{false};
    std::thread m_Thread;
};

}  // namespace trtlab
