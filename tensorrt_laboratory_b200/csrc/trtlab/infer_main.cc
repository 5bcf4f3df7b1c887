// infer.x -- closed-loop inference benchmark CLI, the counterpart of the reference's
// examples/00_TensorRT/infer.cc (flags :68-77, setup :79-122, warm-up + timed run :124-131).
//   infer.x --engine=rn50.plan --contexts=4 --buffers=0 --batch_size=8 --seconds=5 [--replicas=1]
//           [--prethreads=1 --cudathreads=1 --respthreads=3 --runtime=default|unified]
// Plans are produced by `python tools/build_engine.py` (the trtexec step of the reference workflow).
#define B2_WITH_CUDA_RUNTIME 1
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "trtlab/tensorrt/tensorrt.h"

using namespace trtlab;
using namespace trtlab::TensorRT;

static std::string flag(int argc, char** argv, const char* name, const char* dflt) {
    const std::string key = std::string("--") + name + "=";
    for (int i = 1; i < argc; ++i)
        if (std::strncmp(argv[i], key.c_str(), key.size()) == 0) return std::string(argv[i] + key.size());
    return dflt;
}

int main(int argc, char** argv) {
    const std::string engine = flag(argc, argv, "engine", "");
    const int contexts = std::atoi(flag(argc, argv, "contexts", "1").c_str());
    int buffers = std::atoi(flag(argc, argv, "buffers", "0").c_str());
    const int batch = std::atoi(flag(argc, argv, "batch_size", "0").c_str());
    const double seconds = std::atof(flag(argc, argv, "seconds", "5.0").c_str());
    const int replicas = std::atoi(flag(argc, argv, "replicas", "1").c_str());
    const int pre = std::atoi(flag(argc, argv, "prethreads", "1").c_str());
    const int cuda = std::atoi(flag(argc, argv, "cudathreads", "1").c_str());
    const int resp = std::atoi(flag(argc, argv, "respthreads", "3").c_str());
    const std::string runtime_kind = flag(argc, argv, "runtime", "default");
    if (engine.empty()) {
        std::fprintf(stderr, "usage: %s --engine=<plan> [--contexts=N --buffers=M --batch_size=B --seconds=S --replicas=R]\n", argv[0]);
        return 2;
    }
    if (buffers == 0) buffers = 2 * contexts;  // infer.cc:88
    try {
        auto resources = std::make_shared<InferenceManager>(contexts, buffers);
        resources->RegisterThreadPool("pre", std::make_unique<ThreadPool>(size_t(pre)));
        resources->RegisterThreadPool("cuda", std::make_unique<ThreadPool>(size_t(cuda)));
        resources->RegisterThreadPool("post", std::make_unique<ThreadPool>(size_t(resp)));
        std::shared_ptr<Runtime> runtime;
        if (runtime_kind == "unified")
            runtime = std::make_shared<ManagedRuntime>();
        else
            runtime = std::make_shared<StandardRuntime>();
        InferBench::ModelsList models;
        for (int r = 0; r < replicas; ++r) {  // in-process model replicas, infer.cc:118-122
            auto model = runtime->DeserializeEngine(engine);
            resources->RegisterModel(std::to_string(r), model);
            models.push_back(model);
        }
        resources->AllocateResources();
        const uint32_t b = uint32_t(batch > 0 ? batch : models[0]->GetMaxBatchSize());
        InferBench bench(resources);
        bench.Run(models, b, 0.1);  // warm-up, infer.cc:124-126
        std::vector<double> lat;
        auto results = bench.Run(models, b, seconds, 0, &lat);
        auto& r = *results;
        std::printf("Inference Results: %.0f batches in %.3f s; batch_size %u; contexts %d; buffers %d; replicas %d\n",
                    r[kBatchesComputed], r[kWalltime], b, contexts, buffers, replicas);
        std::printf("  inf/sec: %.1f   batches/sec: %.1f   execution time per batch: %.3f ms\n", r[kInferencesPerSecond],
                    r[kBatchesPerSecond], r[kExecutionTimePerBatch] * 1e3);
        std::printf("  request latency p50 %.3f ms  p90 %.3f ms  p99 %.3f ms  max %.3f ms\n", r[kLatencyP50] * 1e3,
                    r[kLatencyP90] * 1e3, r[kLatencyP99] * 1e3, r[kLatencyMax] * 1e3);
        resources->JoinAllThreads();
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "infer.x: %s\n", ex.what());
        return 1;
    }
    return 0;
}
