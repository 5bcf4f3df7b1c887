// CUDA glue of the hot path (include/b200cuda.h): raw allocators, streams, events, async copies.
// Counterpart of trtlab/cuda (reference trtlab/cuda/include/trtlab/cuda/memory/cuda_allocators.h:44-128,
// sync.h:13-62, src/device_guard.cc:36-47, src/device_info.cc:66-132).
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <dlfcn.h>
#include <sched.h>

#include <vector>

#include "../../include/b200cuda.h"
#include "b2_internal.h"

using b2i::fail;

#define B2G_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            cudaGetLastError();                                                                          \
            return fail(_e == cudaErrorMemoryAllocation ? B2_ENOMEM : B2_ECUDA, "%s failed: %s", #expr,  \
                        cudaGetErrorString(_e));                                                         \
        }                                                                                                \
    } while (0)

extern "C" {

int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b2_device_set(int device) {
    B2G_CUDA(cudaSetDevice(device));
    return B2_OK;
}

int b2_device_get(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return d;
}

int b2_device_set_blocking_sync(int blocking) {
    B2G_CUDA(cudaSetDeviceFlags(blocking ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
    return B2_OK;
}
int b2_device_info(int device, char* name, int name_cap, int* cc_major, int* cc_minor, int* sm_count,
                   size_t* total_mem, size_t* l2_bytes) {
    cudaDeviceProp p;
    B2G_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_cap > 0) {
        strncpy(name, p.name, size_t(name_cap) - 1);
        name[name_cap - 1] = 0;
    }
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = size_t(p.l2CacheSize);
    return B2_OK;
}

// ---- GPU <-> CPU affinity (reference trtlab/cuda/src/device_info.cc:66-85, DeviceInfo::Affinity) ------------------
// NVML is loaded lazily with dlopen (no link-time dependency); the CUDA device is matched by PCI bus id, so
// CUDA_VISIBLE_DEVICES remapping does not matter.
int b2_device_cpu_affinity(int device, uint64_t* mask, int n_words) {
    if (!mask || n_words < 1) return fail(B2_EINVAL, "bad mask buffer");
    memset(mask, 0, size_t(n_words) * sizeof(uint64_t));
    typedef int (*init_t)();
    typedef int (*by_pci_t)(const char*, void**);
    typedef int (*affinity_t)(void*, unsigned int, unsigned long*);
    static void* lib = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return B2_OK;
    static init_t init = reinterpret_cast<init_t>(dlsym(lib, "nvmlInit_v2"));
    static by_pci_t by_pci = reinterpret_cast<by_pci_t>(dlsym(lib, "nvmlDeviceGetHandleByPciBusId_v2"));
    static affinity_t affinity = reinterpret_cast<affinity_t>(dlsym(lib, "nvmlDeviceGetCpuAffinity"));
    static bool ok = init && by_pci && affinity && init() == 0;
    if (!ok) return B2_OK;
    char bus[64];
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return fail(B2_EINVAL, "no such device %d", device);
    }
    void* h = nullptr;
    if (by_pci(bus, &h) != 0 || !h) return B2_OK;
    std::vector<unsigned long> words(static_cast<size_t>(n_words), 0ul);
    static_assert(sizeof(unsigned long) == sizeof(uint64_t), "NVML cpu-set words are 64-bit here");
    if (affinity(h, static_cast<unsigned int>(n_words), words.data()) != 0) return B2_OK;
    for (int i = 0; i < n_words; ++i) mask[i] = words[size_t(i)];
    return B2_OK;
}

int b2_bind_thread_to_device(int device, int* n_cpus) {
    if (n_cpus) *n_cpus = 0;
    uint64_t mask[16];
    int rc = b2_device_cpu_affinity(device, mask, 16);
    if (rc) return rc;
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return B2_OK;
    int n = 0;
    for (int cpu = 0; cpu < 1024 && cpu < CPU_SETSIZE; ++cpu)
        if (((mask[cpu / 64] >> (cpu % 64)) & 1u) && CPU_ISSET(cpu, &allowed)) {
            CPU_SET(cpu, &want);
            ++n;
        }
    if (n == 0) return B2_OK;  // NVML silent, or the container's cpuset excludes the GPU's CPUs: leave the thread alone
    if (sched_setaffinity(0, sizeof want, &want) != 0) return B2_OK;
    if (n_cpus) *n_cpus = n;
    return B2_OK;
}

int b2_malloc_device(void** ptr, size_t bytes) {
    if (!ptr) return fail(B2_EINVAL, "null ptr");
    B2G_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return B2_OK;
}
int b2_free_device(void* ptr) {
    B2G_CUDA(cudaFree(ptr));
    return B2_OK;
}
int b2_malloc_host(void** ptr, size_t bytes) {
    if (!ptr) return fail(B2_EINVAL, "null ptr");
    B2G_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));  // mapped: kernels may read it in place
    return B2_OK;
}
int b2_free_host(void* ptr) {
    B2G_CUDA(cudaFreeHost(ptr));
    return B2_OK;
}
int b2_memset_device(void* ptr, int value, size_t bytes, b2_stream_t stream) {
    B2G_CUDA(cudaMemsetAsync(ptr, value, bytes, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}

int b2_stream_create(b2_stream_t* out) {
    if (!out) return fail(B2_EINVAL, "null out");
    cudaStream_t s;
    B2G_CUDA(cudaStreamCreate(&s));
    *out = s;
    return B2_OK;
}
int b2_stream_destroy(b2_stream_t s) {
    B2G_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(s)));
    return B2_OK;
}
int b2_stream_sync(b2_stream_t s) {
    B2G_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(s)));
    return B2_OK;
}
int b2_stream_query(b2_stream_t s) {
    cudaError_t e = cudaStreamQuery(static_cast<cudaStream_t>(s));
    if (e == cudaSuccess) return 0;
    if (e == cudaErrorNotReady) return 1;
    cudaGetLastError();
    fail(B2_ECUDA, "cudaStreamQuery failed: %s", cudaGetErrorString(e));
    return -1;
}

int b2_event_create(b2_event_t* out, int timing) {
    if (!out) return fail(B2_EINVAL, "null out");
    cudaEvent_t e;
    B2G_CUDA(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
    *out = e;
    return B2_OK;
}
int b2_event_destroy(b2_event_t e) {
    B2G_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(e)));
    return B2_OK;
}
int b2_event_record(b2_event_t e, b2_stream_t s) {
    B2G_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(e), static_cast<cudaStream_t>(s)));
    return B2_OK;
}
int b2_event_sync(b2_event_t e) {
    B2G_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(e)));
    return B2_OK;
}
int b2_event_query(b2_event_t e) {
    cudaError_t r = cudaEventQuery(static_cast<cudaEvent_t>(e));
    if (r == cudaSuccess) return 0;
    if (r == cudaErrorNotReady) return 1;
    cudaGetLastError();
    fail(B2_ECUDA, "cudaEventQuery failed: %s", cudaGetErrorString(r));
    return -1;
}
int b2_event_elapsed_ms(b2_event_t start, b2_event_t stop, float* ms) {
    B2G_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
    return B2_OK;
}
int b2_stream_wait_event(b2_stream_t s, b2_event_t e) {
    B2G_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(s), static_cast<cudaEvent_t>(e), 0));
    return B2_OK;
}

int b2_memcpy_h2d(void* dst, const void* src, size_t bytes, b2_stream_t stream) {
    B2G_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}
int b2_memcpy_d2h(void* dst, const void* src, size_t bytes, b2_stream_t stream) {
    B2G_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}
int b2_memcpy_d2d(void* dst, const void* src, size_t bytes, b2_stream_t stream) {
    B2G_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}
int b2_device_sync(void) {
    B2G_CUDA(cudaDeviceSynchronize());
    return B2_OK;
}

int b2_profiler_start(void) {
    B2G_CUDA(cudaProfilerStart());
    return B2_OK;
}
int b2_profiler_stop(void) {
    B2G_CUDA(cudaProfilerStop());
    return B2_OK;
}

}  // extern "C"
