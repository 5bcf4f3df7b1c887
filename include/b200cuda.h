/*
 * b200cuda.h -- C ABI of the CUDA glue the hot path needs around the engine: raw device / pinned-host
 * allocation, streams, events, copies, device selection.  Counterpart of trtlab/cuda in the reference:
 *   cuda_malloc / cuda_malloc_host raw allocators  trtlab/cuda/include/trtlab/cuda/memory/cuda_allocators.h:44-128
 *   device_guard                                   trtlab/cuda/src/device_guard.cc:36-47
 *   cuda_sync<standard_threads|userspace_threads>  trtlab/cuda/include/trtlab/cuda/sync.h:13-62
 *   DeviceInfo                                     trtlab/cuda/src/device_info.cc:66-132
 * Everything returns 0 on success / B2_E* (b200infer.h) on failure with b2_last_error() set.
 */
#ifndef B200CUDA_H_
#define B200CUDA_H_

#include <stddef.h>
#include <stdint.h>

#include "b200infer.h"

#ifdef __cplusplus
extern "C" {
#endif

int b2_device_count(void);                   /* 0 when no driver / no device */
int b2_device_set(int device);               /* cudaSetDevice */
int b2_device_get(void);                     /* current device or -1 */
/* how host threads wait for the current device: 0 = driver default (spin when cores allow), 1 = block on an interrupt
 * (cudaDeviceScheduleBlockingSync) -- for boxes where replicas x pipeline threads outnumber the host cores */
int b2_device_set_blocking_sync(int blocking);
int b2_device_info(int device, char* name, int name_cap, int* cc_major, int* cc_minor, int* sm_count,
                   size_t* total_mem, size_t* l2_bytes);

/* DeviceInfo::Affinity (trtlab/cuda/src/device_info.cc:66-85): the host CPUs closest to `device` (NVML
 * nvmlDeviceGetCpuAffinity), as a bit mask of `n_words` 64-bit words (CPU i = bit i%64 of word i/64).  Returns B2_OK and
 * an all-zero mask when NVML cannot say (not loadable, no topology information). */
int b2_device_cpu_affinity(int device, uint64_t* mask, int n_words);
/* Binds the CALLING thread to that CPU set (intersected with the CPUs the process may use; left unchanged when the
 * intersection is empty).  Pinned host memory allocated afterwards by this thread lands on the GPU's NUMA node (first
 * touch under the default local policy).  *n_cpus (optional) = CPUs in the new mask, 0 = unchanged. */
int b2_bind_thread_to_device(int device, int* n_cpus);

int b2_malloc_device(void** ptr, size_t bytes);      /* cuda_malloc: 256-byte aligned device memory */
int b2_free_device(void* ptr);
int b2_malloc_host(void** ptr, size_t bytes);        /* cuda_malloc_host: pinned, portable */
int b2_free_host(void* ptr);
int b2_memset_device(void* ptr, int value, size_t bytes, b2_stream_t stream);

int b2_stream_create(b2_stream_t* out);              /* blocking stream, as Buffers::Buffers (buffers.cc:42-46) */
int b2_stream_destroy(b2_stream_t s);
int b2_stream_sync(b2_stream_t s);                   /* cuda_sync<standard_threads>::stream_sync */
int b2_stream_query(b2_stream_t s);                  /* 0 done, 1 still running, <0 error (fiber-friendly poll) */

int b2_event_create(b2_event_t* out, int timing);
int b2_event_destroy(b2_event_t e);
int b2_event_record(b2_event_t e, b2_stream_t s);
int b2_event_sync(b2_event_t e);
int b2_event_query(b2_event_t e);                    /* 0 done, 1 pending, <0 error */
int b2_event_elapsed_ms(b2_event_t start, b2_event_t stop, float* ms);
int b2_stream_wait_event(b2_stream_t s, b2_event_t e);

/* async copies on `stream` (host memory should be pinned), cf. Bindings::CopyToDevice/CopyFromDevice
 * (trtlab/tensorrt/src/bindings.cc:128-163) */
int b2_memcpy_h2d(void* dst, const void* src, size_t bytes, b2_stream_t stream);
int b2_memcpy_d2h(void* dst, const void* src, size_t bytes, b2_stream_t stream);
int b2_memcpy_d2d(void* dst, const void* src, size_t bytes, b2_stream_t stream);
int b2_device_sync(void);
/* cudaProfilerStart/Stop: lets `ncu --profile-from-start off` skip plan building and tactic autotuning */
int b2_profiler_start(void);
int b2_profiler_stop(void);

#ifdef __cplusplus
}
#endif
#endif /* B200CUDA_H_ */
