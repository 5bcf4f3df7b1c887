// Microbenchmark: what main-loop rate can a tcgen05 GEMM tile sustain on B200 when its operands stream from L2?
//
// A persistent CTA per (SM x occupancy slot) runs the bare main loop of the conv kernel: warp 0 issues the TMA loads of a
// 128 x 64 activation sub-tile and a BN x 64 weight sub-tile into a STAGES-deep ring, warp 1 issues four
// tcgen05.mma 128 x BN x 16 per stage and releases the stage with tcgen05.commit.  No epilogue, no dependencies: the only
// things that can bound the loop are the L2 -> SM fill rate, the MMA issue rate and the barrier round trips.
// The CTA -> tile map emulates a conv grid: CTA id = mt * NT + nt; the A tile is shared by the NT CTAs of one mt, the B
// tile by every CTA with the same nt.  `mma=0` replaces the MMAs by a plain mbarrier arrive (pure fill rate).
// `cn=2` pairs CTAs (same mt, adjacent nt) in a cluster: each fetches half of the A tile and multicasts it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_feed.x l2_feed.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                 "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24); }

struct Params {
    int iters;      // K-blocks per CTA
    int nt;         // N tiles per M tile (A sharing degree)
    int m_tiles;    // distinct M tiles available in the A buffer
    int n_tiles;    // distinct N tiles available in the B buffer
    int k_blocks;   // 64-wide K blocks per row of the buffers
    int mma;        // 1: consume with tcgen05.mma, 0: plain arrive
    int cn;         // cluster size (1 or 2)
    unsigned long long* out;  // per CTA: cycles
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(128) feed_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    const Params p) {
    constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5;
    const int cn = p.cn;
    uint32_t crank = 0;
    if (cn > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], cn);
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(BN < 32 ? 32 : BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (cn > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    const int id = blockIdx.x;
    const int mt = (id / p.nt) % p.m_tiles;
    const int ntile = (id % p.nt) % p.n_tiles;
    const long long t0 = clock64();
    if (warp == 0) {
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % STAGES;
            if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                const int kc = (i % p.k_blocks) * 64;
                if (cn > 1) {
                    const int rows = 128 / cn;
                    tma_load_2d_mc(&mapA, &full[s], sA + s * A_BYTES + crank * rows * 128, kc, mt * 128 + crank * rows,
                                   static_cast<uint16_t>((1u << cn) - 1));
                } else {
                    tma_load_2d(&mapA, &full[s], sA + s * A_BYTES, kc, mt * 128);
                }
                tma_load_2d(&mapB, &full[s], sB + s * B_BYTES, kc, ntile * BN);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t IDESC = make_idesc(128, BN);
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % STAGES;
            mbar_wait(&full[s], (i / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one_sync()) {
                if (p.mma) {
                    const uint32_t a = smem_u32(sA + s * A_BYTES), b = smem_u32(sB + s * B_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_f16(tmem, make_desc(a + j * 32), make_desc(b + j * 32), IDESC, (i | j) ? 1u : 0u);
                    if (cn > 1) umma_commit_mc(&empty[s], static_cast<uint16_t>((1u << cn) - 1));
                    else umma_commit(&empty[s]);
                } else {
                    if (cn > 1) {
                        for (int r = 0; r < cn; ++r) mbar_arrive_cluster(&empty[s], r);
                    } else {
                        mbar_arrive(&empty[s]);
                    }
                }
            }
            __syncwarp();
        }
        if (elect_one_sync()) {
            if (p.mma) umma_commit(done);
            else mbar_arrive(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t1 = clock64();
    if (cn > 1) {  // hear the peers' last stage releases before leaving
        if (warp == 1) {
            const int first = p.iters > STAGES ? p.iters - STAGES : 0;
            for (int i = first; i < p.iters; ++i) mbar_wait(&empty[i % STAGES], (i / STAGES) & 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (cn > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN < 32 ? 32 : BN) : "memory");
    if (threadIdx.x == 0) p.out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static void make_map(CUtensorMap* m, void* base, uint64_t inner, uint64_t outer, uint32_t box_rows) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", int(r));
        exit(1);
    }
}

template <int BN, int STAGES>
static void run(int per_sm, int nt, int mma, int cn, void* dA, void* dB, int m_tiles, int n_rows, int k_blocks, unsigned long long* d_out,
                int iters) {
    CUtensorMap mapA, mapB;
    make_map(&mapA, dA, uint64_t(k_blocks) * 64, uint64_t(m_tiles) * 128, 128 / cn);
    make_map(&mapB, dB, uint64_t(k_blocks) * 64, uint64_t(n_rows), BN);
    const int smem = STAGES * (128 * 128 + BN * 128) + 256 + 1024;
    cudaFuncSetAttribute(feed_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, feed_kernel<BN, STAGES>, 128, smem);
    if (occ < per_sm) {
        printf("BN=%d st=%d per_sm=%d: occupancy only %d, skipped\n", BN, STAGES, per_sm, occ);
        return;
    }
    Params p{iters, nt, m_tiles, n_rows / BN, k_blocks, mma, cn, d_out};
    const int grid = 148 * per_sm;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cn;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = cn > 1 ? 1 : 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        cudaError_t le = cudaLaunchKernelEx(&cfg, feed_kernel<BN, STAGES>, mapA, mapB, p);
        cudaEventRecord(e1);
        cudaError_t se = cudaEventSynchronize(e1);
        if (le != cudaSuccess || se != cudaSuccess) {
            printf("launch failed: %s / %s\n", cudaGetErrorString(le), cudaGetErrorString(se));
            exit(1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<unsigned long long> cyc(grid);
    cudaMemcpy(cyc.data(), d_out, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto c : cyc) mean += double(c);
    mean /= grid;
    const double bytes = double(grid) * iters * (128 * 128 + BN * 128);
    const double flops = mma ? double(grid) * iters * 2.0 * 128 * BN * 64 : 0.0;
    printf("BN=%3d st=%d ctas/SM=%d NT=%2d cn=%d mma=%d : %.3f ms  fill %.2f TB/s  %.0f TFLOP/s  cyc/kblock %.0f  (B/clk/SM %.1f)\n", BN, STAGES,
           per_sm, nt, cn, mma, best, bytes / best * 1e-9, flops / best * 1e-9, mean / iters,
           double(per_sm) * (128 * 128 + BN * 128) / (mean / iters));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    // A: 512 M tiles x 128 rows x (8 k-blocks x 64) fp16 = 64 MiB ; B: 2048 rows x 512 = 2 MiB -- all L2 resident (126 MB)
    const int m_tiles = 512, k_blocks = 8, n_rows = 2048;
    void *dA, *dB;
    unsigned long long* d_out;
    cudaMalloc(&dA, size_t(m_tiles) * 128 * k_blocks * 64 * 2);
    cudaMalloc(&dB, size_t(n_rows) * k_blocks * 64 * 2);
    cudaMalloc(&d_out, 4096 * sizeof(unsigned long long));
    cudaMemset(dA, 0, size_t(m_tiles) * 128 * k_blocks * 64 * 2);
    cudaMemset(dB, 0, size_t(n_rows) * k_blocks * 64 * 2);
    const int iters = 4000;
    for (int mma = 0; mma <= 1; ++mma) {
        for (int nt : {1, 2, 4, 8}) {
            for (int per_sm : {1, 2}) {
                run<64, 4>(per_sm, nt, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
                run<128, 4>(per_sm, nt, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
                run<256, 3>(per_sm, nt, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
            }
        }
        // more CTAs per SM with small rings (how the one-tile-per-CTA conv kernels meet each other)
        run<64, 2>(3, 4, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
        run<64, 2>(4, 4, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
        run<128, 2>(3, 4, mma, 1, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
        // cluster pairs that multicast the activation tile (compare with NT=2 / NT=4 unicast)
        for (int per_sm : {1, 2}) {
            run<64, 4>(per_sm, 2, mma, 2, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
            run<128, 4>(per_sm, 2, mma, 2, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
            run<128, 4>(per_sm, 4, mma, 2, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
            run<256, 3>(per_sm, 2, mma, 2, dA, dB, m_tiles, n_rows, k_blocks, d_out, iters);
        }
    }
    return 0;
}
