// sm_100a kernels of the B200-native inference engine.
//
//  * conv_f16_tcgen05 -- implicit-GEMM convolution: TMA (tiled or im2col mode) -> 128B-swizzled smem ->
//    tcgen05.mma (UMMA 128xBNx16, fp16 x fp16 -> fp32 in TMEM) -> tcgen05.ld epilogue with
//    bias / residual / ReLU fused -> 128-bit NHWC stores.  One 128xBN output tile per CTA,
//    warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, all 4 warps = epilogue.
//  * SIMT kernels -- fp32-engine reference path (fp64 accumulate) and the non-GEMM operators
//    (layout casts, max/avg pool, FC, softmax).
//
// Replaces the forward pass the reference delegates to TensorRT:
// trtlab/tensorrt/src/workspace.cc:47,52 (enqueueV2) / examples/10_Internals/README.md:50-52.
#include "kernels.h"

#include <stdio.h>

namespace b2k {

// =================================================================================================
// PTX wrappers
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}

// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFFu) == 0) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}

__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// im2col-mode TMA: loads `pixelsPerColumn` consecutive output pixels (wrapping over W, H, N with the
// traversal stride) x `channelsPerPixel` channels; (off_w, off_h) select the filter tap.
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c,
                                                    int w, int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp, PTX ISA "tcgen05 matrix descriptor")
// shared-memory matrix descriptor, K-major operand:
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout
//   layout: 0 = no swizzle (8x16B core matrices; LBO = K-direction core stride, SBO = M/N-direction),
//           2 = SWIZZLE_128B (rows of 128 B, 8-row atoms; SBO = 1024 B, LBO unused)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout & 7) << 61;
    return d;
}
// instruction descriptor, kind::f16: D=f32 (bit4), A=B=f16 (0), both K-major (0), N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// =================================================================================================
// conv_f16_tcgen05
// =================================================================================================
template <int BN>
struct ConvCfg {
    static constexpr int STAGES = (BN == 128) ? 3 : 4;
    static constexpr int A_STAGE = 128 * 64 * 2;  // 16 KiB: 128 rows x 64 K-elements
    static constexpr int B_STAGE = BN * 64 * 2;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static constexpr int SMEM = STAGES * (A_STAGE + B_STAGE) + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int KB>
__global__ void __launch_bounds__(128)
conv_f16_tcgen05(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const ConvArgs p) {
    using Cfg = ConvCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int TPS = 64 / KB;            // TMA sub-tiles (filter taps) per stage; 1 when KB == 64
    constexpr int A_SUB = 128 * KB * 2;     // bytes of one A sub-tile
    constexpr int B_SUB = BN * KB * 2;
    constexpr uint32_t IDESC = make_idesc_f16(128, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_STAGE + Cfg::B_STAGE));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int m0 = blockIdx.y * 128;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int img0 = 0, p0 = 0, q0 = 0;
            if (p.a_mode == A_IM2COL) {
                img0 = m0 / p.HoWo;
                const int rem = m0 - img0 * p.HoWo;
                p0 = rem / p.Wo;
                q0 = rem - p0 * p.Wo;
            }
            const int base_w = q0 * p.stride - p.pad;
            const int base_h = p0 * p.stride - p.pad;
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a_dst = sA + s * Cfg::A_STAGE;
                uint8_t* b_dst = sB + s * Cfg::B_STAGE;
                if (KB == 64) {
                    mbar_expect_tx(&full_bar[s], Cfg::A_STAGE + Cfg::B_STAGE);
                    const int tap = kb / p.cblocks;
                    const int cb = kb - tap * p.cblocks;
                    if (p.a_mode == A_TILED) {
                        tma_load_2d(&mapA, &full_bar[s], a_dst, cb * 64, m0);
                    } else {
                        const int r = tap / p.kw;
                        const int sx = tap - r * p.kw;
                        tma_load_im2col_4d(&mapA, &full_bar[s], a_dst, cb * 64, base_w, base_h, img0,
                                           static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
                    }
                    tma_load_2d(&mapB, &full_bar[s], b_dst, kb * 64, n0);
                } else {
                    int ntaps = p.taps_phys - kb * TPS;
                    ntaps = ntaps > TPS ? TPS : ntaps;
                    mbar_expect_tx(&full_bar[s], ntaps * (A_SUB + B_SUB));
                    for (int t = 0; t < ntaps; ++t) {
                        const int tap = kb * TPS + t;
                        const int tap_a = tap < p.taps ? tap : p.taps - 1;  // padded tap: weights are zero
                        const int r = tap_a / p.kw;
                        const int sx = tap_a - r * p.kw;
                        tma_load_im2col_4d(&mapA, &full_bar[s], a_dst + t * A_SUB, 0, base_w, base_h, img0,
                                           static_cast<uint16_t>(sx), static_cast<uint16_t>(r));
                        tma_load_2d(&mapB, &full_bar[s], b_dst + t * B_SUB, tap * KB, n0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * Cfg::A_STAGE);
                const uint32_t b_addr = smem_u32(sB + s * Cfg::B_STAGE);
                if (KB == 64) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {  // 4 x (K = 16) inside one 128-byte swizzle row
                        const uint64_t ad = make_smem_desc(a_addr + j * 32, 16, 1024, 2);
                        const uint64_t bd = make_smem_desc(b_addr + j * 32, 16, 1024, 2);
                        umma_f16(tmem_base, ad, bd, IDESC, (kb > 0 || j > 0) ? 1u : 0u);
                    }
                } else {
                    int ntaps = p.taps_phys - kb * TPS;
                    ntaps = ntaps > TPS ? TPS : ntaps;
                    for (int j = 0; j < ntaps / 2; ++j) {  // one K=16 step = two 8-channel taps
                        const uint64_t ad = make_smem_desc(a_addr + 2 * j * A_SUB, A_SUB, 128, 0);
                        const uint64_t bd = make_smem_desc(b_addr + 2 * j * B_SUB, B_SUB, 128, 0);
                        umma_f16(tmem_base, ad, bd, IDESC, (kb > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
            }
            umma_commit(accum_bar);  // accumulator complete
        }
        __syncwarp();
    }

    // ================= epilogue: TMEM -> registers -> bias/residual/ReLU -> fp16 NHWC =================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    {
        const int row = warp * 32 + lane;
        const int m = m0 + row;
        const bool valid = m < p.M;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const size_t off = static_cast<size_t>(m) * p.Cout + n0;
        const float* bias = p.bias + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c, v);
            tmem_wait_ld();
            if (valid) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c) + i);
                    f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b4.x;
                    f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
                    f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
                    f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
                }
                if (p.residual != nullptr) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off + c);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint4 rv = __ldg(rp + h);
                        const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 rf = __half22float2(r2[i]);
                            f[8 * h + 2 * i + 0] += rf.x;
                            f[8 * h + 2 * i + 1] += rf.y;
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
                }
                uint4 o[2];
                __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
                for (int i = 0; i < 8; ++i) o2[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                uint4* op = reinterpret_cast<uint4*>(p.out + off + c);
                op[0] = o[0];
                op[1] = o[1];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int KB>
static int launch_one(const ConvLaunch& L, cudaStream_t stream) {
    dim3 grid(L.grid_n, L.grid_m, 1);
    conv_f16_tcgen05<BN, KB><<<grid, 128, ConvCfg<BN>::SMEM, stream>>>(L.mapA, L.mapB, L.args);
    return static_cast<int>(cudaGetLastError());
}

template <int BN, int KB>
static int init_one() {
    return static_cast<int>(cudaFuncSetAttribute(conv_f16_tcgen05<BN, KB>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<BN>::SMEM));
}

bool conv_tile_supported(int bn, int kb) {
    if (kb != 64 && kb != 8) return false;
    return bn == 32 || bn == 64 || bn == 128 || bn == 256;
}

int init_conv_kernels() {
    int e = 0;
    if ((e = init_one<32, 64>())) return e;
    if ((e = init_one<64, 64>())) return e;
    if ((e = init_one<128, 64>())) return e;
    if ((e = init_one<256, 64>())) return e;
    if ((e = init_one<32, 8>())) return e;
    if ((e = init_one<64, 8>())) return e;
    if ((e = init_one<128, 8>())) return e;
    if ((e = init_one<256, 8>())) return e;
    return 0;
}

int launch_conv_f16_tcgen05(const ConvLaunch& L, cudaStream_t stream) {
#define B2_CASE(BN_, KB_) \
    if (L.bn == BN_ && L.kb == KB_) return launch_one<BN_, KB_>(L, stream);
    B2_CASE(32, 64)
    B2_CASE(64, 64)
    B2_CASE(128, 64)
    B2_CASE(256, 64)
    B2_CASE(32, 8)
    B2_CASE(64, 8)
    B2_CASE(128, 8)
    B2_CASE(256, 8)
#undef B2_CASE
    return static_cast<int>(cudaErrorInvalidValue);
}

// =================================================================================================
// SIMT kernels
// =================================================================================================
template <typename T>
struct Acc;
template <>
struct Acc<float> {
    using type = double;  // fp32 engine: fp64 accumulate -> order-independent to ~1e-12
};
template <>
struct Acc<__half> {
    using type = float;
};
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// direct convolution, one thread per output element (co fastest)
template <typename T>
__global__ void conv_simt_kernel(SimtConvArgs a) {
    using acc_t = typename Acc<T>::type;
    const long long total = static_cast<long long>(a.N) * a.Ho * a.Wo * a.Cout_phys;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int co = static_cast<int>(idx % a.Cout_phys);
    long long pix = idx / a.Cout_phys;
    const int wo = static_cast<int>(pix % a.Wo);
    pix /= a.Wo;
    const int ho = static_cast<int>(pix % a.Ho);
    const int n = static_cast<int>(pix / a.Ho);
    T* out = reinterpret_cast<T*>(a.out);
    if (co >= a.Cout) {
        out[idx] = from_f<T>(0.0f);
        return;
    }
    const T* in = reinterpret_cast<const T*>(a.in);
    const T* w = reinterpret_cast<const T*>(a.w) + static_cast<size_t>(co) * a.taps_phys * a.Cin_phys;
    acc_t acc = 0;
    for (int r = 0; r < a.k; ++r) {
        const int hi = ho * a.stride - a.pad + r;
        if (hi < 0 || hi >= a.H) continue;
        for (int s = 0; s < a.k; ++s) {
            const int wi = wo * a.stride - a.pad + s;
            if (wi < 0 || wi >= a.W) continue;
            const T* ip = in + ((static_cast<size_t>(n) * a.H + hi) * a.W + wi) * a.Cin_phys;
            const T* wp = w + static_cast<size_t>(r * a.k + s) * a.Cin_phys;
            for (int c = 0; c < a.Cin; ++c)
                acc += static_cast<acc_t>(to_f(ip[c])) * static_cast<acc_t>(to_f(wp[c]));
        }
    }
    acc_t v = acc + static_cast<acc_t>(a.bias[co]);
    if (a.residual) v += static_cast<acc_t>(to_f(reinterpret_cast<const T*>(a.residual)[idx]));
    if (a.relu && v < 0) v = 0;
    out[idx] = from_f<T>(static_cast<float>(v));
}

int launch_conv_simt(const SimtConvArgs& a, bool half_storage, cudaStream_t stream) {
    const long long total = static_cast<long long>(a.N) * a.Ho * a.Wo * a.Cout_phys;
    const int threads = 128;
    const long long blocks = (total + threads - 1) / threads;
    if (blocks <= 0) return 0;
    if (half_storage)
        conv_simt_kernel<__half><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(a);
    else
        conv_simt_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(a);
    return static_cast<int>(cudaGetLastError());
}

// fp32 NCHW -> NHWC (channel-padded).  One thread per pixel: reads are coalesced per channel plane,
// the write is one contiguous C_phys-element row.
template <typename T>
__global__ void input_cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, int HW,
                                  int C_phys) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * HW) return;
    const int n = static_cast<int>(idx / HW);
    const int px = static_cast<int>(idx - static_cast<long long>(n) * HW);
    const float* s = src + static_cast<size_t>(n) * C * HW + px;
    T* d = dst + static_cast<size_t>(idx) * C_phys;
    for (int c = 0; c < C_phys; ++c) d[c] = from_f<T>(c < C ? __ldg(s + static_cast<size_t>(c) * HW) : 0.0f);
}

// specialisation used by the fp16 path when C_phys == 8: one 16-byte store per pixel
__global__ void input_cast_c8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int N, int C, int HW) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * HW) return;
    const int n = static_cast<int>(idx / HW);
    const int px = static_cast<int>(idx - static_cast<long long>(n) * HW);
    const float* s = src + static_cast<size_t>(n) * C * HW + px;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = c < C ? __ldg(s + static_cast<size_t>(c) * HW) : 0.0f;
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) o2[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    dst[idx] = o;
}

int launch_input_cast(const float* src, void* dst, int N, int C, int H, int W, int C_phys, bool half_storage,
                      cudaStream_t stream) {
    const long long total = static_cast<long long>(N) * H * W;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
    if (half_storage && C_phys == 8 && C <= 8)
        input_cast_c8_kernel<<<blocks, threads, 0, stream>>>(src, reinterpret_cast<uint4*>(dst), N, C, H * W);
    else if (half_storage)
        input_cast_kernel<__half><<<blocks, threads, 0, stream>>>(src, reinterpret_cast<__half*>(dst), N, C, H * W, C_phys);
    else
        input_cast_kernel<float><<<blocks, threads, 0, stream>>>(src, reinterpret_cast<float*>(dst), N, C, H * W, C_phys);
    return static_cast<int>(cudaGetLastError());
}

template <typename T>
__global__ void output_cast_kernel(const T* __restrict__ src, float* __restrict__ dst, int N, int C, int HW,
                                   int C_phys) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * C * HW) return;
    const int px = static_cast<int>(idx % HW);
    const long long t = idx / HW;
    const int c = static_cast<int>(t % C);
    const int n = static_cast<int>(t / C);
    dst[idx] = to_f(src[(static_cast<size_t>(n) * HW + px) * C_phys + c]);
}

int launch_output_cast(const void* src, float* dst, int N, int C, int H, int W, int C_phys, bool half_storage,
                       cudaStream_t stream) {
    const long long total = static_cast<long long>(N) * C * H * W;
    const int threads = 256;
    const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
    if (half_storage)
        output_cast_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(src), dst, N, C, H * W, C_phys);
    else
        output_cast_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(src), dst, N, C, H * W, C_phys);
    return static_cast<int>(cudaGetLastError());
}

// max pool, NHWC, windows clipped to the image (Caffe ceil mode produces partial windows)
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int C, int Ho,
                               int Wo, int k, int stride, int pad) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * C) return;
    const int c = static_cast<int>(idx % C);
    long long t = idx / C;
    const int wo = static_cast<int>(t % Wo);
    t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float m = -INFINITY;
    for (int r = 0; r < k; ++r) {
        const int hi = ho * stride - pad + r;
        if (hi < 0 || hi >= H) continue;
        for (int s = 0; s < k; ++s) {
            const int wi = wo * stride - pad + s;
            if (wi < 0 || wi >= W) continue;
            m = fmaxf(m, to_f(src[((static_cast<size_t>(n) * H + hi) * W + wi) * C + c]));
        }
    }
    dst[idx] = from_f<T>(m);
}

// fp16 NHWC, 8 channels (16 bytes) per thread
__global__ void maxpool_h8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int H, int W,
                                  int C8, int Ho, int Wo, int k, int stride, int pad) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * C8) return;
    const int c = static_cast<int>(idx % C8);
    long long t = idx / C8;
    const int wo = static_cast<int>(t % Wo);
    t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const __half2 ninf = __float2half2_rn(-INFINITY);
    __half2 m[4] = {ninf, ninf, ninf, ninf};
    for (int r = 0; r < k; ++r) {
        const int hi = ho * stride - pad + r;
        if (hi < 0 || hi >= H) continue;
        for (int s = 0; s < k; ++s) {
            const int wi = wo * stride - pad + s;
            if (wi < 0 || wi >= W) continue;
            const uint4 v = __ldg(src + ((static_cast<size_t>(n) * H + hi) * W + wi) * C8 + c);
            const __half2* v2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], v2[i]);
        }
    }
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) o2[i] = m[i];
    dst[idx] = o;
}

int launch_maxpool(const void* src, void* dst, int N, int H, int W, int C_phys, int Ho, int Wo, int k, int stride,
                   int pad, bool half_storage, cudaStream_t stream) {
    const int threads = 256;
    if (half_storage && C_phys % 8 == 0) {
        const long long total = static_cast<long long>(N) * Ho * Wo * (C_phys / 8);
        maxpool_h8_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
            reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), N, H, W, C_phys / 8, Ho, Wo, k, stride, pad);
    } else {
        const long long total = static_cast<long long>(N) * Ho * Wo * C_phys;
        const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
        if (half_storage)
            maxpool_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(src), reinterpret_cast<__half*>(dst), N, H, W, C_phys, Ho, Wo, k, stride, pad);
        else
            maxpool_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(src), reinterpret_cast<float*>(dst), N, H, W, C_phys, Ho, Wo, k, stride, pad);
    }
    return static_cast<int>(cudaGetLastError());
}

template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int HW, int C) {
    using acc_t = typename Acc<T>::type;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    const int c = idx % C;
    const int n = idx / C;
    acc_t acc = 0;
    for (int i = 0; i < HW; ++i) acc += static_cast<acc_t>(to_f(src[(static_cast<size_t>(n) * HW + i) * C + c]));
    dst[idx] = from_f<T>(static_cast<float>(acc / static_cast<acc_t>(HW)));
}

int launch_avgpool(const void* src, void* dst, int N, int HW, int C_phys, bool half_storage, cudaStream_t stream) {
    const int threads = 128;
    const unsigned blocks = static_cast<unsigned>((N * C_phys + threads - 1) / threads);
    if (half_storage)
        avgpool_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(src), reinterpret_cast<__half*>(dst), N, HW, C_phys);
    else
        avgpool_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(src), reinterpret_cast<float*>(dst), N, HW, C_phys);
    return static_cast<int>(cudaGetLastError());
}

// fully connected: one warp per output neuron, all batch rows (<= 8 per pass) share each weight read
template <typename T>
__global__ void fc_kernel(const T* __restrict__ in, const T* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ out, int N, int K, int Cout) {
    using acc_t = typename Acc<T>::type;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Cout) return;
    const T* wr = w + static_cast<size_t>(warp) * K;
    for (int nb = 0; nb < N; nb += 8) {
        acc_t acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0;
        for (int kk = lane; kk < K; kk += 32) {
            const acc_t wv = static_cast<acc_t>(to_f(wr[kk]));
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (nb + i < N) acc[i] += wv * static_cast<acc_t>(to_f(in[static_cast<size_t>(nb + i) * K + kk]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc_t v = acc[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && nb + i < N)
                out[static_cast<size_t>(nb + i) * Cout + warp] = static_cast<float>(v + static_cast<acc_t>(bias[warp]));
        }
    }
}

// fp16 fast path: up to 8 batch rows staged in shared memory, one warp per output neuron, 128-bit loads.
// Requires K % 256 == 0 (each lane owns 8 consecutive K elements per 256-element slab).
template <int ROWS>
__global__ void __launch_bounds__(256)
fc_h8_kernel(const __half* __restrict__ in, const __half* __restrict__ w, const float* __restrict__ bias,
             float* __restrict__ out, int N, int K, int Cout) {
    extern __shared__ uint4 s_in[];  // [ROWS][K/8]
    const int kv = K / 8;
    for (int nb = 0; nb < N; nb += ROWS) {
        const int rows = min(ROWS, N - nb);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * kv; i += blockDim.x)
            s_in[i] = __ldg(reinterpret_cast<const uint4*>(in + static_cast<size_t>(nb) * K) + i);
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int j = blockIdx.x * (blockDim.x >> 5) + warp;
        if (j < Cout) {
            float acc[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
            const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<size_t>(j) * K);
            for (int v = lane; v < kv; v += 32) {
                const uint4 wv = __ldg(wr + v);
                const __half2* w2 = reinterpret_cast<const __half2*>(&wv);
                float2 wf[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) wf[q] = __half22float2(w2[q]);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    if (r < rows) {
                        const uint4 xv = s_in[r * kv + v];
                        const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 xf = __half22float2(x2[q]);
                            acc[r] = fmaf(wf[q].x, xf.x, acc[r]);
                            acc[r] = fmaf(wf[q].y, xf.y, acc[r]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float v = acc[r];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && r < rows) out[static_cast<size_t>(nb + r) * Cout + j] = v + bias[j];
            }
        }
    }
}

int launch_fc(const void* in, const void* w, const float* bias, float* out, int N, int K, int Cout, bool half_storage,
              cudaStream_t stream) {
    if (half_storage && K % 8 == 0 && static_cast<size_t>(K) * 2 * 8 <= 96 * 1024) {
        const int threads = 256;  // 8 warps -> 8 neurons per block
        const unsigned blocks = static_cast<unsigned>((Cout + 7) / 8);
        const size_t smem = static_cast<size_t>(K) * 2 * 8;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(fc_h8_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            attr_set = true;
        }
        fc_h8_kernel<8><<<blocks, threads, smem, stream>>>(reinterpret_cast<const __half*>(in), reinterpret_cast<const __half*>(w), bias, out, N, K, Cout);
        return static_cast<int>(cudaGetLastError());
    }
    const int threads = 128;  // 4 warps
    const unsigned blocks = static_cast<unsigned>((Cout + 3) / 4);
    if (half_storage)
        fc_kernel<__half><<<blocks, threads, 0, stream>>>(reinterpret_cast<const __half*>(in), reinterpret_cast<const __half*>(w), bias, out, N, K, Cout);
    else
        fc_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(in), reinterpret_cast<const float*>(w), bias, out, N, K, Cout);
    return static_cast<int>(cudaGetLastError());
}

// row softmax, one 256-thread block per row
__global__ void softmax_kernel(const float* __restrict__ in, float* __restrict__ out, int C) {
    __shared__ float red[32];
    const float* x = in + static_cast<size_t>(blockIdx.x) * C;
    float* y = out + static_cast<size_t>(blockIdx.x) * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < C; i += blockDim.x) m = fmaxf(m, x[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < nwarp; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s += expf(x[i] - m);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < nwarp; ++i) s += red[i];
    const float inv = 1.0f / s;
    for (int i = threadIdx.x; i < C; i += blockDim.x) y[i] = expf(x[i] - m) * inv;
}

int launch_softmax(const float* in, float* out, int N, int C, cudaStream_t stream) {
    if (N <= 0) return 0;
    softmax_kernel<<<N, 256, 0, stream>>>(in, out, C);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace b2k
