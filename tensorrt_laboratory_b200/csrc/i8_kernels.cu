// INT8 kernels of the B200-native engine (BASELINE configs[2]: ResNet-152 int8; reference examples/ONNX/resnet50/int8.py,
// build.py:63-65 reach INT8 through TensorRT's builder).
//
//  * conv_i8_tcgen05<BN> -- implicit-GEMM convolution on the INT8 tensor path: TMA (tiled or im2col mode, 1-byte elements,
//    one 128-byte swizzle row = 128 channels) -> tcgen05.mma.kind::i8 (UMMA 128 x BN x 32, s8 x s8 -> s32 in TMEM, exact)
//    -> requantising epilogue in fp32, two fused multiply-adds (quantize.py: the CPU oracle reproduces it bit for bit)
//        t = fma(float(acc), m[c], b[c]);  t = fma(float(q_res), r, t);  t = max(t, 0);  q = clip(rint(t), +-127)
//    -> int8 -> 128-byte-swizzled staging tile -> TMA store.  Same warp roles as conv_f16_tcgen05 (warp 0 activation
//    producer, warp 1 MMA issuer, warp 2 TMEM owner, warp 3 weight producer, all four = epilogue), PDL throughout.
//  * quantize_h_to_i8_kernel -- fp16 NHWC -> int8 NHWC (channels zero-padded to the 128-channel rows of the INT8 layout)
//  * avgpool_i8_kernel       -- global average pool: int8 NHWC -> fp16 [N][C]  (exact integer sums)
//  * output_cast_i8_kernel   -- int8 NHWC -> fp32 NCHW binding (dequantised)
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace b2k {

namespace {

// D[tmem] (+)= A[smem desc] * B[smem desc], s8 x s8 -> s32; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor, kind::i8: D = s32 (c_format 2 @4), A = B = signed 8 bit (1 @7, 1 @10), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

constexpr int kI8ASub = 128 * 128;  // 128 rows x 128 K-bytes

__host__ __device__ constexpr int conv_i8_smem_layout_bytes(int bn, int stages, bool residual) {
    return stages * (kI8ASub + bn * 128) + (residual ? 128 * bn : 0) + 256 + 2 * bn * 4 + 1024;
}

}  // namespace

// four s32 -> one word of four s8 (byte i = sat_s8(q_i)): two saturating pack conversions
__device__ __forceinline__ uint32_t pack4_sat_s8(int q0, int q1, int q2, int q3) {
    uint32_t hi, out;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q3), "r"(q2), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(q1), "r"(q0), "r"(hi));
    return out;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(128)
conv_i8_tcgen05(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapOut,
                const __grid_constant__ CUtensorMap mapRes, const I8ConvArgs p) {
    constexpr int A_STAGE = kI8ASub, B_STAGE = BN * 128;
    constexpr int PIPE_BYTES = STAGES * (A_STAGE + B_STAGE);
    constexpr int TILE_BYTES = 128 * BN;     // int8 output / residual tile
    constexpr int NBOX = BN / 128;           // 128-column TMA boxes per tile row
    constexpr int NG = BN / 32;
    static_assert(TILE_BYTES <= PIPE_BYTES, "the output staging tile reuses the pipeline buffers");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const bool has_res = p.has_res != 0;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE;
    uint8_t* sOut = smem;                 // reuses the drained pipeline buffers
    uint8_t* sRes = smem + PIPE_BYTES;
    uint8_t* tail = sRes + (has_res ? TILE_BYTES : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint64_t* res_bar = accum_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
    float* s_m = reinterpret_cast<float*>(tail + 256);
    float* s_b = s_m + BN;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int m0 = blockIdx.y * 128;
    const int nk = p.num_kblocks;
    // real output channels of this N tile, rounded up to the MMA's granularity: the instruction is issued N = nv wide, only
    // nv weight rows are fetched, and the epilogue writes zeros for the rest (what the padded weights would have produced)
    int nv = p.cout_real - n0;
    nv = nv >= BN ? BN : (nv <= 0 ? 32 : ((nv + 31) / 32) * 32);
    const uint32_t b_bytes = static_cast<uint32_t>(nv) * 128u;

    // ---------------- prologue: nothing here depends on the previous kernel's output ----------------
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapOut);
        if (has_res) tma_prefetch_desc(&mapRes);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_init(res_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 3) {  // requantisation constants -> smem (published by the pre-epilogue barrier)
        for (int i = lane; i < BN; i += 32) {
            s_m[i] = __ldg(p.m + n0 + i);
            s_b[i] = __ldg(p.b + n0 + i);
        }
    }

    if (warp == 0) {
        // ================= activation producer =================
        int img0 = 0, p0 = 0, q0 = 0;
        const bool tiled = p.a_mode == A_TILED;
        if (!tiled) {
            img0 = m0 / p.HoWo;
            const int rem = m0 - img0 * p.HoWo;
            p0 = rem / p.Wo;
            q0 = rem - p0 * p.Wo;
        }
        const int base_w = q0 * p.stride_w - p.pad_w;
        const int base_h = p0 * p.stride_h - p.pad_h;
        pdl_wait();
        if (elect_one_sync() && has_res) {
            mbar_expect_tx(res_bar, TILE_BYTES);
#pragma unroll
            for (int b = 0; b < NBOX; ++b) tma_load_2d(&mapRes, res_bar, sRes + b * (128 * 128), n0 + b * 128, m0);
        }
        __syncwarp();
        int cur_cb = 0, cur_r = 0, cur_sx = 0;
        for (int i = 0; i < nk; ++i) {
            const int s = i % STAGES;
            if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full_bar[s], A_STAGE + b_bytes);
                if (tiled)
                    tma_load_2d(&mapA, &full_bar[s], sA + s * A_STAGE, cur_cb * 128, m0);
                else
                    tma_load_im2col_4d(&mapA, &full_bar[s], sA + s * A_STAGE, cur_cb * 128, base_w, base_h, img0,
                                       static_cast<uint16_t>(cur_sx), static_cast<uint16_t>(cur_r));
            }
            __syncwarp();
            if (++cur_cb == p.cblocks) {
                cur_cb = 0;
                if (++cur_sx == p.kw) {
                    cur_sx = 0;
                    ++cur_r;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_i8(128, nv);
        int cur_cb = 0;
        for (int i = 0; i < nk; ++i) {
            const int s = i % STAGES;
            mbar_wait(&full_bar[s], (i / STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
            const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
            // all-zero 32-byte slices at the end of a tap's last channel block contribute nothing: not issued
            const int nj = (cur_cb == p.cblocks - 1) ? p.last_cb_mmas : 4;
            if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // 4 x (K = 32 bytes) inside one 128-byte swizzle row
                    if (j < nj) {
                        const uint64_t ad = make_smem_desc(a_addr + j * 32, 16, 1024, 2);
                        const uint64_t bd = make_smem_desc(b_addr + j * 32, 16, 1024, 2);
                        umma_i8(tmem_base, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            __syncwarp();
            if (++cur_cb == p.cblocks) cur_cb = 0;
        }
        if (elect_one_sync()) umma_commit(accum_bar);
        __syncwarp();
    } else if (warp == 3) {
        // ================= weight producer (constants: no dependency wait) =================
        const uint8_t* src = p.wpacked + static_cast<size_t>(n0 >> 5) * 4096;
        const size_t kstride = static_cast<size_t>(p.Cout >> 5) * 4096;
        for (int i = 0; i < nk; ++i) {
            const int s = i % STAGES;
            if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
            if (elect_one_sync()) bulk_load_1d(&full_bar[s], sB + s * B_STAGE, src + i * kstride, b_bytes);
            __syncwarp();
        }
    }

    // ====== epilogue: TMEM (s32) -> requantise -> int8 -> swizzled staging tile -> TMA store ======
    pdl_wait();
    const int row = warp * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    __syncthreads();  // s_m / s_b visible; every role has left its loop: the pipeline buffers are free
    pdl_launch_dependents();
    if (has_res) mbar_wait(res_bar, 0);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float r = p.r;
    const bool relu = p.relu != 0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g * 32 >= nv) {  // padding channels (CTA-uniform): the result is zero by construction, nothing to read or compute
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = g * 32 + h * 16;
                *reinterpret_cast<uint4*>(sOut + static_cast<uint32_t>((col >> 7) * (128 * 128)) + swz_off<128>(row, (col & 127) >> 4)) =
                    make_uint4(0u, 0u, 0u, 0u);
            }
            continue;
        }
        uint32_t acc[32];
        tmem_ld32(taddr + g * 32, acc);
        tmem_wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // 16 columns = one 16-byte chunk of the int8 row
            const int col = g * 32 + h * 16;
            const int box = col >> 7;
            const int chunk = (col & 127) >> 4;
            const uint32_t so = static_cast<uint32_t>(box * (128 * 128)) + swz_off<128>(row, chunk);
            uint4 rv = make_uint4(0u, 0u, 0u, 0u);
            if (has_res) rv = *reinterpret_cast<const uint4*>(sRes + so);
            const int8_t* rq = reinterpret_cast<const int8_t*>(&rv);
            // q = clip(rint(max(t, relu ? 0 : -inf)), -127, 127) with as few issue slots as the contract allows (this loop is
            // what bounds the wide, short-K layers): the lower clamp and the ReLU are ONE fp32 max before the conversion
            // (rint is monotonic and rint(-127) = -127), the upper clamp is the saturation of the s32 -> s8 pack.
            const float lo = relu ? 0.0f : -127.0f;
            int q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = col + i;
                float t = __fmaf_rn(__int2float_rn(static_cast<int>(acc[h * 16 + i])), s_m[c], s_b[c]);
                if (has_res) t = __fmaf_rn(__int2float_rn(static_cast<int>(rq[i])), r, t);
                q[i] = __float2int_rn(fmaxf(t, lo));  // > 127 (up to INT_MAX for huge t) saturates in the pack
            }
            uint4 o;
            o.x = pack4_sat_s8(q[0], q[1], q[2], q[3]);
            o.y = pack4_sat_s8(q[4], q[5], q[6], q[7]);
            o.z = pack4_sat_s8(q[8], q[9], q[10], q[11]);
            o.w = pack4_sat_s8(q[12], q[13], q[14], q[15]);
            *reinterpret_cast<uint4*>(sOut + so) = o;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, BN);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NBOX; ++b) tma_store_2d(&mapOut, sOut + b * (128 * 128), n0 + b * 128, m0);
        tma_store_commit_and_wait_read();
    }
}

int conv_i8_smem_bytes(int bn, int stages, bool residual) { return conv_i8_smem_layout_bytes(bn, stages, residual); }
bool conv_i8_config_exists(int bn, int stages) {
    return (bn == 128 || bn == 256) && stages >= 1 && stages <= 4 && conv_i8_smem_layout_bytes(bn, stages, true) <= 227 * 1024;
}

#define B2_FOR_EACH_I8(X) X(128, 1) X(128, 2) X(128, 3) X(128, 4) X(256, 1) X(256, 2) X(256, 3)

int init_conv_i8_kernels() {
    int e = 0;
#define B2_I8_INIT(BN_, ST_)                                                                                                  \
    if ((e = static_cast<int>(cudaFuncSetAttribute(conv_i8_tcgen05<BN_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                   conv_i8_smem_layout_bytes(BN_, ST_, true)))))                             \
        return e;
    B2_FOR_EACH_I8(B2_I8_INIT)
#undef B2_I8_INIT
    return 0;
}

int launch_conv_i8_tcgen05(const I8ConvLaunch& L, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(L.grid_n), static_cast<unsigned>(L.grid_m), 1);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = static_cast<size_t>(conv_i8_smem_layout_bytes(L.bn, L.stages, L.args.has_res != 0));
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_pdl() ? 1 : 0;
#define B2_I8_CASE(BN_, ST_) \
    if (L.bn == BN_ && L.stages == ST_) return static_cast<int>(cudaLaunchKernelEx(&cfg, conv_i8_tcgen05<BN_, ST_>, L.mapA, L.mapOut, L.mapRes, L.args));
    B2_FOR_EACH_I8(B2_I8_CASE)
#undef B2_I8_CASE
    return static_cast<int>(cudaErrorInvalidValue);
}

// =================================================================================================
// SIMT helpers of the INT8 path
// =================================================================================================
// fp16 NHWC [P][C_in_phys] -> int8 NHWC [P][C_out_phys]: q = clip(rint(fl(float(h) * inv_s)), +-127), channels >= C zero.
// One thread per 16 output channels (one 16-byte store).
__global__ void quantize_h_to_i8_kernel(const __half* __restrict__ src, int8_t* __restrict__ dst, long long pixels, int C, int C_in_phys,
                                        int C_out_phys, float inv_s) {
    pdl_launch_dependents();
    pdl_wait();
    const int groups = C_out_phys / 16;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= pixels * groups) return;
    const long long px = idx / groups;
    const int c0 = static_cast<int>(idx - px * groups) * 16;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    int8_t* oq = reinterpret_cast<int8_t*>(&o);
    if (c0 < C) {
        const __half* s = src + px * C_in_phys + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (c0 + i < C) {
                int q = __float2int_rn(__fmul_rn(__half2float(s[i]), inv_s));
                q = q < -127 ? -127 : (q > 127 ? 127 : q);
                oq[i] = static_cast<int8_t>(q);
            }
        }
    }
    reinterpret_cast<uint4*>(dst)[idx] = o;
}

int launch_quantize_h_to_i8(const void* src, void* dst, long long pixels, int C, int C_in_phys, int C_out_phys, float inv_s,
                            cudaStream_t stream) {
    if (C_out_phys % 16 || pixels <= 0) return static_cast<int>(cudaErrorInvalidValue);
    const long long total = pixels * (C_out_phys / 16);
    const int threads = 256;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>((total + threads - 1) / threads));
    cfg.blockDim = dim3(threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_pdl() ? 1 : 0;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, quantize_h_to_i8_kernel, static_cast<const __half*>(src), static_cast<int8_t*>(dst), pixels, C,
                                               C_in_phys, C_out_phys, inv_s));
}

// global average pool, int8 NHWC [N][HW][C_in_phys] -> fp16 [N][C_out_phys]: h = fp16(fl(float(sum q) * k)); one thread per channel
__global__ void avgpool_i8_kernel(const int8_t* __restrict__ src, __half* __restrict__ dst, int N, int HW, int C, int C_in_phys,
                                  int C_out_phys, float k) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C_out_phys) return;
    const int c = idx % C_out_phys;
    const int n = idx / C_out_phys;
    int sum = 0;
    if (c < C) {
        const int8_t* s = src + static_cast<size_t>(n) * HW * C_in_phys + c;
        for (int i = 0; i < HW; ++i) sum += static_cast<int>(s[static_cast<size_t>(i) * C_in_phys]);
    }
    dst[idx] = __float2half_rn(__fmul_rn(__int2float_rn(sum), k));
}

int launch_avgpool_i8(const void* src, void* dst, int N, int HW, int C, int C_in_phys, int C_out_phys, float k, cudaStream_t stream) {
    const int threads = 128;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>((N * C_out_phys + threads - 1) / threads));
    cfg.blockDim = dim3(threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_pdl() ? 1 : 0;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, avgpool_i8_kernel, static_cast<const int8_t*>(src), static_cast<__half*>(dst), N, HW, C, C_in_phys,
                                               C_out_phys, k));
}

// int8 NHWC -> fp32 NCHW binding: y = fl(float(q) * s)
__global__ void output_cast_i8_kernel(const int8_t* __restrict__ src, float* __restrict__ dst, int N, int C, int HW, int C_phys, float s) {
    pdl_launch_dependents();
    pdl_wait();
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * C * HW) return;
    const int px = static_cast<int>(idx % HW);
    const long long t = idx / HW;
    const int c = static_cast<int>(t % C);
    const int n = static_cast<int>(t / C);
    dst[idx] = __fmul_rn(__int2float_rn(static_cast<int>(src[(static_cast<size_t>(n) * HW + px) * C_phys + c])), s);
}

int launch_output_cast_i8(const void* src, float* dst, int N, int C, int H, int W, int C_phys, float s, cudaStream_t stream) {
    const long long total = static_cast<long long>(N) * C * H * W;
    const int threads = 256;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>((total + threads - 1) / threads));
    cfg.blockDim = dim3(threads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_pdl() ? 1 : 0;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, output_cast_i8_kernel, static_cast<const int8_t*>(src), dst, N, C, H * W, C_phys, s));
}

}  // namespace b2k
