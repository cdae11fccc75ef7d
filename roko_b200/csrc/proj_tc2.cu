// GRU input projection on tcgen05, second tile shape:  256 x 256 output tile per CTA.
//
// Same math as proj_tc.cu (3xTF32, fp32 accumulation in TMEM; reference op: the gi half of nn.GRU,
// roko/rnn_model.py:57).  proj_tc.cu is bound by what one SM can ingest from L2 (80 KB per 1536
// tensor-cycles: `l1tex__m_xbar2l1tex_read_bytes` 6.6 TB/s chip wide, tensor pipe 49 % busy).  Here one CTA
// owns TWO 128-row accumulators (TMEM columns 0-255 and 256-511) that share every W stage, so a stage of
// 12 MMAs needs 16 KB of activations + 32 KB of W images: 48 KB per 1536 tensor-cycles.
//   k-block = 16 floats = one 64-byte row  ->  K-major SWIZZLE_64B operand images, 3 stages of 64 KB
//   warps 0-3  A producers (256 rows x 64 B per k-block: LDG.128, cvt.rna.tf32 split, st.shared) + epilogue
//   warp 4     one thread: cp.async.bulk of the pre-split, pre-swizzled W images (hi|lo, 32 KB per k-block)
//   warp 5     TMEM alloc; warp-uniform MMA issue, 12 tcgen05.mma.kind::tf32 per k-block
#include <stdlib.h>

#include "common.cuh"

namespace roko {

constexpr int T2_BM = 256, T2_BN = 256, T2_BK = 16;
constexpr int T2_THREADS = 192;
constexpr int T2_STAGES = 3;
constexpr int T2_A_IMG = 128 * T2_BK * 4;               // 8 KB: one 128-row sub-tile, hi or lo
constexpr int T2_W_IMG = T2_BN * T2_BK * 4;             // 16 KB
constexpr int T2_STAGE = 4 * T2_A_IMG + 2 * T2_W_IMG;   // 64 KB: A0hi A0lo A1hi A1lo Whi Wlo
constexpr int T2_SMEM = T2_STAGES * T2_STAGE + 1024 + 256;
constexpr int T2_TMEM_COLS = 512;
constexpr uint32_t T2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t t2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t2_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t2_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t2_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t2_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
// K-major SWIZZLE_64B descriptor: 8-row groups of 64-byte rows (SBO = 512 B), layout type 4
__device__ __forceinline__ uint64_t t2_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void t2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(T2_IDESC), "r"(accumulate), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t t2_elect() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ float t2_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
// byte offset of (row r, 16-byte chunk c) inside a SWIZZLE_64B image of 64-byte rows
__device__ __forceinline__ int t2_off(int r, int c) { return (r >> 3) * 512 + (r & 7) * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

template <int K>
__global__ void __launch_bounds__(T2_THREADS, 1)
proj_tc2_kernel(const float* __restrict__ A, const float* __restrict__ wimg, const float* __restrict__ bias,
                float* __restrict__ C, int M) {
    constexpr int KB = K / T2_BK;
    extern __shared__ unsigned char t2_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)t2_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_STAGES * T2_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const uint32_t sbase = t2_smem_u32(smem);
    const uint32_t bar0 = t2_smem_u32(bars);
    // barrier index: full_a[s] = s, full_w[s] = 3 + s, empty[s] = 6 + s, acc_full = 9
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x, m0 = blockIdx.y * T2_BM;

    if (tid == 0) {
        for (int s = 0; s < T2_STAGES; ++s) {
            t2_mbar_init(BAR(s), 128);
            t2_mbar_init(BAR(3 + s), 1);
            t2_mbar_init(BAR(6 + s), 1);
        }
        t2_mbar_init(BAR(9), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t2_smem_u32(tmem_slot)), "n"(T2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 4) {
        // ------------------------------- A producers ----------------------------------------------
        const int chunk = tid & 3, rr = tid >> 2;                 // 16-byte chunk of the 64-byte row; row in a 32-row group
        const float* arow[8];
        bool valid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rr + 32 * i;                            // 0..255
            valid[i] = (m0 + r) < M;
            arow[i] = A + (size_t)(valid[i] ? m0 + r : 0) * K + chunk * 4;
        }
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(arow[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % T2_STAGES;
            t2_mbar_wait(BAR(6 + s), ((kb / T2_STAGES) & 1) ^ 1);
            unsigned char* st = smem + s * T2_STAGE;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rr + 32 * i;
                unsigned char* ahi = st + (r >> 7) * (2 * T2_A_IMG);      // sub-tile 0: [0,16K), sub-tile 1: [16K,32K)
                const int off = t2_off(r & 127, chunk);
                float4 h, l;
                h.x = t2_tf32(v[i].x); l.x = v[i].x - h.x;
                h.y = t2_tf32(v[i].y); l.y = v[i].y - h.y;
                h.z = t2_tf32(v[i].z); l.z = v[i].z - h.z;
                h.w = t2_tf32(v[i].w); l.w = v[i].w - h.w;
                *reinterpret_cast<float4*>(ahi + off) = h;
                *reinterpret_cast<float4*>(ahi + T2_A_IMG + off) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t2_mbar_arrive(BAR(s));
            if (kb + 1 < KB) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(arow[i] + (kb + 1) * T2_BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // ------------------------------- epilogue --------------------------------------------------
        t2_mbar_wait(BAR(9), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float* brow = bias + n_tile * T2_BN;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            const int m = m0 + sub * 128 + warp * 32 + lane;     // TMEM lane == row of the sub-tile
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * T2_BN);
            float* crow = C + (size_t)(m < M ? m : 0) * GI_N + n_tile * T2_BN;
#pragma unroll 1
            for (int c0 = 0; c0 < T2_BN; c0 += 32) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < M) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(brow + c0 + q * 4));
                        float4 o;
                        o.x = __uint_as_float(r[q * 4 + 0]) + b.x;
                        o.y = __uint_as_float(r[q * 4 + 1]) + b.y;
                        o.z = __uint_as_float(r[q * 4 + 2]) + b.z;
                        o.w = __uint_as_float(r[q * 4 + 3]) + b.w;
                        *reinterpret_cast<float4*>(crow + c0 + q * 4) = o;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ------------------------------- W loader (TMA bulk copies) --------------------------------
        if (lane == 0) {
            const float* src = wimg + (size_t)n_tile * KB * 2 * (T2_BN * T2_BK);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % T2_STAGES;
                t2_mbar_wait(BAR(6 + s), ((kb / T2_STAGES) & 1) ^ 1);
                t2_mbar_expect_tx(BAR(3 + s), 2 * T2_W_IMG);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sbase + s * T2_STAGE + 4 * T2_A_IMG), "l"(src + (size_t)kb * 2 * (T2_BN * T2_BK)),
                               "r"(2 * T2_W_IMG), "r"(BAR(3 + s)) : "memory");
            }
        }
    } else {
        // ------------------------------- MMA issuer (whole warp, uniform) ---------------------------
        if (tmem_d != 0) __trap();                                 // all 512 columns are ours -> base 0
        const uint32_t elected = t2_elect();
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % T2_STAGES;
            const uint32_t ph = (kb / T2_STAGES) & 1;
            t2_mbar_wait(BAR(s), ph);
            t2_mbar_wait(BAR(3 + s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = sbase + s * T2_STAGE;
            const uint32_t w_hi = st + 4 * T2_A_IMG, w_lo = w_hi + T2_W_IMG;
#pragma unroll
            for (int kk = 0; kk < T2_BK / 8; ++kk) {
                const uint64_t dwh = t2_desc(w_hi + kk * 32), dwl = t2_desc(w_lo + kk * 32);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const uint32_t a_hi = st + sub * (2 * T2_A_IMG), a_lo = a_hi + T2_A_IMG;
                    const uint64_t dah = t2_desc(a_hi + kk * 32), dal = t2_desc(a_lo + kk * 32);
                    const uint32_t d = (uint32_t)(sub * T2_BN);
                    t2_mma(d, dal, dwh, (kb | kk) ? 1u : 0u, elected);   // small terms first
                    t2_mma(d, dah, dwl, 1u, elected);
                    t2_mma(d, dah, dwh, 1u, elected);
                }
            }
            if (elected)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(6 + s)) : "memory");
            __syncwarp();
        }
        if (elected)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(9)) : "memory");
        __syncwarp();
    }
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(T2_TMEM_COLS) : "memory");
    }
}

cudaError_t proj_tc2_setup() {
    cudaError_t e = cudaFuncSetAttribute(proj_tc2_kernel<IN0P>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(proj_tc2_kernel<OUT_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM);
}

cudaError_t launch_proj_tc2(const float* A, int K, const float* wimg, const float* bias, float* C, int M,
                            cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    dim3 grid(GI_N / T2_BN, (M + T2_BM - 1) / T2_BM);
    if (K == IN0P) proj_tc2_kernel<IN0P><<<grid, T2_THREADS, T2_SMEM, s>>>(A, wimg, bias, C, M);
    else if (K == OUT_W) proj_tc2_kernel<OUT_W><<<grid, T2_THREADS, T2_SMEM, s>>>(A, wimg, bias, C, M);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace roko
