// GRU input projection on the 5th-generation tensor cores (tcgen05), fp32-accurate by a 3xTF32
// split:   C = A W^T + bias,   A = A_hi + A_lo,  W = W_hi + W_lo  (tf32 each),
//          C ~= A_hi W_hi + A_hi W_lo + A_lo W_hi        (dropped term A_lo W_lo ~ 2^-22 relative)
// accumulated in fp32 in tensor memory (reference op: the gi half of nn.GRU, roko/rnn_model.py:57).
//
// One CTA computes a 128 x 256 tile of C (UMMA M=128, N=256, cta_group::1), walking K in blocks of
// 32 floats (= one 128-byte swizzle atom per row).  Warp roles (192 threads):
//   warps 0-3  A producers: coalesced LDG.128 of the fp32 activations, cvt.rna.tf32 split in
//              registers, st.shared into the K-major SWIZZLE_128B images A_hi / A_lo; afterwards
//              the epilogue: tcgen05.ld (TMEM -> registers), + bias, global stores
//   warp 4     one thread streams the pre-split, pre-swizzled W images (hi|lo, 64 KB per k-block)
//              with a single cp.async.bulk (TMA) per stage
//   warp 5     allocates 256 TMEM columns; one thread issues 12 tcgen05.mma.kind::tf32 per k-block
//              and commits to the stage's "empty" mbarrier
// Two smem stages of 96 KB (A_hi 16 + A_lo 16 + W_hi 32 + W_lo 32).
#include <stdlib.h>

#include "common.cuh"

namespace roko {

constexpr int TC_THREADS = 192;
constexpr int TC_STAGES = 2;
constexpr int A_IMG_BYTES = TC_BM * TC_BK * 4;          // 16 KB
constexpr int W_IMG_BYTES = TC_BN * TC_BK * 4;          // 32 KB
constexpr int STAGE_BYTES = 2 * A_IMG_BYTES + 2 * W_IMG_BYTES;
constexpr int TC_SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;       // the whole tensor memory: the allocation then starts at column 0
// instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, K-major both, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: start>>4 [0,14), LBO [16,30) (unused for
// swizzled K-major, 1), SBO = 1024 B >> 4 [32,46), version 1 [46,48), layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// issued from warp-uniform code (operands stay in uniform registers); only the elected lane executes it
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// CL = CTAs per cluster along the m axis.  The CL CTAs of a cluster compute different row tiles of the
// same 256 output columns, so they need the same W images: each loads 1/CL of every 64 KB k-block with
// one multicast TMA that lands in all CL shared memories, cutting the L2 -> SM traffic per CTA from
// 16 + 64 KB to 16 + 64/CL KB per k-block (the single-CTA version is L2-bandwidth bound).
template <int K, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
proj_tc_kernel(const float* __restrict__ A, const float* __restrict__ wimg, const float* __restrict__ bias,
               float* __restrict__ C, int M) {
    constexpr int KB = K / TC_BK;
    uint32_t cta_rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = smem_u32(bars);
    // barrier index: full_a[s] = s, full_w[s] = 2 + s, empty[s] = 4 + s, acc_full = 6
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x, m0 = blockIdx.y * TC_BM;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(BAR(s), 128);        // every A-producer thread arrives
            mbar_init(BAR(2 + s), 1);      // TMA thread's expect_tx arrive
            mbar_init(BAR(4 + s), CL);     // one tcgen05.commit from every CTA of the cluster
        }
        mbar_init(BAR(6), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) {   // peers' barriers must be initialised before anyone multicasts into them
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 4) {
        // ------------------------------- A producers ----------------------------------------------
        const int chunk = tid & 7, rr = tid >> 3;                 // 16-byte chunk of the row, row in a 16-row group
        const float* arow[8];
        bool valid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rr + 16 * i;
            valid[i] = (m0 + r) < M;
            arow[i] = A + (size_t)(valid[i] ? m0 + r : 0) * K + chunk * 4;
        }
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(arow[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb & 1;
            mbar_wait(BAR(4 + s), ((kb >> 1) & 1) ^ 1);
            unsigned char* ahi = smem + s * STAGE_BYTES;
            unsigned char* alo = ahi + A_IMG_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rr + 16 * i;
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
                float4 h, l;
                h.x = tf32_hi(v[i].x); l.x = v[i].x - h.x;
                h.y = tf32_hi(v[i].y); l.y = v[i].y - h.y;
                h.z = tf32_hi(v[i].z); l.z = v[i].z - h.z;
                h.w = tf32_hi(v[i].w); l.w = v[i].w - h.w;
                *reinterpret_cast<float4*>(ahi + off) = h;
                *reinterpret_cast<float4*>(alo + off) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
            mbar_arrive(BAR(s));
            if (kb + 1 < KB) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(arow[i] + (kb + 1) * TC_BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // ------------------------------- epilogue --------------------------------------------------
        mbar_wait(BAR(6), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane;                        // TMEM lane == tile row
        const int m = m0 + row;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        float* crow = C + (size_t)(m < M ? m : 0) * GI_N + n_tile * TC_BN;
        const float* brow = bias + n_tile * TC_BN;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
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
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == 4) {
        // ------------------------------- W loader (TMA bulk copies) --------------------------------
        if (lane == 0) {
            const float* src = wimg + (size_t)n_tile * KB * 2 * TC_IMG;
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb & 1;
                mbar_wait(BAR(4 + s), ((kb >> 1) & 1) ^ 1);
                mbar_expect_tx(BAR(2 + s), 2 * W_IMG_BYTES);      // the full 64 KB lands here, 1/CL from each CTA
                if (CL == 1) {
                    bulk_g2s(sbase + s * STAGE_BYTES + 2 * A_IMG_BYTES, src + (size_t)kb * 2 * TC_IMG, 2 * W_IMG_BYTES, BAR(2 + s));
                } else {
                    constexpr uint32_t PART = 2 * W_IMG_BYTES / CL;
                    const uint32_t dst = sbase + s * STAGE_BYTES + 2 * A_IMG_BYTES + cta_rank * PART;
                    const unsigned char* g = reinterpret_cast<const unsigned char*>(src + (size_t)kb * 2 * TC_IMG) + cta_rank * PART;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                        ::"r"(dst), "l"(g), "r"(PART), "r"(BAR(2 + s)), "h"((uint16_t)((1u << CL) - 1)) : "memory");
                }
            }
        }
    } else {
        // ------------------------------- MMA issuer (whole warp, uniform) ---------------------------
        if (tmem_d != 0) __trap();                                 // all 512 columns are ours -> base 0
        const uint32_t elected = elect_one();
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb & 1;
            const uint32_t ph = (kb >> 1) & 1;
            mbar_wait(BAR(s), ph);
            mbar_wait(BAR(2 + s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = sbase + s * STAGE_BYTES, a_lo = a_hi + A_IMG_BYTES;
            const uint32_t w_hi = a_lo + A_IMG_BYTES, w_lo = w_hi + W_IMG_BYTES;
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {               // UMMA K = 8 tf32 = 32 bytes inside the atom
                const uint64_t dah = make_desc(a_hi + kk * 32), dal = make_desc(a_lo + kk * 32);
                const uint64_t dwh = make_desc(w_hi + kk * 32), dwl = make_desc(w_lo + kk * 32);
                umma_tf32(0u, dal, dwh, (kb | kk) ? 1u : 0u, elected);   // small terms first
                umma_tf32(0u, dah, dwl, 1u, elected);
                umma_tf32(0u, dah, dwh, 1u, elected);
            }
            if (elected) {                                         // frees the stage (in every CTA of the cluster) when these MMAs retire
                if (CL == 1) umma_commit(BAR(4 + s));
                else
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(BAR(4 + s)), "h"((uint16_t)((1u << CL) - 1)) : "memory");
            }
            __syncwarp();
        }
        if (elected) umma_commit(BAR(6));                          // accumulator complete
        __syncwarp();
    }
    __syncthreads();
    if (CL > 1) {   // nobody leaves while a peer may still multicast into / arrive on this CTA's shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

static int g_cluster = 1;   // measured on B200: 1 -> 0.948 ms, 2 -> 0.991, 4 -> 1.114 (K=512, 2368 windows): not L2 bound

template <int K, int CL>
static cudaError_t tc_set_attr() {
    cudaError_t e = cudaFuncSetAttribute(proj_tc_kernel<K, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    return e;
}

cudaError_t proj_tc_setup() {
    if (const char* c = getenv("ROKO_B200_PROJ_CLUSTER")) g_cluster = atoi(c);
    if (g_cluster != 1 && g_cluster != 2 && g_cluster != 4) g_cluster = 1;
    cudaError_t e = tc_set_attr<IN0P, 1>();
    if (e == cudaSuccess) e = tc_set_attr<OUT_W, 1>();
    if (e == cudaSuccess) e = tc_set_attr<IN0P, 2>();
    if (e == cudaSuccess) e = tc_set_attr<OUT_W, 2>();
    if (e == cudaSuccess) e = tc_set_attr<IN0P, 4>();
    if (e == cudaSuccess) e = tc_set_attr<OUT_W, 4>();
    return e;
}

template <int K, int CL>
static cudaError_t tc_launch(const float* A, const float* wimg, const float* bias, float* C, int M, cudaStream_t s) {
    const int mtiles = (M + TC_BM - 1) / TC_BM;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(GI_N / TC_BN, ((mtiles + CL - 1) / CL) * CL);   // row tiles padded to whole clusters
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = CL; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, proj_tc_kernel<K, CL>, A, wimg, bias, C, M);
}

cudaError_t launch_proj_tc(const float* A, int K, const float* wimg, const float* bias, float* C, int M,
                           cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (K != IN0P && K != OUT_W) return cudaErrorInvalidValue;
    const int cl = (M <= TC_BM) ? 1 : g_cluster;                       // a single row tile has nothing to share
    if (K == IN0P) return cl == 4 ? tc_launch<IN0P, 4>(A, wimg, bias, C, M, s) : cl == 2 ? tc_launch<IN0P, 2>(A, wimg, bias, C, M, s)
                                                                                          : tc_launch<IN0P, 1>(A, wimg, bias, C, M, s);
    return cl == 4 ? tc_launch<OUT_W, 4>(A, wimg, bias, C, M, s) : cl == 2 ? tc_launch<OUT_W, 2>(A, wimg, bias, C, M, s)
                                                                           : tc_launch<OUT_W, 1>(A, wimg, bias, C, M, s);
}

}  // namespace roko
