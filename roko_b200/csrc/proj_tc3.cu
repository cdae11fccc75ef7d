// GRU input projection on tcgen05, persistent form (reference op: the gi half of nn.GRU,
// roko/rnn_model.py:57).  Same math and operand images as proj_tc.cu (128 x 256 tile, 3xTF32, K-major
// SWIZZLE_128B, 2 smem stages), but:
//   * one CTA per SM loops over output tiles, and the 512 TMEM columns hold TWO accumulators, so the
//     epilogue of tile i (TMEM -> registers -> +bias -> global, 128 KB) runs while the tensor core is
//     already multiplying tile i+1.  Measured on proj_tc.cu / proj_tc2.cu: the un-overlapped prologue +
//     epilogue cost as much as a K=256 main loop (0.35 of 0.61 ms at 2368 windows) while skipping a
//     third of the MMAs or the whole operand feed changed the time by < 5 %.
//   * warp roles: 0-3 A producers, 4 TMA W loader, 5 MMA issuer (+TMEM alloc), 6-9 epilogue.
//   * the epilogue transposes each 32 x 32 block through shared memory so global stores are whole 128-byte
//     row segments.
#include <stdlib.h>

#include "common.cuh"

namespace roko {

constexpr int P3_THREADS = 320;
constexpr int P3_STAGES = 2;
constexpr int P3_A_IMG = TC_BM * TC_BK * 4;            // 16 KB
constexpr int P3_W_IMG = TC_BN * TC_BK * 4;            // 32 KB
constexpr int P3_STAGE = 2 * P3_A_IMG + 2 * P3_W_IMG;  // 96 KB
constexpr int P3_EPI_ROW = 36;                          // floats per staged row (144 B: 16-byte aligned, conflict free both ways)
constexpr int P3_EPI_BYTES = 4 * 32 * P3_EPI_ROW * 4;   // one 32 x 32 staging tile per epilogue warp
constexpr int P3_SMEM = P3_STAGES * P3_STAGE + 1024 + 256 + P3_EPI_BYTES;
constexpr int P3_TMEM_COLS = 512;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ uint32_t p3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void p3_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void p3_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p3_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p3_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void p3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: start>>4 [0,14), LBO [16,30) (unused for
// swizzled K-major, 1), SBO = 1024 B >> 4 [32,46), version 1 [46,48), layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t p3_make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// issued from warp-uniform code (operands stay in uniform registers); only the elected lane executes it
__device__ __forceinline__ void p3_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t p3_elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ void p3_umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float p3_tf32_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}


template <int K>
__global__ void __launch_bounds__(P3_THREADS, 1)
proj_tc3_kernel(const float* __restrict__ A, const float* __restrict__ wimg, const float* __restrict__ bias,
                float* __restrict__ C, int M, int ntiles) {
    constexpr int KB = K / TC_BK;
    extern __shared__ unsigned char p3_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)p3_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P3_STAGES * P3_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    float* epi_stage = reinterpret_cast<float*>(smem + P3_STAGES * P3_STAGE + 256);
    const uint32_t sbase = p3_smem_u32(smem);
    const uint32_t bar0 = p3_smem_u32(bars);
    // barriers: full_a[s] = s, full_w[s] = 2+s, empty[s] = 4+s, acc_full[b] = 6+b, acc_empty[b] = 8+b
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < P3_STAGES; ++s) {
            p3_mbar_init(BAR(s), 128);
            p3_mbar_init(BAR(2 + s), 1);
            p3_mbar_init(BAR(4 + s), 1);
            p3_mbar_init(BAR(6 + s), 1);
            p3_mbar_init(BAR(8 + s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p3_smem_u32(tmem_slot)), "n"(P3_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 4) {
        // ------------------------------- A producers ----------------------------------------------
        const int chunk = tid & 7, rr = tid >> 3;
        int it = 0;                                               // k-blocks produced so far (all tiles)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / 3) * TC_BM;
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
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it & 1;
                p3_mbar_wait(BAR(4 + s), ((it >> 1) & 1) ^ 1);
                unsigned char* ahi = smem + s * P3_STAGE;
                unsigned char* alo = ahi + P3_A_IMG;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rr + 16 * i;
                    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
                    float4 h, l;
                    h.x = p3_tf32_hi(v[i].x); l.x = v[i].x - h.x;
                    h.y = p3_tf32_hi(v[i].y); l.y = v[i].y - h.y;
                    h.z = p3_tf32_hi(v[i].z); l.z = v[i].z - h.z;
                    h.w = p3_tf32_hi(v[i].w); l.w = v[i].w - h.w;
                    *reinterpret_cast<float4*>(ahi + off) = h;
                    *reinterpret_cast<float4*>(alo + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                p3_mbar_arrive(BAR(s));
                if (kb + 1 < KB) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        v[i] = valid[i] ? __ldg(reinterpret_cast<const float4*>(arow[i] + (kb + 1) * TC_BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------- W loader (TMA bulk copies) --------------------------------
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const float* src = wimg + (size_t)(tile % 3) * KB * 2 * TC_IMG;
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it & 1;
                    p3_mbar_wait(BAR(4 + s), ((it >> 1) & 1) ^ 1);
                    p3_mbar_expect_tx(BAR(2 + s), 2 * P3_W_IMG);
                    p3_bulk_g2s(sbase + s * P3_STAGE + 2 * P3_A_IMG, src + (size_t)kb * 2 * TC_IMG, 2 * P3_W_IMG, BAR(2 + s));
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------- MMA issuer (whole warp, uniform) ---------------------------
        if (tmem_d != 0) __trap();                                 // all 512 columns are ours -> base 0
        const uint32_t elected = p3_elect_one();
        int it = 0, j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            p3_mbar_wait(BAR(8 + buf), ((j >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = buf * TC_BN;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                p3_mbar_wait(BAR(s), ph);
                p3_mbar_wait(BAR(2 + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = sbase + s * P3_STAGE, a_lo = a_hi + P3_A_IMG;
                const uint32_t w_hi = a_lo + P3_A_IMG, w_lo = w_hi + P3_W_IMG;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                    const uint64_t dah = p3_make_desc(a_hi + kk * 32), dal = p3_make_desc(a_lo + kk * 32);
                    const uint64_t dwh = p3_make_desc(w_hi + kk * 32), dwl = p3_make_desc(w_lo + kk * 32);
                    p3_umma_tf32(d, dal, dwh, (kb | kk) ? 1u : 0u, elected);   // small terms first
                    p3_umma_tf32(d, dah, dwl, 1u, elected);
                    p3_umma_tf32(d, dah, dwh, 1u, elected);
                }
                if (elected) p3_umma_commit(BAR(4 + s));
                __syncwarp();
            }
            if (elected) p3_umma_commit(BAR(6 + buf));             // accumulator complete
            __syncwarp();
        }
    } else {
        // ------------------------------- epilogue warps (6..9) -------------------------------------
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        int j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            const int m0 = (tile / 3) * TC_BM, n_tile = tile % 3;
            p3_mbar_wait(BAR(6 + buf), (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // TMEM lane == tile row: this thread owns row q*32 + lane.  Rows are staged through a 32 x 32
            // shared tile so that every global store instruction writes four whole 128-byte row segments
            // (storing straight from the TMEM layout puts the 32 lanes of an instruction on 32 different rows,
            // 16 bytes each: the epilogue then takes longer than the K=256 main loop it should hide behind).
            const uint32_t taddr = ((uint32_t)(q * 32) << 16) + buf * TC_BN;
            float* T = epi_stage + (warp - 6) * 32 * P3_EPI_ROW;
            const float* brow = bias + n_tile * TC_BN;
            const int rsub = lane >> 3, csub = (lane & 7) * 4;      // read-back role: 4 rows x 8 float4 per instruction
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
#pragma unroll
                for (int qq = 0; qq < 8; ++qq)                       // my row, 32 columns -> staging tile (row-wise STS.128)
                    *reinterpret_cast<float4*>(T + lane * P3_EPI_ROW + qq * 4) =
                        make_float4(__uint_as_float(r[qq * 4 + 0]), __uint_as_float(r[qq * 4 + 1]),
                                    __uint_as_float(r[qq * 4 + 2]), __uint_as_float(r[qq * 4 + 3]));
                __syncwarp();
                const float4 b = __ldg(reinterpret_cast<const float4*>(brow + c0 + csub));
#pragma unroll
                for (int it = 0; it < 8; ++it) {                     // rows 4*it .. 4*it+3, all 32 columns, coalesced
                    const int rr = it * 4 + rsub;
                    const int m = m0 + q * 32 + rr;
                    float4 v = *reinterpret_cast<const float4*>(T + rr * P3_EPI_ROW + csub);
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    if (m < M) *reinterpret_cast<float4*>(C + (size_t)m * GI_N + n_tile * TC_BN + c0 + csub) = v;
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            p3_mbar_arrive(BAR(8 + buf));                          // accumulator may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(P3_TMEM_COLS) : "memory");
    }
}

cudaError_t proj_tc3_setup() {
    cudaError_t e = cudaFuncSetAttribute(proj_tc3_kernel<IN0P>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(proj_tc3_kernel<OUT_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM);
}

cudaError_t launch_proj_tc3(const float* A, int K, const float* wimg, const float* bias, float* C, int M,
                            int num_sms, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    const int ntiles = ((M + TC_BM - 1) / TC_BM) * (GI_N / TC_BN);
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    if (K == IN0P) proj_tc3_kernel<IN0P><<<grid, P3_THREADS, P3_SMEM, s>>>(A, wimg, bias, C, M, ntiles);
    else if (K == OUT_W) proj_tc3_kernel<OUT_W><<<grid, P3_THREADS, P3_SMEM, s>>>(A, wimg, bias, C, M, ntiles);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace roko
