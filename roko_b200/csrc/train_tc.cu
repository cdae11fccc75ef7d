// The two row-streaming products of the training front end on tcgen05 (reference ops: fc1 of
// roko/rnn_model.py:50 in train mode, and the d(embedding output) product of its backward):
//     fc1 :  a1[row][j]  = dropout(relu(b1[j] + sum_r ep[row][r] W1[j][r]))      row = (window, column, channel)
//     dep :  dep[row][r] = sum_j dap[row][j] W1[j][r]
// 576 000 rows per 128-window batch, 0.46 GB in or out each: the kernels stream 128-row blocks of the
// row-major operand through the producer warps (fp32 -> tf32 hi / lo images, K-major SWIZZLE_128B) while the
// small weight operand arrives as pre-split images by bulk copy; 3xTF32 accumulation in TMEM, two accumulators
// so the epilogue of block i overlaps the MMAs of block i+1.  Same machinery as proj_tc3.cu (which see for the
// descriptor and barrier conventions), templated on the tile width, the K extent and the epilogue.
#include "train.cuh"

namespace roko {

constexpr int T_THREADS = 320;
constexpr int T_BM = 128, T_BK = 32;
constexpr int T_A_IMG = T_BM * T_BK * 4;               // 16 KB
constexpr int T_EPI_ROW = 36;
constexpr int T_EPI_BYTES = 4 * 32 * T_EPI_ROW * 4;
__host__ __device__ constexpr int t_stage_bytes(int bn) { return 2 * T_A_IMG + 2 * bn * T_BK * 4; }
__host__ __device__ constexpr int t_stages(int bn) { return bn <= 128 ? 3 : 2; }     // what fits 227 KB
__host__ __device__ constexpr int t_smem_bytes(int bn) { return t_stages(bn) * t_stage_bytes(bn) + 1024 + 256 + T_EPI_BYTES; }

__device__ __forceinline__ uint32_t t_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void t_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t t_make_desc(uint32_t saddr) {      // K-major SWIZZLE_128B, SBO 1024 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void t_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t t_elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ void t_umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float t_tf32_hi(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

enum { TEPI_STORE = 0, TEPI_FC1 = 1 };

// C[m][0..NREAL) (row stride LDC) = epilogue(A[m][0..KREAL) (row stride LDA) x W), W as NT x KB pairs of
// hi / lo images of BN rows x 32 floats ([n tile][k block][hi | lo]); image rows beyond NREAL and columns
// beyond KREAL are zero.  Output tile t covers rows (t / NT) * 128.., columns (t % NT) * BN...
template <int BN, int KB, int NT, int LDA, int KREAL, int NREAL, int LDC, int EPI>
__global__ void __launch_bounds__(T_THREADS, 1)
tc_stream_kernel(const float* __restrict__ A, const float* __restrict__ wimg, const float* __restrict__ bias,
                 float* __restrict__ C, int M, int ntiles, DropCfg drop) {
    static_assert(KREAL % 4 == 0 && NREAL % 4 == 0 && LDA % 4 == 0 && LDC % 4 == 0 && KB * T_BK >= KREAL && NT * BN >= NREAL
                  && 2 * BN <= 512 && (EPI != 1 || NT == 1), "shape");
    constexpr int W_IMG = BN * T_BK * 4;
    constexpr int STAGE = t_stage_bytes(BN);
    constexpr int NS = t_stages(BN);                       // operand pipeline depth
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
    extern __shared__ unsigned char t_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)t_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* epi_stage = reinterpret_cast<float*>(smem + NS * STAGE + 256);
    const uint32_t sbase = t_smem_u32(smem);
    const uint32_t bar0 = t_smem_u32(bars);
    // barriers: full_a[s] = s, full_w[s] = NS+s, empty[s] = 2NS+s, acc_full[b] = 3NS+b, acc_empty[b] = 3NS+2+b
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    constexpr int B_FW = NS, B_EM = 2 * NS, B_AF = 3 * NS, B_AE = 3 * NS + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            t_mbar_init(BAR(s), 128);
            t_mbar_init(BAR(B_FW + s), 1);
            t_mbar_init(BAR(B_EM + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            t_mbar_init(BAR(B_AF + b), 1);
            t_mbar_init(BAR(B_AE + b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 4) {
        // ------------------------------- A producers ----------------------------------------------
        // two k blocks of loads in flight per thread (registers), NS blocks in shared memory
        const int chunk = tid & 7, rr = tid >> 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / NT) * T_BM;
            const float* arow[8];
            bool valid[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rr + 16 * i;
                valid[i] = (m0 + r) < M;
                arow[i] = A + (size_t)(valid[i] ? m0 + r : 0) * LDA + chunk * 4;
            }
            float4 v[2][8];
            auto load_block = [&](float4 (&dst)[8], int kb) {
                const bool kin = kb * T_BK + chunk * 4 < KREAL;             // the last k block may be partly padding
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = (valid[i] && kin) ? __ldg(reinterpret_cast<const float4*>(arow[i] + kb * T_BK))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto store_block = [&](const float4 (&src)[8]) {
                const int s = it % NS;
                t_mbar_wait(BAR(B_EM + s), ((it / NS) & 1) ^ 1);
                unsigned char* ahi = smem + s * STAGE;
                unsigned char* alo = ahi + T_A_IMG;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rr + 16 * i;
                    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
                    float4 h, l;
                    h.x = t_tf32_hi(src[i].x); l.x = src[i].x - h.x;
                    h.y = t_tf32_hi(src[i].y); l.y = src[i].y - h.y;
                    h.z = t_tf32_hi(src[i].z); l.z = src[i].z - h.z;
                    h.w = t_tf32_hi(src[i].w); l.w = src[i].w - h.w;
                    *reinterpret_cast<float4*>(ahi + off) = h;
                    *reinterpret_cast<float4*>(alo + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                t_mbar_arrive(BAR(s));
                ++it;
            };
            load_block(v[0], 0);
            if (KB > 1) load_block(v[1], 1);
#pragma unroll 1
            for (int kb = 0; kb < KB; kb += 2) {
                store_block(v[0]);
                if (kb + 2 < KB) load_block(v[0], kb + 2);
                if (kb + 1 < KB) {
                    store_block(v[1]);
                    if (kb + 3 < KB) load_block(v[1], kb + 3);
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------- W loader (bulk copies) -------------------------------------
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const float* src = wimg + (size_t)(tile % NT) * KB * 2 * (W_IMG / 4);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % NS;
                    t_mbar_wait(BAR(B_EM + s), ((it / NS) & 1) ^ 1);
                    t_mbar_expect_tx(BAR(B_FW + s), 2 * W_IMG);
                    t_bulk_g2s(sbase + s * STAGE + 2 * T_A_IMG, src + (size_t)kb * 2 * (W_IMG / 4), 2 * W_IMG, BAR(B_FW + s));
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------- MMA issuer (whole warp, uniform) ---------------------------
        if (tmem_d != 0) __trap();                                 // all 512 columns are ours -> base 0
        const uint32_t elected = t_elect_one();
        int it = 0, j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            t_mbar_wait(BAR(B_AE + buf), ((j >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = buf * BN;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % NS;
                const uint32_t ph = (it / NS) & 1;
                t_mbar_wait(BAR(s), ph);
                t_mbar_wait(BAR(B_FW + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = sbase + s * STAGE, a_lo = a_hi + T_A_IMG;
                const uint32_t w_hi = a_lo + T_A_IMG, w_lo = w_hi + W_IMG;
#pragma unroll
                for (int kk = 0; kk < T_BK / 8; ++kk) {
                    const uint64_t dah = t_make_desc(a_hi + kk * 32), dal = t_make_desc(a_lo + kk * 32);
                    const uint64_t dwh = t_make_desc(w_hi + kk * 32), dwl = t_make_desc(w_lo + kk * 32);
                    t_umma_tf32(d, dal, dwh, IDESC, (kb | kk) ? 1u : 0u, elected);   // small terms first
                    t_umma_tf32(d, dah, dwl, IDESC, 1u, elected);
                    t_umma_tf32(d, dah, dwh, IDESC, 1u, elected);
                }
                if (elected) t_umma_commit(BAR(B_EM + s));
                __syncwarp();
            }
            if (elected) t_umma_commit(BAR(B_AF + buf));
            __syncwarp();
        }
    } else {
        // ------------------------------- epilogue warps (6..9) -------------------------------------
        const int q = warp & 3;                                    // TMEM lane quarter this warp may read
        int j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            const int m0 = (tile / NT) * T_BM, nbase = (tile % NT) * BN;
            t_mbar_wait(BAR(B_AF + buf), (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = ((uint32_t)(q * 32) << 16) + buf * BN;
            float* T = epi_stage + (warp - 6) * 32 * T_EPI_ROW;
            const int rsub = lane >> 3, csub = (lane & 7) * 4;      // read-back role: 4 rows x 8 float4 per instruction
#pragma unroll 1
            for (int c0 = 0; c0 < BN && nbase + c0 < NREAL; c0 += 32) {
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
                for (int qq = 0; qq < 8; ++qq)
                    *reinterpret_cast<float4*>(T + lane * T_EPI_ROW + qq * 4) =
                        make_float4(__uint_as_float(r[qq * 4 + 0]), __uint_as_float(r[qq * 4 + 1]),
                                    __uint_as_float(r[qq * 4 + 2]), __uint_as_float(r[qq * 4 + 3]));
                __syncwarp();
                const int n = nbase + c0 + csub;                     // first of this lane's 4 columns
                if (n < NREAL) {
                    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EPI == TEPI_FC1) b = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
                    for (int it = 0; it < 8; ++it) {                 // rows 4*it .. 4*it+3, 32 columns, whole row segments
                        const int rr = it * 4 + rsub;
                        const int m = m0 + q * 32 + rr;
                        float4 v = *reinterpret_cast<const float4*>(T + rr * T_EPI_ROW + csub);
                        if (EPI == TEPI_FC1) {
                            const unsigned long long e0 = (unsigned long long)m * NREAL + n;
                            v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f);
                            v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
                            v.x = drop_keep(drop, DROP_FC1, e0) ? v.x * drop.scale : 0.f;
                            v.y = drop_keep(drop, DROP_FC1, e0 + 1) ? v.y * drop.scale : 0.f;
                            v.z = drop_keep(drop, DROP_FC1, e0 + 2) ? v.z * drop.scale : 0.f;
                            v.w = drop_keep(drop, DROP_FC1, e0 + 3) ? v.w * drop.scale : 0.f;
                        }
                        if (m < M) *reinterpret_cast<float4*>(C + (size_t)m * LDC + n) = v;
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            t_mbar_arrive(BAR(B_AE + buf));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512) : "memory");
    }
}

// ---- C += A^T B for row-major A [rows][lda], B [rows][ldb]: the weight-gradient products -----------------
// (dW1 = dap^T ep over 576 000 rows; dW_ih = dgi^T in and dW_hh = dgh^T out over 11 520 rows per layer.)
// Both operands have the REDUCTION index as the row, i.e. the transpose of what a K-major UMMA operand wants.
// Eight producer warps transpose 32-row blocks on the fly: lane = k (row inside the block), one float4 of 4
// output rows per task, written as 4 + 4 scalars (tf32 hi / lo) into the swizzled K-major images -- for a fixed
// output row the 32 lanes fill exactly one 128-byte swizzle row, so the stores are conflict free.  A CTA owns
// one 128 x BN output tile (blockIdx.y, blockIdx.z) and a contiguous range of row blocks (blockIdx.x of
// gridDim.x splits), accumulates in TMEM and adds its tile to global memory once at the end.
struct TnArgs {
    const float* A; int lda; int Mreal;
    const float* B; int ldb; int Nreal;
    float* C; int ldc;
    int rows;
};
constexpr int DW_THREADS = 288;                        // 8 producer warps (0-3 also epilogue) + 1 MMA warp
__host__ __device__ constexpr int dw_stage_bytes(int bn) { return 2 * T_A_IMG + 2 * bn * T_BK * 4; }
__host__ __device__ constexpr int dw_smem_bytes(int bn) { return 2 * dw_stage_bytes(bn) + 1024 + 256; }
constexpr int DW_MAXT = 12;                            // tasks per producer warp: (128 + 256) / 4 quads over 8 warps

template <int BN>
__global__ void __launch_bounds__(DW_THREADS, 1) tn_tc_kernel(const TnArgs g) {
    constexpr int W_IMG = BN * T_BK * 4;
    constexpr int STAGE = dw_stage_bytes(BN);
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
    extern __shared__ unsigned char t_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)t_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t sbase = t_smem_u32(smem);
    const uint32_t bar0 = t_smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };       // full[s] = s, empty[s] = 2+s, acc_full = 4
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int m0 = blockIdx.y * T_BM, n0 = blockIdx.z * BN;
    const int ma = (g.Mreal - m0) < T_BM ? (g.Mreal - m0) : T_BM;        // rows / columns of this tile that exist
    const int nb = (g.Nreal - n0) < BN ? (g.Nreal - n0) : BN;
    const int qa = ma >> 2, ntasks = qa + (nb >> 2);
    // this CTA's range of 32-row blocks
    const int nblocks = (g.rows + T_BK - 1) / T_BK;
    const int per = (nblocks + gridDim.x - 1) / gridDim.x;
    const int kb0 = blockIdx.x * per;
    const int kb1 = (kb0 + per) < nblocks ? (kb0 + per) : nblocks;
    const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { t_mbar_init(BAR(s), 256); t_mbar_init(BAR(2 + s), 1); }
        t_mbar_init(BAR(4), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // image rows the producers never write (beyond ma / nb) feed accumulator rows / columns nobody reads:
    // zero them once so they at least hold finite numbers
    for (int i = tid; i < 2 * STAGE / 16; i += DW_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 8) {
        // ------------------------------- transposing producers --------------------------------------
        for (int it = 0; it < nkb; ++it) {
            const int s = it & 1;
            const int row = (kb0 + it) * T_BK + lane;              // lane = k inside the block
            const bool rv = row < g.rows;
            float4 v[DW_MAXT];
#pragma unroll
            for (int i = 0; i < DW_MAXT; ++i) {
                const int t = warp + 8 * i;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < ntasks && rv)
                    v[i] = t < qa ? __ldg(reinterpret_cast<const float4*>(g.A + (size_t)row * g.lda + m0 + 4 * t))
                                  : __ldg(reinterpret_cast<const float4*>(g.B + (size_t)row * g.ldb + n0 + 4 * (t - qa)));
            }
            t_mbar_wait(BAR(2 + s), ((it >> 1) & 1) ^ 1);
            unsigned char* st = smem + s * STAGE;
#pragma unroll
            for (int i = 0; i < DW_MAXT; ++i) {
                const int t = warp + 8 * i;
                if (t < ntasks) {
                    const bool isA = t < qa;
                    unsigned char* hi = st + (isA ? 0 : 2 * T_A_IMG);
                    unsigned char* lo = hi + (isA ? T_A_IMG : W_IMG);
                    const int r0 = 4 * (isA ? t : t - qa);
                    const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int n = r0 + c;
                        const int off = (n >> 3) * 1024 + (n & 7) * 128 + (((lane >> 2) ^ (n & 7)) << 4) + (lane & 3) * 4;
                        const float h = t_tf32_hi(e[c]);
                        *reinterpret_cast<float*>(hi + off) = h;
                        *reinterpret_cast<float*>(lo + off) = e[c] - h;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t_mbar_arrive(BAR(s));
        }
    } else {
        // ------------------------------- MMA issuer (whole warp, uniform) ---------------------------
        const uint32_t elected = t_elect_one();
        for (int it = 0; it < nkb; ++it) {
            const int s = it & 1;
            t_mbar_wait(BAR(s), (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = sbase + s * STAGE, a_lo = a_hi + T_A_IMG;
            const uint32_t w_hi = a_lo + T_A_IMG, w_lo = w_hi + W_IMG;
#pragma unroll
            for (int kk = 0; kk < T_BK / 8; ++kk) {
                const uint64_t dah = t_make_desc(a_hi + kk * 32), dal = t_make_desc(a_lo + kk * 32);
                const uint64_t dwh = t_make_desc(w_hi + kk * 32), dwl = t_make_desc(w_lo + kk * 32);
                t_umma_tf32(tmem_d, dal, dwh, IDESC, (it | kk) ? 1u : 0u, elected);   // small terms first
                t_umma_tf32(tmem_d, dah, dwl, IDESC, 1u, elected);
                t_umma_tf32(tmem_d, dah, dwh, IDESC, 1u, elected);
            }
            if (elected) t_umma_commit(BAR(2 + s));
            __syncwarp();
        }
        if (elected) t_umma_commit(BAR(4));
        __syncwarp();
    }
    if (warp < 4 && nkb > 0) {
        // ------------------------------- epilogue: TMEM -> global adds ------------------------------
        t_mbar_wait(BAR(4), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int j = warp * 32 + lane;                              // TMEM lane == row inside the tile
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        float* crow = g.C + (size_t)(m0 + j) * g.ldc + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < nb; c0 += 32) {
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
            if (j < ma) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < nb) atomicAdd(crow + c0 + i, __uint_as_float(r[i]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(256) : "memory");
    }
}

// C[Mreal][Nreal] (row stride ldc, zeroed or holding earlier partial sums) += A^T B.  Mreal, Nreal, lda, ldb
// multiples of 4, A and B 16-byte aligned.  bn = 128 or 256: the tile width.
cudaError_t launch_tn_tc(const float* A, int lda, int Mreal, const float* B, int ldb, int Nreal, float* C, int ldc,
                         int rows, int bn, int num_sms, cudaStream_t s) {
    if (rows <= 0 || Mreal <= 0 || Nreal <= 0) return cudaSuccess;
    if ((lda | ldb | Mreal | Nreal) & 3) return cudaErrorInvalidValue;
    const int mt = (Mreal + T_BM - 1) / T_BM, nt = (Nreal + bn - 1) / bn;
    const int nblocks = (rows + T_BK - 1) / T_BK;
    int splits = num_sms / (mt * nt);                    // about one CTA per SM ...
    const int maxs = (nblocks + 7) / 8;                  // ... but at least 8 row blocks each: a tile's epilogue is up to 32 K atomics
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    TnArgs g{A, lda, Mreal, B, ldb, Nreal, C, ldc, rows};
    dim3 grid(splits, mt, nt);
    if (bn == 256) tn_tc_kernel<256><<<grid, DW_THREADS, dw_smem_bytes(256), s>>>(g);
    else if (bn == 128) tn_tc_kernel<128><<<grid, DW_THREADS, dw_smem_bytes(128), s>>>(g);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_dw1_tc(const float* dap, const float* ep, float* dW1, int rows, int num_sms, cudaStream_t s) {
    return launch_tn_tc(dap, FC1, FC1, ep, READS, READS, dW1, READS, rows, 256, num_sms, s);
}

// ---- weight images -------------------------------------------------------------------------------------
// fc1: rows n = j (128, 100 real), k = r (7 blocks of 32, 200 real):  W1[j][r]
// dep: rows n = r (256, 200 real), k = j (4 blocks of 32, 100 real):  W1[j][r]
constexpr int FC1_BN = 128, FC1_KB = 7, DEP_BN = 256, DEP_KB = 4;
constexpr int IMG_FC1_FLOATS = FC1_KB * 2 * FC1_BN * T_BK;        // 57 344
constexpr int IMG_DEP_FLOATS = DEP_KB * 2 * DEP_BN * T_BK;        // 65 536

__global__ void train_images_kernel(const float* __restrict__ W1, float* __restrict__ img) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= IMG_FC1_FLOATS + IMG_DEP_FLOATS) return;
    const bool dep = p >= IMG_FC1_FLOATS;
    int i = dep ? p - IMG_FC1_FLOATS : p;
    const int img_floats = (dep ? DEP_BN : FC1_BN) * T_BK;
    const int kb = i / (2 * img_floats);
    i %= 2 * img_floats;
    const int half = i / img_floats;
    const int ob = (i % img_floats) * 4;                             // byte offset inside the image
    const int rgrp = ob / 1024, within = ob % 1024;
    const int r8 = within / 128, pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
    const int n = rgrp * 8 + r8;
    const int k = kb * T_BK + ((pchunk ^ r8) * 4) + w4;              // undo the 128-byte swizzle
    float v = 0.f;
    if (!dep) { if (n < FC1 && k < READS) v = W1[n * READS + k]; }
    else      { if (n < READS && k < FC1) v = W1[k * READS + n]; }
    const float hi = t_tf32_hi(v);
    img[p] = half == 0 ? hi : v - hi;
}

// d(in) = dgi W_ih over both directions: rows n = input feature (tiles of 256), k = gate column of dgi
// (d*384 + g*128 + j, 24 blocks of 32):  W_ih[d][g*128 + j][n]
constexpr int DIN_BN = 256, DIN_KB = GI_N / T_BK;
__host__ __device__ constexpr int din_ntiles(int l) { return (gru_in(l) + DIN_BN - 1) / DIN_BN; }
__host__ __device__ constexpr int din_img_floats(int l) { return din_ntiles(l) * DIN_KB * 2 * DIN_BN * T_BK; }
__host__ __device__ constexpr int din_img_off(int l) {
    int off = IMG_FC1_FLOATS + IMG_DEP_FLOATS;
    for (int i = 0; i < l; ++i) off += din_img_floats(i);
    return off;
}
constexpr int IMG_TOTAL_FLOATS = din_img_off(LAYERS);

__global__ void din_images_kernel(const float* __restrict__ raw, float* __restrict__ img) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x + din_img_off(0);
    if (p >= IMG_TOTAL_FLOATS) return;
    int l = 0;
    while (l + 1 < LAYERS && p >= din_img_off(l + 1)) ++l;
    const int kin = gru_in(l);
    constexpr int img_floats = DIN_BN * T_BK;
    int i = p - din_img_off(l);
    const int ntile = i / (DIN_KB * 2 * img_floats);
    i %= DIN_KB * 2 * img_floats;
    const int kb = i / (2 * img_floats);
    i %= 2 * img_floats;
    const int half = i / img_floats;
    const int ob = (i % img_floats) * 4;
    const int rgrp = ob / 1024, within = ob % 1024;
    const int r8 = within / 128, pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
    const int n = ntile * DIN_BN + rgrp * 8 + r8;
    const int k = kb * T_BK + ((pchunk ^ r8) * 4) + w4;              // undo the 128-byte swizzle
    const int d = k / G3, q = k - d * G3;
    const float v = n < kin ? raw[raw_wih(l, d) + q * kin + n] : 0.f;
    const float hi = t_tf32_hi(v);
    img[p] = half == 0 ? hi : v - hi;
}

size_t train_tc_image_floats() { return (size_t)IMG_TOTAL_FLOATS; }

cudaError_t train_tc_setup() {
    cudaError_t e = cudaFuncSetAttribute(tc_stream_kernel<FC1_BN, FC1_KB, 1, READS, READS, FC1, FC1, TEPI_FC1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(FC1_BN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_stream_kernel<DEP_BN, DEP_KB, 1, FC1, FC1, READS, READS, TEPI_STORE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(DEP_BN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tn_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes(256));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes(128));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_stream_kernel<DIN_BN, DIN_KB, 2, GI_N, GI_N, IN0, IN0P, TEPI_STORE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(DIN_BN));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_stream_kernel<DIN_BN, DIN_KB, 1, GI_N, GI_N, OUT_W, OUT_W, TEPI_STORE>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(DIN_BN));
}

// raw: the model's flat fp32 weights (state_dict order)
cudaError_t launch_train_images(const float* raw, float* img, cudaStream_t s) {
    const int n = IMG_FC1_FLOATS + IMG_DEP_FLOATS;
    train_images_kernel<<<(n + 255) / 256, 256, 0, s>>>(raw + RAW_W1, img);
    const int nd = IMG_TOTAL_FLOATS - din_img_off(0);
    din_images_kernel<<<(nd + 255) / 256, 256, 0, s>>>(raw, img);
    return cudaGetLastError();
}

// d(in) of GRU layer l = dgi [rows][768] x W_ih (both directions) -> din [rows][gru_inp(l)]
cudaError_t launch_din_tc(int l, const float* dgi, const float* img, float* din, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    DropCfg none{0ull, 0u, 1.f};
    const int mt = (rows + T_BM - 1) / T_BM;
    if (l == 0) {
        const int ntiles = mt * 2, grid = ntiles < num_sms ? ntiles : num_sms;
        tc_stream_kernel<DIN_BN, DIN_KB, 2, GI_N, GI_N, IN0, IN0P, TEPI_STORE>
            <<<grid, T_THREADS, t_smem_bytes(DIN_BN), s>>>(dgi, img + din_img_off(0), nullptr, din, rows, ntiles, none);
    } else {
        const int grid = mt < num_sms ? mt : num_sms;
        tc_stream_kernel<DIN_BN, DIN_KB, 1, GI_N, GI_N, OUT_W, OUT_W, TEPI_STORE>
            <<<grid, T_THREADS, t_smem_bytes(DIN_BN), s>>>(dgi, img + din_img_off(l), nullptr, din, rows, mt, none);
    }
    return cudaGetLastError();
}

// a1 = dropout(relu(ep W1^T + b1)):  ep [rows][200] -> a1 [rows][100]
cudaError_t launch_fc1_tc(const float* ep, const float* img, const float* b1, float* a1, int rows, DropCfg d,
                          int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int ntiles = (rows + T_BM - 1) / T_BM;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    tc_stream_kernel<FC1_BN, FC1_KB, 1, READS, READS, FC1, FC1, TEPI_FC1>
        <<<grid, T_THREADS, t_smem_bytes(FC1_BN), s>>>(ep, img, b1, a1, rows, ntiles, d);
    return cudaGetLastError();
}

// dep = dap W1:  dap [rows][100] -> dep [rows][200]
cudaError_t launch_dep_tc(const float* dap, const float* img, float* dep, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int ntiles = (rows + T_BM - 1) / T_BM;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    DropCfg none{0ull, 0u, 1.f};
    tc_stream_kernel<DEP_BN, DEP_KB, 1, FC1, FC1, READS, READS, TEPI_STORE>
        <<<grid, T_THREADS, t_smem_bytes(DEP_BN), s>>>(dap, img + IMG_FC1_FLOATS, nullptr, dep, rows, ntiles, none);
    return cudaGetLastError();
}

}  // namespace roko
