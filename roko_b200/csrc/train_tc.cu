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
constexpr int T_EPI_BYTES = 26 * 1024;                 // store epilogues: 4 x 32 x 36 floats; d(ep) epilogue: 24 KB of bins + staged codes
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

__device__ __forceinline__ float t_lds(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void t_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }

enum { TEPI_STORE = 0, TEPI_FC1 = 1, TEPI_DE = 2 };
constexpr int EP_VALUES = NCODES * EMB;                // the embedding table, 600 floats
constexpr int DE_SETS = 4;                             // independent bin sets of the TEPI_DE epilogue (column r uses set r & 3)
constexpr int DE_CROW = 208;                           // staged codes: bytes per (window, column) row (200 used)

// C[m][0..NREAL) (row stride LDC) = epilogue(A[m][0..KREAL) (row stride LDA) x W), W as NT x KB pairs of
// hi / lo images of BN rows x 32 floats ([n tile][k block][hi | lo]); image rows beyond NREAL and columns
// beyond KREAL are zero.  Output tile t covers rows (t / NT) * 128.., columns (t % NT) * BN...
//
// AGEN = 1: A is the masked embedding ep (rows = (window, column, channel), k = read), rebuilt from gen's codes and keep
// bits instead of loaded.  EPI = TEPI_DE: the output tile is d(ep); instead of storing it, every row adds its kept
// columns into 12 per-code bins (shared memory, one column of bins per thread) -> dE[code][channel of the row].
template <int BN, int KB, int NT, int LDA, int KREAL, int NREAL, int LDC, int EPI, int AGEN = 0>
__global__ void __launch_bounds__(T_THREADS, 1)
tc_stream_kernel(const float* __restrict__ A, const float* __restrict__ wimg, const float* __restrict__ bias,
                 float* __restrict__ C, int M, int ntiles, DropCfg drop, const EpGen gen) {
    static_assert(KREAL % 4 == 0 && NREAL % 4 == 0 && LDA % 4 == 0 && LDC % 4 == 0 && KB * T_BK >= KREAL && NT * BN >= NREAL
                  && 2 * BN <= 512 && (EPI != 1 || NT == 1), "shape");
    static_assert(!AGEN || (KREAL == READS && LDA == READS), "generated A is the masked embedding");
    static_assert(EPI != TEPI_DE || (NREAL == READS && NT == 1 && DE_SETS * NCODES * 128 * 4 + 2 * 4 * DE_CROW <= T_EPI_BYTES), "d(ep) epilogue");
    static_assert(4 * 32 * T_EPI_ROW * 4 <= T_EPI_BYTES, "store epilogue staging");
    __shared__ float Es[AGEN ? EP_VALUES : 1];             // E * scale
    constexpr int W_IMG = BN * T_BK * 4;
    constexpr int STAGE = t_stage_bytes(BN);
    constexpr int NS = t_stages(BN);                       // operand pipeline depth
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
    extern __shared__ unsigned char t_smem_raw[];
    // 1 KB alignment as an OFFSET into the shared array: the pointer keeps its address space (LDS / STS, not generic LD / ST)
    unsigned char* smem = t_smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(t_smem_raw) & 1023u)) & 1023u);
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
    if constexpr (AGEN) for (int i = tid; i < EP_VALUES; i += T_THREADS) Es[i] = gen.E[i] * gen.scale;
    if constexpr (EPI == TEPI_DE)
        for (int i = tid; i < DE_SETS * NCODES * 128; i += T_THREADS) epi_stage[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (AGEN && warp < 4) {
        // ------------------------------- A producers, operand rebuilt from codes + keep bits ----------
        // thread = 4 consecutive reads (chunk) of rows rr + 16 i; per k block and row one word of codes and one nibble of bits
        const int chunk = tid & 7, rr = tid >> 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m0 = (tile / NT) * T_BM;
            const uint8_t* crow[8];
            const uint32_t* brow[8];
            int ev[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + rr + 16 * i;
                const bool valid = m < M;
                const int bp = valid ? m / EMB : 0;
                ev[i] = valid ? m - bp * EMB : -1;
                crow[i] = gen.xt + (size_t)bp * READS + chunk * 4;
                brow[i] = gen.bits + (size_t)bp * MASK_WORDS;
            }
            uint32_t cw[2][8], kn[2][8];
            auto load_block = [&](uint32_t (&c)[8], uint32_t (&k)[8], int kb) {
                const int k0 = kb * T_BK + chunk * 4;
                const bool kin = k0 < KREAL;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    c[i] = 0u; k[i] = 0u;
                    if (ev[i] >= 0 && kin) {
                        const int idx = ev[i] * READS + k0;                 // multiple of 4: the nibble never straddles a word
                        c[i] = __ldg(reinterpret_cast<const uint32_t*>(crow[i] + kb * T_BK));
                        k[i] = (__ldg(brow[i] + (idx >> 5)) >> (idx & 31)) & 15u;
                    }
                }
            };
            auto store_block = [&](const uint32_t (&c)[8], const uint32_t (&k)[8]) {
                const int s = it % NS;
                t_mbar_wait(BAR(B_EM + s), ((it / NS) & 1) ^ 1);
                unsigned char* ahi = smem + s * STAGE;
                unsigned char* alo = ahi + T_A_IMG;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rr + 16 * i;
                    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
                    const int e = ev[i] < 0 ? 0 : ev[i];
                    float4 v, h, l;
                    v.x = (k[i] & 1u) ? Es[(c[i] & 255u) * EMB + e] : 0.f;
                    v.y = (k[i] & 2u) ? Es[((c[i] >> 8) & 255u) * EMB + e] : 0.f;
                    v.z = (k[i] & 4u) ? Es[((c[i] >> 16) & 255u) * EMB + e] : 0.f;
                    v.w = (k[i] & 8u) ? Es[(c[i] >> 24) * EMB + e] : 0.f;
                    h.x = t_tf32_hi(v.x); l.x = v.x - h.x;
                    h.y = t_tf32_hi(v.y); l.y = v.y - h.y;
                    h.z = t_tf32_hi(v.z); l.z = v.z - h.z;
                    h.w = t_tf32_hi(v.w); l.w = v.w - h.w;
                    *reinterpret_cast<float4*>(ahi + off) = h;
                    *reinterpret_cast<float4*>(alo + off) = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                t_mbar_arrive(BAR(s));
                ++it;
            };
            load_block(cw[0], kn[0], 0);
            if (KB > 1) load_block(cw[1], kn[1], 1);
#pragma unroll 1
            for (int kb = 0; kb < KB; kb += 2) {
                store_block(cw[0], kn[0]);
                if (kb + 2 < KB) load_block(cw[0], kn[0], kb + 2);
                if (kb + 1 < KB) {
                    store_block(cw[1], kn[1]);
                    if (kb + 3 < KB) load_block(cw[1], kn[1], kb + 3);
                }
            }
        }
    } else if (warp < 4) {
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
    } else if (EPI == TEPI_DE) {
        // ------------------------------- epilogue warps (6..9): d(ep) tile -> per-code sums ---------
        // thread = row (one channel e of one (window, column)); column r of the tile belongs to code xt[bp][r] and counts
        // when its keep bit is set.  Bins [set][code][thread] in shared memory: a thread only ever adds into its own column
        // of bins (bank = lane); column r uses set r & 3, so four read-modify-write chains are in flight at a time.
        // After a tile, thread t gathers the bins of targets T = t + 128 k (T = code * 50 + channel; the <= 3 rows of the
        // tile with that channel) into registers that live for the whole kernel -- no shared-memory atomics -- and clears
        // them.  The next tile's keep bits (8 words per row) and codes (the <= 4 (window, column)s a tile spans, staged
        // in shared memory) are fetched while the current tile is processed.
        const int q = warp & 3, tl = q * 32 + lane;
        const uint32_t bins0 = t_smem_u32(epi_stage);                 // cell (set, code, row) at ((set * 12 + code) * 128 + row) * 4
        const uint32_t mybins = bins0 + 4u * tl;
        unsigned char* sc = reinterpret_cast<unsigned char*>(epi_stage + DE_SETS * NCODES * 128);     // [2][4][DE_CROW] codes
        const int nbp = (M + EMB - 1) / EMB;
        constexpr int NTGT = (EP_VALUES + 127) / 128;                 // 5 targets per thread
        float tacc[NTGT];
        int tch[NTGT];                                                // channel of the target, -1 = none
        uint32_t tbin[NTGT];                                          // address of cell (set 0, code of the target, row 0)
#pragma unroll
        for (int k = 0; k < NTGT; ++k) {
            const int T = tl + 128 * k, c = T / EMB;
            tacc[k] = 0.f;
            tch[k] = T < EP_VALUES ? T - c * EMB : -1;
            tbin[k] = bins0 + 4u * (c * 128);
        }
        uint32_t Wn[8];
        uint2 cpre = make_uint2(0u, 0u);
        auto prefetch = [&](int tile) {
            const int m = tile * T_BM + tl;
            const bool valid = tile < ntiles && m < M;
            const int bp = valid ? m / EMB : 0, e = valid ? m - bp * EMB : 0;
            const uint32_t* bw = gen.bits + (size_t)bp * MASK_WORDS + ((e * READS) >> 5);
#pragma unroll
            for (int k = 0; k < 8; ++k) Wn[k] = valid ? __ldg(bw + k) : 0u;
            const int bpq = tl / 26, off = tl - bpq * 26, sbp = (tile * T_BM) / EMB + bpq;      // 4 rows of 26 x 8 bytes
            cpre = make_uint2(0u, 0u);
            if (tile < ntiles && bpq < 4 && off < READS / 8 && sbp < nbp)
                cpre = __ldg(reinterpret_cast<const uint2*>(gen.xt + (size_t)sbp * READS) + off);
        };
        prefetch(blockIdx.x);
        int j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j) {
            const uint32_t buf = j & 1;
            const int m = tile * T_BM + tl;
            const bool valid = m < M;
            const int bp = valid ? m / EMB : 0, e = valid ? m - bp * EMB : 0;
            const int bpl = valid ? bp - (tile * T_BM) / EMB : 0;
            const uint32_t sh = (uint32_t)(e * READS) & 31u;
            uint32_t W[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) W[k] = Wn[k];
            unsigned char* scb = sc + buf * 4 * DE_CROW;
            if (tl < 4 * 26) *reinterpret_cast<uint2*>(scb + (tl / 26) * DE_CROW + (tl % 26) * 8) = cpre;
            asm volatile("bar.sync 1, 128;" ::: "memory");            // codes staged; the previous tile's gather is complete
            prefetch(tile + gridDim.x);
            const unsigned char* crow = scb + bpl * DE_CROW;
            t_mbar_wait(BAR(B_AF + buf), (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = ((uint32_t)(q * 32) << 16) + buf * BN;
#pragma unroll
            for (int c = 0; c < (NREAL + 31) / 32; ++c) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr + (uint32_t)(32 * c)));
                const int ncol = NREAL - 32 * c < 32 ? NREAL - 32 * c : 32;          // 32, or 8 in the last group
                uint32_t keep = __funnelshift_r(W[c], W[c + 1], sh);
                if (ncol < 32) keep &= (1u << ncol) - 1u;
                if (!valid) keep = 0u;
                uint32_t cw[8];
#pragma unroll
                for (int w = 0; w < 8; w += 2) {
                    if (4 * w < ncol) {
                        const uint2 t2 = *reinterpret_cast<const uint2*>(crow + 32 * c + 4 * w);
                        cw[w] = t2.x; cw[w + 1] = t2.y;
                    } else { cw[w] = 0u; cw[w + 1] = 0u; }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (4 * g < ncol) {
                        constexpr uint32_t SET = NCODES * 128 * 4;          // bytes per bin set
                        const uint32_t a0 = mybins + ((cw[g] << 9) & 0x1FE00u);
                        const uint32_t a1 = mybins + SET + ((cw[g] << 1) & 0x1FE00u);
                        const uint32_t a2 = mybins + 2 * SET + ((cw[g] >> 7) & 0x1FE00u);
                        const uint32_t a3 = mybins + 3 * SET + ((cw[g] >> 15) & 0x1FE00u);
                        const float v0 = t_lds(a0), v1 = t_lds(a1), v2 = t_lds(a2), v3 = t_lds(a3);
                        t_sts(a0, v0 + (((keep >> (4 * g)) & 1u) ? __uint_as_float(r[4 * g]) : 0.f));
                        t_sts(a1, v1 + (((keep >> (4 * g + 1)) & 1u) ? __uint_as_float(r[4 * g + 1]) : 0.f));
                        t_sts(a2, v2 + (((keep >> (4 * g + 2)) & 1u) ? __uint_as_float(r[4 * g + 2]) : 0.f));
                        t_sts(a3, v3 + (((keep >> (4 * g + 3)) & 1u) ? __uint_as_float(r[4 * g + 3]) : 0.f));
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            t_mbar_arrive(BAR(B_AE + buf));
            asm volatile("bar.sync 2, 128;" ::: "memory");            // every row's bins are final
            const int off50 = (tile * T_BM) % EMB;                   // channel of the tile's first row
#pragma unroll
            for (int k = 0; k < NTGT; ++k) {
                if (tch[k] >= 0) {
                    int row = tch[k] - off50;
                    if (row < 0) row += EMB;
#pragma unroll
                    for (int rep = 0; rep < 3; ++rep, row += EMB) {
                        if (row < T_BM) {
#pragma unroll
                            for (int st = 0; st < DE_SETS; ++st) {
                                const uint32_t a = tbin[k] + 4u * row + st * (NCODES * 128 * 4);
                                tacc[k] += t_lds(a);
                                t_sts(a, 0.f);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NTGT; ++k)
            if (tch[k] >= 0 && tacc[k] != 0.f) atomicAdd(gen.dE + tl + 128 * k, tacc[k] * gen.scale);
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
// Eight producer warps build the images of one 32-row block at a time: a thread takes (output column, 4 consecutive
// rows), so the transposition happens in the loads (4-byte, lanes on consecutive columns: coalesced) and each image
// gets one 16-byte store (tf32 hi / lo).  A CTA owns one 128 x BN output tile (blockIdx.y, blockIdx.z) and a contiguous
// range of row blocks (blockIdx.x of gridDim.x splits), accumulates in TMEM and adds its tile to global memory once at
// the end (staged through shared memory: row-contiguous atomics).
// One launch covers a LIST of such products over the same rows (blockIdx.y = tile): the six dW_ih and dW_hh tiles of both
// directions of a GRU layer go out together, so the fixed cost per launch (operand-stage clearing, tensor-memory
// allocation, the atomics of the tile epilogue) is paid once per layer and every CTA gets a longer row range.
struct TnTile {                // C[ma][nb] += A[:, 0..ma)^T B[:, 0..nb): pointers already at the tile's first column / element
    const float* A; const float* B; float* C;
    int lda, ldb, ldc;
    int ma, nb;                // <= 128, <= 256
    int n_mma;                 // MMA width for this tile: 128 or 256 (>= nb)
};
constexpr int TN_MAX_TILES = 18;
struct TnArgs {
    TnTile tile[TN_MAX_TILES];
    int rows;
    EpGen gen;                 // BGEN: B is the masked embedding (rows = (window, column, channel), 200 columns), rebuilt not loaded
};
constexpr int DW_THREADS = 288;                        // 8 producer warps (0-3 also epilogue) + 1 MMA warp
__host__ __device__ constexpr int dw_stage_bytes(int bn) { return 2 * T_A_IMG + 2 * bn * T_BK * 4; }
__host__ __device__ constexpr int dw_smem_bytes(int bn) { return 2 * dw_stage_bytes(bn) + 1024 + 256; }

template <int BN, int BGEN = 0>
__global__ void __launch_bounds__(DW_THREADS, 1) tn_tc_kernel(const __grid_constant__ TnArgs g) {
    __shared__ float Es[BGEN ? EP_VALUES : 1];             // E * scale
    constexpr int W_IMG = BN * T_BK * 4;
    constexpr int STAGE = dw_stage_bytes(BN);
    const TnTile& tl_ = g.tile[blockIdx.y];
    const float* const tA = tl_.A;
    const float* const tB = tl_.B;
    float* const tC = tl_.C;
    const int lda = tl_.lda, ldb = tl_.ldb, ldc = tl_.ldc;
    const uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tl_.n_mma >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
    extern __shared__ unsigned char t_smem_raw[];
    // 1 KB alignment as an OFFSET into the shared array: the pointer keeps its address space (LDS / STS, not generic LD / ST)
    unsigned char* smem = t_smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(t_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t sbase = t_smem_u32(smem);
    const uint32_t bar0 = t_smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };       // full[s] = s, empty[s] = 2+s, acc_full = 4
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int ma = tl_.ma, nb = tl_.nb;                                  // rows / columns of this tile that exist
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
    if constexpr (BGEN) for (int i = tid; i < EP_VALUES; i += DW_THREADS) Es[i] = g.gen.E[i] * g.gen.scale;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(t_smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (BGEN && warp < 8) {
        // ------------------------------- producers, B = masked embedding rebuilt in place -----------
        // Nothing to transpose on the B side: thread r (< 200) owns image row r = one read and computes its 32 values of
        // the block (rows = consecutive channels of one or two (window, column)s) from the read's code(s) and the
        // per-read keep words (bitsT) -> eight 16-byte stores per image.  The A side (dap, row-major [row][100]) is read
        // as (column j, 4 consecutive rows): lanes = consecutive j, coalesced 4-byte loads, one 16-byte store per image.
        // Everything that does not depend on the block is computed once; loads of block it + 1 are issued before block
        // it is stored.
        constexpr int NA = (FC1 * 8 + 255) / 256;                       // 4 A tasks per thread
        int a_off[NA], a_k[NA];
        const float* a_ptr[NA];
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const int task = tid + 256 * u, kq = task / FC1, j = task - kq * FC1;
            a_k[u] = task < FC1 * 8 ? 4 * kq : (1 << 28);              // out-of-range tasks fail every row test
            a_off[u] = (j >> 3) * 1024 + (j & 7) * 128 + (((kq & 7) ^ (j & 7)) << 4);
            a_ptr[u] = tA + ((size_t)kb0 * T_BK + 4 * (kq & 7)) * FC1 + j;
        }
        const bool b_on = tid < READS;
        const int r = b_on ? tid : 0;
        const int b_off = 2 * T_A_IMG + (r >> 3) * 1024 + (r & 7) * 128;
        const int nbp = g.rows / EMB;
        struct Loads { float a[NA][4]; uint2 wA, wB; uint32_t cA, cB; };
        auto load_block = [&](Loads& L, int it) {
            const int k0 = (kb0 + it) * T_BK;
#pragma unroll
            for (int u = 0; u < NA; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    L.a[u][i] = (k0 + a_k[u] + i < g.rows) ? __ldg(a_ptr[u] + ((size_t)it * T_BK + i) * FC1) : 0.f;
            const int bp0 = k0 / EMB;
            L.wA = L.wB = make_uint2(0u, 0u);
            L.cA = L.cB = 0u;
            if (b_on && bp0 < nbp) {
                L.wA = __ldg(reinterpret_cast<const uint2*>(g.gen.bitsT) + (size_t)bp0 * READS + r);
                L.cA = __ldg(g.gen.xt + (size_t)bp0 * READS + r);
            }
            if (b_on && bp0 + 1 < nbp) {
                L.wB = __ldg(reinterpret_cast<const uint2*>(g.gen.bitsT) + (size_t)(bp0 + 1) * READS + r);
                L.cB = __ldg(g.gen.xt + (size_t)(bp0 + 1) * READS + r);
            }
        };
        auto store_block = [&](const Loads& L, int it) {
            const int s = it & 1, k0 = (kb0 + it) * T_BK;
            t_mbar_wait(BAR(2 + s), ((it >> 1) & 1) ^ 1);
            unsigned char* st = smem + s * STAGE;
#pragma unroll
            for (int u = 0; u < NA; ++u) {
                if (a_k[u] < T_BK) {
                    float4 h, l;
                    h.x = t_tf32_hi(L.a[u][0]); l.x = L.a[u][0] - h.x;
                    h.y = t_tf32_hi(L.a[u][1]); l.y = L.a[u][1] - h.y;
                    h.z = t_tf32_hi(L.a[u][2]); l.z = L.a[u][2] - h.z;
                    h.w = t_tf32_hi(L.a[u][3]); l.w = L.a[u][3] - h.w;
                    *reinterpret_cast<float4*>(st + a_off[u]) = h;
                    *reinterpret_cast<float4*>(st + T_A_IMG + a_off[u]) = l;
                }
            }
            if (b_on) {
                const int bp0 = k0 / EMB, e0 = k0 - bp0 * EMB;
                const int n1 = EMB - e0;                                // rows of the block inside (window, column) bp0 (>= 32: all)
                const unsigned long long mA = ((unsigned long long)L.wA.y << 32) | L.wA.x;
                const unsigned long long mB = ((unsigned long long)L.wB.y << 32) | L.wB.x;
                uint32_t keep = (uint32_t)(mA >> e0);                   // channels >= 50 hold zeros
                if (n1 < T_BK) keep |= (uint32_t)mB << n1;
                const int baseA = (int)L.cA * EMB + e0, baseB = (int)L.cB * EMB - n1;
#pragma unroll
                for (int kq = 0; kq < 8; ++kq) {
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = 4 * kq + i;
                        v[i] = ((keep >> k) & 1u) ? Es[(k < n1 ? baseA : baseB) + k] : 0.f;
                    }
                    const int off = b_off + ((kq ^ (r & 7)) << 4);
                    float4 h, l;
                    h.x = t_tf32_hi(v[0]); l.x = v[0] - h.x;
                    h.y = t_tf32_hi(v[1]); l.y = v[1] - h.y;
                    h.z = t_tf32_hi(v[2]); l.z = v[2] - h.z;
                    h.w = t_tf32_hi(v[3]); l.w = v[3] - h.w;
                    *reinterpret_cast<float4*>(st + off) = h;
                    *reinterpret_cast<float4*>(st + W_IMG + off) = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t_mbar_arrive(BAR(s));
        };
        Loads L0, L1;
        if (nkb > 0) load_block(L0, 0);
#pragma unroll 1
        for (int it = 0; it < nkb; it += 2) {
            if (it + 1 < nkb) load_block(L1, it + 1);
            store_block(L0, it);
            if (it + 1 < nkb) {
                if (it + 2 < nkb) load_block(L0, it + 2);
                store_block(L1, it + 1);
            }
        }
    } else if (warp < 8) {
        // ------------------------------- producers, both operands loaded ------------------------------
        // A and B are row-major with the reduction index as the row, the transpose of a K-major image.  Thread = (column
        // c, 4 consecutive rows): four 4-byte loads with lanes on consecutive columns (coalesced), then ONE 16-byte store
        // per image -- no scalar transposing stores.  Loads of block it + 1 are issued before block it is stored.
        constexpr int NA = T_BM * 8 / 256, NB = BN * 8 / 256;          // 4 and 4 / 8 tasks per thread
        int a_off[NA], a_k[NA], b_off[NB], b_k[NB];
        const float* a_ptr[NA];
        const float* b_ptr[NB];
#pragma unroll
        for (int u = 0; u < NA; ++u) {
            const int task = tid + 256 * u, kq = task / T_BM, c = task - kq * T_BM;
            a_k[u] = c < ma ? 4 * kq : (1 << 28);                        // columns past the tile fail every row test
            a_off[u] = (c >> 3) * 1024 + (c & 7) * 128 + ((kq ^ (c & 7)) << 4);
            a_ptr[u] = tA + ((size_t)kb0 * T_BK + 4 * kq) * lda + c;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int task = tid + 256 * u, kq = task / BN, c = task - kq * BN;
            b_k[u] = c < nb ? 4 * kq : (1 << 28);
            b_off[u] = 2 * T_A_IMG + (c >> 3) * 1024 + (c & 7) * 128 + ((kq ^ (c & 7)) << 4);
            b_ptr[u] = tB + ((size_t)kb0 * T_BK + 4 * kq) * ldb + c;
        }
        struct Loads { float a[NA][4]; float b[NB][4]; };
        auto load_block = [&](Loads& L, int it) {
            const int k0 = (kb0 + it) * T_BK;
#pragma unroll
            for (int u = 0; u < NA; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    L.a[u][i] = (k0 + a_k[u] + i < g.rows) ? __ldg(a_ptr[u] + ((size_t)it * T_BK + i) * lda) : 0.f;
#pragma unroll
            for (int u = 0; u < NB; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    L.b[u][i] = (k0 + b_k[u] + i < g.rows) ? __ldg(b_ptr[u] + ((size_t)it * T_BK + i) * ldb) : 0.f;
        };
        auto split_store = [&](unsigned char* hi, unsigned char* lo, const float (&v)[4]) {
            float4 h, l;
            h.x = t_tf32_hi(v[0]); l.x = v[0] - h.x;
            h.y = t_tf32_hi(v[1]); l.y = v[1] - h.y;
            h.z = t_tf32_hi(v[2]); l.z = v[2] - h.z;
            h.w = t_tf32_hi(v[3]); l.w = v[3] - h.w;
            *reinterpret_cast<float4*>(hi) = h;
            *reinterpret_cast<float4*>(lo) = l;
        };
        auto store_block = [&](const Loads& L, int it) {
            const int s = it & 1;
            t_mbar_wait(BAR(2 + s), ((it >> 1) & 1) ^ 1);
            unsigned char* st = smem + s * STAGE;
#pragma unroll
            for (int u = 0; u < NA; ++u)
                if (a_k[u] < T_BK) split_store(st + a_off[u], st + T_A_IMG + a_off[u], L.a[u]);
#pragma unroll
            for (int u = 0; u < NB; ++u)
                if (b_k[u] < T_BK) split_store(st + b_off[u], st + W_IMG + b_off[u], L.b[u]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            t_mbar_arrive(BAR(s));
        };
        Loads L0, L1;
        if (nkb > 0) load_block(L0, 0);
#pragma unroll 1
        for (int it = 0; it < nkb; it += 2) {
            if (it + 1 < nkb) load_block(L1, it + 1);
            store_block(L0, it);
            if (it + 1 < nkb) {
                if (it + 2 < nkb) load_block(L0, it + 2);
                store_block(L1, it + 1);
            }
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
        // 32-column chunks go through shared memory (the operand stages are free once the last MMA has completed) so that
        // a warp's 32 atomics of one instruction hit 128 consecutive bytes of one row instead of 32 different rows.
        t_mbar_wait(BAR(4), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        float* T = reinterpret_cast<float*>(smem) + warp * 32 * T_EPI_ROW;
        const int mrows = ma - warp * 32;                            // rows of this warp's quarter that exist
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
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
                *reinterpret_cast<float4*>(T + lane * T_EPI_ROW + qq * 4) =
                    make_float4(__uint_as_float(r[qq * 4 + 0]), __uint_as_float(r[qq * 4 + 1]),
                                __uint_as_float(r[qq * 4 + 2]), __uint_as_float(r[qq * 4 + 3]));
            __syncwarp();
            if (c0 + lane < nb) {
                float* cc = tC + (size_t)(warp * 32) * ldc + c0 + lane;
#pragma unroll 4
                for (int row = 0; row < 32; ++row)
                    if (row < mrows) atomicAdd(cc + (size_t)row * ldc, T[row * T_EPI_ROW + lane]);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(256) : "memory");
    }
}

static int tn_splits(int ntiles, int rows, int num_sms) {
    const int nblocks = (rows + T_BK - 1) / T_BK;
    int splits = num_sms / ntiles;                       // about one CTA per SM ...
    const int maxs = (nblocks + 7) / 8;                  // ... but at least 8 row blocks each: a tile's epilogue is up to 32 K atomics
    if (splits > maxs) splits = maxs;
    return splits < 1 ? 1 : splits;
}

// every tile: C (zeroed or holding earlier partial sums) += A^T B over the same `rows`
cudaError_t launch_tn_tiles(const TnTile* tiles, int ntiles, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0 || ntiles <= 0) return cudaSuccess;
    if (ntiles > TN_MAX_TILES) return cudaErrorInvalidValue;
    TnArgs g{};
    for (int i = 0; i < ntiles; ++i) {
        g.tile[i] = tiles[i];
        if (tiles[i].ma < 1 || tiles[i].ma > T_BM || tiles[i].nb < 1 || tiles[i].nb > 256) return cudaErrorInvalidValue;
        g.tile[i].n_mma = tiles[i].nb <= 128 ? 128 : 256;
    }
    g.rows = rows;
    tn_tc_kernel<256><<<dim3(tn_splits(ntiles, rows, num_sms), ntiles), DW_THREADS, dw_smem_bytes(256), s>>>(g);
    return cudaGetLastError();
}

// C[Mreal][Nreal] (row stride ldc, zeroed or holding earlier partial sums) += A^T B
cudaError_t launch_tn_tc(const float* A, int lda, int Mreal, const float* B, int ldb, int Nreal, float* C, int ldc,
                         int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0 || Mreal <= 0 || Nreal <= 0) return cudaSuccess;
    TnTile t[TN_MAX_TILES];
    int n = 0;
    for (int m0 = 0; m0 < Mreal; m0 += T_BM)
        for (int n0 = 0; n0 < Nreal; n0 += 256) {
            if (n == TN_MAX_TILES) return cudaErrorInvalidValue;
            t[n++] = TnTile{A + m0, B + n0, C + (size_t)m0 * ldc + n0, lda, ldb, ldc,
                            Mreal - m0 < T_BM ? Mreal - m0 : T_BM, Nreal - n0 < 256 ? Nreal - n0 : 256, 0};
        }
    return launch_tn_tiles(t, n, rows, num_sms, s);
}

// dW_ih = dgi_d^T in and dW_hh = dgh_prev_d^T out_d, both directions of GRU layer l: 12 (18 for layer 0) tiles, one launch
cudaError_t launch_gru_dw(int l, const float* dgi, const float* in, const float* dghp, const float* out, float* grad_raw,
                          int rows, int num_sms, cudaStream_t s) {
    const int in_w = gru_in(l), in_ld = gru_inp(l);
    TnTile t[TN_MAX_TILES];
    int n = 0;
    for (int dir = 0; dir < 2; ++dir)
        for (int m0 = 0; m0 < G3; m0 += T_BM) {
            for (int n0 = 0; n0 < in_w; n0 += 256) {
                if (n == TN_MAX_TILES) return cudaErrorInvalidValue;
                t[n++] = TnTile{dgi + dir * G3 + m0, in + n0, grad_raw + raw_wih(l, dir) + (size_t)m0 * in_w + n0, GI_N, in_ld, in_w,
                                T_BM, in_w - n0 < 256 ? in_w - n0 : 256, 0};
            }
            if (n == TN_MAX_TILES) return cudaErrorInvalidValue;
            t[n++] = TnTile{dghp + dir * G3 + m0, out + dir * HID, grad_raw + raw_whh(l, dir) + (size_t)m0 * HID, GI_N, OUT_W, HID,
                            T_BM, HID, 0};
        }
    return launch_tn_tiles(t, n, rows, num_sms, s);
}

cudaError_t launch_dw1_tc(const float* dap, const float* ep, float* dW1, int rows, int num_sms, cudaStream_t s) {
    return launch_tn_tc(dap, FC1, FC1, ep, READS, READS, dW1, READS, rows, num_sms, s);
}

// dW1 = dap^T ep with ep rebuilt from (codes, keep bits): one 100 x 200 tile, the row range split over every SM
cudaError_t launch_dw1_gen(const float* dap, EpGen gen, float* dW1, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    TnArgs g{};
    g.tile[0] = TnTile{dap, nullptr, dW1, FC1, READS, READS, FC1, READS, 256};
    g.rows = rows;
    g.gen = gen;
    tn_tc_kernel<256, 1><<<dim3(tn_splits(1, rows, num_sms), 1), DW_THREADS, dw_smem_bytes(256), s>>>(g);
    return cudaGetLastError();
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
    e = cudaFuncSetAttribute(tn_tc_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem_bytes(256));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_stream_kernel<FC1_BN, FC1_KB, 1, READS, READS, FC1, FC1, TEPI_FC1, 1>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(FC1_BN));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_stream_kernel<DEP_BN, DEP_KB, 1, FC1, FC1, READS, READS, TEPI_DE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, t_smem_bytes(DEP_BN));
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
            <<<grid, T_THREADS, t_smem_bytes(DIN_BN), s>>>(dgi, img + din_img_off(0), nullptr, din, rows, ntiles, none, EpGen{});
    } else {
        const int grid = mt < num_sms ? mt : num_sms;
        tc_stream_kernel<DIN_BN, DIN_KB, 1, GI_N, GI_N, OUT_W, OUT_W, TEPI_STORE>
            <<<grid, T_THREADS, t_smem_bytes(DIN_BN), s>>>(dgi, img + din_img_off(l), nullptr, din, rows, mt, none, EpGen{});
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
        <<<grid, T_THREADS, t_smem_bytes(FC1_BN), s>>>(ep, img, b1, a1, rows, ntiles, d, EpGen{});
    return cudaGetLastError();
}

// dep = dap W1:  dap [rows][100] -> dep [rows][200]
cudaError_t launch_dep_tc(const float* dap, const float* img, float* dep, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int ntiles = (rows + T_BM - 1) / T_BM;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    DropCfg none{0ull, 0u, 1.f};
    tc_stream_kernel<DEP_BN, DEP_KB, 1, FC1, FC1, READS, READS, TEPI_STORE>
        <<<grid, T_THREADS, t_smem_bytes(DEP_BN), s>>>(dap, img + IMG_FC1_FLOATS, nullptr, dep, rows, ntiles, none, EpGen{});
    return cudaGetLastError();
}

// a1 = dropout(relu(ep W1^T + b1)) with ep rebuilt from (codes, keep bits)
cudaError_t launch_fc1_gen(EpGen gen, const float* img, const float* b1, float* a1, int rows, DropCfg d, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int ntiles = (rows + T_BM - 1) / T_BM;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    tc_stream_kernel<FC1_BN, FC1_KB, 1, READS, READS, FC1, FC1, TEPI_FC1, 1>
        <<<grid, T_THREADS, t_smem_bytes(FC1_BN), s>>>(nullptr, img, b1, a1, rows, ntiles, d, gen);
    return cudaGetLastError();
}

// dE += scale * sum over kept (row, read) of (dap W1)[row][read], filed by the read's code: d(ep) never leaves the SM
cudaError_t launch_dep_de(const float* dap, const float* img, EpGen gen, int rows, int num_sms, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int ntiles = (rows + T_BM - 1) / T_BM;
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    DropCfg none{0ull, 0u, 1.f};
    tc_stream_kernel<DEP_BN, DEP_KB, 1, FC1, FC1, READS, READS, TEPI_DE>
        <<<grid, T_THREADS, t_smem_bytes(DEP_BN), s>>>(dap, img + IMG_FC1_FLOATS, nullptr, nullptr, rows, ntiles, none, gen);
    return cudaGetLastError();
}

}  // namespace roko
