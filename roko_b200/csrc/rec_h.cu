// Recurrent half of one bidirectional GRU layer on tcgen05, fp16-split operands (reference
// roko/rnn_model.py:57; gate math SURVEY.md App. B.3).  Successor of rec_tc.cu (3xTF32), same contract.
//
// Per step and direction   D[384 x 32] = W_hh[384 x 128] . H^T[128 x 32]   for the 32 windows a CTA owns,
// evaluated as  W_hi h_hi + W_hi h_lo + W_lo h_hi  in fp16 with fp32 accumulation (tc.cuh).
//
// Measured on B200 (round 2): a tcgen05.mma with N <= 64 costs ~35 cycles whatever its N or operand type, so
// the serial chain of this kernel is bound by the NUMBER of MMAs per step, not by their size.  Hence:
//   * fp16 operands: K = 16 per instruction instead of tf32's 8 halves the count;
//   * the two products that share the A operand W_hi are ONE instruction: the B operand is the 64-row image
//     [h_hi (32 windows) ; h_lo (32 windows)], so  W_hi . [h_hi ; h_lo]  lands in 64 accumulator columns
//     (hi.hi in 0..31, hi.lo in 32..63) for the price of one N = 32 MMA;  W_lo . h_hi  accumulates onto columns 0..31.
//     48 MMAs per step instead of rec_tc's 144;
//   * W_hi (fp16, 3 x [128 x 128]) lives in TENSOR MEMORY as the A operand (192 columns), W_lo in shared memory
//     (96 KB, K-major swizzled, one bulk copy per CTA lifetime); D = 3 gates x 64 columns;
//   * the gate tiles are multiplied in the order r, n, z and committed one by one, so the 512 gate threads run r's sigmoid under
//     the n tile's MMAs and the tanh (which needs r) under the z tile's; only z's sigmoid and the state update follow the last
//     MMA.  (Round-2 trace, scripts/ubench/rec_trace.cu: with r and z first and n last, 1.35 k cycles of gate math trailed the
//     last MMA of every 3.9 k-cycle step.)
// Gate threads: thread = (hidden unit j = TMEM lane, 8 windows).  They read their unit's pre-activations with
// tcgen05.ld, add the two partial columns, keep h in registers, write the layer output (fp32) and the scaled fp16
// hi / lo split of h into the K-major swizzled H image for the next step.  The input projections gi of a step (32 windows x 384
// floats) are staged into shared memory by a loader warp with one TMA bulk copy per window, two steps ahead.
// mbarriers:  h_ready (512 gate threads -> MMA warp),  d_r / d_n / d_z (tcgen05.commit -> gate threads),  w (W_lo landed),
//             gi_full / gi_empty per stage (loader warp <-> gate threads).
#include "common.cuh"
#include "tc.cuh"

namespace roko {

using namespace tc;

#ifdef ROKO_TRACE      // scripts/ubench/rec_trace.cu: clock64 stamps of CTA 0's gate thread 0 and MMA lane 0
__device__ long long roko_trace[4096];
#define RTRACE(slot, s) do { if (blockIdx.x == 0 && (s) >= 20 && (s) < 24) roko_trace[((s) - 20) * 8 + (slot)] = clock64(); } while (0)
#else
#define RTRACE(slot, s) do { } while (0)
#endif

constexpr int RH_N = 32;                         // windows per CTA pass
constexpr int RH_GATE_THREADS = 512;             // 16 warps: TMEM lane quarter = warp & 3, window octet = warp >> 2
constexpr int RH_WPT = RH_N / 4;                 // windows per gate thread (8)
constexpr int RH_THREADS = RH_GATE_THREADS + 64;             // + MMA issuer warp + gi loader warp
constexpr int RH_TMEM_COLS = 512;
constexpr int RH_A_HI = 0, RH_D0 = 3 * (HID / 2);                // columns: W_hi 0..191, D 192..383 (gate tile mt at 192 + 64 mt)
constexpr int RH_WLO_BYTES = G3 * HID * 2;                        // 98 304: [gate tile 3][k atom 2][128 rows][128 B]
constexpr int RH_H_BYTES = 2 * (2 * RH_N) * 128;                  // 16 384: [k atom 2][64 rows: h_hi 0..31, h_lo 32..63][128 B]
constexpr int RH_GI_STAGES = 2;
constexpr int RH_GI_BYTES = RH_N * G3 * 4;                        // 49 152: one step's gi of the CTA's 32 windows, [window][3 j + gate]
constexpr int RH_SMEM = RH_WLO_BYTES + RH_H_BYTES + RH_GI_STAGES * RH_GI_BYTES + 1024 /*align*/ + 128;
static_assert(RH_SMEM <= 232448, "shared memory budget");
constexpr uint32_t RH_ID64 = idesc_f16(128, 2 * RH_N), RH_ID32 = idesc_f16(128, RH_N);
constexpr float RH_INV = 1.f / (W_SCALE * H_SCALE);
static_assert(RH_D0 + 3 * 2 * RH_N <= RH_TMEM_COLS, "tensor memory budget");

__global__ void __launch_bounds__(RH_THREADS, 1)
rec_h_kernel(const float* __restrict__ gi, const float* __restrict__ rh16_d0, float* __restrict__ out, int nwin) {
    extern __shared__ unsigned char rh_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)rh_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* s_wlo = smem;
    unsigned char* s_h = smem + RH_WLO_BYTES;
    unsigned char* s_gi = s_h + RH_H_BYTES;                         // [stage][window][384]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_gi + RH_GI_STAGES * RH_GI_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const uint32_t bar_h = smem_u32(bars), bar_r = bar_h + 8, bar_n = bar_h + 16, bar_z = bar_h + 24, bar_w = bar_h + 32;
    const uint32_t bar_gfull = bar_h + 40, bar_gempty = bar_h + 56;  // [stage]: gi landed (bulk-copy bytes) / gi consumed (512 gate threads)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x & 1;
    const float* wimg = rh16_d0 + (size_t)dir * RH16_DIR;

    if (tid == 0) {
        mbar_init(bar_h, RH_GATE_THREADS);
        mbar_init(bar_r, 1);
        mbar_init(bar_n, 1);
        mbar_init(bar_z, 1);
        mbar_init(bar_w, 1);
        for (int g = 0; g < RH_GI_STAGES; ++g) { mbar_init(bar_gfull + 8 * g, 1); mbar_init(bar_gempty + 8 * g, RH_GATE_THREADS); }
        mbar_init_fence();
    }
    if (warp == RH_GATE_THREADS / 32) tmem_alloc<RH_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- one-time operand residency ------------------------------------------------------------------------
    if (tid == RH_GATE_THREADS) {                                   // W_lo image -> shared memory (bulk copies)
        mbar_expect_tx(bar_w, RH_WLO_BYTES);
        for (int c = 0; c < 3; ++c)
            bulk_g2s(smem_u32(s_wlo) + c * 32768, reinterpret_cast<const unsigned char*>(wimg + RH16_W / 2) + c * 32768, 32768, bar_w);
    }
    if (warp < 4) {                                                 // W_hi -> tensor memory (lane = gate row, 2 fp16 per column)
        const int row = warp * 32 + lane;
#pragma unroll 1
        for (int mt = 0; mt < 3; ++mt) {
            const uint4* src = reinterpret_cast<const uint4*>(wimg + ((size_t)mt * HID + row) * (HID / 2));
#pragma unroll
            for (int c0 = 0; c0 < HID / 2; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 f = __ldg(src + c0 / 4 + q);
                    v[q * 4 + 0] = f.x; v[q * 4 + 1] = f.y; v[q * 4 + 2] = f.z; v[q * 4 + 3] = f.w;
                }
                ROKO_TMEM_ST32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(RH_A_HI + mt * (HID / 2) + c0), v);
            }
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int npass = (nwin + RH_N - 1) / RH_N;

    if (warp < RH_GATE_THREADS / 32) {
        // ================================ gate threads ==============================================
        const int q = warp & 3, oct = warp >> 2;
        const int j = q * 32 + lane;                                // hidden unit == TMEM lane == k index of H
        const float bhn = wimg[RH16_W + j];
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + RH_D0 + oct * RH_WPT;
        // H image element (row, k = j): k atom j >> 6, rows 0..31 hold h_hi of window row, rows 32..63 h_lo
        const uint32_t hs = smem_u32(s_h) + (uint32_t)(j >> 6) * (2 * RH_N * 128);
        uint32_t hoff[RH_WPT];
#pragma unroll
        for (int b = 0; b < RH_WPT; ++b) hoff[b] = sw128_off((uint32_t)(oct * RH_WPT + b), (uint32_t)(j & 63));
        constexpr uint32_t LO_ROWS = (RH_N >> 3) * 1024;            // byte distance from row r to row r + 32 of the image
        const int dt = dir ? -1 : 1;
        uint32_t ph = 0, gstep = 0;                                 // gstep: steps done by this CTA (all passes): stage = gstep % 2
        for (int pass = blockIdx.x >> 1; pass < npass; pass += gridDim.x >> 1) {
            const int t0 = dir ? COLS - 1 : 0;
            // one 32-bit element offset per window: out offset o = (w * 90 + t) * 256 + dir * 128 + j, and the gate-interleaved
            // gi layout makes its offset exactly 3 o  ((w * 90 + t) * 768 + dir * 384 + 3 j)
            unsigned oofs[RH_WPT];
            bool valid[RH_WPT];
            float hprev[RH_WPT];
#pragma unroll
            for (int b = 0; b < RH_WPT; ++b) {
                const int w = pass * RH_N + oct * RH_WPT + b;
                valid[b] = w < nwin;
                const int wl = valid[b] ? w : nwin - 1;             // rows past the batch re-read the last window; never stored
                oofs[b] = (unsigned)(wl * COLS + t0) * OUT_W + dir * HID + j;
                hprev[b] = 0.f;
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + hoff[b]), "h"((unsigned short)0) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + hoff[b] + LO_ROWS), "h"((unsigned short)0) : "memory");
            }
            fence_async_smem();
            mbar_arrive(bar_h);                                     // H = 0 is in place
#pragma unroll 1
            for (int s = 0; s < COLS; ++s) {
                // The gate tiles arrive in the order r, n, z, each committed on its own: r's sigmoid runs under the n tile's MMAs,
                // tanh (which needs r) under the z tile's MMAs, and only z's sigmoid + the state update remain after the last MMA.
                uint32_t a0[RH_WPT], a1[RH_WPT];
                // this step's gi of my windows: staged by the loader warp (bulk copies, two steps ahead), [window][3 j + gate]
                const uint32_t gstage = gstep % RH_GI_STAGES;
                const float* gs = reinterpret_cast<const float*>(s_gi + gstage * RH_GI_BYTES) + (oct * RH_WPT) * G3 + j * 3;
                mbar_wait(bar_gfull + 8 * gstage, (gstep / RH_GI_STAGES) & 1);
                mbar_wait(bar_r, ph);
                tc_fence_after();
                if (tid == 0) RTRACE(0, s);
                ROKO_TMEM_LD8(a0, t_lane);                          // r: hi.hi + lo.hi
                ROKO_TMEM_LD8(a1, t_lane + RH_N);                   // r: hi.lo
                tmem_wait_ld();
                float rr[RH_WPT], nn[RH_WPT];
#pragma unroll
                for (int b = 0; b < RH_WPT; ++b) {
                    const float xr = fmaf(__uint_as_float(a0[b]) + __uint_as_float(a1[b]), RH_INV, gs[b * G3]);
                    rr[b] = rcpf(1.f + ex2f(-1.4426950408889634f * xr));
                }
                if (tid == 0) RTRACE(1, s);
                mbar_wait(bar_n, ph);
                tc_fence_after();
                if (tid == 0) RTRACE(2, s);
                ROKO_TMEM_LD8(a0, t_lane + 4 * RH_N);               // n
                ROKO_TMEM_LD8(a1, t_lane + 5 * RH_N);
                tmem_wait_ld();
#pragma unroll
                for (int b = 0; b < RH_WPT; ++b) {
                    const float xn = fmaf(rr[b], fmaf(__uint_as_float(a0[b]) + __uint_as_float(a1[b]), RH_INV, bhn), gs[b * G3 + 2]);
                    nn[b] = fmaf(2.f, rcpf(1.f + ex2f(-2.8853900817779268f * xn)), -1.f);
                }
                mbar_wait(bar_z, ph); ph ^= 1;
                tc_fence_after();
                ROKO_TMEM_LD8(a0, t_lane + 2 * RH_N);               // z
                ROKO_TMEM_LD8(a1, t_lane + 3 * RH_N);
                tmem_wait_ld();
#pragma unroll
                for (int b = 0; b < RH_WPT; ++b) {
                    const float xz = fmaf(__uint_as_float(a0[b]) + __uint_as_float(a1[b]), RH_INV, gs[b * G3 + 1]);
                    const float z = rcpf(1.f + ex2f(-1.4426950408889634f * xz));
                    const float h = fmaf(z, hprev[b] - nn[b], nn[b]);
                    hprev[b] = h;
                    unsigned short hi, lo;
                    split_f16(h * H_SCALE, hi, lo);
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + hoff[b]), "h"(hi) : "memory");
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + hoff[b] + LO_ROWS), "h"(lo) : "memory");
                    if (valid[b]) out[oofs[b]] = h;
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(bar_h);                                 // h_t is in shared memory, D has been consumed
                if (tid == 0) RTRACE(3, s);
                mbar_arrive(bar_gempty + 8 * gstage);               // this stage of gi may be refilled
                ++gstep;
#pragma unroll
                for (int b = 0; b < RH_WPT; ++b) oofs[b] += dt * OUT_W;
            }
        }
    } else if (warp == RH_GATE_THREADS / 32 + 1) {
        // ================================ gi loader: one bulk copy per window and step ==============
        // gi[w][t][dir * 384 ..] is 1 536 contiguous bytes; lane = window.  Two steps are in flight, so the gate threads never wait
        // on HBM (with the one-step register prefetch this replaced, `long_scoreboard` was the top stall at full load).
        uint32_t gstep = 0;
        for (int pass = blockIdx.x >> 1; pass < npass; pass += gridDim.x >> 1) {
            const int w = pass * RH_N + lane;
            const int wl = w < nwin ? w : nwin - 1;                 // windows past the batch re-read the last one (never stored)
            const int dt = dir ? -1 : 1;
            const float* src = gi + ((size_t)wl * COLS + (dir ? COLS - 1 : 0)) * GI_N + dir * G3;
            for (int s = 0; s < COLS; ++s, ++gstep, src += dt * GI_N) {
                const uint32_t st = gstep % RH_GI_STAGES;
                mbar_wait(bar_gempty + 8 * st, ((gstep / RH_GI_STAGES) & 1) ^ 1);
                if (lane == 0) mbar_expect_tx(bar_gfull + 8 * st, RH_GI_BYTES);
                __syncwarp();
                bulk_g2s(smem_u32(s_gi) + st * RH_GI_BYTES + lane * (G3 * 4), src, G3 * 4, bar_gfull + 8 * st);
            }
        }
    } else {
        // ================================ MMA issuer (whole warp, uniform) ==========================
        if (tmem != 0) __trap();                                    // all 512 columns are ours -> base 0: literal TMEM addresses
        mbar_wait(bar_w, 0);                                        // W_lo image has landed
        const uint32_t a_lo = smem_u32(s_wlo), b_img = smem_u32(s_h);
        const uint32_t elected = elect_one();
        uint32_t ph_h = 0;
        for (int pass = blockIdx.x >> 1; pass < npass; pass += gridDim.x >> 1) {
#pragma unroll 1
            for (int s = 0; s <= COLS; ++s) {
                mbar_wait(bar_h, ph_h); ph_h ^= 1;
                if (s == COLS) break;                               // the last arrival only closes the pass
                tc_fence_after();
                if (lane == 0) RTRACE(4, s);
#pragma unroll
                for (int part = 0; part < 3; ++part) {              // gate tiles in the order r, n, z, each committed on its own
                    const int mt = part == 0 ? 0 : (part == 1 ? 2 : 1);
                    const uint32_t d = RH_D0 + mt * (2 * RH_N);
#pragma unroll
                    for (int kk = 0; kk < HID / 16; ++kk) {
                        const uint64_t db = desc_sw128(b_img + (uint32_t)(kk >> 2) * (2 * RH_N * 128) + (uint32_t)(kk & 3) * 32);
                        const uint64_t da = desc_sw128(a_lo + (uint32_t)mt * 32768 + (uint32_t)(kk >> 2) * 16384 + (uint32_t)(kk & 3) * 32);
                        mma_f16_ts(d, RH_A_HI + mt * (HID / 2) + kk * 8, db, RH_ID64, kk ? 1u : 0u, elected);   // W_hi . [h_hi ; h_lo]
                        mma_f16_ss(d, da, db, RH_ID32, 1u, elected);                                            // W_lo . h_hi  -> columns 0..31
                    }
                    mma_commit(part == 0 ? bar_r : (part == 1 ? bar_n : bar_z), elected);
                }
                if (lane == 0) RTRACE(5, s);
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == RH_GATE_THREADS / 32) {
        tc_fence_after();
        tmem_dealloc<RH_TMEM_COLS>(tmem);
    }
}


cudaError_t rec_h_setup() {
    return cudaFuncSetAttribute(rec_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RH_SMEM);
}

cudaError_t launch_rec_h(const float* gi, const float* rh16_d0, float* out, int nwin, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int npass = (nwin + RH_N - 1) / RH_N;
    const int pairs = num_sms / 2;
    const int grid = 2 * (npass < pairs ? npass : pairs);
    rec_h_kernel<<<grid, RH_THREADS, RH_SMEM, s>>>(gi, rh16_d0, out, nwin);
    return cudaGetLastError();
}

}  // namespace roko
