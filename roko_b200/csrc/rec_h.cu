// Recurrent half of one bidirectional GRU layer on tcgen05, fp16-split operands (reference
// roko/rnn_model.py:57; gate math SURVEY.md App. B.3).  Successor of rec_tc.cu (3xTF32), same contract.
//
// Per step and direction   D[384 x N] = W_hh[384 x 128] . H^T[128 x N]   for the N windows of a group,
// evaluated as  W_lo h_hi + W_hi h_lo + W_hi h_hi  in fp16 with fp32 accumulation (tc.cuh).  What the
// fp16 split buys over the tf32 one, and how the CTA is organised around it:
//   * an fp16 MMA covers K = 16 per instruction (tf32: 8) at the same cost, so a step needs 72 MMAs
//     instead of 144;
//   * W_hi AND W_lo (fp16: 96 KB each) both fit in TENSOR MEMORY as A operands (384 of the 512 columns);
//     the tf32 kernel had to read its 192 KB W_lo image from shared memory on every step, which made a
//     third of its MMAs shared-memory-bandwidth bound (41 vs 19 cycles);
//   * a CTA owns 32 windows of one direction as TWO half-groups of 16 (UMMA N = 16), each with its own
//     accumulator (3 gates x 16 columns) and its own H images in shared memory.  The half-groups run
//     one step out of phase: while the 512 gate threads apply sigmoid/tanh to half-group A (MUFU-bound:
//     5 MUFU per unit and window), the tensor core multiplies half-group B.  The serial chain per step
//     is therefore max(gates, MMA) instead of gates + MMA.
// Gate threads: thread = (hidden unit j = TMEM lane, 4 windows).  They read their unit's r/z/n
// pre-activations with tcgen05.ld, keep h in registers, write the layer output (fp32) and the scaled
// fp16 hi/lo split of h into the K-major swizzled H images for the next step.
// mbarriers per half-group g:  h_ready[g] (512 gate threads -> MMA warp),  d_ready[g] (tcgen05.commit -> gate threads).
#include "common.cuh"
#include "tc.cuh"

namespace roko {

using namespace tc;

constexpr int RH_NG = 16;                        // windows per half-group (UMMA N)
constexpr int RH_CTA_WIN = 2 * RH_NG;            // windows per CTA pass
constexpr int RH_GATE_THREADS = 512;             // 16 warps: TMEM lane quarter = warp & 3, window quad = warp >> 2
constexpr int RH_WPT = RH_NG / 4;                // windows per gate thread and half-group
constexpr int RH_THREADS = RH_GATE_THREADS + 32;
constexpr int RH_TMEM_COLS = 512;
constexpr int RH_A_HI = 0, RH_A_LO = 3 * (HID / 2), RH_D0 = 6 * (HID / 2);   // columns: 0, 192, 384
constexpr int RH_H_IMG = RH_NG * HID * 2;        // 4 096 B: one hi (or lo) image of a half-group, [k atom 2][16 rows][128 B]
constexpr int RH_SMEM = 2 * 2 * RH_H_IMG + 1024 /*align*/ + 64;
constexpr uint32_t RH_IDESC = idesc_f16(128, RH_NG);
constexpr float RH_INV = 1.f / (W_SCALE * H_SCALE);
static_assert(RH_D0 + 2 * 3 * RH_NG <= RH_TMEM_COLS, "tensor memory budget");

__global__ void __launch_bounds__(RH_THREADS, 1)
rec_h_kernel(const float* __restrict__ gi, const float* __restrict__ rh16_d0, float* __restrict__ out, int nwin) {
    extern __shared__ unsigned char rh_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)rh_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* s_h = smem;                                     // [half-group][hi|lo] images
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * RH_H_IMG);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t bar0 = smem_u32(bars);
    // bar_h[g] = bar0 + 8 g,  bar_d[g] = bar0 + 16 + 8 g

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x & 1;
    const float* wimg = rh16_d0 + (size_t)dir * RH16_DIR;

    if (tid == 0) {
        mbar_init(bar0, RH_GATE_THREADS);
        mbar_init(bar0 + 8, RH_GATE_THREADS);
        mbar_init(bar0 + 16, 1);
        mbar_init(bar0 + 24, 1);
        mbar_init_fence();
    }
    if (warp == RH_GATE_THREADS / 32) tmem_alloc<RH_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- one-time operand residency: W_hi, W_lo -> tensor memory (lane = gate row, 2 fp16 per column) ----
    if (warp < 4) {
        const int row = warp * 32 + lane;
#pragma unroll 1
        for (int img = 0; img < 6; ++img) {                        // img = gate tile * 2 + (hi | lo)
            const int mt = img >> 1, lo = img & 1;
            const uint4* src = reinterpret_cast<const uint4*>(wimg + ((size_t)img * HID + row) * (HID / 2));
#pragma unroll
            for (int c0 = 0; c0 < HID / 2; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 f = __ldg(src + c0 / 4 + q);
                    v[q * 4 + 0] = f.x; v[q * 4 + 1] = f.y; v[q * 4 + 2] = f.z; v[q * 4 + 3] = f.w;
                }
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((lo ? RH_A_LO : RH_A_HI) + mt * (HID / 2) + c0);
                ROKO_TMEM_ST32(taddr, v);
            }
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int npass = (nwin + RH_CTA_WIN - 1) / RH_CTA_WIN;

    if (warp < RH_GATE_THREADS / 32) {
        // ================================ gate threads ==============================================
        const int q = warp & 3, sub = warp >> 2;
        const int j = q * 32 + lane;                                // hidden unit == TMEM lane == k index of H
        const float bhn = wimg[RH16_W + j];
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + RH_D0 + sub * RH_WPT;
        // H image element (row = window in half-group, k = j): k atom j >> 6, 16-byte chunk (j & 63) >> 3, swizzled by row & 7
        uint32_t hoff[RH_WPT];
#pragma unroll
        for (int b = 0; b < RH_WPT; ++b) {
            const uint32_t row = (uint32_t)(sub * RH_WPT + b);
            hoff[b] = (uint32_t)(j >> 6) * (RH_NG * 128) + sw128_off(row, (uint32_t)(j & 63));
        }
        const uint32_t hs = smem_u32(s_h);
        const int dt = dir ? -1 : 1;
        uint32_t ph_d0 = 0, ph_d1 = 0;
        for (int pass = blockIdx.x >> 1; pass < npass; pass += gridDim.x >> 1) {
            const int t0 = dir ? COLS - 1 : 0;
            // one 32-bit element offset per window: out offset o = (w * 90 + t) * 256 + dir * 128 + j, and the gate-interleaved
            // gi layout makes its offset exactly 3 o  ((w * 90 + t) * 768 + dir * 384 + 3 j)
            unsigned oofs[2][RH_WPT];
            bool valid[2][RH_WPT];
            float hprev[2][RH_WPT], g_r[2][RH_WPT], g_z[2][RH_WPT], g_n[2][RH_WPT];
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int b = 0; b < RH_WPT; ++b) {
                    const int w = pass * RH_CTA_WIN + g * RH_NG + sub * RH_WPT + b;
                    valid[g][b] = w < nwin;
                    const int wl = valid[g][b] ? w : nwin - 1;      // rows past the batch re-read the last window; never stored
                    oofs[g][b] = (unsigned)(wl * COLS + t0) * OUT_W + dir * HID + j;
                    hprev[g][b] = 0.f;
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + (2 * g) * RH_H_IMG + hoff[b]), "h"((unsigned short)0) : "memory");
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + (2 * g + 1) * RH_H_IMG + hoff[b]), "h"((unsigned short)0) : "memory");
                    const float* gp = gi + 3u * oofs[g][b];
                    g_r[g][b] = __ldg(gp); g_z[g][b] = __ldg(gp + 1); g_n[g][b] = __ldg(gp + 2);
                }
            fence_async_smem();
            mbar_arrive(bar0);                                      // H = 0 is in place for both half-groups
            mbar_arrive(bar0 + 8);
#pragma unroll 1
            for (int s = 0; s < COLS; ++s) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (g == 0) { mbar_wait(bar0 + 16, ph_d0); ph_d0 ^= 1; }
                    else        { mbar_wait(bar0 + 24, ph_d1); ph_d1 ^= 1; }
                    tc_fence_after();
                    uint32_t dr[RH_WPT], dz[RH_WPT], dn[RH_WPT];
                    const uint32_t ta = t_lane + g * (3 * RH_NG);
                    ROKO_TMEM_LD4(dr, ta);
                    ROKO_TMEM_LD4(dz, ta + RH_NG);
                    ROKO_TMEM_LD4(dn, ta + 2 * RH_NG);
                    tmem_wait_ld();
#pragma unroll
                    for (int b = 0; b < RH_WPT; ++b) {
                        // r and z share one reciprocal: 1 / ((1 + ea)(1 + eb)); inputs clamped so the product stays finite
                        const float xr = fmaxf(fmaf(__uint_as_float(dr[b]), RH_INV, g_r[g][b]), -40.f);
                        const float xz = fmaxf(fmaf(__uint_as_float(dz[b]), RH_INV, g_z[g][b]), -40.f);
                        const float ea = 1.f + ex2f(-1.4426950408889634f * xr);
                        const float eb = 1.f + ex2f(-1.4426950408889634f * xz);
                        const float rc = rcpf(ea * eb);
                        const float r = rc * eb, z = rc * ea;
                        const float xn = fmaf(r, fmaf(__uint_as_float(dn[b]), RH_INV, bhn), g_n[g][b]);
                        const float n = fmaf(2.f, rcpf(1.f + ex2f(-2.8853900817779268f * xn)), -1.f);
                        const float h = fmaf(z, hprev[g][b] - n, n);
                        hprev[g][b] = h;
                        unsigned short hi, lo;
                        split_f16(h * H_SCALE, hi, lo);
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + (2 * g) * RH_H_IMG + hoff[b]), "h"(hi) : "memory");
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(hs + (2 * g + 1) * RH_H_IMG + hoff[b]), "h"(lo) : "memory");
                        if (valid[g][b]) out[oofs[g][b]] = h;
                    }
                    tc_fence_before();
                    fence_async_smem();
                    mbar_arrive(bar0 + 8 * g);                      // h_t is in shared memory, D has been consumed
                    if (s + 1 < COLS) {                             // lands while the tensor core runs this half-group's next step
#pragma unroll
                        for (int b = 0; b < RH_WPT; ++b) {
                            oofs[g][b] += dt * OUT_W;
                            const float* gp = gi + 3u * oofs[g][b];
                            g_r[g][b] = __ldg(gp); g_z[g][b] = __ldg(gp + 1); g_n[g][b] = __ldg(gp + 2);
                        }
                    }
                }
            }
        }
    } else {
        // ================================ MMA issuer (whole warp, uniform) ==========================
        if (tmem != 0) __trap();                                    // all 512 columns are ours -> base 0: literal TMEM addresses
        const uint32_t b_base = smem_u32(s_h);
        const uint32_t elected = elect_one();
        uint32_t ph_h0 = 0, ph_h1 = 0;
        for (int pass = blockIdx.x >> 1; pass < npass; pass += gridDim.x >> 1) {
#pragma unroll 1
            for (int s = 0; s <= COLS; ++s) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (g == 0) { mbar_wait(bar0, ph_h0); ph_h0 ^= 1; }
                    else        { mbar_wait(bar0 + 8, ph_h1); ph_h1 ^= 1; }
                    if (s == COLS) continue;                        // the last arrival only closes the pass
                    tc_fence_after();
                    const uint32_t b_hi = b_base + (2 * g) * RH_H_IMG, b_lo = b_hi + RH_H_IMG;
#pragma unroll
                    for (int mt = 0; mt < 3; ++mt) {
                        const uint32_t d = RH_D0 + g * (3 * RH_NG) + mt * RH_NG;
#pragma unroll
                        for (int kk = 0; kk < HID / 16; ++kk) {
                            const uint32_t koff = (uint32_t)(kk >> 2) * (RH_NG * 128) + (uint32_t)(kk & 3) * 32;
                            const uint64_t dbh = desc_sw128(b_hi + koff), dbl = desc_sw128(b_lo + koff);
                            const uint32_t a_hi = RH_A_HI + mt * (HID / 2) + kk * 8, a_lo = RH_A_LO + mt * (HID / 2) + kk * 8;
                            mma_f16_ts(d, a_lo, dbh, RH_IDESC, kk ? 1u : 0u, elected);   // W_lo h_hi   (small terms first)
                            mma_f16_ts(d, a_hi, dbl, RH_IDESC, 1u, elected);             // W_hi h_lo
                            mma_f16_ts(d, a_hi, dbh, RH_IDESC, 1u, elected);             // W_hi h_hi
                        }
                    }
                    mma_commit(bar0 + 16 + 8 * g, elected);
                    __syncwarp();
                }
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
    const int npass = (nwin + RH_CTA_WIN - 1) / RH_CTA_WIN;
    const int pairs = num_sms / 2;
    const int grid = 2 * (npass < pairs ? npass : pairs);
    rec_h_kernel<<<grid, RH_THREADS, RH_SMEM, s>>>(gi, rh16_d0, out, nwin);
    return cudaGetLastError();
}

}  // namespace roko
