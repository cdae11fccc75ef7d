// Recurrent half of one bidirectional GRU layer on the tcgen05 tensor cores
// (reference roko/rnn_model.py:57; gate math SURVEY.md App. B.3).
//
// Per step and direction the recurrence needs  D[384 x N] = W_hh[384 x 128] . H^T[128 x N]  for the N
// windows a CTA owns.  fp32 accuracy comes from the same 3xTF32 split as the projection:
//     D = W_hi h_hi + W_hi h_lo + W_lo h_hi                      (fp32 accumulation in TMEM)
// The point of the design is where the operands live for the CTA's whole lifetime:
//     W_hi  (tf32, 3 x [128 x 128])  in TENSOR MEMORY as the MMA's A operand       384 columns
//     W_lo  (tf32, 192 KB)           in SHARED MEMORY (K-major, 128B swizzle), loaded once by TMA
//     D     ([128 x 32] x 3 gates)   in tensor memory                               96 columns
//     H     (h_hi, h_lo: [32 windows x 128] each, K-major swizzled)  in shared memory, rewritten
//           every step by the gate threads
// so a step is 144 tcgen05.mma (M=128, N=32, K=8) issued by one thread, then 256 gate threads read
// their unit's r/z/n pre-activations with tcgen05.ld (TMEM lane == hidden unit), apply the gates with
// MUFU sigmoid/tanh, keep h in registers, write the layer output and the tf32-split h back to shared
// memory for the next step.  Two mbarriers alternate: h_ready (gate threads -> MMA thread) and
// d_ready (tcgen05.commit -> gate threads).
#include <stdlib.h>

#include "common.cuh"

namespace roko {

constexpr int RT_N = 32;                        // windows per CTA group (UMMA N)
constexpr int RT_GATE_THREADS = 512;            // 16 warps: TMEM lane quarter = warp % 4, window octet = warp / 4
constexpr int RT_WPT = RT_N / 4;                // windows per gate thread (8)
constexpr int RT_THREADS = RT_GATE_THREADS + 32;
constexpr int RT_WLO_BYTES = G3 * HID * 4;      // 196 608
constexpr int RT_H_BYTES = RT_N * HID * 4;      // 16 384 per hi / lo image
constexpr int RT_SMEM_BYTES = RT_WLO_BYTES + 2 * RT_H_BYTES + 1024 /*align*/ + 64 /*barriers*/;
constexpr int RT_TMEM_COLS = 512;
constexpr int RT_D_COL = 3 * HID;               // D starts after the three W_hi tiles
constexpr uint32_t RT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(RT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t rt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void rt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t rt_desc(uint32_t saddr) {   // K-major SWIZZLE_128B, SBO = 1024 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// The MMA warp runs warp-uniform code (all 32 lanes compute the same descriptors, so they live in
// uniform registers); only the instruction itself is predicated on the elected lane.
__device__ __forceinline__ void rt_mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t acc, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(adesc), "l"(bdesc), "r"(RT_IDESC), "r"(acc), "r"(elected) : "memory");
}
__device__ __forceinline__ void rt_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t acc, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(RT_IDESC), "r"(acc), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t rt_elect() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ float rt_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rt_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rt_sigmoid(float v) { return rt_rcp(1.f + rt_ex2(-1.4426950408889634f * v)); }
__device__ __forceinline__ float rt_tanh(float v) { return fmaf(2.f, rt_rcp(1.f + rt_ex2(-2.8853900817779268f * v)), -1.f); }
__device__ __forceinline__ float rt_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

#define RT_TMEM_LD8(r, taddr)                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), \
                   "=r"(r[7])                                                                         \
                 : "r"(taddr))
__device__ __forceinline__ void rt_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// whi: [dir][384][128] tf32-rounded W_hh, row major.  wlo: [dir] shared-memory image of W_hh - whi.
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_kernel(const float* __restrict__ gi, const float* __restrict__ whi0, const float* __restrict__ wlo0,
              size_t dir_stride, const float* __restrict__ bhn0, float* __restrict__ out, int nwin) {
    extern __shared__ unsigned char rt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)rt_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* s_wlo = smem;                                 // 192 KB  [mt][katom][128 rows][128 B]
    unsigned char* s_hhi = smem + RT_WLO_BYTES;                  // 16 KB   [katom][32 rows][128 B]
    unsigned char* s_hlo = s_hhi + RT_H_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_hlo + RT_H_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const uint32_t bar_w = rt_smem_u32(bars), bar_h = bar_w + 8, bar_d = bar_w + 16, bar_d2 = bar_w + 24;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.x & 1;
    const float* whi = whi0 + dir * dir_stride;
    const float* wlo = wlo0 + dir * dir_stride;

    if (tid == 0) {
        rt_mbar_init(bar_w, 1);
        rt_mbar_init(bar_h, RT_GATE_THREADS);
        rt_mbar_init(bar_d, 1);
        rt_mbar_init(bar_d2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == RT_GATE_THREADS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(tmem_slot)), "n"(RT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // ---- one-time operand residency ---------------------------------------------------------------
    if (tid == RT_GATE_THREADS) {                                 // W_lo image -> shared memory (TMA)
        rt_mbar_expect_tx(bar_w, RT_WLO_BYTES);
        for (int c = 0; c < 3; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(rt_smem_u32(s_wlo) + c * 65536), "l"(reinterpret_cast<const unsigned char*>(wlo) + c * 65536),
                           "r"(65536), "r"(bar_w) : "memory");
    }
    if (warp < 4) {                                               // W_hi -> tensor memory (lane = gate row)
        const int row = warp * 32 + lane;
        for (int mt = 0; mt < 3; ++mt) {
            const float4* src = reinterpret_cast<const float4*>(whi + (size_t)(mt * HID + row) * HID);
            for (int c0 = 0; c0 < HID; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 f = __ldg(src + c0 / 4 + q);
                    v[q * 4 + 0] = __float_as_uint(f.x); v[q * 4 + 1] = __float_as_uint(f.y);
                    v[q * 4 + 2] = __float_as_uint(f.z); v[q * 4 + 3] = __float_as_uint(f.w);
                }
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * HID + c0);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                    ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                    : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int ngroups = (nwin + RT_N - 1) / RT_N;
    uint32_t ph_h = 0, ph_d = 0;                                  // mbarrier phase parities

    if (warp < RT_GATE_THREADS / 32) {
        // ================================ gate threads ==============================================
        // thread = (hidden unit j = TMEM lane, octet of windows).  All per-window addresses are
        // compile-time offsets from three per-thread bases.  gi loads of rows past the batch are not
        // guarded: the launcher only uses this kernel when the scratch buffer behind gi is large enough
        // for them to stay in bounds, and those windows' results are never stored.
        const int q = warp & 3, oct = warp >> 2;
        const int j = q * 32 + lane;                              // hidden unit == TMEM lane
        const float bhn = (bhn0 + dir * dir_stride)[j];
        const uint32_t t_lane = ((uint32_t)(q * 32) << 16) + RT_D_COL + oct * RT_WPT;
        // H images: k-atom q, row = oct*8 + b (so row & 7 == b, row >> 3 == oct), 16-byte chunk lane>>2
        const uint32_t hs_base = (uint32_t)q * (RT_N * 128) + (uint32_t)oct * 1024 + (uint32_t)(lane & 3) * 4;
        const uint32_t hhi_s = rt_smem_u32(s_hhi) + hs_base, hlo_s = rt_smem_u32(s_hlo) + hs_base;
        const uint32_t lq = (uint32_t)(lane >> 2);
        const int dt = dir ? -1 : 1;
        for (int grp = blockIdx.x >> 1; grp < ngroups; grp += gridDim.x >> 1) {
            const int w0 = grp * RT_N + oct * RT_WPT;
            const int nvalid = nwin - w0;                         // windows b < nvalid exist
            int t = dir ? COLS - 1 : 0;
            const float* gp = gi + ((size_t)w0 * COLS + t) * GI_N + dir * G3 + j * 3;
            float* op = out + ((size_t)w0 * COLS + t) * OUT_W + dir * HID + j;
            float hprev[RT_WPT], g_r[RT_WPT], g_z[RT_WPT], g_n[RT_WPT];
#pragma unroll
            for (int b = 0; b < RT_WPT; ++b) {
                hprev[b] = 0.f;
                const uint32_t off = (uint32_t)b * 128 + ((lq ^ (uint32_t)b) << 4);
                rt_sts(hhi_s + off, 0.f);
                rt_sts(hlo_s + off, 0.f);
                g_r[b] = __ldg(gp + b * (COLS * GI_N)); g_z[b] = __ldg(gp + b * (COLS * GI_N) + 1);
                g_n[b] = __ldg(gp + b * (COLS * GI_N) + 2);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            rt_mbar_arrive(bar_h);                                // H = 0 is in place
            for (int s = 0; s < COLS; ++s) {
                rt_mbar_wait(bar_d, ph_d); ph_d ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t dr[RT_WPT], dz[RT_WPT], dn[RT_WPT];
                RT_TMEM_LD8(dr, t_lane);
                RT_TMEM_LD8(dz, t_lane + RT_N);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float rr[RT_WPT], zz[RT_WPT];                     // r, z while the n-gate MMAs are still running
#pragma unroll
                for (int b = 0; b < RT_WPT; ++b) {
                    rr[b] = rt_sigmoid(g_r[b] + __uint_as_float(dr[b]));
                    zz[b] = rt_sigmoid(g_z[b] + __uint_as_float(dz[b]));
                }
                rt_mbar_wait(bar_d2, ph_d ^ 1);                   // same phase sequence as bar_d
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                RT_TMEM_LD8(dn, t_lane + 2 * RT_N);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int b = 0; b < RT_WPT; ++b) {
                    const float r = rr[b], z = zz[b];
                    const float n = rt_tanh(g_n[b] + r * (__uint_as_float(dn[b]) + bhn));
                    const float h = fmaf(z, hprev[b] - n, n);
                    hprev[b] = h;
                    const uint32_t off = (uint32_t)b * 128 + ((lq ^ (uint32_t)b) << 4);
                    const float hi = rt_tf32(h);
                    rt_sts(hhi_s + off, hi);
                    rt_sts(hlo_s + off, h - hi);
                    if (b < nvalid) op[b * (COLS * OUT_W)] = h;
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rt_mbar_arrive(bar_h);                            // h_t is in shared memory, D has been consumed
                gp += dt * GI_N; op += dt * OUT_W;
                if (s + 1 < COLS) {                               // lands while the tensor core runs step s+1
#pragma unroll
                    for (int b = 0; b < RT_WPT; ++b) {
                        g_r[b] = __ldg(gp + b * (COLS * GI_N)); g_z[b] = __ldg(gp + b * (COLS * GI_N) + 1);
                        g_n[b] = __ldg(gp + b * (COLS * GI_N) + 2);
                    }
                }
            }
        }
    } else {
        // ================================ MMA issuer (whole warp, uniform) ==========================
        // All 512 TMEM columns are ours, so the allocation starts at column 0 / lane 0: using literal
        // TMEM addresses keeps every MMA operand in the uniform datapath.
        if (tmem != 0) __trap();
        rt_mbar_wait(bar_w, 0);                                   // W_lo image has landed
        const uint32_t a_lo = rt_smem_u32(s_wlo), b_hi = rt_smem_u32(s_hhi), b_lo = rt_smem_u32(s_hlo);
        const uint32_t elected = rt_elect();
        for (int grp = blockIdx.x >> 1; grp < ngroups; grp += gridDim.x >> 1) {
            for (int s = 0; s < COLS + 1; ++s) {
                rt_mbar_wait(bar_h, ph_h); ph_h ^= 1;
                if (s == COLS) break;                             // the last arrival only closes the group
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // r and z tiles first and committed on their own, so the gate threads can start the two
                // sigmoids while the n tile is still being multiplied
#pragma unroll
                for (int part = 0; part < 2; ++part) {
#pragma unroll
                    for (int kk = 0; kk < HID / 8; ++kk) {
                        const uint32_t koff = (uint32_t)(kk >> 2) * (RT_N * 128) + (uint32_t)(kk & 3) * 32;
                        const uint64_t dbh = rt_desc(b_hi + koff), dbl = rt_desc(b_lo + koff);
#pragma unroll
                        for (int mt = part ? 2 : 0; mt < (part ? 3 : 2); ++mt) {
                            const uint32_t d = RT_D_COL + mt * RT_N;
                            const uint64_t dal = rt_desc(a_lo + mt * 65536 + (kk >> 2) * 16384 + (kk & 3) * 32);
                            const uint32_t a_hi = (uint32_t)(mt * HID + kk * 8);
                            rt_mma_ss(d, dal, dbh, kk ? 1u : 0u, elected);   // W_lo h_hi   (small terms first)
                            rt_mma_ts(d, a_hi, dbl, 1u, elected);            // W_hi h_lo
                            rt_mma_ts(d, a_hi, dbh, 1u, elected);            // W_hi h_hi
                        }
                    }
                    if (elected)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     ::"r"(part ? bar_d2 : bar_d) : "memory");
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == RT_GATE_THREADS / 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(RT_TMEM_COLS) : "memory");
    }
}

cudaError_t rec_tc_setup() {
    return cudaFuncSetAttribute(rec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM_BYTES);
}

cudaError_t launch_rec_tc(const float* gi, const float* whi_d0, const float* wlo_d0, size_t dir_stride,
                          const float* bhn_d0, float* out, int nwin, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int ngroups = (nwin + RT_N - 1) / RT_N;
    const int pairs = num_sms / 2;
    const int grid = 2 * (ngroups < pairs ? ngroups : pairs);
    rec_tc_kernel<<<grid, RT_THREADS, RT_SMEM_BYTES, s>>>(gi, whi_d0, wlo_d0, dir_stride, bhn_d0, out, nwin);
    return cudaGetLastError();
}

}  // namespace roko
