// Small inline-PTX vocabulary shared by the fp16-split tcgen05 kernels (rec_h.cu, proj_h.cu, front_tc.cu):
// mbarriers, bulk copies, UMMA shared-memory descriptors, tcgen05.mma / ld / st / commit, TMEM allocation.
//
// Precision scheme of those kernels ("3xFP16"): an fp32 product  x . w  is evaluated as
//     x_lo w_hi + x_hi w_lo + x_hi w_hi,      x_hi = fp16(x),  x_lo = fp16(x - x_hi)      (same for w)
// with fp32 accumulation in tensor memory.  fp16 carries 11 significant bits like tf32, so the result
// has the accuracy of the 3xTF32 split (about 2^-22 relative per product), but kind::f16 MMAs run at twice
// the tf32 rate and move half the operand bytes.  fp16's narrow exponent is handled by power-of-two
// operand scales (weights x 256, activations x 256 or x 16) that the epilogues undo exactly.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace roko {
namespace tc {

constexpr float W_SCALE = 256.f;        // every fp16-split weight image holds 256 w
constexpr float H_SCALE = 256.f;        // hidden states (|h| <= 1)
constexpr float U_SCALE = 16.f;         // front-end output (ReLU, unbounded above; |u| < 4094 representable)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// The wait carries a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of spinning.  Measured on front_tc.cu: without it 18 % of all issued instructions were the polling
// loops of waiting warps (BRA / SYNCS.TRYWAIT / YIELD), taken from the issue slots of the warps doing the work.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 128 bytes (64 fp16), SWIZZLE_128B: 8-row groups of
// 1024 B, 16-byte chunk c of row r stored at chunk position c ^ (r & 7).
// descriptor: start >> 4 [0,14) | LBO (unused for swizzled K-major) [16,30) | SBO = 1024 >> 4 [32,46) |
// version 1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of fp16 element (row, k) inside one [rows x 64] K-major SWIZZLE_128B image
__host__ __device__ constexpr uint32_t sw128_off(uint32_t row, uint32_t k) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + (k & 7u) * 2u;
}

// instruction descriptor, kind::f16: D = fp32 [4,6) = 1, A = B = fp16 (format 0), both K-major, N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// All MMA helpers are meant for warp-uniform code: every lane computes the same operands (they stay in
// uniform registers) and only the instruction itself is predicated on the elected lane.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(elected) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(elected) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar, uint32_t elected) {
    if (elected)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int COLS_>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS_) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS_>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS_) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define ROKO_TMEM_LD4(r, taddr)                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"          \
                 : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]) : "r"(taddr))
#define ROKO_TMEM_LD8(r, taddr)                                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"            \
                 : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]),       \
                   "=r"((r)[6]), "=r"((r)[7]) : "r"(taddr))
#define ROKO_TMEM_LD16(r, taddr)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                            \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
                 : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]),   \
                   "=r"((r)[7]), "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]),               \
                   "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15]) : "r"(taddr))
#define ROKO_TMEM_LD32(r, taddr)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]),      \
                   "=r"((r)[7]), "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]),   \
                   "=r"((r)[14]), "=r"((r)[15]), "=r"((r)[16]), "=r"((r)[17]), "=r"((r)[18]), "=r"((r)[19]),               \
                   "=r"((r)[20]), "=r"((r)[21]), "=r"((r)[22]), "=r"((r)[23]), "=r"((r)[24]), "=r"((r)[25]),               \
                   "=r"((r)[26]), "=r"((r)[27]), "=r"((r)[28]), "=r"((r)[29]), "=r"((r)[30]), "=r"((r)[31])               \
                 : "r"(taddr))
#define ROKO_TMEM_ST32(taddr, v)                                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                          \
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                               \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                       \
                 ::"r"(taddr), "r"((v)[0]), "r"((v)[1]), "r"((v)[2]), "r"((v)[3]), "r"((v)[4]), "r"((v)[5]), "r"((v)[6]),  \
                   "r"((v)[7]), "r"((v)[8]), "r"((v)[9]), "r"((v)[10]), "r"((v)[11]), "r"((v)[12]), "r"((v)[13]),           \
                   "r"((v)[14]), "r"((v)[15]), "r"((v)[16]), "r"((v)[17]), "r"((v)[18]), "r"((v)[19]), "r"((v)[20]),        \
                   "r"((v)[21]), "r"((v)[22]), "r"((v)[23]), "r"((v)[24]), "r"((v)[25]), "r"((v)[26]), "r"((v)[27]),        \
                   "r"((v)[28]), "r"((v)[29]), "r"((v)[30]), "r"((v)[31]) : "memory")
#define ROKO_TMEM_ST8(taddr, v)                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"    \
                 ::"r"(taddr), "r"((v)[0]), "r"((v)[1]), "r"((v)[2]), "r"((v)[3]), "r"((v)[4]),      \
                   "r"((v)[5]), "r"((v)[6]), "r"((v)[7]) : "memory")

// fp16 hi / lo split of a (pre-scaled) fp32 value; returns the two halves as raw 16-bit patterns
__device__ __forceinline__ void split_f16(float v, unsigned short& hi, unsigned short& lo) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}
// two values -> packed hi word (v0 in the low half) and packed lo word: one packed convert, two unpacks, two
// subtractions, one packed convert (6 instructions for two values)
__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// relu(v0), relu(v1) -> packed hi / lo words in 6 instructions: the hi halves are converted with round-toward-zero and
// the ReLU clamp in one instruction, so that v - hi >= 0 wherever v > 0 and v - hi = v < 0 wherever v < 0; the second
// ReLU-clamped conversion then yields the lo halves for positive v and 0 for negative v.  (Truncation instead of
// rounding costs one bit: the pair carries 21 instead of 22 significant bits.)
__device__ __forceinline__ void relu_split_f16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - f.y), "f"(v0 - f.x));
}

__device__ __forceinline__ float ex2f(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rcpf(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

}  // namespace tc
}  // namespace roko
