// Probe of tcgen05 cta_group::2 semantics on B200: one CTA pair computes D[256 x 128] = A[256 x 64] . B[128 x 64]^T (fp16 -> fp32).
// CTA r of the pair holds rows 128 r .. 128 r + 127 of A and rows 64 r .. 64 r + 63 of B (half of N) in its own shared memory;
// the leader CTA issues the MMAs; each CTA reads its 128 rows of D from its own tensor memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench/cta2_gemm scripts/ubench/cta2_gemm.cu
//   cta2_gemm <mode>     mode bit 0: tcgen05.alloc.cta_group::2 issued by both CTAs (1) or by the leader only (0)
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../roko_b200/csrc/tc.cuh"

using namespace roko::tc;

constexpr int M = 256, N = 128, K = 64;
// Result on B200: correct in both modes.  The full CTA-pair projection built on it (proj_h2_kernel, git history: the commit before
// this file's last change) was also correct but slower than the single-CTA kernel (0.526 / 0.298 vs 0.422 / 0.257 ms per layer).

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
cta2_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int mode) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                 // 128 rows x 128 B
    unsigned char* sB = smem + 16384;         // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();

    if (tid == 0) { mbar_init(smem_u32(bar), 1); mbar_init_fence(); }
    // operands -> swizzled K-major images (thread = row)
    if (tid < 128) {
        const int row = tid;
        for (int k = 0; k < K; ++k)
            *reinterpret_cast<__half*>(sA + sw128_off(row, k)) = A[(size_t)(rank * 128 + row) * K + k];
        if (row < 64)
            for (int k = 0; k < K; ++k)
                *reinterpret_cast<__half*>(sB + sw128_off(row, k)) = B[(size_t)(rank * 64 + row) * K + k];
    }
    fence_async_smem();
    if (warp == 4 && ((mode & 1) || rank == 0)) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = (mode & 1) || rank == 0 ? *slot : 0u;

    if (warp == 4 && rank == 0) {
        const uint32_t elected = elect_one();
        constexpr uint32_t IDESC = idesc_f16(M, N);
#pragma unroll
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint64_t da = desc_sw128(smem_u32(sA) + kk * 32), db = desc_sw128(smem_u32(sB) + kk * 32);
            asm volatile(
                "{\n\t.reg .pred p, pe;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 pe, %5, 0;\n\t"
                "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(kk ? 1u : 0u), "r"(elected) : "memory");
        }
        if (elected)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
        __syncwarp();
    }
    if (warp < 4) {
        mbar_wait(smem_u32(bar), 0);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            ROKO_TMEM_LD32(r, taddr + c0);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync();
    if (warp == 4 && ((mode & 1) || rank == 0)) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
    }
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { fA[i] = (rand() % 17 - 8) / 8.f; hA[i] = __float2half(fA[i]); }
    for (int i = 0; i < N * K; ++i) { fB[i] = (rand() % 13 - 6) / 4.f; hB[i] = __float2half(fB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    cudaFuncSetAttribute(cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cta2_kernel<<<2, 160, 32768>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
            const double err = fabs(ref - hD[m * N + n]);
            if (!(err <= 1e-3)) { if (bad < 5) printf("  D[%d][%d] = %g, want %g\n", m, n, hD[m * N + n], ref); ++bad; }
            if (err > worst || err != err) worst = err;
        }
    printf("mode %d: max error %g, %d of %d wrong\n", mode, worst, bad, M * N);
    return bad != 0;
}
