// Microbenchmark: legacy warp-level mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16).
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, unsigned seed) {
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    a0 &= 0x3f800000; a1 &= 0x3f800000; a2 &= 0x3f800000; a3 &= 0x3f800000; b0 &= 0x3f800000; b1 &= 0x3f800000;
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 1.2345f) out[threadIdx.x] = s;
}

template <int KIND, int NACC>
void run(float* out, int threads, const char* name) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, NACC><<<148, threads>>>(out, 10, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND, NACC><<<148, threads>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)iters * NACC * (threads / 32) * 148;
    const double macs = mmas * 16 * 8 * (KIND == 0 ? 8 : 16);
    const double cyc_per_sm = ms * 1e-3 * 1.92e9;
    printf("%-22s warps/SM=%2d acc=%d: %.3f ms  %.1f MAC/cycle/SM  %.1f dense TFLOP/s  (%.2f cycles per mma per SM)\n", name,
           threads / 32, NACC, ms, macs / 148 / cyc_per_sm, 2 * macs / (ms * 1e-3) / 1e12, cyc_per_sm / (mmas / 148));
}

int main() {
    float* out;
    cudaMalloc(&out, 4096);
    run<0, 8>(out, 512, "tf32 m16n8k8");
    run<0, 8>(out, 256, "tf32 m16n8k8");
    run<0, 4>(out, 512, "tf32 m16n8k8");
    run<0, 2>(out, 512, "tf32 m16n8k8");
    run<1, 8>(out, 512, "bf16 m16n8k16");
    run<1, 4>(out, 512, "bf16 m16n8k16");
    return 0;
}
