// Microbenchmark: what FFMA rate can the recurrent kernel's inner loop reach on sm_100a?
// 512 threads, 96 weights in registers, h broadcast from shared (LDS.128), NB windows.
//   mode 0: scalar FFMA   acc += w * h      (2 fresh register reads + 1 reused per FFMA)
//   mode 1: packed FFMA2  fma.rn.f32x2      (pairs of k)
//   mode 2: scalar FFMA, 1 fresh register operand (acc = acc * a + b)   -> nominal peak
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int NB, int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ wg, float* __restrict__ out, int steps) {
    __shared__ __align__(16) float hs[2][NB][136];
    const int tid = threadIdx.x, kq = tid & 3;
    for (int i = tid; i < 2 * NB * 136; i += 512) (&hs[0][0][0])[i] = 0.001f * (i & 15);
    __syncthreads();
    float total = 0.f;
    if (MODE == 0) {
        float w[96];
#pragma unroll
        for (int i = 0; i < 96; ++i) w[i] = wg[i * 512 + tid];
        for (int s = 0; s < steps; ++s) {
            const int cur = s & 1;
            float acc[NB][3];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float4 hv = *reinterpret_cast<const float4*>(&hs[cur][b][16 * i + 4 * kq]);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 0], hv.x, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 1], hv.y, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 2], hv.z, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 3], hv.w, acc[b][g]);
                    }
                }
#pragma unroll
            for (int b = 0; b < NB; ++b) total += acc[b][0] + acc[b][1] + acc[b][2];
            if (total == 1.2345f) hs[cur ^ 1][0][tid & 127] = total;   // keeps the loop alive, never taken
        }
    } else if (MODE == 1) {
        unsigned long long w2[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) w2[i] = pack2(wg[(2 * i) * 512 + tid], wg[(2 * i + 1) * 512 + tid]);
        for (int s = 0; s < steps; ++s) {
            const int cur = s & 1;
            unsigned long long acc[NB][3];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(&hs[cur][b][16 * i + 4 * kq]);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[b][g] = ffma2(w2[g * 16 + i * 2 + 0], hv.x, acc[b][g]);
                        acc[b][g] = ffma2(w2[g * 16 + i * 2 + 1], hv.y, acc[b][g]);
                    }
                }
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int g = 0; g < 3; ++g) { float x, y; unpack2(acc[b][g], x, y); total += x + y; }
            if (total == 1.2345f) hs[cur ^ 1][0][tid & 127] = total;
        }
    } else {
        float acc[16];
        const float a = wg[tid], bb = wg[tid + 512];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = (float)(tid + i);
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int r = 0; r < 6 * NB; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, bb);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) total += acc[i];
    }
    if (total == 1.2345f) out[tid] = total;
}

template <int NB, int MODE>
int run(const float* w, float* out, const char* name) {
    const int steps = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NB, MODE><<<148, 512>>>(w, out, 100);
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<NB, MODE><<<148, 512>>>(w, out, steps);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double us_step = ms * 1e3 / steps;
    const double fma_per_step_thread = 96.0 * NB;
    const double tf = 2.0 * fma_per_step_thread * 512 * 148 * steps / (ms * 1e-3) / 1e12;
    printf("%-28s NB=%d  %.3f us/step  ~%.0f cyc/step @1.92GHz  (ideal %d)  %.1f TFLOP/s\n", name, NB, us_step,
           us_step * 1920.0, 96 * NB * 4, tf);
    return 0;
}

int main() {
    float *w, *out;
    CHECK(cudaMalloc(&w, 96 * 512 * 4));
    CHECK(cudaMalloc(&out, 512 * 4));
    CHECK(cudaMemset(w, 0, 96 * 512 * 4));
    run<1, 0>(w, out, "scalar FFMA (rec loop)");
    run<2, 0>(w, out, "scalar FFMA (rec loop)");
    run<4, 0>(w, out, "scalar FFMA (rec loop)");
    run<1, 1>(w, out, "packed FFMA2 (rec loop)");
    run<2, 1>(w, out, "packed FFMA2 (rec loop)");
    run<4, 1>(w, out, "packed FFMA2 (rec loop)");
    run<2, 2>(w, out, "acc=acc*a+b (1 fresh operand)");
    return 0;
}
