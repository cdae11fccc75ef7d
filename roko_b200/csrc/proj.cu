// GRU input projection (the "gi" half of nn.GRU, reference roko/rnn_model.py:57):
//     C[m][n] = sum_k A[m][k] * W[n][k] + bias[n]      m = window*90 + t,  n = d*384 + j*3 + g
// for both directions at once (N = 768).  A is the front-end output (K = 512, zero padded from
// 500) for layer 0 and the previous layer's [fwd ; bwd] output (K = 256) for layers 1-2.
//
// FP32 FFMA SGEMM: 128x128x8 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double
// buffering through transposed shared tiles (row stride 132 keeps both the transposing stores and
// the LDS.128 fragment reads conflict free).
#include "common.cuh"

namespace roko {

constexpr int BM = 128, BN = 128, BK = 8, PJ_THREADS = 256, TS = 132;

template <int K>
__global__ void __launch_bounds__(PJ_THREADS, 2)
proj_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
            float* __restrict__ C, int M) {
    __shared__ __align__(16) float As[2][BK][TS];
    __shared__ __align__(16) float Bs[2][BK][TS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lrow = tid >> 1, lk = (tid & 1) * 4;
    const bool avalid = (m0 + lrow) < M;
    const float4* Ap = reinterpret_cast<const float4*>(A + (size_t)(avalid ? m0 + lrow : 0) * K + lk);
    const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)(n0 + lrow) * K + lk);

    float4 ra = avalid ? __ldg(Ap) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 rb = __ldg(Wp);
    As[0][lk + 0][lrow] = ra.x; As[0][lk + 1][lrow] = ra.y; As[0][lk + 2][lrow] = ra.z; As[0][lk + 3][lrow] = ra.w;
    Bs[0][lk + 0][lrow] = rb.x; Bs[0][lk + 1][lrow] = rb.y; Bs[0][lk + 2][lrow] = rb.z; Bs[0][lk + 3][lrow] = rb.w;
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    constexpr int KT = K / BK;
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) {
            ra = avalid ? __ldg(Ap + (kt + 1) * 2) : make_float4(0.f, 0.f, 0.f, 0.f);
            rb = __ldg(Wp + (kt + 1) * 2);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            const int nb = buf ^ 1;
            As[nb][lk + 0][lrow] = ra.x; As[nb][lk + 1][lrow] = ra.y; As[nb][lk + 2][lrow] = ra.z; As[nb][lk + 3][lrow] = ra.w;
            Bs[nb][lk + 0][lrow] = rb.x; Bs[nb][lk + 1][lrow] = rb.y; Bs[nb][lk + 2][lrow] = rb.z; Bs[nb][lk + 3][lrow] = rb.w;
        }
        __syncthreads();
    }

    const float4 bia0 = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
    const float4 bia1 = *reinterpret_cast<const float4*>(bias + n0 + 64 + tx * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m < M) {
            float* crow = C + (size_t)m * GI_N + n0;
            *reinterpret_cast<float4*>(crow + tx * 4) =
                make_float4(acc[i][0] + bia0.x, acc[i][1] + bia0.y, acc[i][2] + bia0.z, acc[i][3] + bia0.w);
            *reinterpret_cast<float4*>(crow + 64 + tx * 4) =
                make_float4(acc[i][4] + bia1.x, acc[i][5] + bia1.y, acc[i][6] + bia1.z, acc[i][7] + bia1.w);
        }
    }
}

cudaError_t launch_proj(const float* A, int K, const float* W, const float* bias, float* C, int M,
                        cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    dim3 grid(GI_N / BN, (M + BM - 1) / BM);
    if (K == IN0P) proj_kernel<IN0P><<<grid, PJ_THREADS, 0, s>>>(A, W, bias, C, M);
    else if (K == OUT_W) proj_kernel<OUT_W><<<grid, PJ_THREADS, 0, s>>>(A, W, bias, C, M);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace roko
