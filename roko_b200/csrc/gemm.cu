// General fp32 GEMM of the training path (reference roko/train.py:46-53 runs these products through
// autograd): C[M][N] (op)= sum_k A(m,k) * B(k,n), any M, N, K, every operand a strided view.
//
//   KA = true : A(m,k) = A[m*lda + k]   (k contiguous)      KA = false: A(m,k) = A[k*lda + m]
//   KB = true : B(k,n) = B[n*ldb + k]   (k contiguous)      KB = false: B(k,n) = B[k*ldb + n]
//
// FP32 FFMA, 128x128x8 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double
// buffering (same engine as proj.cu).  gridDim.z splits K; split launches add with atomics into a
// zeroed C.  Epilogues: store, accumulate, atomic add, and fc1's bias + ReLU + dropout.
#include "train.cuh"

namespace roko {

constexpr int GM = 128, GN = 128, GK = 8, G_THREADS = 256, GTS = 132;

template <bool KCONTIG>
__device__ __forceinline__ void tile_load(const float* __restrict__ P, int ld, int row0, int nrows, int k0,
                                          int kend, bool vec, int tid, float (&r)[4]) {
    r[0] = r[1] = r[2] = r[3] = 0.f;
    if (KCONTIG) {
        const int row = row0 + (tid >> 1), k = k0 + (tid & 1) * 4;
        if (row < nrows && k < kend) {
            const float* p = P + (size_t)row * ld + k;
            if (vec && k + 3 < kend) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p));
                r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (k + i < kend) r[i] = __ldg(p + i);
            }
        }
    } else {
        const int k = k0 + (tid >> 5), row = row0 + (tid & 31) * 4;
        if (k < kend && row < nrows) {
            const float* p = P + (size_t)k * ld + row;
            if (vec && row + 3 < nrows) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p));
                r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (row + i < nrows) r[i] = __ldg(p + i);
            }
        }
    }
}

template <bool KCONTIG>
__device__ __forceinline__ void tile_store(float (*T)[GTS], int tid, const float (&r)[4]) {
    if (KCONTIG) {
        const int row = tid >> 1, k = (tid & 1) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) T[k + i][row] = r[i];
    } else {
        *reinterpret_cast<float4*>(&T[tid >> 5][(tid & 31) * 4]) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

template <bool KA, bool KB, int EPI>
__global__ void __launch_bounds__(G_THREADS, 2) sgemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][GK][GTS];
    __shared__ __align__(16) float Bs[2][GK][GTS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = (kbeg + g.kchunk) < g.K ? (kbeg + g.kchunk) : g.K;

    float ra[4], rb[4];
    tile_load<KA>(g.A, g.lda, m0, g.M, kbeg, kend, g.vecA, tid, ra);
    tile_load<KB>(g.B, g.ldb, n0, g.N, kbeg, kend, g.vecB, tid, rb);
    tile_store<KA>(As[0], tid, ra);
    tile_store<KB>(Bs[0], tid, rb);
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int KT = (kend - kbeg + GK - 1) / GK;
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) {
            tile_load<KA>(g.A, g.lda, m0, g.M, kbeg + (kt + 1) * GK, kend, g.vecA, tid, ra);
            tile_load<KB>(g.B, g.ldb, n0, g.N, kbeg + (kt + 1) * GK, kend, g.vecB, tid, rb);
        }
#pragma unroll
        for (int k = 0; k < GK; ++k) {
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
            tile_store<KA>(As[buf ^ 1], tid, ra);
            tile_store<KB>(Bs[buf ^ 1], tid, rb);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            float* c = g.C + (size_t)m * g.ldc + n;
            float v = acc[i][j];
            if (EPI == EPI_STORE) {
                *c = v;
            } else if (EPI == EPI_ACC) {
                *c += v;
            } else if (EPI == EPI_ATOMIC) {
                atomicAdd(c, v);
            } else {   // EPI_FC1: relu(v + b1[n]) then dropout site 1, element index m*N + n
                v = fmaxf(v + __ldg(g.bias + n), 0.f);
                v = drop_keep(g.drop, DROP_FC1, (unsigned long long)m * g.N + n) ? v * g.drop.scale : 0.f;
                *c = v;
            }
        }
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <bool KA, bool KB>
static cudaError_t launch_kab(const GemmArgs& g, int epi, dim3 grid, cudaStream_t s) {
    switch (epi) {
        case EPI_STORE: sgemm_kernel<KA, KB, EPI_STORE><<<grid, G_THREADS, 0, s>>>(g); break;
        case EPI_ACC: sgemm_kernel<KA, KB, EPI_ACC><<<grid, G_THREADS, 0, s>>>(g); break;
        case EPI_ATOMIC: sgemm_kernel<KA, KB, EPI_ATOMIC><<<grid, G_THREADS, 0, s>>>(g); break;
        case EPI_FC1: sgemm_kernel<KA, KB, EPI_FC1><<<grid, G_THREADS, 0, s>>>(g); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// splits > 1 requires EPI_ATOMIC (C zeroed by the caller); splits <= 0 picks enough K chunks to
// put about two waves of CTAs on the machine.
cudaError_t launch_gemm(GemmArgs g, bool ka, bool kb, int epi, int splits, int num_sms, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
    const int tiles = ((g.M + GM - 1) / GM) * ((g.N + GN - 1) / GN);
    if (splits <= 0) {
        splits = epi == EPI_ATOMIC ? (4 * num_sms + tiles - 1) / tiles : 1;
        const int maxs = (g.K + 255) / 256;              // at least 256 of K per CTA
        if (splits > maxs) splits = maxs;
        if (splits < 1) splits = 1;
    }
    if (splits > 1 && epi != EPI_ATOMIC) return cudaErrorInvalidValue;
    int kchunk = (g.K + splits - 1) / splits;
    kchunk = (kchunk + GK - 1) / GK * GK;
    splits = (g.K + kchunk - 1) / kchunk;
    g.kchunk = kchunk;
    g.vecA = (g.lda % 4 == 0) && aligned16(g.A);
    g.vecB = (g.ldb % 4 == 0) && aligned16(g.B);
    dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM, splits);
    if (ka && kb) return launch_kab<true, true>(g, epi, grid, s);
    if (ka && !kb) return launch_kab<true, false>(g, epi, grid, s);
    if (!ka && kb) return launch_kab<false, true>(g, epi, grid, s);
    return launch_kab<false, false>(g, epi, grid, s);
}

}  // namespace roko
