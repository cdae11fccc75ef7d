// General fp32 GEMM of the training path (reference roko/train.py:46-53 runs these products through
// autograd): C[M][N] (op)= sum_k A(m,k) * B(k,n), any M, N, K, every operand a strided view.
//
//   KA = true : A(m,k) = A[m*lda + k]   (k contiguous)      KA = false: A(m,k) = A[k*lda + m]
//   KB = true : B(k,n) = B[n*ldb + k]   (k contiguous)      KB = false: B(k,n) = B[k*ldb + n]
//
// 3xTF32 on mma.sync.m16n8k8 (fp32-level accuracy, see split_tf32), 128x128x8 tiles, 256 threads,
// register-prefetch double buffering through k-major shared tiles.  gridDim.z splits K; split
// launches add with atomics into a zeroed C.  Epilogues: store, accumulate, atomic add, and fc1's
// bias + ReLU + dropout.
#include <stdlib.h>

#include "train.cuh"

namespace roko {

constexpr int GM = 128, GN = 128, GK = 8, G_THREADS = 256;
constexpr int GTS = 136;     // row stride = 8 (mod 32) floats: the mma fragment reads (k = lane&3, row = lane>>2) hit 32 banks

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

// tf32 split: x = hi + lo with hi = rna(x) on 10 mantissa bits; the tensor core reads the top 19 bits of lo
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 3xTF32 products of two m-tiles x two n-tiles, term by term: consecutive MMAs on one accumulator are four
// issues apart, so the tensor pipe never waits on the accumulate dependency (small terms first).
__device__ __forceinline__ void mma3_block(float (&acc)[2][8][4], int jp, const uint32_t (&ah)[2][4],
                                           const uint32_t (&al)[2][4], const uint32_t (&bh)[2][2],
                                           const uint32_t (&bl)[2][2]) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int i = 0; i < 2; ++i) mma_tf32(acc[i][jp * 2 + jj], al[i], bh[jj]);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int i = 0; i < 2; ++i) mma_tf32(acc[i][jp * 2 + jj], ah[i], bl[jj]);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int i = 0; i < 2; ++i) mma_tf32(acc[i][jp * 2 + jj], ah[i], bh[jj]);
}

// accumulator fragment (i, j, c): row mrow + 16 i + 8 (c >> 1), column ncol + 8 j + (c & 1)
template <int EPI>
__device__ __forceinline__ void gemm_epilogue(const GemmArgs& g, const float (&acc)[2][8][4], int mrow, int ncol) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = mrow + i * 16 + h * 8;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = ncol + j * 8 + c;
                    if (n >= g.N) continue;
                    float* cp = g.C + (size_t)m * g.ldc + n;
                    float v = acc[i][j][h * 2 + c];
                    if (EPI == EPI_STORE) {
                        *cp = v;
                    } else if (EPI == EPI_ACC) {
                        *cp += v;
                    } else if (EPI == EPI_ATOMIC) {
                        atomicAdd(cp, v);
                    } else {   // EPI_FC1: relu(v + b1[n]) then dropout site 1, element index m*N + n
                        v = fmaxf(v + __ldg(g.bias + n), 0.f);
                        v = drop_keep(g.drop, DROP_FC1, (unsigned long long)m * g.N + n) ? v * g.drop.scale : 0.f;
                        *cp = v;
                    }
                }
            }
        }
    }
}

// 8 warps as 4 (m) x 2 (n): a warp owns 32 x 64 of the 128 x 128 tile = 2 x 8 m16n8k8 accumulators.
// Every product is 3xTF32 (lo*hi + hi*lo + hi*hi, small terms first): fp32-level accuracy on the tensor pipe.
template <bool KA, bool KB, int EPI>
__global__ void __launch_bounds__(G_THREADS, 2) sgemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][GK][GTS];
    __shared__ __align__(16) float Bs[2][GK][GTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = (kbeg + g.kchunk) < g.K ? (kbeg + g.kchunk) : g.K;

    float ra[4], rb[4];
    tile_load<KA>(g.A, g.lda, m0, g.M, kbeg, kend, g.vecA, tid, ra);
    tile_load<KB>(g.B, g.ldb, n0, g.N, kbeg, kend, g.vecB, tid, rb);
    tile_store<KA>(As[0], tid, ra);
    tile_store<KB>(Bs[0], tid, rb);
    __syncthreads();

    float acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

    const int KT = (kend - kbeg + GK - 1) / GK;
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) {
            tile_load<KA>(g.A, g.lda, m0, g.M, kbeg + (kt + 1) * GK, kend, g.vecA, tid, ra);
            tile_load<KB>(g.B, g.ldb, n0, g.N, kbeg + (kt + 1) * GK, kend, g.vecB, tid, rb);
        }
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = wm + i * 16 + gq;
            split_tf32(As[buf][tq][m], ah[i][0], al[i][0]);
            split_tf32(As[buf][tq][m + 8], ah[i][1], al[i][1]);
            split_tf32(As[buf][tq + 4][m], ah[i][2], al[i][2]);
            split_tf32(As[buf][tq + 4][m + 8], ah[i][3], al[i][3]);
        }
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            uint32_t bh[2][2], bl[2][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int n = wn + (jp * 2 + jj) * 8 + gq;
                split_tf32(Bs[buf][tq][n], bh[jj][0], bl[jj][0]);
                split_tf32(Bs[buf][tq + 4][n], bh[jj][1], bl[jj][1]);
            }
            mma3_block(acc, jp, ah, al, bh, bl);
        }
        if (kt + 1 < KT) {
            tile_store<KA>(As[buf ^ 1], tid, ra);
            tile_store<KB>(Bs[buf ^ 1], tid, rb);
        }
        __syncthreads();
    }

    gemm_epilogue<EPI>(g, acc, m0 + wm + gq, n0 + wn + tq * 2);
}

// ---- streamed variant: cp.async, 32-deep k tiles, 3 stages ------------------------------------------
// Same warp tiling and 3xTF32 arithmetic; operand tiles keep the operand's own contiguity so 16-byte
// asynchronous copies can fill them: a k-contiguous operand lands as [128 rows][36] (fragment reads hit
// banks 4*row + k), a row-contiguous one as [32 k][136] (banks 8*k + row).  About 100 KB of loads are in
// flight per CTA, which is what the 0.5 GB front-end products need to approach the HBM rate.
// Requirements (checked by the launcher, otherwise the kernel above runs): 16-byte aligned operands with
// leading dimensions and K that are multiples of 4, and a multiple-of-4 extent along a row-contiguous axis.
constexpr int FK = 32, F_STAGES = 3;
constexpr int F_KC_LD = 36, F_RC_LD = 136;
template <bool KC> __host__ __device__ constexpr int f_tile_floats() { return KC ? 128 * F_KC_LD : FK * F_RC_LD; }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <bool KC>
__device__ __forceinline__ void f_tile_fill(float* tile, const float* __restrict__ P, int ld, int row0, int nrows,
                                            int k0, int kend, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tid + i * G_THREADS;                  // 1024 chunks of 16 bytes
        if (KC) {
            const int row = c >> 3, k = k0 + (c & 7) * 4;
            const bool ok = (row0 + row) < nrows && k < kend;
            cp_async16(tile + row * F_KC_LD + (c & 7) * 4, ok ? P + (size_t)(row0 + row) * ld + k : P, ok ? 16 : 0);
        } else {
            const int kk = c >> 5, row = row0 + (c & 31) * 4;
            const bool ok = (k0 + kk) < kend && row < nrows;
            cp_async16(tile + kk * F_RC_LD + (c & 31) * 4, ok ? P + (size_t)(k0 + kk) * ld + row : P, ok ? 16 : 0);
        }
    }
}

template <bool KC>
__device__ __forceinline__ float f_at(const float* tile, int row, int k) {
    return KC ? tile[row * F_KC_LD + k] : tile[k * F_RC_LD + row];
}

template <bool KA, bool KB, int EPI>
__global__ void __launch_bounds__(G_THREADS, 2) sgemm_stream_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) float dyn[];
    constexpr int TA = f_tile_floats<KA>(), TB = f_tile_floats<KB>();
    float* At = dyn;
    float* Bt = dyn + F_STAGES * TA;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = (kbeg + g.kchunk) < g.K ? (kbeg + g.kchunk) : g.K;
    const int KT = (kend - kbeg + FK - 1) / FK;

#pragma unroll
    for (int st = 0; st < F_STAGES - 1; ++st) {
        if (st < KT) {
            f_tile_fill<KA>(At + st * TA, g.A, g.lda, m0, g.M, kbeg + st * FK, kend, tid);
            f_tile_fill<KB>(Bt + st * TB, g.B, g.ldb, n0, g.N, kbeg + st * FK, kend, tid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

    for (int kt = 0; kt < KT; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(F_STAGES - 2) : "memory");
        __syncthreads();                                    // tile kt landed for everyone; tile kt-1's buffer is free
        {
            const int nx = kt + F_STAGES - 1;
            if (nx < KT) {
                f_tile_fill<KA>(At + (nx % F_STAGES) * TA, g.A, g.lda, m0, g.M, kbeg + nx * FK, kend, tid);
                f_tile_fill<KB>(Bt + (nx % F_STAGES) * TB, g.B, g.ldb, n0, g.N, kbeg + nx * FK, kend, tid);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const float* a = At + (kt % F_STAGES) * TA;
        const float* b = Bt + (kt % F_STAGES) * TB;
#pragma unroll
        for (int ks = 0; ks < FK / 8; ++ks) {
            const int kb = ks * 8;
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm + i * 16 + gq;
                split_tf32(f_at<KA>(a, m, kb + tq), ah[i][0], al[i][0]);
                split_tf32(f_at<KA>(a, m + 8, kb + tq), ah[i][1], al[i][1]);
                split_tf32(f_at<KA>(a, m, kb + tq + 4), ah[i][2], al[i][2]);
                split_tf32(f_at<KA>(a, m + 8, kb + tq + 4), ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bh[2][2], bl[2][2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int n = wn + (jp * 2 + jj) * 8 + gq;
                    split_tf32(f_at<KB>(b, n, kb + tq), bh[jj][0], bl[jj][0]);
                    split_tf32(f_at<KB>(b, n, kb + tq + 4), bh[jj][1], bl[jj][1]);
                }
                mma3_block(acc, jp, ah, al, bh, bl);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    gemm_epilogue<EPI>(g, acc, m0 + wm + gq, n0 + wn + tq * 2);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <bool KA, bool KB, int EPI>
constexpr int stream_smem_bytes() { return F_STAGES * (f_tile_floats<KA>() + f_tile_floats<KB>()) * (int)sizeof(float); }

template <bool KA, bool KB, int EPI>
static cudaError_t launch_stream(const GemmArgs& g, dim3 grid, cudaStream_t s) {
    sgemm_stream_kernel<KA, KB, EPI><<<grid, G_THREADS, stream_smem_bytes<KA, KB, EPI>(), s>>>(g);
    return cudaGetLastError();
}

template <bool KA, bool KB>
static cudaError_t setup_kab() {
    cudaError_t e = cudaFuncSetAttribute(sgemm_stream_kernel<KA, KB, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         stream_smem_bytes<KA, KB, EPI_STORE>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgemm_stream_kernel<KA, KB, EPI_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   stream_smem_bytes<KA, KB, EPI_ACC>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgemm_stream_kernel<KA, KB, EPI_ATOMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   stream_smem_bytes<KA, KB, EPI_ATOMIC>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sgemm_stream_kernel<KA, KB, EPI_FC1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   stream_smem_bytes<KA, KB, EPI_FC1>());
    return e;
}

// opt the streamed kernels into > 48 KB of dynamic shared memory on the CURRENT device (called per model_create)
cudaError_t gemm_setup() {
    cudaError_t e = setup_kab<true, true>();
    if (e == cudaSuccess) e = setup_kab<true, false>();
    if (e == cudaSuccess) e = setup_kab<false, true>();
    if (e == cudaSuccess) e = setup_kab<false, false>();
    return e;
}

template <bool KA, bool KB>
static cudaError_t launch_kab(const GemmArgs& g, int epi, dim3 grid, cudaStream_t s, bool stream_ok) {
    if (stream_ok) {
        switch (epi) {
            case EPI_STORE: return launch_stream<KA, KB, EPI_STORE>(g, grid, s);
            case EPI_ACC: return launch_stream<KA, KB, EPI_ACC>(g, grid, s);
            case EPI_ATOMIC: return launch_stream<KA, KB, EPI_ATOMIC>(g, grid, s);
            case EPI_FC1: return launch_stream<KA, KB, EPI_FC1>(g, grid, s);
            default: return cudaErrorInvalidValue;
        }
    }
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
    kchunk = (kchunk + FK - 1) / FK * FK;                // multiple of both engines' k tiles
    splits = (g.K + kchunk - 1) / kchunk;
    g.kchunk = kchunk;
    g.vecA = (g.lda % 4 == 0) && aligned16(g.A);
    g.vecB = (g.ldb % 4 == 0) && aligned16(g.B);
    static const bool no_stream = getenv("ROKO_B200_GEMM_NOSTREAM") != nullptr;      // A/B switch
    const bool stream_ok = !no_stream && g.vecA && g.vecB && g.K % 4 == 0 && g.K >= 64 && (ka || g.M % 4 == 0)
                           && (kb || g.N % 4 == 0);
    dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM, splits);
    if (ka && kb) return launch_kab<true, true>(g, epi, grid, s, stream_ok);
    if (ka && !kb) return launch_kab<true, false>(g, epi, grid, s, stream_ok);
    if (!ka && kb) return launch_kab<false, true>(g, epi, grid, s, stream_ok);
    return launch_kab<false, false>(g, epi, grid, s, stream_ok);
}

}  // namespace roko
