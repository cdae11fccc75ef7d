// Front end of the roko network (reference roko/rnn_model.py:47-56): embedding gather, read-axis
// fc1 + ReLU, fc2 + ReLU, flatten to the GRU input  u[b][p][10*e + k].
//
// The reference materialises E[x] (3.6 MB / window) and contracts the 200 reads with a dense
// GEMM.  Because the embedding has only 12 rows the contraction factorises exactly
// (SURVEY.md App. B.2):
//     M[p][j][c] = sum_{r : x[r][p] == c} W1[j][r]                 (adds only)
//     a[p][e][j] = relu(b1[j] + sum_c M[p][j][c] * E[c][e])
//     g[p][e][k] = relu(b2[k] + sum_j W2[k][j] * a[p][e][j])
// One CTA owns a window (18 000 contiguous bytes, staged once in shared memory) and walks its
// 90 columns in chunks of CC:
//   phase 1  one warp per column: counting sort of the 200 reads by code -> per-code read lists,
//            each padded to a multiple of 4 with the index of an all-zero W1T row
//   phase 2  thread = (column, 4 consecutive j): sums W1T rows over each list (LDS.128 + 4 FADD
//            per read) -> M in shared memory
//   phase 3  thread = (column, e): a and g entirely in registers; W2/b1/b2 come from the kernel
//            parameter constant bank so every FFMA takes its weight operand for free
#include "common.cuh"

namespace roko {

constexpr int FR_THREADS = 512;
constexpr int CC = 8;             // columns per chunk: phase 3 uses 64 threads per column
constexpr int LIST_LEN = 240;     // 200 reads + up to 3 pads for each of 12 codes
constexpr int JQ = FC1 / 4;

struct FrontSmem {
    float w1t[W1T_ROWS * FC1];                 // 80 400 B  [r][j], row 200 zero
    float ms[CC][FC1][NCODES];                 // 38 400 B
    alignas(16) uint8_t xs[READS * COLS];      // 18 000 B  the window, [read][col]
    alignas(16) uint8_t lists[CC][LIST_LEN];
    int starts[CC][16];
};

__device__ __forceinline__ void sort_column(FrontSmem& S, int cl, int p, int lane, int* status) {
    uint32_t codes[7];
    int cnt[NCODES];
#pragma unroll
    for (int c = 0; c < NCODES; ++c) cnt[c] = 0;
    bool bad = false;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        int r = it * 32 + lane;
        uint32_t code = r < READS ? S.xs[r * COLS + p] : 255u;
        bad |= (r < READS && code >= NCODES);
        codes[it] = code;
#pragma unroll
        for (int c = 0; c < NCODES; ++c) cnt[c] += __popc(__ballot_sync(0xffffffffu, code == (uint32_t)c));
    }
    if (bad) atomicOr(status, 1);              // nn.Embedding would raise IndexError (CPU) / assert (CUDA)
    int base[NCODES], end[NCODES];
    int run = 0;
#pragma unroll
    for (int c = 0; c < NCODES; ++c) {
        base[c] = run;
        if (lane == 0) S.starts[cl][c] = run;
        run += (cnt[c] + 3) & ~3;
        end[c] = run;
    }
    if (lane == 0) S.starts[cl][NCODES] = run;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        uint32_t code = codes[it];
#pragma unroll
        for (int c = 0; c < NCODES; ++c) {
            uint32_t m = __ballot_sync(0xffffffffu, code == (uint32_t)c);
            if (code == (uint32_t)c) S.lists[cl][base[c] + __popc(m & lt)] = (uint8_t)(it * 32 + lane);
            base[c] += __popc(m);
        }
    }
#pragma unroll
    for (int c = 0; c < NCODES; ++c)
        if (lane < end[c] - base[c]) S.lists[cl][base[c] + lane] = (uint8_t)READS;   // zero row
}

__device__ __forceinline__ void build_m(FrontSmem& S, int cl, int jq) {
    const float4* w4 = reinterpret_cast<const float4*>(S.w1t);
    float4 acc[NCODES];
#pragma unroll
    for (int c = 0; c < NCODES; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i0 = S.starts[cl][c], i1 = S.starts[cl][c + 1];
        for (int i = i0; i < i1; i += 4) {
            const uint32_t q = *reinterpret_cast<const uint32_t*>(&S.lists[cl][i]);
            const float4 v0 = w4[(q & 0xffu) * JQ + jq];
            const float4 v1 = w4[((q >> 8) & 0xffu) * JQ + jq];
            const float4 v2 = w4[((q >> 16) & 0xffu) * JQ + jq];
            const float4 v3 = w4[(q >> 24) * JQ + jq];
            a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
            a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
            a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
            a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
        }
        acc[c] = a;
    }
    float4* m = reinterpret_cast<float4*>(&S.ms[cl][4 * jq][0]);
    m[0] = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
    m[1] = make_float4(acc[4].x, acc[5].x, acc[6].x, acc[7].x);
    m[2] = make_float4(acc[8].x, acc[9].x, acc[10].x, acc[11].x);
    m[3] = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
    m[4] = make_float4(acc[4].y, acc[5].y, acc[6].y, acc[7].y);
    m[5] = make_float4(acc[8].y, acc[9].y, acc[10].y, acc[11].y);
    m[6] = make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z);
    m[7] = make_float4(acc[4].z, acc[5].z, acc[6].z, acc[7].z);
    m[8] = make_float4(acc[8].z, acc[9].z, acc[10].z, acc[11].z);
    m[9] = make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w);
    m[10] = make_float4(acc[4].w, acc[5].w, acc[6].w, acc[7].w);
    m[11] = make_float4(acc[8].w, acc[9].w, acc[10].w, acc[11].w);
}

__global__ void __launch_bounds__(FR_THREADS, 1)
front_kernel(const __grid_constant__ FrontConst P, const uint8_t* __restrict__ x,
             const float* __restrict__ packed, float* __restrict__ u, int nwin, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem& S = *reinterpret_cast<FrontSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    {   // W1T (with its zero row) stays resident for every window this CTA processes
        const float4* src = reinterpret_cast<const float4*>(packed + PK_W1T);
        float4* dst = reinterpret_cast<float4*>(S.w1t);
        for (int i = tid; i < W1T_ROWS * FC1 / 4; i += FR_THREADS) dst[i] = src[i];
    }
    const int cl3 = tid >> 6, e = tid & 63;
    float Ee[NCODES];
#pragma unroll
    for (int c = 0; c < NCODES; ++c) Ee[c] = e < EMB ? packed[PK_E + c * EMB + e] : 0.f;

    for (int w = blockIdx.x; w < nwin; w += gridDim.x) {
        __syncthreads();
        {   // the window: 18 000 contiguous bytes, 16-byte vector loads
            const int4* src = reinterpret_cast<const int4*>(x + (size_t)w * (READS * COLS));
            int4* dst = reinterpret_cast<int4*>(S.xs);
            for (int i = tid; i < READS * COLS / 16; i += FR_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        for (int p0 = 0; p0 < COLS; p0 += CC) {
            const int ncol = min(CC, COLS - p0);
            if (warp < ncol) sort_column(S, warp, p0 + warp, lane, status);
            __syncthreads();
            if (tid < ncol * JQ) build_m(S, tid / JQ, tid % JQ);
            __syncthreads();
            if (cl3 < ncol) {
                float* urow = u + ((size_t)w * COLS + p0 + cl3) * IN0P;
                if (e < EMB) {
                    float g[FC2];
#pragma unroll
                    for (int k = 0; k < FC2; ++k) g[k] = P.b2[k];
                    const float4* mrow = reinterpret_cast<const float4*>(&S.ms[cl3][0][0]);
#pragma unroll
                    for (int j = 0; j < FC1; ++j) {
                        const float4 m0 = mrow[j * 3], m1 = mrow[j * 3 + 1], m2 = mrow[j * 3 + 2];
                        float a = P.b1[j];
                        a = fmaf(m0.x, Ee[0], a); a = fmaf(m0.y, Ee[1], a);
                        a = fmaf(m0.z, Ee[2], a); a = fmaf(m0.w, Ee[3], a);
                        a = fmaf(m1.x, Ee[4], a); a = fmaf(m1.y, Ee[5], a);
                        a = fmaf(m1.z, Ee[6], a); a = fmaf(m1.w, Ee[7], a);
                        a = fmaf(m2.x, Ee[8], a); a = fmaf(m2.y, Ee[9], a);
                        a = fmaf(m2.z, Ee[10], a); a = fmaf(m2.w, Ee[11], a);
                        a = fmaxf(a, 0.f);
#pragma unroll
                        for (int k = 0; k < FC2; ++k) g[k] = fmaf(P.W2[k * FC1 + j], a, g[k]);
                    }
                    float2* dst = reinterpret_cast<float2*>(urow + e * FC2);
#pragma unroll
                    for (int k = 0; k < FC2; k += 2)
                        dst[k / 2] = make_float2(fmaxf(g[k], 0.f), fmaxf(g[k + 1], 0.f));
                } else if (e < EMB + (IN0P - IN0) / 4) {          // zero the k-padding of the row
                    reinterpret_cast<float4*>(urow + IN0)[e - EMB] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}

cudaError_t front_setup() {
    return cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FrontSmem));
}

cudaError_t launch_front(const FrontConst& fc, const uint8_t* x, const float* packed, float* u, int nwin,
                         int* status, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    int grid = nwin < num_sms ? nwin : num_sms;
    front_kernel<<<grid, FR_THREADS, sizeof(FrontSmem), s>>>(fc, x, packed, u, nwin, status);
    return cudaGetLastError();
}

}  // namespace roko
