// Front end of the roko network (reference roko/rnn_model.py:47-56): embedding gather, read-axis
// fc1 + ReLU, fc2 + ReLU, flatten to the GRU input  u[b][p][10*e + k].
//
// The reference materialises E[x] (3.6 MB / window) and contracts the 200 reads with a dense
// GEMM.  Because the embedding has only 12 rows the contraction factorises exactly
// (SURVEY.md App. B.2):
//     M[p][j][c] = sum_{r : x[r][p] == c} W1[j][r]                 (adds only)
//     a[p][e][j] = relu(b1[j] + sum_c M[p][j][c] * E[c][e])
//     g[p][e][k] = relu(b2[k] + sum_j W2[k][j] * a[p][e][j])
//
// One persistent CTA per SM walks windows.  A window (18 000 contiguous bytes) arrives by a single
// TMA bulk copy, double buffered so the next window lands while this one computes.  Inside a
// window every WARP owns whole columns and runs them end to end with warp-level sync only:
//   sort      counting sort of the column's 200 reads by code (packed-byte histograms + shuffle
//             scan) -> per-code read lists, padded to a multiple of 4 with the index of an all-zero
//             W1T row
//   gather    lane = 4 consecutive j (25 lanes): sum W1T rows over each list (LDS.128 + 4 FADD per
//             read) -> the column's M (100 x 12) in the warp's private shared-memory slab
//   a, g      lane = two embedding dims (e, e+25): a and g entirely in registers; W2/b1/b2 are
//             constant-bank operands (kernel parameters), M rows are warp-broadcast LDS.128
// The 16 warps of a CTA drift apart, so the latency-bound gather of some warps overlaps the
// FFMA-bound a/g stage of others without any block-level barrier (an earlier producer/consumer
// split on named barriers serialised: 0.10 + 0.10 = 0.21 ms per 128 windows).
#include "common.cuh"

namespace roko {

constexpr int FR_THREADS = 512;
constexpr int FR_WARPS = FR_THREADS / 32;
constexpr int LIST_LEN = 240;                 // 200 reads + up to 3 pads for each of 12 codes
constexpr int JQ = FC1 / 4;                   // 25 gather lanes
constexpr int EH = EMB / 2;                   // 25 a/g lanes, two embedding dims each

struct FrontWarp {                            // private to one warp
    float m[FC1][NCODES];                     // 4 800 B
    alignas(16) uint8_t list[LIST_LEN];       //   240 B
    uint8_t start[16];
};

struct FrontSmem {
    float w1t[W1T_ROWS * FC1];                // 80 400 B  [r][j], row 200 zero
    FrontWarp wp[FR_WARPS];                   // 80 896 B
    alignas(16) float w2t[FC1][12];           //  4 800 B  [j] -> W2[0..9][j], b1[j], 0
    alignas(16) uint8_t xs[2][READS * COLS];  // 36 000 B  double-buffered window, [read][col]
    alignas(8) unsigned long long xbar[2];    // TMA arrival barriers of the two window buffers
};

__device__ __forceinline__ uint32_t fr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Counting sort of one column's 200 reads by code, one warp per column.  Lane l owns the 7
// consecutive reads 7l..7l+6, keeps a per-code histogram packed as bytes in three 32-bit words
// (4 codes each; a warp total never exceeds 200 < 256), and a 5-step shuffle scan turns the
// histograms into write positions.  Reads stay in ascending order inside every code's list.
__device__ __forceinline__ uint32_t byte_of(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t code) {
    const uint32_t w = code < 4 ? w0 : (code < 8 ? w1 : w2);
    return (w >> ((code & 3u) * 8u)) & 0xffu;
}

__device__ __forceinline__ void sort_column(FrontWarp& W, const uint8_t* xs, int p, int lane, int* status) {
    uint32_t codes[7];
    uint32_t h0 = 0, h1 = 0, h2 = 0;               // my histogram
    uint32_t rank[7];                              // rank of read i among my earlier reads of the same code
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int r = lane * 7 + i;
        uint32_t code = r < READS ? xs[r * COLS + p] : 15u;
        bad |= (r < READS && code >= NCODES);
        if (code >= NCODES) code = 15u;            // 15 -> no list
        codes[i] = code;
        rank[i] = byte_of(h0, h1, h2, code);
        const uint32_t inc = 1u << ((code & 3u) * 8u);
        h0 += code < 4 ? inc : 0u;
        h1 += (code >= 4 && code < 8) ? inc : 0u;
        h2 += (code >= 8 && code < 12) ? inc : 0u;
    }
    if (bad) atomicOr(status, 1);                  // nn.Embedding would raise IndexError (CPU) / assert (CUDA)
    uint32_t s0 = h0, s1 = h1, s2 = h2;            // inclusive scan over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o),
                       t2 = __shfl_up_sync(0xffffffffu, s2, o);
        if (lane >= o) { s0 += t0; s1 += t1; s2 += t2; }
    }
    const uint32_t tot0 = __shfl_sync(0xffffffffu, s0, 31), tot1 = __shfl_sync(0xffffffffu, s1, 31),
                   tot2 = __shfl_sync(0xffffffffu, s2, 31);
    const uint32_t e0 = s0 - h0, e1 = s1 - h1, e2 = s2 - h2;      // exclusive prefix of my lane
    // list starts, each list padded to a multiple of 4; packed the same way (start <= 236 fits a byte)
    uint32_t st0 = 0, st1 = 0, st2 = 0, run = 0;
#pragma unroll
    for (int c = 0; c < NCODES; ++c) {
        const uint32_t cnt = byte_of(tot0, tot1, tot2, c);
        const uint32_t sh = (c & 3) * 8;
        if (c < 4) st0 |= run << sh; else if (c < 8) st1 |= run << sh; else st2 |= run << sh;
        if (lane == 0) W.start[c] = (uint8_t)run;
        const uint32_t end = run + ((cnt + 3u) & ~3u);
        if (lane < (int)(end - run - cnt)) W.list[run + cnt + lane] = (uint8_t)READS;   // pad -> zero row
        run = end;
    }
    if (lane == 0) W.start[NCODES] = (uint8_t)run;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint32_t code = codes[i];
        if (code < NCODES) {
            const uint32_t pos = byte_of(st0, st1, st2, code) + byte_of(e0, e1, e2, code) + rank[i];
            W.list[pos] = (uint8_t)(lane * 7 + i);
        }
    }
}

__device__ __forceinline__ void build_m(const float* w1t, FrontWarp& W, int jq) {
    const float4* w4 = reinterpret_cast<const float4*>(w1t);
    float4 acc[NCODES];
    auto add4 = [&](float4& a, uint32_t q) {       // four list entries -> four W1T rows
        const float4 v0 = w4[(q & 0xffu) * JQ + jq];
        const float4 v1 = w4[((q >> 8) & 0xffu) * JQ + jq];
        const float4 v2 = w4[((q >> 16) & 0xffu) * JQ + jq];
        const float4 v3 = w4[(q >> 24) * JQ + jq];
        a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
        a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
        a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
        a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
    };
#pragma unroll
    for (int c = 0; c < NCODES; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i0 = W.start[c], i1 = W.start[c + 1];
#pragma unroll 1
        for (int i = i0; i < i1; i += 4) add4(a, *reinterpret_cast<const uint32_t*>(&W.list[i]));
        acc[c] = a;
    }
    float4* m = reinterpret_cast<float4*>(&W.m[4 * jq][0]);
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
    constexpr uint32_t WIN_BYTES = READS * COLS;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fr_smem_u32(&S.xbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fr_smem_u32(&S.xbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // W1T (with its zero row) stays resident for every window this CTA processes
        const float4* src = reinterpret_cast<const float4*>(packed + PK_W1T);
        float4* dst = reinterpret_cast<float4*>(S.w1t);
        for (int i = tid; i < W1T_ROWS * FC1 / 4; i += FR_THREADS) dst[i] = src[i];
    }
    for (int i = tid; i < FC1 * 12; i += FR_THREADS) {
        const int j = i / 12, k = i % 12;
        S.w2t[j][k] = k < FC2 ? P.W2[k * FC1 + j] : (k == FC2 ? P.b1[j] : 0.f);
    }
    __syncthreads();
    auto fetch_window = [&](int w, int buf) {      // one thread: TMA bulk copy of a whole window
        const uint32_t bar = fr_smem_u32(&S.xbar[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(WIN_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(fr_smem_u32(S.xs[buf])), "l"(x + (size_t)w * WIN_BYTES), "r"(WIN_BYTES), "r"(bar) : "memory");
    };
    if (tid == 0 && (int)blockIdx.x < nwin) fetch_window(blockIdx.x, 0);

    FrontWarp& W = S.wp[warp];
    // a/g stage: lane l < 25 owns embedding dims l and l + 25.  Lanes 25..31 run the same loop on a
    // duplicate index and store nothing: keeping the loop in warp-uniform control flow lets the
    // compiler index W2/b1 in the constant bank through the uniform datapath with a rolled loop
    // (a fully unrolled body is 88 KB of code and thrashed the instruction cache: 2.1 stalled warps
    // per issue on "no instruction").
    const int el = lane < EH ? lane : lane - 8;
    float E0[NCODES], E1[NCODES];
#pragma unroll
    for (int c = 0; c < NCODES; ++c) {
        E0[c] = packed[PK_E + c * EMB + el];
        E1[c] = packed[PK_E + c * EMB + el + EH];
    }

    int it = 0;
    for (int w = blockIdx.x; w < nwin; w += gridDim.x, ++it) {
        const int xb = it & 1;
        {   // wait for this window's bytes
            const uint32_t bar = fr_smem_u32(&S.xbar[xb]), parity = (it >> 1) & 1;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n\t"
                "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
        }
        // every warp is past the previous window (barrier at the bottom), so its buffer is free
        if (tid == 0 && w + (int)gridDim.x < nwin) fetch_window(w + gridDim.x, xb ^ 1);

        for (int p = warp; p < COLS; p += FR_WARPS) {
            sort_column(W, S.xs[xb], p, lane, status);
            __syncwarp();
            if (lane < JQ) build_m(S.w1t, W, lane);
            __syncwarp();
            float* urow = u + ((size_t)w * COLS + p) * IN0P;
            {
                float g0[FC2], g1[FC2];
#pragma unroll
                for (int k = 0; k < FC2; ++k) g0[k] = g1[k] = P.b2[k];
                const float4* mrow = reinterpret_cast<const float4*>(&W.m[0][0]);
                const float4* wrow = reinterpret_cast<const float4*>(&S.w2t[0][0]);
#pragma unroll 4
                for (int j = 0; j < FC1; ++j) {
                    const float4 m0 = mrow[j * 3], m1 = mrow[j * 3 + 1], m2 = mrow[j * 3 + 2];
                    const float4 w0 = wrow[j * 3], w1 = wrow[j * 3 + 1], w2 = wrow[j * 3 + 2];   // W2[0..9][j], b1[j]
                    // two embedding dims share every load; per dim three independent 4-term chains
                    float a0 = fmaf(m0.x, E0[0], w2.z), b0 = fmaf(m0.x, E1[0], w2.z);
                    float a1 = m1.x * E0[4], b1 = m1.x * E1[4];
                    float a2 = m2.x * E0[8], b2 = m2.x * E1[8];
                    a0 = fmaf(m0.y, E0[1], a0); b0 = fmaf(m0.y, E1[1], b0);
                    a1 = fmaf(m1.y, E0[5], a1); b1 = fmaf(m1.y, E1[5], b1);
                    a2 = fmaf(m2.y, E0[9], a2); b2 = fmaf(m2.y, E1[9], b2);
                    a0 = fmaf(m0.z, E0[2], a0); b0 = fmaf(m0.z, E1[2], b0);
                    a1 = fmaf(m1.z, E0[6], a1); b1 = fmaf(m1.z, E1[6], b1);
                    a2 = fmaf(m2.z, E0[10], a2); b2 = fmaf(m2.z, E1[10], b2);
                    a0 = fmaf(m0.w, E0[3], a0); b0 = fmaf(m0.w, E1[3], b0);
                    a1 = fmaf(m1.w, E0[7], a1); b1 = fmaf(m1.w, E1[7], b1);
                    a2 = fmaf(m2.w, E0[11], a2); b2 = fmaf(m2.w, E1[11], b2);
                    const float a = fmaxf(a0 + (a1 + a2), 0.f), b = fmaxf(b0 + (b1 + b2), 0.f);
                    g0[0] = fmaf(w0.x, a, g0[0]); g1[0] = fmaf(w0.x, b, g1[0]);
                    g0[1] = fmaf(w0.y, a, g0[1]); g1[1] = fmaf(w0.y, b, g1[1]);
                    g0[2] = fmaf(w0.z, a, g0[2]); g1[2] = fmaf(w0.z, b, g1[2]);
                    g0[3] = fmaf(w0.w, a, g0[3]); g1[3] = fmaf(w0.w, b, g1[3]);
                    g0[4] = fmaf(w1.x, a, g0[4]); g1[4] = fmaf(w1.x, b, g1[4]);
                    g0[5] = fmaf(w1.y, a, g0[5]); g1[5] = fmaf(w1.y, b, g1[5]);
                    g0[6] = fmaf(w1.z, a, g0[6]); g1[6] = fmaf(w1.z, b, g1[6]);
                    g0[7] = fmaf(w1.w, a, g0[7]); g1[7] = fmaf(w1.w, b, g1[7]);
                    g0[8] = fmaf(w2.x, a, g0[8]); g1[8] = fmaf(w2.x, b, g1[8]);
                    g0[9] = fmaf(w2.y, a, g0[9]); g1[9] = fmaf(w2.y, b, g1[9]);
                }
                if (lane < EH) {
                    float2* d0 = reinterpret_cast<float2*>(urow + lane * FC2);
                    float2* d1 = reinterpret_cast<float2*>(urow + (lane + EH) * FC2);
#pragma unroll
                    for (int k = 0; k < FC2; k += 2) {
                        d0[k / 2] = make_float2(fmaxf(g0[k], 0.f), fmaxf(g0[k + 1], 0.f));
                        d1[k / 2] = make_float2(fmaxf(g1[k], 0.f), fmaxf(g1[k + 1], 0.f));
                    }
                } else if (lane < EH + 3) {                        // zero the k-padding of the row
                    reinterpret_cast<float4*>(urow + IN0)[lane - EH] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            __syncwarp();                                          // M slab is reused by the next column
        }
        __syncthreads();                                           // window done: its xs buffer may be refilled
    }
}

cudaError_t front_setup() {
    return cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FrontSmem) + 128);
}

cudaError_t launch_front(const FrontConst& fc, const uint8_t* x, const float* packed, float* u, int nwin,
                         int* status, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    int grid = nwin < num_sms ? nwin : num_sms;
    front_kernel<<<grid, FR_THREADS, sizeof(FrontSmem) + 128, s>>>(fc, x, packed, u, nwin, status);
    return cudaGetLastError();
}

}  // namespace roko
