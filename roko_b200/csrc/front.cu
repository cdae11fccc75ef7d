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
//   a, g      two chained warp-level tensor-core GEMMs (mma.sync m16n8k8 tf32, 3xTF32 split):
//             a = relu(b1 + E^T M^T) feeds g = relu(b2 + a W2^T) without leaving registers -- the C
//             fragment of the first is the A fragment of the second under a permuted contraction order
// Odd warps sort all their columns up front so the two halves of the CTA run out of phase.  Measured
// on B200 (128 windows): sort+gather alone 0.076 ms, a/g alone 0.104 ms, together 0.151 ms -- the two
// barely overlap because legacy mma.sync holds the sub-partition's issue port for ~8.5 cycles per
// instruction (480 MAC/clk/SM, scripts/ubench/mma_rate.cu); the next step is to move a/g to tcgen05.
#include "common.cuh"

namespace roko {

constexpr int FR_THREADS = 512;
constexpr int FR_WARPS = FR_THREADS / 32;
constexpr int LIST_LEN = 240;                 // 200 reads + up to 3 pads for each of 12 codes
constexpr int JQ = FC1 / 4;                   // 25 gather lanes
constexpr int FC1P = 104;                     // fc1 width padded to 13 mma n-tiles

constexpr int FR_MAXCOLS = (COLS + FR_WARPS - 1) / FR_WARPS;   // columns a warp owns per window (6)
struct FrontList {
    alignas(16) uint8_t list[LIST_LEN];       //   240 B
    uint8_t start[16];
};
struct FrontWarp {                            // private to one warp
    float m[FC1][NCODES];                     // 4 800 B
    FrontList cols[FR_MAXCOLS];               // 1 536 B  sorted read lists of this warp's columns
};

struct FrontSmem {
    float w1t[W1T_ROWS * FC1];                // 80 400 B  [r][j], row 200 zero
    FrontWarp wp[FR_WARPS];                   // 80 896 B
    // operands of the tensor-core a/g stage (tf32 hi / lo, zero padded):
    alignas(16) float efrag[2][2][2][2][32][4];   //  8 192 B  E^T as mma A fragments [m-tile pair][m-tile][k-step][hi|lo][lane][reg]
    alignas(16) float w2h[8][FC1P];               //  3 328 B  W2[k][j] for k = 0..7 (tf32 hi), j padded to 104
    alignas(16) float w2l[8][FC1P];               //  3 328 B  (tf32 lo)
    alignas(16) float w2f[2][FC1P];               //    832 B  W2[8][j], W2[9][j] in full fp32
    alignas(16) float b1p[FC1P];                  //    416 B
    float b2p[16];
    alignas(16) uint8_t xs[READS * COLS];     // 18 000 B  the window, [read][col] (single buffer: the per-warp lists took the space)
    alignas(8) unsigned long long xbar;       // TMA arrival barrier of the window buffer
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

__device__ __forceinline__ void sort_column(FrontList& W, const uint8_t* xs, int p, int lane, int* status) {
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

__device__ __forceinline__ void build_m(const float* w1t, const FrontList& W, float (*mslab)[NCODES], int jq) {
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
    float4* m = reinterpret_cast<float4*>(&mslab[4 * jq][0]);
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
front_kernel(const uint8_t* __restrict__ x,
             const float* __restrict__ packed, float* __restrict__ u, int nwin, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrontSmem& S = *reinterpret_cast<FrontSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t WIN_BYTES = READS * COLS;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fr_smem_u32(&S.xbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // W1T (with its zero row) stays resident for every window this CTA processes
        const float4* src = reinterpret_cast<const float4*>(packed + PK_W1T);
        float4* dst = reinterpret_cast<float4*>(S.w1t);
        for (int i = tid; i < W1T_ROWS * FC1 / 4; i += FR_THREADS) dst[i] = src[i];
    }
    auto tf32r = [](float v) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v)); return __uint_as_float(u); };
    for (int i = tid; i < 2 * 2 * 2 * 2 * 32 * 4; i += FR_THREADS) {       // E^T in mma.m16n8k8 A-fragment order
        const int q = i & 3, ln = (i >> 2) & 31, hl = (i >> 7) & 1, ks = (i >> 8) & 1, mtl = (i >> 9) & 1, mp = (i >> 10) & 1;
        const int g = ln >> 2, t = ln & 3;
        const int c = 8 * ks + t + ((q & 2) ? 4 : 0);                      // a0,a1: col t   a2,a3: col t+4
        const int e = 16 * (2 * mp + mtl) + g + ((q & 1) ? 8 : 0);         // a0,a2: row g   a1,a3: row g+8
        const float v = (c < NCODES && e < EMB) ? packed[PK_E + c * EMB + e] : 0.f;
        const float hi = tf32r(v);
        (&S.efrag[0][0][0][0][0][0])[i] = hl ? v - hi : hi;
    }
    for (int i = tid; i < FC2 * FC1P; i += FR_THREADS) {
        const int k = i / FC1P, j = i % FC1P;
        const float v = j < FC1 ? packed[PK_W2 + k * FC1 + j] : 0.f;
        if (k < 8) {
            const float hi = tf32r(v);
            S.w2h[k][j] = hi;
            S.w2l[k][j] = v - hi;
        } else {
            S.w2f[k - 8][j] = v;
        }
    }
    for (int i = tid; i < FC1P; i += FR_THREADS) S.b1p[i] = i < FC1 ? packed[PK_B1 + i] : 0.f;
    if (tid < 16) S.b2p[tid] = tid < FC2 ? packed[PK_B2 + tid] : 0.f;
    __syncthreads();
    auto fetch_window = [&](int w) {               // one thread: TMA bulk copy of a whole window
        const uint32_t bar = fr_smem_u32(&S.xbar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(WIN_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(fr_smem_u32(S.xs)), "l"(x + (size_t)w * WIN_BYTES), "r"(WIN_BYTES), "r"(bar) : "memory");
    };
    if (tid == 0 && (int)blockIdx.x < nwin) fetch_window(blockIdx.x);
    __syncthreads();

    FrontWarp& W = S.wp[warp];
    // a/g stage on the tensor pipe (warp-level mma.sync m16n8k8 tf32, fp32-accurate 3xTF32 split):
    //   stage A   a[e][j] = relu(b1[j] + sum_c E^T[e][c] M^T[c][j])      M = e (4 tiles of 16), N = j (13 tiles of 8), K = c (2 steps)
    //   stage G   g[e][k] = relu(b2[k] + sum_j a[e][j] W2^T[j][k])       M = e, N = k (2 tiles of 8), K = j (one stage-A n-tile per step)
    // The C fragment of stage A is reused as the A fragment of stage G by reading the contraction index in
    // the order j = 8nt + {0,2,4,6,1,3,5,7}: C holds columns (2t, 2t+1), A wants columns (t, t+4).
    const int fg = lane >> 2, ft = lane & 3;
    int it = 0;
    for (int w = blockIdx.x; w < nwin; w += gridDim.x, ++it) {
        {   // wait for this window's bytes
            const uint32_t bar = fr_smem_u32(&S.xbar), parity = it & 1;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n\t"
                "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
        }

        // Static column ownership (p = warp, warp + 16, ...).  The two halves of the CTA run out of phase on
        // purpose: odd warps sort ALL their columns before the first gather, even warps sort each column
        // just before it.  The gather is bound by shared-memory wavefronts and the a/g stage by the tensor
        // pipe; with every warp in the same phase at the same time the two costs simply add
        // (measured: 0.053 + 0.090 + sort 0.010 + fixed 0.014 = 0.159 ms per 128 windows).
        const bool presort = warp & 1;
        if (presort)
            for (int ci = 0, p = warp; p < COLS; p += FR_WARPS, ++ci) sort_column(W.cols[ci], S.xs, p, lane, status);
        for (int ci = 0, p = warp; p < COLS; p += FR_WARPS, ++ci) {
            if (!presort) sort_column(W.cols[ci], S.xs, p, lane, status);
            __syncwarp();
            if (lane < JQ) build_m(S.w1t, W.cols[ci], W.m, lane);
            __syncwarp();
            float* urow = u + ((size_t)w * COLS + p) * IN0P;
            auto split = [](float v, uint32_t& hi, uint32_t& lo) {
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
                lo = __float_as_uint(v - __uint_as_float(hi));
            };
            auto mma = [](float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            };
#pragma unroll 1
            for (int mp = 0; mp < 2; ++mp) {                       // two passes of two 16-row tiles of e
                uint32_t eh[2][2][4], el[2][2][4];
#pragma unroll
                for (int mtl = 0; mtl < 2; ++mtl)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint4 h4 = *reinterpret_cast<const uint4*>(&S.efrag[mp][mtl][ks][0][lane][0]);
                        const uint4 l4 = *reinterpret_cast<const uint4*>(&S.efrag[mp][mtl][ks][1][lane][0]);
                        eh[mtl][ks][0] = h4.x; eh[mtl][ks][1] = h4.y; eh[mtl][ks][2] = h4.z; eh[mtl][ks][3] = h4.w;
                        el[mtl][ks][0] = l4.x; el[mtl][ks][1] = l4.y; el[mtl][ks][2] = l4.z; el[mtl][ks][3] = l4.w;
                    }
                // k = 0..7 of stage G run on the tensor pipe; k = 8, 9 would need a second n-tile that is 3/4
                // padding, so they are accumulated with plain FFMAs from the same a values (4 partial sums per
                // tile, reduced over the 4 lanes of a row group at the end of the column)
                float gacc[2][4], g89[2][4];
                {
                    const float bx = S.b2p[2 * ft], by = S.b2p[2 * ft + 1];
#pragma unroll
                    for (int mtl = 0; mtl < 2; ++mtl) {
                        gacc[mtl][0] = bx; gacc[mtl][1] = by; gacc[mtl][2] = bx; gacc[mtl][3] = by;
                        g89[mtl][0] = g89[mtl][1] = g89[mtl][2] = g89[mtl][3] = 0.f;
                    }
                }
#pragma unroll 1
                for (int nt = 0; nt < FC1P / 8; ++nt) {
                    // B fragments of stage A: M^T[c][j], j = 8nt + g, c = t, t+4, t+8 (c >= 12 is zero padding)
                    const int jr = 8 * nt + fg;
                    const bool jv = jr < FC1;
                    uint32_t m0h, m0l, m4h, m4l, m8h, m8l;
                    split(jv ? W.m[jr][ft] : 0.f, m0h, m0l);
                    split(jv ? W.m[jr][ft + 4] : 0.f, m4h, m4l);
                    split(jv ? W.m[jr][ft + 8] : 0.f, m8h, m8l);
                    const float2 b1v = *reinterpret_cast<const float2*>(&S.b1p[8 * nt + 2 * ft]);
                    // B fragments of stage G: W2[k = 8kt + g][j = 8nt + 2t, 2t + 1]
                    const uint2 wh = *reinterpret_cast<const uint2*>(&S.w2h[fg][8 * nt + 2 * ft]);
                    const uint2 wl = *reinterpret_cast<const uint2*>(&S.w2l[fg][8 * nt + 2 * ft]);
                    const float2 w8 = *reinterpret_cast<const float2*>(&S.w2f[0][8 * nt + 2 * ft]);   // W2[8][j], full fp32
                    const float2 w9 = *reinterpret_cast<const float2*>(&S.w2f[1][8 * nt + 2 * ft]);   // W2[9][j]
#pragma unroll
                    for (int mtl = 0; mtl < 2; ++mtl) {
                        float c[4] = {0.f, 0.f, 0.f, 0.f};
                        mma(c, el[mtl][0], m0h, m4h); mma(c, eh[mtl][0], m0l, m4l); mma(c, eh[mtl][0], m0h, m4h);
                        mma(c, el[mtl][1], m8h, 0u);  mma(c, eh[mtl][1], m8l, 0u);  mma(c, eh[mtl][1], m8h, 0u);
                        // c0:(g, 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)  ->  A of stage G: a0:(g, t) a1:(g+8, t) a2:(g, t+4) a3:(g+8, t+4)
                        const float a00 = fmaxf(c[0] + b1v.x, 0.f), a10 = fmaxf(c[2] + b1v.x, 0.f);   // (g, 2t)  (g+8, 2t)
                        const float a01 = fmaxf(c[1] + b1v.y, 0.f), a11 = fmaxf(c[3] + b1v.y, 0.f);   // (g, 2t+1) (g+8, 2t+1)
                        uint32_t ah[4], al[4];
                        split(a00, ah[0], al[0]);
                        split(a10, ah[1], al[1]);
                        split(a01, ah[2], al[2]);
                        split(a11, ah[3], al[3]);
                        mma(gacc[mtl], al, wh.x, wh.y);
                        mma(gacc[mtl], ah, wl.x, wl.y);
                        mma(gacc[mtl], ah, wh.x, wh.y);
                        g89[mtl][0] = fmaf(w8.x, a00, fmaf(w8.y, a01, g89[mtl][0]));   // k = 8, row g
                        g89[mtl][1] = fmaf(w9.x, a00, fmaf(w9.y, a01, g89[mtl][1]));   // k = 9, row g
                        g89[mtl][2] = fmaf(w8.x, a10, fmaf(w8.y, a11, g89[mtl][2]));   // k = 8, row g+8
                        g89[mtl][3] = fmaf(w9.x, a10, fmaf(w9.y, a11, g89[mtl][3]));   // k = 9, row g+8
                    }
                }
#pragma unroll
                for (int mtl = 0; mtl < 2; ++mtl) {
                    const int e = 16 * (2 * mp + mtl) + fg, k = 2 * ft;
                    if (e < EMB)
                        *reinterpret_cast<float2*>(urow + e * FC2 + k) = make_float2(fmaxf(gacc[mtl][0], 0.f), fmaxf(gacc[mtl][1], 0.f));
                    if (e + 8 < EMB)
                        *reinterpret_cast<float2*>(urow + (e + 8) * FC2 + k) = make_float2(fmaxf(gacc[mtl][2], 0.f), fmaxf(gacc[mtl][3], 0.f));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {                      // finish k = 8, 9 over the 4 lanes that share rows g, g+8
                        g89[mtl][i] += __shfl_xor_sync(0xffffffffu, g89[mtl][i], 1);
                        g89[mtl][i] += __shfl_xor_sync(0xffffffffu, g89[mtl][i], 2);
                    }
                    if (ft == 0) {
                        const float b8 = S.b2p[8], b9 = S.b2p[9];
                        if (e < EMB)
                            *reinterpret_cast<float2*>(urow + e * FC2 + 8) = make_float2(fmaxf(g89[mtl][0] + b8, 0.f), fmaxf(g89[mtl][1] + b9, 0.f));
                        if (e + 8 < EMB)
                            *reinterpret_cast<float2*>(urow + (e + 8) * FC2 + 8) = make_float2(fmaxf(g89[mtl][2] + b8, 0.f), fmaxf(g89[mtl][3] + b9, 0.f));
                    }
                }
            }
            if (lane < (IN0P - IN0) / 4)                               // zero the k-padding of the row
                reinterpret_cast<float4*>(urow + IN0)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();                                          // M slab is reused by the next column
        }
        __syncthreads();                                           // every warp has sorted its columns: refill the buffer
        if (tid == 0 && w + (int)gridDim.x < nwin) fetch_window(w + gridDim.x);
    }
}

cudaError_t front_setup() {
    return cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FrontSmem) + 128);
}

cudaError_t launch_front(const uint8_t* x, const float* packed, float* u, int nwin,
                         int* status, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    int grid = nwin < num_sms ? nwin : num_sms;
    front_kernel<<<grid, FR_THREADS, sizeof(FrontSmem) + 128, s>>>(x, packed, u, nwin, status);
    return cudaGetLastError();
}

}  // namespace roko
