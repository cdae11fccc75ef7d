// Small kernels of the training path (everything that is not a GEMM or a recurrence).
//
// Train-mode front end, reference roko/rnn_model.py:47-56 with dropout active: the embedding output
// is masked per (read, column, channel) BEFORE fc1 mixes the reads, so the one-hot factorisation of
// the inference kernel (front.cu) does not apply; the masked embedding is materialised once, in the
// layout fc1 wants ([window][column][channel][read], read contiguous), and kept for the backward.
#include "train.cuh"

namespace roko {

constexpr int TR_THREADS = 256;
constexpr int PE = EMB * READS;           // 10 000 embedding outputs per (window, column)

// ep[(b,p,e)][r] = dropout(E[x[b][r][p]][e])                       rnn_model.py:47-48 (+ permute)
// The keep bits of the block's 10 000 outputs are filed as 313 words (MASK_WORDS per block) for the backward.
__global__ void __launch_bounds__(TR_THREADS)
embed_drop_kernel(const uint8_t* __restrict__ x, const float* __restrict__ E, float* __restrict__ ep,
                  uint32_t* __restrict__ bits, uint8_t* __restrict__ xt, uint32_t* __restrict__ bitsT, DropCfg d,
                  int* __restrict__ status) {
    __shared__ float Es[NCODES * EMB];
    __shared__ uint8_t codes[READS];
    __shared__ uint8_t kept[PE];                                        // keep flags, [channel][read], for the transposed file
    const int bp = blockIdx.x, b = bp / COLS, p = bp - b * COLS, tid = threadIdx.x;
    for (int i = tid; i < NCODES * EMB; i += TR_THREADS) Es[i] = E[i];
    if (tid < READS) {
        uint8_t c = x[((size_t)b * READS + tid) * COLS + p];
        if (c >= NCODES) { atomicOr(status, 1); c = 0; }
        codes[tid] = c;
        if (xt) xt[(size_t)bp * READS + tid] = c;
    }
    __syncthreads();
    float* dst = ep + (size_t)bp * PE;
    uint32_t* bdst = bits + (size_t)bp * MASK_WORDS;
    for (int idx = tid; idx < MASK_WORDS * 32; idx += TR_THREADS) {      // whole warps: the ballot needs all 32 lanes
        bool keep = false;
        if (idx < PE) {
            const int e = idx / READS, r = idx - e * READS;
            const float v = Es[codes[r] * EMB + e];
            const unsigned long long i0 = (((unsigned long long)b * READS + r) * COLS + p) * EMB + e;
            keep = drop_keep(d, DROP_EMB, i0);
            if (ep) dst[idx] = keep ? v * d.scale : 0.f;
            kept[idx] = keep;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, keep);
        if ((tid & 31) == 0) bdst[idx >> 5] = word;
    }
    if (bitsT) {
        __syncthreads();
        if (tid < READS) {
            unsigned long long w = 0ull;
#pragma unroll 10
            for (int e = 0; e < EMB; ++e) w |= (unsigned long long)kept[e * READS + tid] << e;
            reinterpret_cast<uint2*>(bitsT)[(size_t)bp * READS + tid] = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
        }
    }
}

// u[m][e*10+k] = dropout(relu(b2[k] + sum_j W2[k][j] a1[(m,e)][j]))   rnn_model.py:53-56
// 96-row tiles staged with 16-byte loads; thread = (row, half of the 10 outputs): per 4 inputs one 16-byte read of its
// row (lanes = consecutive rows: conflict free) and five broadcast 16-byte reads of W2 for 20 FMA.
constexpr int F2_ROWS = 64;                                // fc2_bwd tile
constexpr int F2F_ROWS = 96, F2F_THREADS = 2 * F2F_ROWS;
__global__ void __launch_bounds__(F2F_THREADS)
fc2_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ W2, const float* __restrict__ b2,
               float* __restrict__ u, int rows50, DropCfg d) {
    __shared__ __align__(16) float as[F2F_ROWS][FC1];
    __shared__ __align__(16) float ws[FC2][FC1];
    __shared__ float bs[FC2];
    const int tid = threadIdx.x, row0 = blockIdx.x * F2F_ROWS;
    for (int i = tid; i < FC2 * FC1; i += F2F_THREADS) (&ws[0][0])[i] = W2[i];
    if (tid < FC2) bs[tid] = b2[tid];
    const int nrow = (rows50 - row0) < F2F_ROWS ? (rows50 - row0) : F2F_ROWS;
    const float4* src = reinterpret_cast<const float4*>(a1 + (size_t)row0 * FC1);      // 96 * 100 floats per tile: 16-byte aligned
    float4* dst = reinterpret_cast<float4*>(&as[0][0]);
#pragma unroll
    for (int it = 0; it < (F2F_ROWS * FC1 / 4 + F2F_THREADS - 1) / F2F_THREADS; ++it) {
        const int i = tid + it * F2F_THREADS;
        if (i < F2F_ROWS * FC1 / 4) dst[i] = i < nrow * (FC1 / 4) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int half = tid / F2F_ROWS, row = tid - half * F2F_ROWS, kb = half * (FC2 / 2);
    if (row < nrow) {
        float acc[FC2 / 2];
#pragma unroll
        for (int k = 0; k < FC2 / 2; ++k) acc[k] = bs[kb + k];
#pragma unroll 5
        for (int q = 0; q < FC1 / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(&as[row][4 * q]);
#pragma unroll
            for (int k = 0; k < FC2 / 2; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(&ws[kb + k][4 * q]);
                acc[k] = fmaf(a.x, w.x, acc[k]);             // j ascending, as the scalar loop it replaces
                acc[k] = fmaf(a.y, w.y, acc[k]);
                acc[k] = fmaf(a.z, w.z, acc[k]);
                acc[k] = fmaf(a.w, w.w, acc[k]);
            }
        }
        const int g = row0 + row, m = g / EMB, e = g - m * EMB;
#pragma unroll
        for (int k = 0; k < FC2 / 2; ++k) {
            float v = fmaxf(acc[k], 0.f);
            v = drop_keep(d, DROP_FC2, (unsigned long long)g * FC2 + kb + k) ? v * d.scale : 0.f;
            u[(size_t)m * IN0P + e * FC2 + kb + k] = v;
        }
    }
}

// Backward of fc2 + its ReLU/dropout and of fc1's ReLU/dropout, one pass over a1:
//   dg[k]  = du[(m,e),k] * scale * [u > 0]                 (u > 0  <=>  kept and pre-activation > 0)
//   dW2   += dg (x) a1,  db2 += dg
//   dap[j] = (sum_k dg[k] W2[k][j]) * scale * [a1[j] > 0]   written over a1
// Thread (j, half) keeps column j of W2 and of the dW2 partial sums in registers and walks 32 rows
// of the 64-row tile: per row one a1 value, the row's 10 dg (broadcast), 20 FMA.
__global__ void __launch_bounds__(TR_THREADS)
fc2_bwd_kernel(const float* __restrict__ du, const float* __restrict__ u, float* __restrict__ a1,
               const float* __restrict__ W2, float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ db1,
               int rows50, float scale) {
    __shared__ __align__(16) float as[F2_ROWS][FC1];
    __shared__ __align__(16) float dgs[F2_ROWS][12];        // 10 used, padded for LDS.128
    const int tid = threadIdx.x, j = tid & 127, half = tid >> 7;
    const bool active = j < FC1;
    float w2c[FC2], accw[FC2];
#pragma unroll
    for (int k = 0; k < FC2; ++k) { w2c[k] = active ? W2[k * FC1 + j] : 0.f; accw[k] = 0.f; }
    float accb = 0.f, accd = 0.f;                           // db2 partial (threads 0..9), db1[j] partial: column sum of d(a1)
    const int ntiles = (rows50 + F2_ROWS - 1) / F2_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * F2_ROWS;
        const int nrow = (rows50 - row0) < F2_ROWS ? (rows50 - row0) : F2_ROWS;
        __syncthreads();                                    // previous tile's readers are done
        {   // 64 * 100 floats per tile: 16-byte loads, all of a thread's 6-7 in flight
            const float4* src = reinterpret_cast<const float4*>(a1 + (size_t)row0 * FC1);
            float4* dst = reinterpret_cast<float4*>(&as[0][0]);
#pragma unroll
            for (int it = 0; it < (F2_ROWS * FC1 / 4 + TR_THREADS - 1) / TR_THREADS; ++it) {
                const int i = tid + it * TR_THREADS;
                if (i < F2_ROWS * FC1 / 4) dst[i] = i < nrow * (FC1 / 4) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int i = tid; i < F2_ROWS * 12; i += TR_THREADS) {
            const int row = i / 12, k = i - row * 12;
            float v = 0.f;
            if (row < nrow && k < FC2) {
                const int g = row0 + row, m = g / EMB, e = g - m * EMB;
                const size_t off = (size_t)m * IN0P + e * FC2 + k;
                v = u[off] > 0.f ? du[off] * scale : 0.f;
            }
            dgs[row][k] = v;
        }
        __syncthreads();
        if (tid < FC2)
            for (int row = 0; row < F2_ROWS; ++row) accb += dgs[row][tid];
        if (active) {
            const int rbeg = half * (F2_ROWS / 2);
#pragma unroll 4
            for (int row = rbeg; row < rbeg + F2_ROWS / 2; ++row) {
                const float a = as[row][j];
                const float4 d0 = *reinterpret_cast<const float4*>(&dgs[row][0]);
                const float4 d1 = *reinterpret_cast<const float4*>(&dgs[row][4]);
                const float4 d2 = *reinterpret_cast<const float4*>(&dgs[row][8]);
                const float dg[FC2] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
                float da = 0.f;
#pragma unroll
                for (int k = 0; k < FC2; ++k) {
                    accw[k] = fmaf(dg[k], a, accw[k]);
                    da = fmaf(dg[k], w2c[k], da);
                }
                const float dpre = a > 0.f ? da * scale : 0.f;     // rows past the end hold a = 0
                accd += dpre;
                if (row < nrow) a1[(size_t)(row0 + row) * FC1 + j] = dpre;
            }
        }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < FC2; ++k) atomicAdd(dW2 + k * FC1 + j, accw[k]);
        if (db1) atomicAdd(db1 + j, accd);
    }
    if (tid < FC2) atomicAdd(db2 + tid, accb);
}

// out = dropout(in) with the mask of `site` (forward between GRU layers, rnn_model.py:41, and its backward)
__global__ void __launch_bounds__(TR_THREADS)
drop_apply_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, unsigned int site, DropCfg d) {
    for (size_t i = (size_t)blockIdx.x * TR_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * TR_THREADS)
        out[i] = drop_keep(d, site, i) ? in[i] * d.scale : 0.f;
}

__global__ void __launch_bounds__(TR_THREADS)
drop_mask_kernel(unsigned int site, size_t n, uint8_t* __restrict__ out, DropCfg d) {
    for (size_t i = (size_t)blockIdx.x * TR_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * TR_THREADS)
        out[i] = drop_keep(d, site, i) ? 1 : 0;
}

// out[n] += sum_m A[m*lda + n]   (bias gradients)
__global__ void __launch_bounds__(TR_THREADS)
colsum_kernel(const float* __restrict__ A, int lda, int rows, int ncols, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cl;
    float acc = 0.f;
    if (n < ncols)
        for (int m = blockIdx.y * 8 + rl; m < rows; m += gridDim.y * 8) acc += A[(size_t)m * lda + n];
    part[rl][cl] = acc;
    __syncthreads();
    if (rl == 0 && n < ncols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][cl];
        atomicAdd(out + n, t);
    }
}

// dE[c][e] += sum over (b, r, p) with x[b][r][p] == c of mask * scale * dep[(b,p,e)][r]    (embedding backward)
// Per (window, column): the 50 x 200 tile of dep is staged with the forward's keep bits applied, the 200
// reads are counting-sorted by code (stable, ballot based), and worker (c, e) adds tile[e][r] over the
// reads of code c -- 10 000 adds per tile.  Partial sums stay in registers across tiles.
__global__ void __launch_bounds__(TR_THREADS)
embed_grad_kernel(const float* __restrict__ dep, const uint8_t* __restrict__ x, const uint32_t* __restrict__ bits,
                  float* __restrict__ dE, int nwin, float scale) {
    __shared__ float tile[EMB][READS + 1];
    __shared__ uint8_t codes[READS];
    __shared__ uint8_t order[READS];                       // reads sorted by code
    __shared__ int cbeg[NCODES + 1];
    __shared__ int ccount[NCODES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc[3] = {0.f, 0.f, 0.f};                        // workers q = tid + 256 i < 600:  c = q / 50, e = q % 50
    const int nbp = nwin * COLS;
    for (int bp = blockIdx.x; bp < nbp; bp += gridDim.x) {
        const int b = bp / COLS, p = bp - b * COLS;
        __syncthreads();
        const float* src = dep + (size_t)bp * PE;
        const uint32_t* bsrc = bits + (size_t)bp * MASK_WORDS;
#pragma unroll 8
        for (int i = tid; i < PE; i += TR_THREADS) {        // 8 loads in flight per thread
            const int ee = i / READS;
            const bool keep = (bsrc[i >> 5] >> (i & 31)) & 1u;
            tile[ee][i - ee * READS] = keep ? src[i] * scale : 0.f;
        }
        if (tid < READS) {
            const uint8_t c = x[((size_t)b * READS + tid) * COLS + p];
            codes[tid] = c < NCODES ? c : 0;
        }
        __syncthreads();
        // stable counting sort: warp w ranks the reads of codes w and w + 8
        for (int c = warp; c < NCODES; c += TR_THREADS / 32) {
            int n = 0;
            for (int r0 = 0; r0 < READS; r0 += 32) {
                const int r = r0 + lane;
                n += __popc(__ballot_sync(0xffffffffu, r < READS && codes[r] == c));
            }
            if (lane == 0) ccount[c] = n;
        }
        __syncthreads();
        if (tid == 0) {
            int o = 0;
            for (int c = 0; c < NCODES; ++c) { cbeg[c] = o; o += ccount[c]; }
            cbeg[NCODES] = o;
        }
        __syncthreads();
        for (int c = warp; c < NCODES; c += TR_THREADS / 32) {
            int o = cbeg[c];
            for (int r0 = 0; r0 < READS; r0 += 32) {
                const int r = r0 + lane;
                const bool hit = r < READS && codes[r] == c;
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (hit) order[o + __popc(m & ((1u << lane) - 1u))] = (uint8_t)r;
                o += __popc(m);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int q = tid + i * TR_THREADS;
            if (q < NCODES * EMB) {
                const int c = q / EMB, e = q - c * EMB;
                float a = acc[i];
                for (int k = cbeg[c]; k < cbeg[c + 1]; ++k) a += tile[e][order[k]];
                acc[i] = a;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int q = tid + i * TR_THREADS;
        if (q < NCODES * EMB) atomicAdd(dE + q, acc[i]);
    }
}

// Bias gradients of one GRU layer, both directions, in one pass over dgi [rows][768] and dghn [rows][256]:
//   b_ih[d] = colsum(dgi_d)          b_hh[d] = (colsum(dgi_d)[r, z], colsum(dghn_d))
__global__ void __launch_bounds__(TR_THREADS)
gru_bias_grad_kernel(const float* __restrict__ dgi, const float* __restrict__ dghn, int rows,
                     float* __restrict__ bih0, float* __restrict__ bhh0, float* __restrict__ bih1,
                     float* __restrict__ bhh1) {
    __shared__ float part[8][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cl;                  // 0..767 dgi, 768..1023 dghn
    const bool from_gi = col < GI_N;
    const float* src = from_gi ? dgi + col : dghn + (col - GI_N);
    const int ld = from_gi ? GI_N : OUT_W;
    float acc = 0.f;
    for (int m = blockIdx.y * 8 + rl; m < rows; m += gridDim.y * 8) acc += src[(size_t)m * ld];
    part[rl][cl] = acc;
    __syncthreads();
    if (rl == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][cl];
        if (from_gi) {
            const int d = col / G3, q = col - d * G3;
            atomicAdd((d ? bih1 : bih0) + q, t);
            if (q < 2 * HID) atomicAdd((d ? bhh1 : bhh0) + q, t);
        } else {
            const int c = col - GI_N, d = c / HID, jj = c - d * HID;
            atomicAdd((d ? bhh1 : bhh0) + 2 * HID + jj, t);
        }
    }
}

static int grid_for(size_t n, int cap) {
    size_t g = (n + TR_THREADS - 1) / TR_THREADS;
    return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

cudaError_t launch_embed_drop(const uint8_t* x, const float* E, float* ep, uint32_t* bits, uint8_t* xt, uint32_t* bitsT, int nwin,
                              DropCfg d, int* status, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    embed_drop_kernel<<<nwin * COLS, TR_THREADS, 0, s>>>(x, E, ep, bits, xt, bitsT, d, status);
    return cudaGetLastError();
}

cudaError_t launch_fc2_fwd(const float* a1, const float* W2, const float* b2, float* u, int rows50, DropCfg d,
                           cudaStream_t s) {
    if (rows50 <= 0) return cudaSuccess;
    fc2_fwd_kernel<<<(rows50 + F2F_ROWS - 1) / F2F_ROWS, F2F_THREADS, 0, s>>>(a1, W2, b2, u, rows50, d);
    return cudaGetLastError();
}

cudaError_t launch_fc2_bwd(const float* du, const float* u, float* a1_dap, const float* W2, float* dW2, float* db2, float* db1,
                           int rows50, float scale, int num_sms, cudaStream_t s) {
    if (rows50 <= 0) return cudaSuccess;
    const int ntiles = (rows50 + F2_ROWS - 1) / F2_ROWS;
    const int grid = ntiles < 4 * num_sms ? ntiles : 4 * num_sms;
    fc2_bwd_kernel<<<grid, TR_THREADS, 0, s>>>(du, u, a1_dap, W2, dW2, db2, db1, rows50, scale);
    return cudaGetLastError();
}

cudaError_t launch_drop_apply(const float* in, float* out, size_t n, unsigned int site, DropCfg d, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    drop_apply_kernel<<<grid_for(n, 148 * 16), TR_THREADS, 0, s>>>(in, out, n, site, d);
    return cudaGetLastError();
}

cudaError_t launch_drop_mask(unsigned int site, size_t n, uint8_t* out, DropCfg d, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    drop_mask_kernel<<<grid_for(n, 148 * 16), TR_THREADS, 0, s>>>(site, n, out, d);
    return cudaGetLastError();
}

cudaError_t launch_colsum(const float* A, int lda, int rows, int ncols, float* out, cudaStream_t s) {
    if (rows <= 0 || ncols <= 0) return cudaSuccess;
    int gy = (rows + 63) / 64;
    if (gy > 256) gy = 256;
    colsum_kernel<<<dim3((ncols + 31) / 32, gy), TR_THREADS, 0, s>>>(A, lda, rows, ncols, out);
    return cudaGetLastError();
}

cudaError_t launch_embed_grad(const float* dep, const uint8_t* x, const uint32_t* bits, float* dE, int nwin,
                              float scale, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int nbp = nwin * COLS;
    const int grid = nbp < 4 * num_sms ? nbp : 4 * num_sms;         // 41 KB of shared memory each: 4 fit an SM
    embed_grad_kernel<<<grid, TR_THREADS, 0, s>>>(dep, x, bits, dE, nwin, scale);
    return cudaGetLastError();
}

cudaError_t launch_gru_bias_grad(const float* dgi, const float* dghn, int rows, float* bih0, float* bhh0,
                                 float* bih1, float* bhh1, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    int gy = (rows + 63) / 64;
    if (gy > 64) gy = 64;
    gru_bias_grad_kernel<<<dim3((GI_N + OUT_W) / 32, gy), TR_THREADS, 0, s>>>(dgi, dghn, rows, bih0, bhh0, bih1, bhh1);
    return cudaGetLastError();
}

}  // namespace roko
