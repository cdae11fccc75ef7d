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
__global__ void __launch_bounds__(TR_THREADS)
embed_drop_kernel(const uint8_t* __restrict__ x, const float* __restrict__ E, float* __restrict__ ep,
                  DropCfg d, int* __restrict__ status) {
    __shared__ float Es[NCODES * EMB];
    __shared__ uint8_t codes[READS];
    const int bp = blockIdx.x, b = bp / COLS, p = bp - b * COLS, tid = threadIdx.x;
    for (int i = tid; i < NCODES * EMB; i += TR_THREADS) Es[i] = E[i];
    if (tid < READS) {
        uint8_t c = x[((size_t)b * READS + tid) * COLS + p];
        if (c >= NCODES) { atomicOr(status, 1); c = 0; }
        codes[tid] = c;
    }
    __syncthreads();
    float* dst = ep + (size_t)bp * PE;
    for (int idx = tid; idx < PE; idx += TR_THREADS) {
        const int e = idx / READS, r = idx - e * READS;
        const float v = Es[codes[r] * EMB + e];
        const unsigned long long i0 = (((unsigned long long)b * READS + r) * COLS + p) * EMB + e;
        dst[idx] = drop_keep(d, DROP_EMB, i0) ? v * d.scale : 0.f;
    }
}

// u[m][e*10+k] = dropout(relu(b2[k] + sum_j W2[k][j] a1[(m,e)][j]))   rnn_model.py:53-56
constexpr int F2_ROWS = 64;
__global__ void __launch_bounds__(TR_THREADS)
fc2_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ W2, const float* __restrict__ b2,
               float* __restrict__ u, int rows50, DropCfg d) {
    __shared__ float as[F2_ROWS][FC1];
    __shared__ float ws[FC2][FC1];
    __shared__ float bs[FC2];
    const int tid = threadIdx.x, row0 = blockIdx.x * F2_ROWS;
    for (int i = tid; i < FC2 * FC1; i += TR_THREADS) (&ws[0][0])[i] = W2[i];
    if (tid < FC2) bs[tid] = b2[tid];
    const int nrow = (rows50 - row0) < F2_ROWS ? (rows50 - row0) : F2_ROWS;
    for (int i = tid; i < F2_ROWS * FC1; i += TR_THREADS)
        (&as[0][0])[i] = i < nrow * FC1 ? a1[(size_t)row0 * FC1 + i] : 0.f;
    __syncthreads();
    const int row = tid >> 2, kq = tid & 3;
    if (row < nrow) {
        const int g = row0 + row, m = g / EMB, e = g - m * EMB;
        for (int k = kq; k < FC2; k += 4) {
            float acc = bs[k];
#pragma unroll 4
            for (int j = 0; j < FC1; ++j) acc = fmaf(as[row][j], ws[k][j], acc);
            acc = fmaxf(acc, 0.f);
            acc = drop_keep(d, DROP_FC2, (unsigned long long)g * FC2 + k) ? acc * d.scale : 0.f;
            u[(size_t)m * IN0P + e * FC2 + k] = acc;
        }
    }
}

// Backward of fc2 + its ReLU/dropout and of fc1's ReLU/dropout, one pass over a1:
//   dg[k]  = du[(m,e),k] * scale * [u > 0]                 (u > 0  <=>  kept and pre-activation > 0)
//   dW2   += dg (x) a1,  db2 += dg
//   dap[j] = (sum_k dg[k] W2[k][j]) * scale * [a1[j] > 0]   written over a1
__global__ void __launch_bounds__(TR_THREADS)
fc2_bwd_kernel(const float* __restrict__ du, const float* __restrict__ u, float* __restrict__ a1,
               const float* __restrict__ W2, float* __restrict__ dW2, float* __restrict__ db2, int rows50,
               float scale) {
    __shared__ float as[F2_ROWS][FC1];
    __shared__ float ws[FC2][FC1];
    __shared__ float dgs[F2_ROWS][FC2];
    const int tid = threadIdx.x;
    for (int i = tid; i < FC2 * FC1; i += TR_THREADS) (&ws[0][0])[i] = W2[i];
    float accw[4] = {0.f, 0.f, 0.f, 0.f};
    float accb = 0.f;
    int ok[4], oj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = tid + i * TR_THREADS;
        ok[i] = o < FC2 * FC1 ? o / FC1 : -1;
        oj[i] = o % FC1;
    }
    const int ntiles = (rows50 + F2_ROWS - 1) / F2_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * F2_ROWS;
        const int nrow = (rows50 - row0) < F2_ROWS ? (rows50 - row0) : F2_ROWS;
        __syncthreads();                                    // previous tile's readers are done
        for (int i = tid; i < F2_ROWS * FC1; i += TR_THREADS)
            (&as[0][0])[i] = i < nrow * FC1 ? a1[(size_t)row0 * FC1 + i] : 0.f;
        for (int i = tid; i < F2_ROWS * FC2; i += TR_THREADS) {
            const int row = i / FC2, k = i - row * FC2;
            float v = 0.f;
            if (row < nrow) {
                const int g = row0 + row, m = g / EMB, e = g - m * EMB;
                const size_t off = (size_t)m * IN0P + e * FC2 + k;
                v = u[off] > 0.f ? du[off] * scale : 0.f;
            }
            dgs[row][k] = v;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (ok[i] < 0) continue;
            float a = accw[i];
            for (int row = 0; row < F2_ROWS; ++row) a = fmaf(dgs[row][ok[i]], as[row][oj[i]], a);
            accw[i] = a;
        }
        if (tid < FC2)
            for (int row = 0; row < F2_ROWS; ++row) accb += dgs[row][tid];
        for (int i = tid; i < nrow * FC1; i += TR_THREADS) {
            const int row = i / FC1, j = i - row * FC1;
            float da = 0.f;
#pragma unroll
            for (int k = 0; k < FC2; ++k) da = fmaf(dgs[row][k], ws[k][j], da);
            a1[(size_t)row0 * FC1 + i] = as[row][j] > 0.f ? da * scale : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (ok[i] >= 0) atomicAdd(dW2 + tid + i * TR_THREADS, accw[i]);
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
__global__ void __launch_bounds__(TR_THREADS)
embed_grad_kernel(const float* __restrict__ dep, const uint8_t* __restrict__ x, float* __restrict__ dE, int nwin,
                  DropCfg d) {
    __shared__ float tile[EMB][READS + 1];
    __shared__ uint8_t codes[READS];
    __shared__ float tab[NCODES * EMB];
    const int tid = threadIdx.x;
    const int e = tid / 5, rl = tid - e * 5;               // 250 workers: channel e, reads rl, rl+5, ...
    float acc[NCODES];
#pragma unroll
    for (int c = 0; c < NCODES; ++c) acc[c] = 0.f;
    for (int i = tid; i < NCODES * EMB; i += TR_THREADS) tab[i] = 0.f;
    const int nbp = nwin * COLS;
    for (int bp = blockIdx.x; bp < nbp; bp += gridDim.x) {
        const int b = bp / COLS, p = bp - b * COLS;
        __syncthreads();
        const float* src = dep + (size_t)bp * PE;
        for (int i = tid; i < PE; i += TR_THREADS) {
            const int ee = i / READS;
            tile[ee][i - ee * READS] = src[i];
        }
        if (tid < READS) {
            const uint8_t c = x[((size_t)b * READS + tid) * COLS + p];
            codes[tid] = c < NCODES ? c : 0;
        }
        __syncthreads();
        if (e < EMB) {
            for (int r = rl; r < READS; r += 5) {
                const unsigned long long i0 = (((unsigned long long)b * READS + r) * COLS + p) * EMB + e;
                const float v = drop_keep(d, DROP_EMB, i0) ? tile[e][r] * d.scale : 0.f;
                const int c = codes[r];
#pragma unroll
                for (int cc = 0; cc < NCODES; ++cc) acc[cc] += (c == cc) ? v : 0.f;
            }
        }
    }
    __syncthreads();
    if (e < EMB) {
#pragma unroll
        for (int c = 0; c < NCODES; ++c) atomicAdd(&tab[c * EMB + e], acc[c]);
    }
    __syncthreads();
    for (int i = tid; i < NCODES * EMB; i += TR_THREADS) atomicAdd(dE + i, tab[i]);
}

static int grid_for(size_t n, int cap) {
    size_t g = (n + TR_THREADS - 1) / TR_THREADS;
    return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

cudaError_t launch_embed_drop(const uint8_t* x, const float* E, float* ep, int nwin, DropCfg d, int* status,
                              cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    embed_drop_kernel<<<nwin * COLS, TR_THREADS, 0, s>>>(x, E, ep, d, status);
    return cudaGetLastError();
}

cudaError_t launch_fc2_fwd(const float* a1, const float* W2, const float* b2, float* u, int rows50, DropCfg d,
                           cudaStream_t s) {
    if (rows50 <= 0) return cudaSuccess;
    fc2_fwd_kernel<<<(rows50 + F2_ROWS - 1) / F2_ROWS, TR_THREADS, 0, s>>>(a1, W2, b2, u, rows50, d);
    return cudaGetLastError();
}

cudaError_t launch_fc2_bwd(const float* du, const float* u, float* a1_dap, const float* W2, float* dW2, float* db2,
                           int rows50, float scale, int num_sms, cudaStream_t s) {
    if (rows50 <= 0) return cudaSuccess;
    const int ntiles = (rows50 + F2_ROWS - 1) / F2_ROWS;
    const int grid = ntiles < 2 * num_sms ? ntiles : 2 * num_sms;
    fc2_bwd_kernel<<<grid, TR_THREADS, 0, s>>>(du, u, a1_dap, W2, dW2, db2, rows50, scale);
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

cudaError_t launch_embed_grad(const float* dep, const uint8_t* x, float* dE, int nwin, DropCfg d, int num_sms,
                              cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int nbp = nwin * COLS;
    const int grid = nbp < 2 * num_sms ? nbp : 2 * num_sms;
    embed_grad_kernel<<<grid, TR_THREADS, 0, s>>>(dep, x, dE, nwin, d);
    return cudaGetLastError();
}

}  // namespace roko
