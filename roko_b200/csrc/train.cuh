// Declarations shared by the training-path translation units (gemm.cu, train.cu, rec_bwd.cu,
// train_api.cu).  The training forward is the reference's train-mode forward
// (roko/rnn_model.py:46-59 with the four dropout sites :29,:32,:35,:41 active) and its backward is
// what autograd derives for roko/train.py:46-53.
#pragma once
#include "common.cuh"

namespace roko {

// ---- dropout: counter-based masks, reproducible from (seed, site, element index) ---------------
// Site element indices follow the reference tensors' own row-major order:
//   DROP_EMB  (B,200,90,50)   DROP_FC1 (B,90,50,100)   DROP_FC2 (B,90,50,10)   DROP_GRU0/1 (B,90,256)
enum { DROP_EMB = 0, DROP_FC1 = 1, DROP_FC2 = 2, DROP_GRU0 = 3, DROP_GRU1 = 4, DROP_SITES = 5 };

struct DropCfg {
    unsigned long long seed;
    unsigned int thresh;     // drop when hash < thresh;  thresh = round(p * 2^32), 0 = keep everything
    float scale;             // 1 / (1 - p)
};

__host__ __device__ __forceinline__ unsigned int drop_hash(unsigned long long seed, unsigned int site,
                                                           unsigned long long idx) {
    // stream keys (k1, k2): splitmix64 finaliser of (seed, site) -- uniform per launch, hoisted out of every loop
    unsigned long long z = seed ^ ((unsigned long long)(site + 1) * 0xD1B54A32D192ED03ull);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    // per element: a 32-bit multiply-xorshift finaliser of (idx + k1), whitened with k2.  32-bit arithmetic only (9
    // instructions against ~45 for a 64-bit splitmix per element, which made the fc1 epilogue and the embedding pass
    // hash-bound); element indices are < 2^32 for every site up to the 1024-window batch limit.
    unsigned int x = (unsigned int)idx + (unsigned int)(z >> 32);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x ^ (unsigned int)z;
}
__host__ __device__ __forceinline__ bool drop_keep(const DropCfg& d, unsigned int site, unsigned long long idx) {
    return d.thresh == 0u || drop_hash(d.seed, site, idx) >= d.thresh;
}

// ---- general GEMM (gemm.cu) --------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_ACC = 1, EPI_ATOMIC = 2, EPI_FC1 = 3 };
struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    const float* bias;       // EPI_FC1
    DropCfg drop;            // EPI_FC1
    int kchunk;              // filled by the launcher
    bool vecA, vecB;         // filled by the launcher
};
cudaError_t launch_gemm(GemmArgs g, bool ka, bool kb, int epi, int splits, int num_sms, cudaStream_t s);
cudaError_t gemm_setup();

// ---- element-wise / small kernels (train.cu) -----------------------------------------------------
constexpr int MASKT_WORDS = 2 * READS;   // transposed keep bits of one (window, column): 64 bits (50 used) per read
constexpr int MASK_WORDS = 320;          // keep bits of one (window, column)'s 10 000 embedding outputs (313 words used)
// ep may be NULL (keep bits and codes only); xt [nwin*90][200] receives the validated codes of each (window, column) and
// bitsT [nwin*90][200][2] the keep bits transposed (per read, bit = channel), or NULL
cudaError_t launch_embed_drop(const uint8_t* x, const float* E, float* ep, uint32_t* bits, uint8_t* xt, uint32_t* bitsT, int nwin,
                              DropCfg d, int* status, cudaStream_t s);
cudaError_t launch_fc2_fwd(const float* a1, const float* W2, const float* b2, float* u, int rows50, DropCfg d,
                           cudaStream_t s);
// a1_dap: a1 in, d(fc1 pre-activation) out (in place); db1 += its column sums
cudaError_t launch_fc2_bwd(const float* du, const float* u, float* a1_dap, const float* W2, float* dW2, float* db2, float* db1,
                           int rows50, float scale, int num_sms, cudaStream_t s);
cudaError_t launch_drop_apply(const float* in, float* out, size_t n, unsigned int site, DropCfg d, cudaStream_t s);
cudaError_t launch_colsum(const float* A, int lda, int rows, int ncols, float* out, cudaStream_t s);
cudaError_t launch_embed_grad(const float* dep, const uint8_t* x, const uint32_t* bits, float* dE, int nwin,
                              float scale, int num_sms, cudaStream_t s);
cudaError_t launch_gru_bias_grad(const float* dgi, const float* dghn, int rows, float* bih0, float* bhh0,
                                 float* bih1, float* bhh1, cudaStream_t s);
cudaError_t launch_drop_mask(unsigned int site, size_t n, uint8_t* out, DropCfg d, cudaStream_t s);

// ---- row-streaming front-end products on tcgen05 (train_tc.cu) -------------------------------------
// The masked embedding ep[(bp, e)][r] = keep ? E[xt[bp][r]][e] * scale : 0 takes 600 distinct values per seed: instead of
// storing its 0.46 GB per 128 windows, the products that consume it rebuild their operand tiles from the codes (1 byte per
// 50 values) and the keep bits (bit e*200 + r of the (window, column)'s MASK_WORDS words).
struct EpGen {
    const uint8_t* xt;
    const uint32_t* bits;
    const uint32_t* bitsT;   // the same keep bits filed per read: [window*90 + column][200][2 words], bit e of the pair
    const float* E;          // [12][50] embedding table (fp32, as loaded)
    float scale;             // 1 / (1 - p)
    float* dE;               // embedding gradient [12][50] (fused d(ep) product only)
};
size_t train_tc_image_floats();
cudaError_t train_tc_setup();
cudaError_t launch_train_images(const float* raw, float* img, cudaStream_t s);
cudaError_t launch_din_tc(int l, const float* dgi, const float* img, float* din, int rows, int num_sms, cudaStream_t s);
cudaError_t launch_tn_tc(const float* A, int lda, int Mreal, const float* B, int ldb, int Nreal, float* C, int ldc,
                         int rows, int num_sms, cudaStream_t s);
struct TnTile;      // train_tc.cu: one 128 x <= 256 tile of a C += A^T B product
// the weight gradients of GRU layer l (dW_ih, dW_hh, both directions) in one launch
cudaError_t launch_gru_dw(int l, const float* dgi, const float* in, const float* dghp, const float* out, float* grad_raw,
                          int rows, int num_sms, cudaStream_t s);
cudaError_t launch_fc1_tc(const float* ep, const float* img, const float* b1, float* a1, int rows, DropCfg d,
                          int num_sms, cudaStream_t s);
cudaError_t launch_dep_tc(const float* dap, const float* img, float* dep, int rows, int num_sms, cudaStream_t s);
cudaError_t launch_dw1_tc(const float* dap, const float* ep, float* dW1, int rows, int num_sms, cudaStream_t s);
// the same three products with ep regenerated from (codes, keep bits); the d(ep) product reduces straight into dE
cudaError_t launch_fc1_gen(EpGen gen, const float* img, const float* b1, float* a1, int rows, DropCfg d, int num_sms, cudaStream_t s);
cudaError_t launch_dw1_gen(const float* dap, EpGen gen, float* dW1, int rows, int num_sms, cudaStream_t s);
cudaError_t launch_dep_de(const float* dap, const float* img, EpGen gen, int rows, int num_sms, cudaStream_t s);

// ---- recurrence (rec.cu forward with gate saving, rec_bwd.cu) -------------------------------------
// gates: [row][dir][j] float4 (r, z, n, W_hn h + b_hn)
cudaError_t launch_rec_train(const float* gi, const float* whh_d0, size_t dir_stride, const float* bhn_d0,
                             float* out, float4* gates, int nwin, int num_sms, cudaStream_t s);
// dgi [row][768] (n = d*384 + g*128 + j), dghn [row][256], dgh_prev [row][768] (the recurrent-side
// gate gradients filed under the row whose output was that step's h_prev; zero where there is none)
cudaError_t launch_rec_bwd(const float* dout, const float4* gates, const float* out, const float* whh_raw_d0,
                           size_t raw_dir_stride, float* dgi, float* dghn, float* dgh_prev, int nwin, int num_sms,
                           cudaStream_t s);
cudaError_t rec_bwd_setup();

}  // namespace roko
