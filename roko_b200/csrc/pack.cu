// Weight packing: raw state_dict order -> the layouts the kernels read.
// One thread per packed float; runs once per load_state_dict (reference inference.py:95).
#include "common.cuh"
#include "tc.cuh"

namespace roko {

__device__ __forceinline__ float pack_one(const float* __restrict__ raw, int p, int* __restrict__ status) {
    if (p < PK_W1T) return p < NCODES * EMB ? raw[RAW_E + p] : 0.f;
    if (p < PK_B1) {                                   // W1T[r][j] = fc1.weight[j][r]; row 200 = 0
        int i = p - PK_W1T;
        if (i >= W1T_ROWS * FC1) return 0.f;
        int r = i / FC1, j = i % FC1;
        return r < READS ? raw[RAW_W1 + j * READS + r] : 0.f;
    }
    if (p < PK_W2) { int i = p - PK_B1; return i < FC1 ? raw[RAW_B1 + i] : 0.f; }
    if (p < PK_B2) { int i = p - PK_W2; return i < FC2 * FC1 ? raw[RAW_W2 + i] : 0.f; }
    if (p < PK_GRU) { int i = p - PK_B2; return i < FC2 ? raw[RAW_B2 + i] : 0.f; }
    if (p < PK_W4) {
        int l = 0;
        while (l + 1 < LAYERS && p >= pk_layer(l + 1)) ++l;
        const int kin = gru_in(l), kp = gru_inp(l);
        if (p < pk_bgi(l)) {                           // WIH[n][k], n = d*384 + j*3 + g
            int i = p - pk_wih(l);
            if (i >= GI_N * kp) return 0.f;
            int n = i / kp, k = i % kp;
            int d = n / G3, j = (n % G3) / 3, g = (n % G3) % 3;
            return k < kin ? raw[raw_wih(l, d) + (g * HID + j) * kin + k] : 0.f;
        }
        if (p < pk_whh(l, 0)) {                        // BGI[n] = b_ih + (r,z: b_hh)
            int n = p - pk_bgi(l);
            if (n >= GI_N) return 0.f;
            int d = n / G3, j = (n % G3) / 3, g = (n % G3) % 3;
            float v = raw[raw_bih(l, d) + g * HID + j];
            if (g < 2) v += raw[raw_bhh(l, d) + g * HID + j];
            return v;
        }
        int d = p >= pk_whh(l, 1) ? 1 : 0;
        if (p < pk_bhn(l, d)) {                        // WHH register image [idx][tid]
            int i = p - pk_whh(l, d);
            if (i >= WHH_REGS * REC_THREADS) return 0.f;
            int idx = i / REC_THREADS, tid = i % REC_THREADS;
            int j = tid >> 2, kq = tid & 3;
            int g = idx / 32, rem = idx % 32, ii = rem / 4, q = rem % 4;
            int k = 16 * ii + 4 * kq + q;
            return raw[raw_whh(l, d) + (g * HID + j) * HID + k];
        }
        int j = p - pk_bhn(l, d);
        return j < HID ? raw[raw_bhh(l, d) + 2 * HID + j] : 0.f;
    }
    if (p < PK_B4) { int i = p - PK_W4; return i < CLASSES * OUT_W ? raw[RAW_W4 + i] : 0.f; }
    if (p < pk_wtc(0)) { int i = p - PK_B4; return i < CLASSES ? raw[RAW_B4 + i] : 0.f; }
    if (p >= pk_wh16(0)) {     // fp16-split operands of the default kernels (tc.cuh): two halves per packed word
        auto pair = [&](float v0, float v1, bool lo) -> float {
            v0 *= tc::W_SCALE; v1 *= tc::W_SCALE;
            if (fabsf(v0) > 65000.f || fabsf(v1) > 65000.f) atomicOr(status, 2);       // outside fp16's range: see roko_b200_model_check
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half a = lo ? __float2half_rn(v0 - __half2float(h0)) : h0;
            const __half b = lo ? __float2half_rn(v1 - __half2float(h1)) : h1;
            return __uint_as_float((uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16));
        };
        if (p >= PK_FT_W1HI) {                         // front_tc.cu operands (scales: W1 x 64, W2 x 256)
            auto pair_s = [&](float v0, float v1, float scale, bool lo) -> float {
                v0 *= scale; v1 *= scale;
                if (fabsf(v0) > 65000.f || fabsf(v1) > 65000.f) atomicOr(status, 2);
                const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                const __half a = lo ? __float2half_rn(v0 - __half2float(h0)) : h0;
                const __half b = lo ? __float2half_rn(v1 - __half2float(h1)) : h1;
                return __uint_as_float((uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16));
            };
            auto w1 = [&](int j, int r) { return (j < FC1 && r < READS) ? raw[RAW_W1 + j * READS + r] : 0.f; };
            if (p < PK_FT_W1LO) {                      // [j][word c]: r = 2c, 2c+1, hi halves
                const int i = p - PK_FT_W1HI, j = i / (FT_K1 / 2), c = i % (FT_K1 / 2);
                return pair_s(w1(j, 2 * c), w1(j, 2 * c + 1), 64.f, false);
            }
            if (p < PK_FT_W2) {                        // [k atom][row j][64 B], SWIZZLE_64B: chunk position = chunk ^ ((row >> 1) & 3)
                const int ob = (p - PK_FT_W1LO) * 4;
                const int atom = ob / (128 * 64), within = ob % (128 * 64);
                const int j = (within / 512) * 8 + (within % 512) / 64;
                const int pchunk = (within % 64) / 16, w4 = (within % 16) / 4;
                const int r = atom * 32 + ((pchunk ^ ((j >> 1) & 3)) * 8) + w4 * 2;
                return pair_s(w1(j, r), w1(j, r + 1), 64.f, true);
            }
            {                                          // W2: [k atom of 64 j][32 rows: hi of k = row (0..15), lo of k = row - 16][128 B], SWIZZLE_128B; j = 100 carries b2
                const int ob = (p - PK_FT_W2) * 4;
                const int atom = ob / (32 * 128), within = ob % (32 * 128);
                const int row = (within / 1024) * 8 + (within % 1024) / 128;
                const int pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
                const int j = atom * 64 + ((pchunk ^ (row & 7)) * 8) + w4 * 2;
                const int k = row & 15;
                auto w2 = [&](int kk, int jj) {
                    if (kk >= FC2) return 0.f;
                    if (jj < FC1) return raw[RAW_W2 + kk * FC1 + jj];
                    return jj == FC1 ? raw[RAW_B2 + kk] : 0.f;
                };
                return pair_s(w2(k, j), w2(k, j + 1), 256.f, row >= 16);
            }
        }
        if (p >= pk_rh16(0, 0)) {                      // rec_h.cu: W_hh hi (tensor-memory rows), lo (swizzled shared-memory image), b_hn
            int i = p - pk_rh16(0, 0);
            const int ld = i / RH16_DIR, l = ld / 2, d = ld % 2;
            i %= RH16_DIR;
            if (i >= RH16_W) return raw[raw_bhh(l, d) + 2 * HID + (i - RH16_W)];
            const bool lo = i >= RH16_W / 2;
            i %= RH16_W / 2;
            const int mt = i / (HID * (HID / 2));
            i %= HID * (HID / 2);
            int row, k;
            if (!lo) { row = i / (HID / 2); k = 2 * (i % (HID / 2)); }
            else {
                const int ob = i * 4, katom = ob / 16384, within = ob % 16384;
                row = (within / 1024) * 8 + (within % 1024) / 128;
                const int pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
                k = katom * 64 + ((pchunk ^ (row & 7)) * 8) + w4 * 2;
            }
            const float* w = raw + raw_whh(l, d) + (mt * HID + row) * HID + k;
            return pair(w[0], w[1], lo);
        }
        int l = 0;                                     // proj_h.cu: W_ih images, 128-byte swizzle, 64-element k blocks
        while (l + 1 < LAYERS && p >= pk_wh16(l + 1)) ++l;
        const int kin = gru_in(l), kblocks = gru_inp(l) / H16_BK;
        int i = p - pk_wh16(l);
        const int ntile = i / (kblocks * 2 * H16_IMG);
        i %= kblocks * 2 * H16_IMG;
        const int kb = i / (2 * H16_IMG);
        i %= 2 * H16_IMG;
        const bool lo = i >= H16_IMG;
        const int ob = (i % H16_IMG) * 4;                // byte offset inside the 32 KB image
        const int rgrp = ob / 1024, within = ob % 1024;
        const int r8 = within / 128, pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
        const int k = kb * H16_BK + ((pchunk ^ r8) * 8) + w4 * 2;
        const int n = ntile * TC_BN + rgrp * 8 + r8;
        const int d = n / G3, j = (n % G3) / 3, g = (n % G3) % 3;
        const float* w = raw + raw_wih(l, d) + (g * HID + j) * kin;
        return pair(k < kin ? w[k] : 0.f, k + 1 < kin ? w[k + 1] : 0.f, lo);
    }
    if (p >= pk_rtc(0, 0)) {   // tensor-core recurrence operands (rec_tc.cu)
        int i = p - pk_rtc(0, 0);
        const int ld = i / RTC_DIR, l = ld / 2, d = ld % 2;
        i %= RTC_DIR;
        if (i >= 2 * RTC_W) return raw[raw_bhh(l, d) + 2 * HID + (i - 2 * RTC_W)];      // b_hn
        int row, k;
        const bool lo = i >= RTC_W;
        if (!lo) { row = i / HID; k = i % HID; }
        else {
            const int ob = (i - RTC_W) * 4;
            const int mt = ob / 65536, rem = ob % 65536, katom = rem / 16384, rem2 = rem % 16384;
            const int rgrp = rem2 / 1024, within = rem2 % 1024;
            const int r8 = within / 128, pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
            row = mt * HID + rgrp * 8 + r8;
            k = katom * 32 + ((pchunk ^ r8) * 4) + w4;
        }
        const float v = raw[raw_whh(l, d) + row * HID + k];
        unsigned hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        return lo ? v - hi : hi;
    }
    {   // tf32 hi/lo images of W_ih in the swizzled shared-memory layout of proj_tc.cu
        int l = 0;
        while (l + 1 < LAYERS && p >= pk_wtc(l + 1)) ++l;
        const int kin = gru_in(l), kblocks = gru_inp(l) / TC_BK;
        int i = p - pk_wtc(l);
        const int ntile = i / (kblocks * 2 * TC_IMG);
        i %= kblocks * 2 * TC_IMG;
        const int kb = i / (2 * TC_IMG);
        i %= 2 * TC_IMG;
        const int half = i / TC_IMG;
        const int ob = (i % TC_IMG) * 4;                 // byte offset inside the 32 KB image
        const int rgrp = ob / 1024, within = ob % 1024;
        const int r8 = within / 128, pchunk = (within % 128) / 16, w4 = (within % 16) / 4;
        const int r = rgrp * 8 + r8;
        const int kk = ((pchunk ^ r8) * 4) + w4;         // undo the 128B swizzle: logical 16B chunk
        const int k = kb * TC_BK + kk;
        const int n = ntile * TC_BN + r;
        const int d = n / G3, j = (n % G3) / 3, g = (n % G3) % 3;
        const float v = k < kin ? raw[raw_wih(l, d) + (g * HID + j) * kin + k] : 0.f;
        unsigned hi_bits;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_bits) : "f"(v));
        const float hi = __uint_as_float(hi_bits);
        return half == 0 ? hi : v - hi;
    }
}

__global__ void pack_kernel(const float* __restrict__ raw, float* __restrict__ packed, int* __restrict__ status) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < PK_TOTAL) packed[p] = pack_one(raw, p, status);
}

// int64 codes (what the reference caller passes, inference.py:113) -> uint8; flags codes > 11
__global__ void narrow_i64_kernel(const long long* __restrict__ x64, uint8_t* __restrict__ x8, size_t n,
                                  int* __restrict__ status) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) {
        long long v = x64[i];
        bad |= (v < 0 || v >= NCODES);
        x8[i] = (uint8_t)(v < 0 || v > 255 ? 255 : v);
    }
    if (bad) atomicOr(status, 1);
}

// FP32 FFMA peak: 16 independent accumulator chains per thread, no memory traffic.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* sink, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) sink[0] = s;
}

cudaError_t measure_fp32_peak(double* tflops) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    float* sink = nullptr;
    e = cudaMalloc(&sink, sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int iters = 1 << 14, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(t0);
        ffma_peak_kernel<<<blocks, 256>>>(sink, 0.999f, 0.001f, iters);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        double tf = 2.0 * 16.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    *tflops = best;
    return e;
}

cudaError_t launch_pack(const float* raw, float* packed, int* status, cudaStream_t s) {
    pack_kernel<<<(PK_TOTAL + 255) / 256, 256, 0, s>>>(raw, packed, status);
    return cudaGetLastError();
}

cudaError_t launch_narrow_i64(const long long* x64, uint8_t* x8, size_t n, int* status, cudaStream_t s) {
    int blocks = (int)((n + 1023) / 1024);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    narrow_i64_kernel<<<blocks, 256, 0, s>>>(x64, x8, n, status);
    return cudaGetLastError();
}

}  // namespace roko
