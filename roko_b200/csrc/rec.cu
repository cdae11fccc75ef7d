// Recurrent half of one bidirectional GRU layer (reference roko/rnn_model.py:57; gate math
// SURVEY.md App. B.3):
//     r = sigmoid(gi_r + W_hr h)          (b_ir + b_hr already folded into gi by the projection)
//     z = sigmoid(gi_z + W_hz h)
//     n = tanh   (gi_n + r * (W_hn h + b_hn))
//     h = (1 - z) * n + z * h
// 90 dependent steps per direction.  Persistent design: a CTA is bound to one direction and keeps
// that direction's whole W_hh (384 x 128 fp32 = 192 KB) in REGISTERS for its lifetime -- 512
// threads x 96 registers.  Thread (j, kq) owns hidden unit j's three gate rows restricted to the
// interleaved k-slice {16 i + 4 kq + q}; per step it reads its slice of h from shared memory as
// 8 conflict-free LDS.128 per window, issues 96 FFMA per window, and a 2-stage shuffle butterfly
// over the 4 kq lanes finishes the dot products.  Lane kq then owns window kq of the group: gate
// math, h kept in a register, new h to shared (double buffered -> one barrier per step) and to
// the layer output.  gi for step s+1 is prefetched while step s computes.
#include <stdlib.h>

#include "train.cuh"

namespace roko {

constexpr int HS_STRIDE = HID + 8;     // +8 floats: the 4 kq lanes' h stores land in distinct banks

// sigmoid / tanh on the MUFU pipe: ex2.approx (<= 2 ulp) + rcp.approx (<= 1 ulp); absolute error
// ~3e-7, the size of fp32 rounding noise in the reference's own CPU evaluation.
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
#ifdef ROKO_ACCURATE_GATES
__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float tanh_f(float v) { return tanhf(v); }
#else
__device__ __forceinline__ float sigmoid_f(float v) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * v)); }
__device__ __forceinline__ float tanh_f(float v) {
    return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.8853900817779268f * v)), -1.f);
}
#endif

// SAVE (training forward): also file (r, z, n, W_hn h + b_hn) per (row, direction, unit) for the backward.
template <int NB, bool SAVE>
__global__ void __launch_bounds__(REC_THREADS, 1)
rec_kernel(const float* __restrict__ gi, const float* __restrict__ whh0, size_t dir_stride,
           const float* __restrict__ bhn0, float* __restrict__ out, int nwin, float4* __restrict__ gates) {
    static_assert(NB == 1 || NB == 2 || NB == 4, "group size");
    __shared__ __align__(16) float hs[2][NB][HS_STRIDE];
    const int tid = threadIdx.x, j = tid >> 2, kq = tid & 3;
    const int dir = blockIdx.x & 1;
    const float* whh = whh0 + dir * dir_stride;

    float w[WHH_REGS];
#pragma unroll
    for (int i = 0; i < WHH_REGS; ++i) w[i] = whh[i * REC_THREADS + tid];
    const float bhn = (bhn0 + dir * dir_stride)[j];

    const int ngroups = (nwin + NB - 1) / NB;
    for (int grp = blockIdx.x >> 1; grp < ngroups; grp += gridDim.x >> 1) {
        const int b0 = grp * NB;
        const bool mine = kq < NB && (b0 + kq) < nwin;      // this lane finishes window b0+kq
        const int row0 = (b0 + (mine ? kq : 0)) * COLS;
        int t = dir ? COLS - 1 : 0;
        const int dt = dir ? -1 : 1;
        // 32-bit element offsets (a chunk is far below 2^31 floats): fewer live registers
        unsigned gofs = (unsigned)(row0 + t) * GI_N + dir * G3 + j * 3;
        unsigned oofs = (unsigned)(row0 + t) * OUT_W + dir * HID + j;
        unsigned sofs = ((unsigned)(row0 + t) * 2 + dir) * HID + j;
        const int gstep = dt * GI_N, ostep = dt * OUT_W;

        __syncthreads();                                     // previous group's readers are done
        for (int i = tid; i < 2 * NB * HS_STRIDE; i += REC_THREADS) (&hs[0][0][0])[i] = 0.f;
        float hprev = 0.f;
        float g_r = 0.f, g_z = 0.f, g_n = 0.f;               // gi of the step about to run
        if (mine) { g_r = gi[gofs]; g_z = gi[gofs + 1]; g_n = gi[gofs + 2]; }
        __syncthreads();

        for (int s = 0; s < COLS; ++s) {
            const int cur = s & 1;
            float acc[NB][3];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float4 hv = *reinterpret_cast<const float4*>(&hs[cur][b][16 * i + 4 * kq]);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 0], hv.x, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 1], hv.y, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 2], hv.z, acc[b][g]);
                        acc[b][g] = fmaf(w[g * 32 + i * 4 + 3], hv.w, acc[b][g]);
                    }
                }
            }
            // finish the dot products over the 4 kq lanes; lane kq ends up with window kq's sums
            float a_r, a_z, a_n;
            if (NB == 4) {
                // reduce-scatter: 9 shuffles instead of a 24-shuffle all-reduce
                const bool hi1 = kq & 1, hi2 = kq & 2;
                float k0[3], k1[3];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    // stage 1 (xor 1): keep windows of my parity {kq&1, 2+(kq&1)}, send the others
                    const float s0 = hi1 ? acc[0][g] : acc[1][g];
                    const float s1 = hi1 ? acc[2][g] : acc[3][g];
                    k0[g] = (hi1 ? acc[1][g] : acc[0][g]) + __shfl_xor_sync(0xffffffffu, s0, 1);
                    k1[g] = (hi1 ? acc[3][g] : acc[2][g]) + __shfl_xor_sync(0xffffffffu, s1, 1);
                }
                // stage 2 (xor 2): keep window kq = (kq&1) + 2*(kq>>1)
                {
                    const float sr = hi2 ? k0[0] : k1[0], sz = hi2 ? k0[1] : k1[1], sn = hi2 ? k0[2] : k1[2];
                    a_r = (hi2 ? k1[0] : k0[0]) + __shfl_xor_sync(0xffffffffu, sr, 2);
                    a_z = (hi2 ? k1[1] : k0[1]) + __shfl_xor_sync(0xffffffffu, sz, 2);
                    a_n = (hi2 ? k1[2] : k0[2]) + __shfl_xor_sync(0xffffffffu, sn, 2);
                }
            } else {
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[b][g] += __shfl_xor_sync(0xffffffffu, acc[b][g], 1);
                        acc[b][g] += __shfl_xor_sync(0xffffffffu, acc[b][g], 2);
                    }
                a_r = acc[0][0]; a_z = acc[0][1]; a_n = acc[0][2];
                if (NB == 2 && kq == 1) { a_r = acc[NB - 1][0]; a_z = acc[NB - 1][1]; a_n = acc[NB - 1][2]; }
            }
            if (mine) {
                const float r = sigmoid_f(g_r + a_r);
                const float z = sigmoid_f(g_z + a_z);
                const float n = tanh_f(g_n + r * (a_n + bhn));
                const float h = fmaf(z, hprev - n, n);       // (1 - z) * n + z * h
                hprev = h;
                hs[cur ^ 1][kq][j] = h;
                out[oofs] = h;
                if (SAVE) { gates[sofs] = make_float4(r, z, n, a_n + bhn); sofs += dt * 2 * HID; }
                gofs += gstep; oofs += ostep;
                if (s + 1 < COLS) { g_r = gi[gofs]; g_z = gi[gofs + 1]; g_n = gi[gofs + 2]; }   // a full step ahead
            }
            __syncthreads();
        }
    }
}

constexpr double REC_C0 = 600.0, REC_C1 = 450.0;
static int g_force_nb = 0;

cudaError_t rec_setup() {
    const char* e = getenv("ROKO_B200_REC_NB");              // tuning / profiling override
    g_force_nb = e ? atoi(e) : 0;
    return cudaSuccess;
}

cudaError_t launch_rec(const float* gi, const float* whh_d0, size_t dir_stride, const float* bhn_d0,
                       float* out, int nwin, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    // CTAs come in (fwd, bwd) pairs; pick the largest group size that still fills the machine
    // cost model (cycles per step ~ C0 + C1*nb, measured) x rounds each CTA pair has to run
    const int pairs = num_sms / 2;
    int nb = 1;
    double best = 1e30;
    for (int cand = 1; cand <= 4; cand *= 2) {
        const int groups = (nwin + cand - 1) / cand;
        const int rounds = (groups + pairs - 1) / pairs;
        const double cost = rounds * (REC_C0 + REC_C1 * cand);
        if (cost < best) { best = cost; nb = cand; }
    }
    if (g_force_nb == 1 || g_force_nb == 2 || g_force_nb == 4) nb = g_force_nb;
    const int ngroups = (nwin + nb - 1) / nb;
    const int grid = 2 * (ngroups < pairs ? ngroups : pairs);
    if (nb == 1) rec_kernel<1, false><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, nullptr);
    else if (nb == 2) rec_kernel<2, false><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, nullptr);
    else rec_kernel<4, false><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, nullptr);
    return cudaGetLastError();
}

int rec_pick_nb(int nwin, int num_sms) {
    if (g_force_nb == 1 || g_force_nb == 2 || g_force_nb == 4) return g_force_nb;
    const int pairs = num_sms / 2;
    int nb = 1;
    double best = 1e30;
    for (int cand = 1; cand <= 4; cand *= 2) {
        const int groups = (nwin + cand - 1) / cand;
        const int rounds = (groups + pairs - 1) / pairs;
        const double cost = rounds * (REC_C0 + REC_C1 * cand);
        if (cost < best) { best = cost; nb = cand; }
    }
    return nb;
}

cudaError_t launch_rec_train(const float* gi, const float* whh_d0, size_t dir_stride, const float* bhn_d0,
                             float* out, float4* gates, int nwin, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int pairs = num_sms / 2;
    const int nb = rec_pick_nb(nwin, num_sms);
    const int ngroups = (nwin + nb - 1) / nb;
    const int grid = 2 * (ngroups < pairs ? ngroups : pairs);
    if (nb == 1) rec_kernel<1, true><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, gates);
    else if (nb == 2) rec_kernel<2, true><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, gates);
    else rec_kernel<4, true><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin, gates);
    return cudaGetLastError();
}

}  // namespace roko
