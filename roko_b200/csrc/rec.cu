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
#include "common.cuh"

namespace roko {

constexpr int HS_STRIDE = HID + 8;     // +8 floats: the 4 kq lanes' h stores land in distinct banks

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

template <int NB>
__global__ void __launch_bounds__(REC_THREADS, 1)
rec_kernel(const float* __restrict__ gi, const float* __restrict__ whh0, size_t dir_stride,
           const float* __restrict__ bhn0, float* __restrict__ out, int nwin) {
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
        const size_t row0 = (size_t)(b0 + (mine ? kq : 0)) * COLS;
        const float* gp = gi + row0 * GI_N + dir * G3 + j * 3;
        float* op = out + row0 * OUT_W + dir * HID + j;

        __syncthreads();                                     // previous group's readers are done
        for (int i = tid; i < 2 * NB * HS_STRIDE; i += REC_THREADS) (&hs[0][0][0])[i] = 0.f;
        float hprev = 0.f;
        int t = dir ? COLS - 1 : 0;
        const int dt = dir ? -1 : 1;
        float g_r = 0.f, g_z = 0.f, g_n = 0.f;
        if (mine) { const float* p = gp + (size_t)t * GI_N; g_r = p[0]; g_z = p[1]; g_n = p[2]; }
        __syncthreads();

        for (int s = 0; s < COLS; ++s) {
            const int cur = s & 1;
            float n_r = 0.f, n_z = 0.f, n_n = 0.f;           // prefetch gi of the next step
            if (mine && s + 1 < COLS) {
                const float* p = gp + (size_t)(t + dt) * GI_N;
                n_r = p[0]; n_z = p[1]; n_n = p[2];
            }
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
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    acc[b][g] += __shfl_xor_sync(0xffffffffu, acc[b][g], 1);
                    acc[b][g] += __shfl_xor_sync(0xffffffffu, acc[b][g], 2);
                }
            float a_r = acc[0][0], a_z = acc[0][1], a_n = acc[0][2];
#pragma unroll
            for (int b = 1; b < NB; ++b)
                if (kq == b) { a_r = acc[b][0]; a_z = acc[b][1]; a_n = acc[b][2]; }
            if (mine) {
                const float r = sigmoid_acc(g_r + a_r);
                const float z = sigmoid_acc(g_z + a_z);
                const float n = tanhf(g_n + r * (a_n + bhn));
                const float h = (1.f - z) * n + z * hprev;
                hprev = h;
                hs[cur ^ 1][kq][j] = h;
                op[(size_t)t * OUT_W] = h;
            }
            g_r = n_r; g_z = n_z; g_n = n_n;
            t += dt;
            __syncthreads();
        }
    }
}

cudaError_t rec_setup() { return cudaSuccess; }

cudaError_t launch_rec(const float* gi, const float* whh_d0, size_t dir_stride, const float* bhn_d0,
                       float* out, int nwin, int num_sms, cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    // CTAs come in (fwd, bwd) pairs; pick the largest group size that still fills the machine
    const int pairs = num_sms / 2;
    int nb = 1;
    if (nwin >= 2 * pairs) nb = 2;
    if (nwin >= 8 * pairs) nb = 4;
    const int ngroups = (nwin + nb - 1) / nb;
    const int grid = 2 * (ngroups < pairs ? ngroups : pairs);
    if (nb == 1) rec_kernel<1><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin);
    else if (nb == 2) rec_kernel<2><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin);
    else rec_kernel<4><<<grid, REC_THREADS, 0, s>>>(gi, whh_d0, dir_stride, bhn_d0, out, nwin);
    return cudaGetLastError();
}

}  // namespace roko
