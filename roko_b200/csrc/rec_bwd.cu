// Backward through time of one bidirectional GRU layer (what autograd derives for nn.GRU,
// reference roko/rnn_model.py:57 under roko/train.py:52).  With the forward's saved gates
// (r, z, n, q = W_hn h_prev + b_hn) and outputs, one step with incoming dh = dout_t + carry is
//     dn  = dh (1 - z)            dz  = dh (h_prev - n)           carry' = dh z + W_hh^T dgh
//     dnp = dn (1 - n^2)          dzp = dz z (1 - z)
//     drp = dnp q r (1 - r)       dq  = dnp r
//     dgi = (drp, dzp, dnp)       dgh = (drp, dzp, dq)
// Same persistent layout as the forward (rec.cu): a CTA is bound to one direction and keeps W_hh
// (transposed use: 384-long contraction per hidden unit) in registers, 512 threads x 96; thread
// (jp, k8) owns hidden units 2jp, 2jp+1 and the interleaved slice {32 i + 4 k8 + c} of the gate axis, so
// each dgh value read from shared memory feeds two FMAs (the forward gets three per h value); a
// reduce-scatter over the 8 slice lanes leaves every (window, unit) sum on the lane that finishes it.
// The steps run in the reverse of the forward's order.
//
// Outputs:  dgi      [row][768]  n = d*384 + g*128 + j   (feeds dW_ih, db_ih, dX)
//           dghn     [row][256]  the n-gate entry of dgh  (b_hh gradient)
//           dgh_prev [row][768]  dgh of the step whose h_prev is this row's output; zero rows where
//                                no step follows (feeds dW_hh = dgh_prev^T out)
#include "train.cuh"

namespace roko {

constexpr int DG_STRIDE = G3 + 8;      // +8 floats: the 4 kq lanes' stores land in distinct banks
int rec_pick_nb(int nwin, int num_sms);

template <int NB>
__global__ void __launch_bounds__(REC_THREADS, 1)
rec_bwd_kernel(const float* __restrict__ dout, const float4* __restrict__ gates, const float* __restrict__ out,
               const float* __restrict__ whh0, size_t dir_stride, float* __restrict__ dgi,
               float* __restrict__ dghn, float* __restrict__ dgh_prev, int nwin) {
    static_assert(NB == 1 || NB == 2 || NB == 4, "group size");
    __shared__ __align__(16) float ds[2][NB][DG_STRIDE];
    // thread (jp, k8): hidden units 2jp and 2jp+1, gate-axis slice {32 i + 4 k8 + c}: every dgh value it reads
    // from shared memory feeds two FMAs
    const int tid = threadIdx.x, jp = tid >> 3, k8 = tid & 7;
    const int dir = blockIdx.x & 1;
    const float* whh = whh0 + dir * dir_stride;            // raw W_hh of this direction: [384][128]

    float w0[48], w1[48];
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 p = *reinterpret_cast<const float2*>(whh + (32 * i + 4 * k8 + c) * HID + 2 * jp);
            w0[i * 4 + c] = p.x;
            w1[i * 4 + c] = p.y;
        }
    // after the reduce-scatter over the 8 slice lanes, lane k8 = (b2 b1 b0) holds one (window, unit) sum:
    //   NB = 4: window 2 b2 + b1, unit b0        NB = 2: window b2, unit b1 (b0 duplicates)
    //   NB = 1: unit b2 (b1, b0 duplicate)
    const int b2 = (k8 >> 2) & 1, b1 = (k8 >> 1) & 1, b0 = k8 & 1;
    const int wb = NB == 4 ? 2 * b2 + b1 : NB == 2 ? b2 : 0;
    const int o = NB == 4 ? b0 : NB == 2 ? b1 : b2;
    const bool role = NB == 4 ? true : NB == 2 ? b0 == 0 : (k8 & 3) == 0;
    const int j = 2 * jp + o;

    const int ngroups = (nwin + NB - 1) / NB;
    for (int grp = blockIdx.x >> 1; grp < ngroups; grp += gridDim.x >> 1) {
        const int b0w = grp * NB;
        const bool mine = role && (b0w + wb) < nwin;
        const int row0 = (b0w + (mine ? wb : 0)) * COLS;
        // forward direction d ran t = 0..89 (d = 0) or 89..0 (d = 1); walk it backwards
        int t = dir ? 0 : COLS - 1;
        const int dt = dir ? 1 : -1;                         // t of the next step of THIS loop; h_prev sits at t + dt

        __syncthreads();
        for (int i = tid; i < 2 * NB * DG_STRIDE; i += REC_THREADS) (&ds[0][0][0])[i] = 0.f;
        float carry = 0.f;
        float4 gt = make_float4(0.f, 0.f, 0.f, 0.f);
        float dov = 0.f, hp = 0.f;
        if (mine) {
            const unsigned row = row0 + t;
            gt = gates[(row * 2 + dir) * HID + j];
            dov = dout[row * OUT_W + dir * HID + j];
            hp = out[(row + dt) * OUT_W + dir * HID + j];    // COLS >= 2: the first loop step always has a predecessor
            // the row no step follows (forward's last step) contributes nothing to dW_hh
            dgh_prev[row * GI_N + dir * G3 + j] = 0.f;
            dgh_prev[row * GI_N + dir * G3 + HID + j] = 0.f;
            dgh_prev[row * GI_N + dir * G3 + 2 * HID + j] = 0.f;
        }
        __syncthreads();

        for (int s = 0; s < COLS; ++s) {
            const int nxt = s & 1;
            float dh = 0.f, zz = 0.f;
            if (mine) {
                const unsigned row = row0 + t;
                const float r = gt.x, z = gt.y, n = gt.z, q = gt.w;
                dh = carry + dov;
                zz = z;
                const float dn = dh * (1.f - z);
                const float dz = dh * (hp - n);
                const float dnp = dn * (1.f - n * n);
                const float dzp = dz * z * (1.f - z);
                const float drp = dnp * q * r * (1.f - r);
                const float dq = dnp * r;
                float* gi_row = dgi + row * GI_N + dir * G3 + j;
                gi_row[0] = drp; gi_row[HID] = dzp; gi_row[2 * HID] = dnp;
                dghn[row * OUT_W + dir * HID + j] = dq;
                if (s + 1 < COLS) {
                    float* gp = dgh_prev + (row + dt) * GI_N + dir * G3 + j;
                    gp[0] = drp; gp[HID] = dzp; gp[2 * HID] = dq;
                }
                ds[nxt][wb][j] = drp; ds[nxt][wb][HID + j] = dzp; ds[nxt][wb][2 * HID + j] = dq;
                // a full step ahead: operands of the next loop step
                if (s + 1 < COLS) {
                    const unsigned rn = row + dt;
                    gt = gates[(rn * 2 + dir) * HID + j];
                    dov = dout[rn * OUT_W + dir * HID + j];
                    hp = (s + 2 < COLS) ? out[(rn + dt) * OUT_W + dir * HID + j] : 0.f;
                }
            }
            __syncthreads();
            float v[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b) v[b][0] = v[b][1] = 0.f;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float4 dv = *reinterpret_cast<const float4*>(&ds[nxt][b][32 * i + 4 * k8]);
                    v[b][0] = fmaf(w0[i * 4 + 0], dv.x, v[b][0]); v[b][1] = fmaf(w1[i * 4 + 0], dv.x, v[b][1]);
                    v[b][0] = fmaf(w0[i * 4 + 1], dv.y, v[b][0]); v[b][1] = fmaf(w1[i * 4 + 1], dv.y, v[b][1]);
                    v[b][0] = fmaf(w0[i * 4 + 2], dv.z, v[b][0]); v[b][1] = fmaf(w1[i * 4 + 2], dv.z, v[b][1]);
                    v[b][0] = fmaf(w0[i * 4 + 3], dv.w, v[b][0]); v[b][1] = fmaf(w1[i * 4 + 3], dv.w, v[b][1]);
                }
            }
            // reduce-scatter over the 8 slice lanes (xor 4, 2, 1): each stage keeps the half this lane finishes
            float a;
            if (NB == 4) {
                float k1[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int oo = 0; oo < 2; ++oo) {
                        const float keep = b2 ? v[2 + h][oo] : v[h][oo], send = b2 ? v[h][oo] : v[2 + h][oo];
                        k1[h][oo] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                float k2[2];
#pragma unroll
                for (int oo = 0; oo < 2; ++oo) {
                    const float keep = b1 ? k1[1][oo] : k1[0][oo], send = b1 ? k1[0][oo] : k1[1][oo];
                    k2[oo] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                a = (b0 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? k2[0] : k2[1], 1);
            } else if (NB == 2) {
                float k1[2];
#pragma unroll
                for (int oo = 0; oo < 2; ++oo) {
                    const float keep = b2 ? v[NB - 1][oo] : v[0][oo], send = b2 ? v[0][oo] : v[NB - 1][oo];
                    k1[oo] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                a = (b1 ? k1[1] : k1[0]) + __shfl_xor_sync(0xffffffffu, b1 ? k1[0] : k1[1], 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
            } else {
                a = (b2 ? v[0][1] : v[0][0]) + __shfl_xor_sync(0xffffffffu, b2 ? v[0][0] : v[0][1], 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
            }
            carry = fmaf(dh, zz, a);
            t += dt;
        }
    }
}

cudaError_t rec_bwd_setup() { return cudaSuccess; }

cudaError_t launch_rec_bwd(const float* dout, const float4* gates, const float* out, const float* whh_raw_d0,
                           size_t raw_dir_stride, float* dgi, float* dghn, float* dgh_prev, int nwin, int num_sms,
                           cudaStream_t s) {
    if (nwin <= 0) return cudaSuccess;
    const int pairs = num_sms / 2;
    const int nb = rec_pick_nb(nwin, num_sms);
    const int ngroups = (nwin + nb - 1) / nb;
    const int grid = 2 * (ngroups < pairs ? ngroups : pairs);
    if (nb == 1) rec_bwd_kernel<1><<<grid, REC_THREADS, 0, s>>>(dout, gates, out, whh_raw_d0, raw_dir_stride, dgi, dghn, dgh_prev, nwin);
    else if (nb == 2) rec_bwd_kernel<2><<<grid, REC_THREADS, 0, s>>>(dout, gates, out, whh_raw_d0, raw_dir_stride, dgi, dghn, dgh_prev, nwin);
    else rec_bwd_kernel<4><<<grid, REC_THREADS, 0, s>>>(dout, gates, out, whh_raw_d0, raw_dir_stride, dgi, dghn, dgh_prev, nwin);
    return cudaGetLastError();
}

}  // namespace roko
