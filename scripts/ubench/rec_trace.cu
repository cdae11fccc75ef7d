// Timeline probe for the tensor-core recurrence (rec_h.cu): clock64 stamps of CTA 0 for steps 20..23.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DROKO_TRACE -o scripts/ubench/rec_trace scripts/ubench/rec_trace.cu
// slots per step: 0 gate thread 0: r tile ready   1 r sigmoid done   2 n tile ready   3 h written + arrived
//                 4 MMA warp: h ready, issue starts  5 MMA warp: 48 MMAs issued + both commits
// (the ping-pong variant this probe was first written for lives in git history: commit df7c646, rec_h2_kernel)
#include <stdio.h>
#include "../../roko_b200/csrc/rec_h.cu"

int main() {
    using namespace roko;
    const int nwin = 64;
    float *gi, *w, *out;
    cudaMalloc(&gi, (size_t)nwin * COLS * GI_N * 4); cudaMemset(gi, 0, (size_t)nwin * COLS * GI_N * 4);
    cudaMalloc(&w, (size_t)2 * RH16_DIR * 4); cudaMemset(w, 0, (size_t)2 * RH16_DIR * 4);
    cudaMalloc(&out, (size_t)nwin * COLS * OUT_W * 4);
    if (rec_h_setup() != cudaSuccess) { printf("setup failed\n"); return 1; }
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = launch_rec_h(gi, w, out, nwin, 148, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long t[32];
    cudaMemcpyFromSymbol(t, roko_trace, sizeof(t));
    const long long t0 = t[4];
    for (int s = 0; s < 4; ++s) {
        const long long* p = t + s * 8;
        printf("step %d: mma issue-start %6lld issue-end %6lld | gates r-ready %6lld r-done %6lld n-ready %6lld arrived %6lld\n",
               20 + s, p[4] - t0, p[5] - t0, p[0] - t0, p[1] - t0, p[2] - t0, p[3] - t0);
    }
    return 0;
}
