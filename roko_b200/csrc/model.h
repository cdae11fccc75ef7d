// Private definition of the opaque roko_b200_model handle (include/roko_b200.h), shared by api.cu
// and train_api.cu.
#pragma once
#include "common.cuh"

constexpr int NSLOT = 3;
constexpr int ROKO_ERRBUF = 512;
char* roko_b200_errbuf();         // the calling thread's roko_b200_last_error() buffer (api.cu)

struct roko_b200_model {
    int device = 0;
    int num_sms = 148;
    float* packed = nullptr;
    float* raw_stage = nullptr;     // device copy of the raw (state_dict order) weights; the training backward reads them
    float* raw_al = nullptr;        // the same with 2 floats of padding before the GRU section: every tensor 16-byte aligned
    float* train_img = nullptr;     // tf32 hi / lo images of fc1.weight for the tcgen05 training products (train_tc.cu)
    int train_tc = 3;               // ROKO_B200_TRAIN_TC: which training products run on tcgen05 (train_tc.cu) instead of the
                                    // generic GEMM: >= 1 fc1 and d(ep), >= 2 dW1, >= 3 the GRU d(in), >= 4 dW_ih / dW_hh
                                    // (4 is measured slower: 11 520-row reductions leave too little work per 32 K-atomic tile epilogue)
    int* status = nullptr;          // device flag word, bit 0: code outside 0..11
    bool loaded = false;
    int use_tc = 3;                 // projection: 3 = persistent tcgen05, double-buffered accumulators (proj_tc3.cu, default);
                                    // ROKO_B200_PROJ=tc2 -> 256x256 tile, tc1 -> 128x256 tile, ffma -> FFMA SGEMM
    int superbatch = 2368;          // windows per device pass of infer_host (148 SMs x 16; ROKO_B200_SUPERBATCH)
    int rec_tc_min = 256;           // chunks of at least this many windows use the tcgen05 recurrence (ROKO_B200_REC_TC_MIN; 0 = never)
    roko::FrontConst fc;
    struct Slot {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        uint8_t* x = nullptr;
        uint8_t* labels = nullptr;
        float* logits = nullptr;
        void* ws = nullptr;
    } slot[NSLOT];
    int slot_cap = 0;
};

// offset of raw element `off` inside raw_al (RAW_GRU is 2 mod 4 and every later tensor size is a multiple of 4)
constexpr int RAW_AL_PAD = 2;
static_assert((roko::RAW_GRU + RAW_AL_PAD) % 4 == 0 && roko::RAW_W1 % 4 == 0 && roko::RAW_W2 % 4 == 0, "raw_al alignment");
__host__ __device__ constexpr int raw_al_off(int off) { return off >= roko::RAW_GRU ? off + RAW_AL_PAD : off; }

// Input projection of GRU layer `l` with the kernel the model is configured for (see use_tc).
inline cudaError_t proj_dispatch(const roko_b200_model* m, const float* in, int l, float* gi, int rows,
                                 cudaStream_t s) {
    using namespace roko;
    const float* pk = m->packed;
    if (m->use_tc == 3) return launch_proj_tc3(in, gru_inp(l), pk + pk_wtc(l), pk + pk_bgi(l), gi, rows, m->num_sms, s);
    if (m->use_tc == 2) return launch_proj_tc2(in, gru_inp(l), pk + pk_wt2(l), pk + pk_bgi(l), gi, rows, s);
    if (m->use_tc) return launch_proj_tc(in, gru_inp(l), pk + pk_wtc(l), pk + pk_bgi(l), gi, rows, s);
    return launch_proj(in, gru_inp(l), pk + pk_wih(l), pk + pk_bgi(l), gi, rows, s);
}
