// libb200chan.so - C ABI (include/b200chan.h) over the hand-written sm_100a kernels.
// Host runtime: one handle = one CUDA device + streams + the streaming state of one wideband stream.
#include "../../include/b200chan.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <new>
#include <vector>

#include "common.cuh"
#include "ddc_bank.cuh"
#include "ddc_mma.cuh"
#include "ddc_mma2.cuh"
#include "demod.cuh"
#include "fft_logpow.cuh"
#include "fft_scan.cuh"
#include "fft_frame.cuh"
#include "pfb_fm.cuh"
#include "pfb_fm_tma.cuh"
#include "pfb_fm_ws.cuh"
#include "pfb_cl.cuh"
#include "pfb_fm1.cuh"
#include "internal.h"

using namespace rcb;

namespace {

constexpr int kVersion = 100;       // 0.1.0
constexpr int kDdcHistCap = 1 << 14;  // samples of wideband history kept for the DDC bank (>= max ntaps-1)
constexpr int kStages = 3;          // host<->device pipeline depth of the e2e path
constexpr size_t kChunkSamples = 1u << 22;

struct DdcChan {
    int id = 0;
    int decim = 1, ntaps = 0, out_mask = RCB_OUT_IQ;
    float gain = 1.f;
    double center_freq = 0, samp_rate = 1;
    std::vector<float> taps;
    float2* d_ctaps_rev = nullptr;
    float4* d_ctaps4_rev = nullptr;
    float2* d_out_iq = nullptr;
    float* d_out_fm = nullptr;
    float2* d_prev = nullptr;  // [2]: FM carry, double buffered
    int prev_cur = 0;
    size_t out_cap = 0;
    size_t nout_last = 0;
    uint64_t start_sample = 0;  // stream position of the newest sample of output 0
    uint64_t i_next = 0;        // next output index
    double cyc = 0;             // frac(f0*D/fs)
    double phase_base = 0;      // phase (cycles) at output i_base
    uint64_t i_base = 0;
    uint64_t taps_ver = 0;      // changes with every upload of the composite taps (open / retune / set_taps)
};

struct Stage {
    float2* d_in = nullptr;
    float* d_fm = nullptr;
    float2* d_iq = nullptr;
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    bool used = false;
};

}  // namespace

struct rcb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    char err[512] = {0};
    rcb_stats_t stats{};
    void* l2_scratch = nullptr;
    size_t l2_scratch_bytes = 0;

    // ---- PFB ----
    struct {
        bool configured = false;
        int N = 0, L = 0, P = 0, R = 0, mode = 0;
        float gain = 1.f;
        float* d_taps = nullptr;
        float2* d_tw = nullptr;
        float2* d_tw_tma = nullptr;  // dense swizzled table for pfb_fm_tma_kernel
        float* d_taps_kc = nullptr;  // [PT][N] column-major taps for the time-blocked FIR (P > 1)
        int PT = 1;                  // compiled taps-per-arm of the fast FM kernel (P rounded up to 2^k)
        bool use_tma = false;

        float2* d_hist[2] = {nullptr, nullptr};
        int hist_cur = 0;
        float2* d_zeros = nullptr;  // one all-zero row
        int* d_counter = nullptr;   // dynamic work counter of the TMA kernel
        int* d_counter2 = nullptr;  // pfb_fm1_kernel: ping-pong pair (each launch zeroes the other one)
        unsigned fm1_launches = 0;
        float2* d_ys = nullptr;  // generic path scratch [N][T+1]
        size_t ys_cap = 0;
        int blocks_per_sm = 0;
        size_t smem = 0;
        bool taps_smem = true;
        int variant = 0;  // RCB_PFB_VARIANT: tuning experiments, read only by -DRCB_EXPERIMENTS builds (never shipped)
        float4* d_tw4 = nullptr;  // pfb_cl_kernel twiddles [CS][R/2][16]
        // fused integer ingest (rcb_pfb_set_input_format): raw copy of the last P input rows + how many of them are real
        int in_fmt = 0;
        float in_off = 0.f, in_scale = 1.f;
        void* d_hist_raw[2] = {nullptr, nullptr};
        int hist_valid = 0;
        float2* d_conv = nullptr;      // complex64 staging for kernels without the fused conversion
        size_t conv_cap = 0;
        // multi-stream launches (rcb_pfb_process_multi): parameter blocks of all streams, owned by the leader handle
        unsigned char* d_multi = nullptr;
        unsigned char* h_multi = nullptr;   // pinned, kMultiSlots slots
        size_t multi_cap = 0;               // streams
        unsigned long long multi_no = 0;
        cudaEvent_t ev_mslot[4] = {nullptr, nullptr, nullptr, nullptr};
        cudaEvent_t ev_multi = nullptr;     // per handle: "my stream has reached this point"
        int* d_mcounters = nullptr;
        // > 16 taps per arm, FM only, 1024 channels: time-blocked arm FIR (pfb_arm_fir_kernel) + the one-tap kernel
        bool use_bigp = false;
        float* d_taps_unit = nullptr;  // all-ones taps of the second pass
        float2* d_u = nullptr;         // filtered frames of one chunk
        size_t u_cap = 0;
        float2* d_u_prev = nullptr;    // filtered frame before the block (FM carry of the second pass)
        float* d_taps_gen = nullptr;   // generic-kernel tables, also built for the fast shapes (fallback for
        float2* d_tw_gen = nullptr;    // output buffers the sector-store / TMA kernels cannot address)
        bool use_cl = false;      // FM only, N in {256, 1024}, <= 16 taps per arm: cluster / register-window kernel
        int oblock_log2 = 0;  // rcb_pfb_set_out_block
        Stage st[kStages];
        size_t chunk_frames = 0;
    } pfb;

    // ---- DDC ----
    struct {
        std::map<int, DdcChan> chans;
        int next_id = 1;
        float2* d_hist[2] = {nullptr, nullptr};
        int hist_cur = 0;
        uint64_t n_consumed = 0;
        float2* d_in = nullptr;
        size_t in_cap = 0;
        DdcChanDev* d_chans = nullptr;      // current slot of the rings below
        DdcChanDev* h_chans = nullptr;
        DdcChanDev* d_chans_ring = nullptr;
        DdcChanDev* h_chans_ring = nullptr;  // pinned
        size_t chans_cap = 0;
        DdcGroupDev* d_groups = nullptr;
        DdcGroupDev* h_groups = nullptr;
        DdcGroupDev* d_groups_ring = nullptr;
        DdcGroupDev* h_groups_ring = nullptr;  // pinned
        size_t groups_cap = 0;
        unsigned long long slot_no = 0;
        cudaEvent_t ev_slot[4] = {nullptr, nullptr, nullptr, nullptr};
        float2* d_in2[2] = {nullptr, nullptr};   // host-input staging (chunked)
        void* d_raw2[2] = {nullptr, nullptr};    // wire-format staging (rcb_ddc_set_input_format)
        size_t raw_cap2[2] = {0, 0};             // bytes
        int in_fmt = 0;
        float in_off = 0.f, in_scale = 1.f;
        size_t in_cap2[2] = {0, 0};
        cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
        bool in_used[2] = {false, false};
        // tensor-core path (ddc_mma_kernel): group descriptors (pinned ring like the two above) and the packed B operand
        bool use_lone = true;     // frame-per-lane kernel for lone channels (experiments build: RCB_DDC_LONE=0 disables)
        bool use_mma = true;      // rcb_ddc_set_tensor_cores
        int mma_gen = 2;          // 2: ddc_mma2_kernel (A operand in TMEM), 1: ddc_mma_kernel (A operand in shared memory)
        int mma_nseg = 3;         // generation 1: main accumulators
        int mma_seg_len = 16;     // generation 2: k-chunks per accumulation segment
        DdcMmaGroupDev* d_mgroups_ring = nullptr;
        DdcMmaGroupDev* h_mgroups_ring = nullptr;  // pinned
        size_t mgroups_cap = 0;
        float* d_mma_b = nullptr;
        size_t mma_b_cap = 0;     // floats
        uint64_t mma_launches = 0;
        cudaStream_t s_aux = nullptr;      // ddc_head_kernel runs beside ddc_mma*_kernel (fork / join events)
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        std::vector<long long> mma_b_sig;  // what d_mma_b holds (bucket shapes, channel ids, tap versions): unchanged ->
                                           // the pack kernel is skipped
    } ddc;

    // ---- FFT ----
    FftState fft;

    // generic staging for small host-side helpers
    void* d_tmp[2] = {nullptr, nullptr};
    size_t d_tmp_cap[2] = {0, 0};

    void* post_state = nullptr;  // K6 chains (postdemod_api.cu)
};

// ---- cross-translation-unit accessors (internal.h) ----
int rcb_internal_device(rcb_t* h) { return h->device; }
cudaStream_t rcb_internal_stream(rcb_t* h) { return h->stream; }
int rcb_internal_fail(rcb_t* h, cudaError_t e, const char* what) {
    if (h) snprintf(h->err, sizeof(h->err), "%s: %s", what, cudaGetErrorString(e));
    return RCB_ECUDA;
}
void rcb_internal_count(rcb_t* h, int launches, size_t h2d_bytes, size_t d2h_bytes) {
    h->stats.kernel_launches += (uint64_t)launches;
    h->stats.h2d_bytes += h2d_bytes;
    h->stats.d2h_bytes += d2h_bytes;
}
void** rcb_internal_post_slot(rcb_t* h) { return &h->post_state; }

namespace {

int fail_cuda(rcb_t* h, cudaError_t e, const char* what) {
    if (h) snprintf(h->err, sizeof(h->err), "%s: %s", what, cudaGetErrorString(e));
    return RCB_ECUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call); \
    } while (0)
#define CKL(h)                                                       \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, "kernel launch"); \
        (h)->stats.kernel_launches++;                                \
    } while (0)

int ensure_tmp(rcb_t* h, int which, size_t bytes) {
    if (h->d_tmp_cap[which] >= bytes) return RCB_OK;
    if (h->d_tmp[which]) cudaFree(h->d_tmp[which]);
    h->d_tmp[which] = nullptr;
    h->d_tmp_cap[which] = 0;
    CK(cudaMalloc(&h->d_tmp[which], bytes));
    h->d_tmp_cap[which] = bytes;
    return RCB_OK;
}

// frac(cyc * k) computed exactly (cyc is an exact binary rational, k an integer)
double frac_mul(double cyc, uint64_t k) {
    if (cyc == 0.0 || k == 0) return 0.0;
    int e;
    const double m = frexp(cyc, &e);                       // cyc = m * 2^e, m in [0.5,1)
    const uint64_t mant = (uint64_t)ldexp(m, 53);          // exact
    const int shift = 53 - e;                              // cyc = mant * 2^-shift, shift >= 53
    unsigned __int128 prod = (unsigned __int128)mant * (unsigned __int128)k;
    if (shift < 128) {
        const unsigned __int128 mask = (((unsigned __int128)1) << shift) - 1;
        prod &= mask;
    }
    const long double v = (long double)(uint64_t)(prod >> 64) * 18446744073709551616.0L + (long double)(uint64_t)prod;
    double f = (double)ldexpl(v, -shift);
    if (f >= 1.0) f -= 1.0;
    return f;
}

// ------------------------------------------------------------------------------------------------
// PFB kernel dispatch
// ------------------------------------------------------------------------------------------------
template <int R, int MODE, bool TS, int PT, int PF = 0>
int pfb_launch_t(rcb_t* h, const PfbParams& p, bool query_only) {
    using G = PfbGeom<R>;
    auto kern = pfb_fm_kernel<R, MODE, TS, PT, PF>;
    const size_t smem = G::smem_bytes(p.P, TS);
    if (query_only) {
        // cudaFuncSetAttribute is per function and per DEVICE and a later, smaller value lowers the limit: several handles
        // (other sources, pfb-mode bin engines) share a device, so the opt-in is tracked per device and only ever raised
        static size_t attr_max[64] = {};
        size_t& cur = attr_max[h->device & 63];
        if (smem > cur) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::THREADS, smem));
        h->pfb.blocks_per_sm = std::max(nb, 1);
        h->pfb.smem = smem;
        return RCB_OK;
    }
    const int NI = (p.T + G::FPI - 1) / G::FPI;
    const int grid = std::max(1, std::min(NI, h->pfb.blocks_per_sm * h->sm_count));
    kern<<<grid, G::THREADS, smem, h->stream>>>(p);
    CKL(h);
    return RCB_OK;
}
template <int R, int MODE>
int pfb_launch_rm(rcb_t* h, const PfbParams& p, bool q) {
    if (p.P == 1) {
        if (h->pfb.variant == 1 && R == 32 && MODE == PFB_OUT_FM) return pfb_launch_t<R, MODE, true, 1, 1>(h, p, q);
        return pfb_launch_t<R, MODE, true, 1>(h, p, q);
    }
    if (h->pfb.taps_smem) return pfb_launch_t<R, MODE, true, 0>(h, p, q);
    return pfb_launch_t<R, MODE, false, 0>(h, p, q);
}
template <int R>
int pfb_launch_r(rcb_t* h, const PfbParams& p, bool q) {
    switch (h->pfb.mode) {
        case RCB_OUT_IQ: return pfb_launch_rm<R, PFB_OUT_IQ>(h, p, q);
        case RCB_OUT_FM: return pfb_launch_rm<R, PFB_OUT_FM>(h, p, q);
        default: return pfb_launch_rm<R, PFB_OUT_IQ | PFB_OUT_FM>(h, p, q);
    }
}
template <int R, int W = 8, bool PK = true, int PT = 1, int MODE = PFB_OUT_FM, bool OB8 = false>
int pfb_launch_tma(rcb_t* h, const PfbParams& p, bool query_only) {
    using G = PfbTmaGeom<R, W, MODE>;
    auto kern = pfb_fm_tma_kernel<R, W, PK, PT, MODE, OB8>;
    const size_t smem = G::smem_bytes;
    if (query_only) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::THREADS, smem));
        h->pfb.blocks_per_sm = std::max(nb, 1);
        h->pfb.smem = smem;
        return RCB_OK;
    }
    if (OB8) {  // layout variants are picked per launch: opt in to the shared-memory size once
        static bool attr_dev[64] = {};
        bool& attr_set = attr_dev[h->device & 63];
        if (!attr_set) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
    }
    const int NI = (p.T + G::FPI - 1) / G::FPI;
    const int grid = std::max(1, std::min(NI, h->pfb.blocks_per_sm * h->sm_count));
    PfbParams q = p;
    q.twiddle = h->pfb.d_tw_tma;
    q.taps_kc = h->pfb.d_taps_kc;
    q.work_counter = h->pfb.d_counter;
    CK(cudaMemsetAsync(h->pfb.d_counter, 0, sizeof(int), h->stream));
    kern<<<grid, G::THREADS, smem, h->stream>>>(q);
    CKL(h);
    return RCB_OK;
}

// warp-specialised producer / consumer kernel (FM only, 2..16 taps per arm)
template <int R, int PT, int PW = 8, int MODE = PFB_OUT_FM>
int pfb_launch_ws(rcb_t* h, const PfbParams& p, bool query_only) {
    using G = PfbWsGeom<R, PW, MODE>;
    auto kern = pfb_fm_ws_kernel<R, PT, PW, MODE>;
    const size_t smem = G::smem_bytes(PT);
    if (query_only) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->pfb.blocks_per_sm = 1;
        h->pfb.smem = smem;
        return RCB_OK;
    }
    const int NI = (p.T + G::FPI - 1) / G::FPI;
    const int grid = std::max(1, std::min(NI, h->sm_count));
    PfbParams q = p;
    q.twiddle = h->pfb.d_tw_tma;
    q.taps_kc = h->pfb.d_taps_kc;
    kern<<<grid, G::THREADS, smem, h->stream>>>(q);
    CKL(h);
    return RCB_OK;
}

// cluster / register-window kernel (pfb_cl.cuh).  Returns 1 when the TMA descriptors cannot describe this call
// (misaligned buffers or strides, time blocks shorter than 16 frames): the caller then runs the round-1 kernels.
template <int R, int PT>
int pfb_launch_cl_t(rcb_t* h, const float2* d_x, const float2* d_hist, size_t frames, float* d_fm, size_t ostride) {
    using G = PfbClGeom<R>;
    auto& s = h->pfb;
    constexpr int N = G::N;
    const int kb = s.oblock_log2;
    if (kb > 0 && kb < 4) return 1;
    CUtensorMap tm_x, tm_hist, tm_out;
    {
        const uint64_t dims[3] = {(uint64_t)2 * R, (uint64_t)R, (uint64_t)frames};
        const uint64_t str[2] = {(uint64_t)2 * R * 4, (uint64_t)N * 8};
        const uint32_t box[3] = {32, (uint32_t)R, 2};  // a pair of rows of the CTA's 16 R columns
        if (!tmap_encode_f32(&tm_x, 3, d_x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
        const uint64_t hd[3] = {(uint64_t)2 * R, (uint64_t)R, (uint64_t)s.P};
        if (!tmap_encode_f32(&tm_hist, 3, d_hist, hd, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    }
    PfbClParams p{};
    if (kb == 0) {
        const uint64_t dims[3] = {(uint64_t)frames, (uint64_t)R, (uint64_t)R};
        const uint64_t str[2] = {(uint64_t)ostride * 4, (uint64_t)ostride * 4 * R};
        const uint32_t box[3] = {16, 16, (uint32_t)G::M2W};  // one FIR warp's slice of the tile
        if (ostride < frames || !tmap_encode_f32(&tm_out, 3, d_fm, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
        p.out_rank = 3;
    } else {
        const uint64_t blk = (uint64_t)1 << kb;
        const uint64_t dims[4] = {blk, (uint64_t)R, (uint64_t)R, ((uint64_t)frames + blk - 1) >> kb};
        const uint64_t str[3] = {blk * 4, blk * 4 * R, blk * 4 * N};
        const uint32_t box[4] = {16, 16, (uint32_t)G::M2W, 1};
        if (!tmap_encode_f32(&tm_out, 4, d_fm, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
        p.out_rank = 4;
    }
    p.taps_kc = s.d_taps_kc;
    p.tw4 = s.d_tw4;
    p.out_fm = d_fm;
    p.ostride = (long long)ostride;
    p.oblock_log2 = kb;
    p.T = (int)frames;
    p.P = s.P;
    p.N = N;
    p.gain = s.gain;
    p.work_counter = s.d_counter;
    auto kern = pfb_cl_kernel<R, PT>;
    static bool attr_dev[64] = {};
    bool& attr_set = attr_dev[h->device & 63];
    if (!attr_set) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes));
        attr_set = true;
    }
    const int NI = (int)((frames + 15) / 16);
    const int ncl = std::max(1, std::min(NI, h->sm_count / G::CS));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ncl * G::CS));
    cfg.blockDim = dim3(G::THREADS);
    cfg.dynamicSmemBytes = G::smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G::CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, tm_x, tm_hist, tm_out, p));
    h->stats.kernel_launches++;
    return RCB_OK;
}
template <int R>
int pfb_launch_cl_r(rcb_t* h, const float2* d_x, const float2* d_hist, size_t frames, float* d_fm, size_t ostride) {
    switch (h->pfb.PT) {
        case 1: return pfb_launch_cl_t<R, 1>(h, d_x, d_hist, frames, d_fm, ostride);
        case 2: return pfb_launch_cl_t<R, 2>(h, d_x, d_hist, frames, d_fm, ostride);
        case 4: return pfb_launch_cl_t<R, 4>(h, d_x, d_hist, frames, d_fm, ostride);
        case 8: return pfb_launch_cl_t<R, 8>(h, d_x, d_hist, frames, d_fm, ostride);
        case 16: return pfb_launch_cl_t<R, 16>(h, d_x, d_hist, frames, d_fm, ostride);
    }
    return 1;
}
int pfb_launch_cl(rcb_t* h, const float2* d_x, const float2* d_hist, size_t frames, float* d_fm, size_t ostride) {
    if (h->pfb.R == 32) return pfb_launch_cl_r<32>(h, d_x, d_hist, frames, d_fm, ostride);
    return 1;  // (the kernel also supports N = 256 as a single CTA; the round-1 kernels measured faster there)
}

// single-tap 1024-channel FM kernel with the TMA tile store (pfb_fm1.cuh).  Returns 1 when the output cannot be
// described by a tensor map (misaligned base / stride): the caller falls back.
int pfb_launch_fm1(rcb_t* h, const PfbParams& p0, size_t frames, float* d_fm, size_t ostride) {
    using G = PfbFm1Geom;
    auto& s = h->pfb;
    const int kb = s.oblock_log2;
    CUtensorMap tm_out;
    int out_rank;
    if (kb == 0) {
        const uint64_t dims[2] = {(uint64_t)frames, (uint64_t)G::N};
        const uint64_t str[1] = {(uint64_t)ostride * 4};
        const uint32_t box[2] = {8, 256};
        if (ostride < frames || !tmap_encode_f32(&tm_out, 2, d_fm, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
        out_rank = 2;
    } else {
        const uint64_t blk = (uint64_t)1 << kb;
        const uint64_t dims[3] = {blk, (uint64_t)G::N, ((uint64_t)frames + blk - 1) >> kb};
        const uint64_t str[2] = {blk * 4, blk * 4 * G::N};
        const uint32_t box[3] = {8, 256, 1};
        if (!tmap_encode_f32(&tm_out, 3, d_fm, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
        out_rank = 3;
    }
    typedef void (*fm1_fn)(const CUtensorMap, const PfbParams, const int);
    static const fm1_fn kerns[4] = {pfb_fm1_kernel<0>, pfb_fm1_kernel<1>, pfb_fm1_kernel<2>, pfb_fm1_kernel<3>};
    static bool attr_dev[64][4] = {};
    static int per_sm[64][4] = {};
    const int di = h->device & 63, fi = p0.in_fmt & 3;  // (the format of THIS launch's input: a converted block is complex64)
    fm1_fn kern = kerns[fi];
    if (!attr_dev[di][fi]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::THREADS, G::smem_bytes));
        per_sm[di][fi] = std::max(nb, 1);
        attr_dev[di][fi] = true;
    }
    PfbParams q = p0;
    q.twiddle = s.d_tw_tma;
    // ping-pong work counters: this launch uses one and zeroes the other for the next launch (stream order)
    q.work_counter = s.d_counter2 + (s.fm1_launches & 1);
    q.next_counter = s.d_counter2 + ((s.fm1_launches + 1) & 1);
    ++s.fm1_launches;
    const int NI = (int)((frames + G::FPI - 1) / G::FPI);
    const int grid = std::max(1, std::min(NI, per_sm[di][fi] * h->sm_count));
    kern<<<grid, G::THREADS, G::smem_bytes, h->stream>>>(tm_out, q, out_rank);
    CKL(h);
    return RCB_OK;
}

// fast kernel dispatch: taps per arm (1, 2, 4, 8, 16) x output mode
template <int R, int MODE>
int pfb_launch_tma_pm(rcb_t* h, const PfbParams& p, bool q) {
    switch (h->pfb.PT) {
        case 1: return pfb_launch_tma<R, 8, true, 1, MODE>(h, p, q);
        case 2: return pfb_launch_tma<R, 8, true, 2, MODE>(h, p, q);
        case 4: return pfb_launch_tma<R, 8, true, 4, MODE>(h, p, q);
        case 8: return pfb_launch_tma<R, 8, true, 8, MODE>(h, p, q);
        case 16: return pfb_launch_tma<R, 8, true, 16, MODE>(h, p, q);
    }
    return RCB_EUNSUPPORTED;
}
template <int R>
int pfb_launch_tma_r(rcb_t* h, const PfbParams& p, bool q) {
    const int v = h->pfb.variant;
    if (h->pfb.mode == RCB_OUT_FM) {
        if (R == 32) {
            // 1024 channels, FM only: measured variants.  8+ taps per arm run one 16-warp CTA
            // per SM with 16-frame FIR tasks (fewer history re-reads); RCB_PFB_VARIANT=8 forces 2 x 8 warps.
            if (h->pfb.PT == 1 && v == 16) return pfb_launch_tma<32, 16, true, 1>(h, p, q);
            if (h->pfb.PT == 1 && !q && p.oblock_log2 == 3 && p.out_fm) return pfb_launch_tma<32, 8, true, 1, PFB_OUT_FM, true>(h, p, q);
            if (h->pfb.PT == 1 && v == 3) return pfb_launch_tma<32, 8, false, 1>(h, p, q);  // scalar-arithmetic v5 kernel
            if (h->pfb.PT == 16 && v == 16) return pfb_launch_tma<32, 16, true, 16>(h, p, q);
            if (h->pfb.PT == 8 && v == 16) return pfb_launch_tma<32, 16, true, 8>(h, p, q);
        }
        // 8+ taps per arm: warp-specialised producer / consumer kernel (RCB_PFB_VARIANT=8 / 16: the phase-serial
        // 2 x 8-warp / 16-warp variants)
        if (R == 32 && v != 8 && v != 16) {
            if (h->pfb.PT == 16) return (v == 12) ? pfb_launch_ws<R, 16, 12>(h, p, q) : pfb_launch_ws<R, 16>(h, p, q);
            if (h->pfb.PT == 8) return (v == 12) ? pfb_launch_ws<R, 8, 12>(h, p, q) : pfb_launch_ws<R, 8>(h, p, q);
        }
        return pfb_launch_tma_pm<R, PFB_OUT_FM>(h, p, q);
    }
    // IQ outputs, 1024 channels, 8+ taps per arm: the warp-specialised kernel as well (208 KB: two frame-buffer sets +
    // a ring of Y); RCB_PFB_VARIANT=8: the phase-serial one-CTA variant
    if (R == 32 && v != 8 && (h->pfb.PT == 16 || h->pfb.PT == 8)) {
        if (h->pfb.mode == RCB_OUT_IQ)
            return h->pfb.PT == 16 ? pfb_launch_ws<R, 16, 8, PFB_OUT_IQ>(h, p, q) : pfb_launch_ws<R, 8, 8, PFB_OUT_IQ>(h, p, q);
        return h->pfb.PT == 16 ? pfb_launch_ws<R, 16, 8, PFB_OUT_IQ | PFB_OUT_FM>(h, p, q)
                               : pfb_launch_ws<R, 8, 8, PFB_OUT_IQ | PFB_OUT_FM>(h, p, q);
    }
    if (h->pfb.mode == RCB_OUT_IQ) return pfb_launch_tma_pm<R, PFB_OUT_IQ>(h, p, q);
    return pfb_launch_tma_pm<R, PFB_OUT_IQ | PFB_OUT_FM>(h, p, q);
}

int pfb_launch_fast(rcb_t* h, const PfbParams& p, bool q) {
    if (h->pfb.use_tma) {
        switch (h->pfb.R) {
            case 8: return pfb_launch_tma_r<8>(h, p, q);
            case 16: return pfb_launch_tma_r<16>(h, p, q);
            case 32: return pfb_launch_tma_r<32>(h, p, q);
        }
    }
    switch (h->pfb.R) {
        case 8: return pfb_launch_r<8>(h, p, q);
        case 16: return pfb_launch_r<16>(h, p, q);
        case 32: return pfb_launch_r<32>(h, p, q);
    }
    return RCB_EUNSUPPORTED;
}

void pfb_free_stages(rcb_t* h) {
    auto& s = h->pfb;
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->s_in);
    cudaStreamSynchronize(h->s_out);
    for (auto& st : s.st) {
        cudaFree(st.d_in);
        cudaFree(st.d_fm);
        cudaFree(st.d_iq);
        if (st.ev_in) cudaEventDestroy(st.ev_in);
        if (st.ev_k) cudaEventDestroy(st.ev_k);
        if (st.ev_out) cudaEventDestroy(st.ev_out);
        st = Stage{};
    }
    s.chunk_frames = 0;
}

void pfb_free(rcb_t* h) {
    auto& s = h->pfb;
    cudaFree(s.d_taps);
    cudaFree(s.d_tw);
    cudaFree(s.d_tw_tma);
    s.d_tw_tma = nullptr;
    cudaFree(s.d_taps_kc);
    s.d_taps_kc = nullptr;
    cudaFree(s.d_tw4);
    s.d_tw4 = nullptr;
    cudaFree(s.d_taps_gen);
    cudaFree(s.d_tw_gen);
    s.d_taps_gen = nullptr;
    s.d_tw_gen = nullptr;
    cudaFree(s.d_multi);
    if (s.h_multi) cudaFreeHost(s.h_multi);
    cudaFree(s.d_mcounters);
    s.d_multi = nullptr;
    s.h_multi = nullptr;
    s.d_mcounters = nullptr;
    s.multi_cap = 0;
    for (int i = 0; i < 4; ++i)
        if (s.ev_mslot[i]) {
            cudaEventDestroy(s.ev_mslot[i]);
            s.ev_mslot[i] = nullptr;
        }
    if (s.ev_multi) {
        cudaEventDestroy(s.ev_multi);
        s.ev_multi = nullptr;
    }
    cudaFree(s.d_taps_unit);
    cudaFree(s.d_u);
    cudaFree(s.d_u_prev);
    s.d_taps_unit = nullptr;
    s.d_u = nullptr;
    s.d_u_prev = nullptr;
    s.u_cap = 0;
    s.use_bigp = false;
    cudaFree(s.d_hist_raw[0]);
    cudaFree(s.d_hist_raw[1]);
    s.d_hist_raw[0] = s.d_hist_raw[1] = nullptr;
    cudaFree(s.d_conv);
    s.d_conv = nullptr;
    s.conv_cap = 0;
    s.in_fmt = 0;
    s.in_off = 0.f;
    s.in_scale = 1.f;
    s.hist_valid = 0;
    s.use_cl = false;

    cudaFree(s.d_hist[0]);
    cudaFree(s.d_hist[1]);
    cudaFree(s.d_ys);
    cudaFree(s.d_zeros);
    s.d_zeros = nullptr;
    cudaFree(s.d_counter);
    s.d_counter = nullptr;
    cudaFree(s.d_counter2);
    s.d_counter2 = nullptr;
    s.fm1_launches = 0;
    pfb_free_stages(h);
    s.d_taps = nullptr;
    s.d_tw = nullptr;
    s.d_hist[0] = s.d_hist[1] = nullptr;
    s.d_ys = nullptr;
    s.ys_cap = 0;
    s.chunk_frames = 0;
    s.configured = false;
}

// one launch over device-resident input; advances the streaming history
size_t pfb_in_bytes_per_sample(int fmt) { return fmt == 0 ? 8 : (fmt == RCB_FMT_S16 ? 4 : 2); }

// one launch over device-resident input (complex64, or raw integer I/Q when an input format is set); advances the
// streaming history
int pfb_run_device(rcb_t* h, const void* d_in, size_t frames, float2* d_iq, float* d_fm, size_t ostride) {
    auto& s = h->pfb;
    if (frames == 0) return RCB_OK;
    const bool raw = (s.in_fmt != 0);
    const size_t nsamp = frames * (size_t)s.N;
    PfbParams p{};
    p.x = (const float2*)d_in;
    p.hist = s.d_hist[s.hist_cur];
    p.taps = s.d_taps;
    p.twiddle = s.d_tw;
    p.zeros = s.d_zeros;
    p.out_fm = (s.mode & RCB_OUT_FM) ? d_fm : nullptr;
    p.out_iq = (s.mode & RCB_OUT_IQ) ? d_iq : nullptr;
    p.ostride = (long long)ostride;
    p.oblock_log2 = s.oblock_log2;
    p.T = (int)frames;
    p.P = s.P;
    p.N = s.N;
    p.gain = s.gain;
    p.in_fmt = s.in_fmt;
    p.in_off = s.in_off;
    p.in_scale = s.in_scale;
    p.hist_raw = s.d_hist_raw[s.hist_cur];
    p.hist_valid = s.hist_valid;
    int cl_rc = 1;
    // 1024 channels, FM only: one tap per arm -> the TMA-store variant of the round-1 headline kernel (2 CTAs / SM, 16
    // FFT warps; reads raw integer I/Q directly); 2..8 taps per arm -> the cluster / register-window kernel; 16 taps per
    // arm -> the round-1 warp-specialised kernel (measured, scripts/exp/sweep_pfb.py).
    // (experiment builds: RCB_PFB_VARIANT=21 runs the cluster kernel for every tap count)
    bool hist_in_kernel = false;
    if (s.use_cl && s.R == 32 && s.PT == 1 && s.variant != 21) {
        PfbParams q1 = p;
        if (!raw) q1.hist_out = s.d_hist[s.hist_cur ^ 1];   // P = 1: the kernel saves the block's last row itself
        cl_rc = pfb_launch_fm1(h, q1, frames, d_fm, ostride);
        if (cl_rc != RCB_OK && cl_rc != 1) return cl_rc;
        hist_in_kernel = (cl_rc == RCB_OK && !raw);
    }
    const float2* d_x = (const float2*)d_in;
    if (cl_rc == 1 && raw) {
        // every other kernel takes complex64: convert this block once (K5) and go on as usual
        if (s.conv_cap < nsamp) {
            cudaFree(s.d_conv);
            s.d_conv = nullptr;
            s.conv_cap = 0;
            CK(cudaMalloc(&s.d_conv, nsamp * sizeof(float2)));
            s.conv_cap = nsamp;
        }
        const unsigned grid = (unsigned)((nsamp + 1023) / 1024);
        if (s.in_fmt == RCB_FMT_U8)
            convert_iq_kernel<uint8_t><<<grid, 256, 0, h->stream>>>((const uint8_t*)d_in, s.d_conv, (long long)nsamp, s.in_off, s.in_scale);
        else if (s.in_fmt == RCB_FMT_S8)
            convert_iq_kernel<int8_t><<<grid, 256, 0, h->stream>>>((const int8_t*)d_in, s.d_conv, (long long)nsamp, s.in_off, s.in_scale);
        else
            convert_iq_kernel<int16_t><<<grid, 256, 0, h->stream>>>((const int16_t*)d_in, s.d_conv, (long long)nsamp, s.in_off, s.in_scale);
        CKL(h);
        d_x = s.d_conv;
        p.x = d_x;
        p.in_fmt = 0;
    }
    if (cl_rc == 1 && s.use_cl && (s.PT <= 8 || s.variant == 21)) {
        cl_rc = pfb_launch_cl(h, d_x, s.d_hist[s.hist_cur], frames, d_fm, ostride);
        if (cl_rc != RCB_OK && cl_rc != 1) return cl_rc;
    }
    if (cl_rc == 1 && s.use_bigp) {
        // > 16 taps per arm: u = arm FIR of the block (time-blocked, shared-memory tiles), then FFT + FM demod of u as a
        // one-tap channelizer with unit taps; chunks of 2^26 samples bound the temporary
        const size_t chunk = std::max<size_t>(128, ((size_t)1 << 26) / (size_t)s.N);
        const size_t ucap = std::min(frames, chunk) * (size_t)s.N;
        if (s.u_cap < ucap) {
            cudaFree(s.d_u);
            s.d_u = nullptr;
            s.u_cap = 0;
            CK(cudaMalloc(&s.d_u, ucap * sizeof(float2)));
            s.u_cap = ucap;
        }
        const size_t smem = (size_t)(128 + s.P - 1) * 16 * sizeof(float2) + (size_t)s.P * 16 * sizeof(float);
        static size_t fir_attr[64] = {};
        size_t& cur = fir_attr[h->device & 63];
        if (smem > cur) {
            CK(cudaFuncSetAttribute(pfb_arm_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
        const size_t blk = s.oblock_log2 ? ((size_t)1 << s.oblock_log2) : 0;
        bool ok = true;
        for (size_t t0 = 0; t0 < frames && ok; t0 += chunk) {
            const size_t cnt = std::min(chunk, frames - t0);
            dim3 grid((unsigned)(s.N / 16), (unsigned)((cnt + 127) / 128));
            pfb_arm_fir_kernel<<<grid, 256, smem, h->stream>>>(d_x, s.d_hist[s.hist_cur], s.d_taps_kc, s.d_u, s.N, s.P,
                                                             (long long)frames, (long long)t0, (long long)cnt);
            CKL(h);
            PfbParams q = p;
            q.x = s.d_u;
            q.hist = s.d_u_prev;
            q.taps = s.d_taps_unit;
            q.P = 1;
            q.T = (int)cnt;
            q.in_fmt = 0;
            float* o = blk ? d_fm + (t0 / blk) * (size_t)s.N * blk : d_fm + t0;
            q.out_fm = o;
            const int rc = pfb_launch_fm1(h, q, cnt, o, ostride);
            if (rc == 1 && t0 == 0) {
                ok = false;   // output not addressable by a tensor map: the round-1 kernel runs the whole call
                break;
            }
            if (rc) return rc;
            CK(cudaMemcpyAsync(s.d_u_prev, s.d_u + (cnt - 1) * (size_t)s.N, (size_t)s.N * sizeof(float2), cudaMemcpyDeviceToDevice,
                               h->stream));
        }
        if (ok) cl_rc = RCB_OK;
    }
    // the round-1 fast kernels emit 32-byte sector stores (st.global.v8 / two of them per complex row piece): they need
    // 32-byte aligned output rows; anything else takes the generic path
    const bool aligned32 = (s.oblock_log2 > 0 || (ostride % 8) == 0) && (!d_fm || ((uintptr_t)d_fm & 31) == 0) &&
                           (!d_iq || ((uintptr_t)d_iq & 31) == 0);
    if (cl_rc == RCB_OK) {
    } else if (s.R && aligned32) {
        int rc = pfb_launch_fast(h, p, false);
        if (rc) return rc;
    } else {
        const size_t need = (size_t)s.N * (frames + 1);
        if (s.ys_cap < need) {
            cudaFree(s.d_ys);
            s.d_ys = nullptr;
            s.ys_cap = 0;
            CK(cudaMalloc(&s.d_ys, need * sizeof(float2)));
            s.ys_cap = need;
        }
        const int grid = (int)std::min<size_t>(frames + 1, (size_t)h->sm_count * 8);
        const int threads = std::min(256, std::max(32, ((s.N + 31) / 32) * 32));
        if (s.R) {  // a fast shape on the generic path: its own table layouts
            p.taps = s.d_taps_gen;
            p.twiddle = s.d_tw_gen;
        }
        pfb_generic_kernel<<<grid, threads, 2 * (size_t)s.N * sizeof(float2), h->stream>>>(p, s.d_ys, (long long)frames + 1);
        CKL(h);
        dim3 g2((unsigned)((frames + 255) / 256), (unsigned)s.N);
        pfb_generic_emit_kernel<<<g2, 256, 0, h->stream>>>(s.d_ys, (long long)frames + 1, p.out_iq, p.out_fm,
                                                           p.ostride, p.T, p.gain, p.oblock_log2, p.N);
        CKL(h);
    }
    // hist <- last P rows of (hist ++ x): complex64 always (every kernel but the fused-ingest one reads it), plus the raw
    // copy when an input format is set
    const long long cap = (long long)s.P * s.N;
    const int nxt = s.hist_cur ^ 1;
    if (raw) {
        hist_update_convert_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, h->stream>>>(
            s.d_hist[s.hist_cur], d_in, s.in_fmt, s.in_off, s.in_scale, (long long)nsamp, s.d_hist[nxt], cap);
        CKL(h);
        const long long u = (long long)(pfb_in_bytes_per_sample(s.in_fmt) / 2);  // 16-bit units per sample
        hist_update_raw_kernel<<<(unsigned)((cap * u + 255) / 256), 256, 0, h->stream>>>(
            (const unsigned short*)s.d_hist_raw[s.hist_cur], (const unsigned short*)d_in, (long long)nsamp * u,
            (unsigned short*)s.d_hist_raw[nxt], cap * u);
        CKL(h);
        s.hist_valid = (int)std::min<long long>((long long)s.P, (long long)s.hist_valid + (long long)frames);
    } else if (!hist_in_kernel) {
        hist_update_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, h->stream>>>(s.d_hist[s.hist_cur], d_x,
                                                                              (long long)nsamp, s.d_hist[nxt], cap);
        CKL(h);
    }
    s.hist_cur = nxt;
    h->stats.samples_in += frames * (uint64_t)s.N;
    h->stats.channel_samples += frames * (uint64_t)s.N;
    return RCB_OK;
}

int pfb_ensure_stages(rcb_t* h) {
    auto& s = h->pfb;
    if (s.chunk_frames) return RCB_OK;
    size_t cf = std::max<size_t>(8, kChunkSamples / (size_t)s.N);
    cf = (cf + 31) / 32 * 32;
    if (s.oblock_log2) {  // whole time blocks per chunk (a block larger than the default chunk becomes the chunk)
        const size_t blk = (size_t)1 << s.oblock_log2;
        cf = std::max(blk, cf / blk * blk);
    }
    for (auto& st : s.st) {
        CK(cudaMalloc(&st.d_in, cf * s.N * pfb_in_bytes_per_sample(s.in_fmt)));
        if (s.mode & RCB_OUT_FM) CK(cudaMalloc(&st.d_fm, cf * s.N * sizeof(float)));
        if (s.mode & RCB_OUT_IQ) CK(cudaMalloc(&st.d_iq, cf * s.N * sizeof(float2)));
        CK(cudaEventCreateWithFlags(&st.ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st.ev_k, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st.ev_out, cudaEventDisableTiming));
        st.used = false;
    }
    s.chunk_frames = cf;
    return RCB_OK;
}

}  // namespace

// =================================================================================================
// library / handle
// =================================================================================================
extern "C" int rcb_version(void) { return kVersion; }

extern "C" const char* rcb_strerror(int st) {
    switch (st) {
        case RCB_OK: return "ok";
        case RCB_EINVAL: return "invalid argument";
        case RCB_ENOMEM: return "out of memory";
        case RCB_ECUDA: return "CUDA error";
        case RCB_ESTATE: return "invalid state / call order";
        case RCB_ENODEV: return "no such CUDA device";
        case RCB_ERANGE: return "out of range";
        case RCB_EUNSUPPORTED: return "unsupported shape";
    }
    return "unknown status";
}

extern "C" int rcb_device_count(int* count) {
    if (!count) return RCB_EINVAL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return RCB_ENODEV;
    }
    *count = n;
    return RCB_OK;
}

extern "C" int rcb_open(int device, rcb_t** out) {
    if (!out) return RCB_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return RCB_ENODEV;
    }
    if (device < 0 || device >= n) return RCB_ENODEV;
    rcb_t* h = new (std::nothrow) rcb_ctx();
    if (!h) return RCB_ENOMEM;
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->t0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->t1);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        delete h;
        return RCB_ECUDA;
    }
    *out = h;
    return RCB_OK;
}

extern "C" int rcb_close(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    pfb_free(h);
    rcb_post_free_all(h);
    for (auto& kv : h->ddc.chans) {
        cudaFree(kv.second.d_ctaps_rev);
        cudaFree(kv.second.d_ctaps4_rev);
        cudaFree(kv.second.d_out_iq);
        cudaFree(kv.second.d_out_fm);
        cudaFree(kv.second.d_prev);
    }
    cudaFree(h->ddc.d_hist[0]);
    cudaFree(h->ddc.d_hist[1]);
    cudaFree(h->ddc.d_in);
    cudaFree(h->ddc.d_chans_ring);
    if (h->ddc.h_chans_ring) cudaFreeHost(h->ddc.h_chans_ring);
    cudaFree(h->ddc.d_groups_ring);
    if (h->ddc.h_groups_ring) cudaFreeHost(h->ddc.h_groups_ring);
    cudaFree(h->ddc.d_mgroups_ring);
    if (h->ddc.h_mgroups_ring) cudaFreeHost(h->ddc.h_mgroups_ring);
    cudaFree(h->ddc.d_mma_b);
    if (h->ddc.s_aux) cudaStreamDestroy(h->ddc.s_aux);
    if (h->ddc.ev_fork) cudaEventDestroy(h->ddc.ev_fork);
    if (h->ddc.ev_join) cudaEventDestroy(h->ddc.ev_join);
    for (int i = 0; i < 4; ++i)
        if (h->ddc.ev_slot[i]) cudaEventDestroy(h->ddc.ev_slot[i]);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->ddc.d_in2[i]);
        cudaFree(h->ddc.d_raw2[i]);
        if (h->ddc.ev_in[i]) cudaEventDestroy(h->ddc.ev_in[i]);
        if (h->ddc.ev_done[i]) cudaEventDestroy(h->ddc.ev_done[i]);
    }
    fft_free(h->fft);
    cudaFree(h->d_tmp[0]);
    cudaFree(h->d_tmp[1]);
    cudaFree(h->l2_scratch);
    cudaEventDestroy(h->t0);
    cudaEventDestroy(h->t1);
    cudaStreamDestroy(h->stream);
    cudaStreamDestroy(h->s_in);
    cudaStreamDestroy(h->s_out);
    delete h;
    return RCB_OK;
}

extern "C" int rcb_sync(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_in));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->s_out));
    return RCB_OK;
}

extern "C" const char* rcb_last_error(rcb_t* h) { return h ? h->err : "null handle"; }

extern "C" int rcb_stats(rcb_t* h, rcb_stats_t* out) {
    if (!h || !out) return RCB_EINVAL;
    *out = h->stats;
    return RCB_OK;
}

extern "C" int rcb_device_name(rcb_t* h, char* buf, size_t cap, int* sm_count) {
    if (!h) return RCB_EINVAL;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    if (buf && cap) {
        strncpy(buf, prop.name, cap - 1);
        buf[cap - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return RCB_OK;
}

// =================================================================================================
// memory / timing helpers
// =================================================================================================
extern "C" int rcb_dev_alloc(rcb_t* h, size_t bytes, void** p) {
    if (!h || !p) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return RCB_ENOMEM;
    }
    CK(e);
    return RCB_OK;
}
extern "C" int rcb_dev_free(rcb_t* h, void* p) {
    if (!h) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaFree(p));
    return RCB_OK;
}
extern "C" int rcb_host_alloc(rcb_t* h, size_t bytes, void** p) {
    if (!h || !p) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return RCB_ENOMEM;
    }
    CK(e);
    return RCB_OK;
}
extern "C" int rcb_host_free(rcb_t* h, void* p) {
    if (!h) return RCB_EINVAL;
    CK(cudaFreeHost(p));
    return RCB_OK;
}
extern "C" int rcb_memcpy(rcb_t* h, void* dst, const void* src, size_t bytes, int kind) {
    if (!h || (!dst && bytes) || (!src && bytes)) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    cudaMemcpyKind k;
    switch (kind) {
        case RCB_COPY_H2D: k = cudaMemcpyHostToDevice; h->stats.h2d_bytes += bytes; break;
        case RCB_COPY_D2H: k = cudaMemcpyDeviceToHost; h->stats.d2h_bytes += bytes; break;
        case RCB_COPY_D2D: k = cudaMemcpyDeviceToDevice; break;
        default: return RCB_EINVAL;
    }
    CK(cudaMemcpyAsync(dst, src, bytes, k, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RCB_OK;
}
extern "C" int rcb_memset(rcb_t* h, void* dev, int value, size_t bytes) {
    if (!h || !dev) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(dev, value, bytes, h->stream));
    return RCB_OK;
}
extern "C" int rcb_l2_flush(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    if (!h->l2_scratch) {
        int l2 = 0;
        CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
        h->l2_scratch_bytes = std::max<size_t>((size_t)l2 * 2, (size_t)256 << 20);
        CK(cudaMalloc(&h->l2_scratch, h->l2_scratch_bytes));
    }
    CK(cudaMemsetAsync(h->l2_scratch, 0x5a, h->l2_scratch_bytes, h->stream));
    return RCB_OK;
}
extern "C" int rcb_timer_start(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->t0, h->stream));
    return RCB_OK;
}
extern "C" int rcb_timer_stop(rcb_t* h, float* ms) {
    if (!h || !ms) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->t1, h->stream));
    CK(cudaEventSynchronize(h->t1));
    CK(cudaEventElapsedTime(ms, h->t0, h->t1));
    return RCB_OK;
}

// Bare host<->device copy rate of this handle's device: `iters` rounds of an H2D copy of h2d_bytes and a D2H copy of
// d2h_bytes running CONCURRENTLY on the two copy streams from / to pinned memory, no kernels.  This is the ceiling of
// every host-facing (e2e) figure: bench.py reports e2e as a fraction of it.
extern "C" int rcb_copy_ceiling(rcb_t* h, size_t h2d_bytes, size_t d2h_bytes, int iters, double* h2d_gbs, double* d2h_gbs,
                                double* wall_s) {
    if (!h || iters < 1 || (!h2d_bytes && !d2h_bytes)) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    void *hin = nullptr, *hout = nullptr, *din = nullptr, *dout = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, f0 = nullptr, f1 = nullptr;
    int rc = RCB_OK;
    auto cleanup = [&]() {
        if (hin) cudaFreeHost(hin);
        if (hout) cudaFreeHost(hout);
        cudaFree(din);
        cudaFree(dout);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (f0) cudaEventDestroy(f0);
        if (f1) cudaEventDestroy(f1);
    };
#define CCK(call)                                   \
    do {                                            \
        cudaError_t e__ = (call);                   \
        if (e__ != cudaSuccess) {                   \
            rc = fail_cuda(h, e__, #call);          \
            cleanup();                              \
            return rc;                              \
        }                                           \
    } while (0)
    if (h2d_bytes) {
        CCK(cudaHostAlloc(&hin, h2d_bytes, cudaHostAllocDefault));
        memset(hin, 1, h2d_bytes);  // first touch on the calling thread's NUMA node
        CCK(cudaMalloc(&din, h2d_bytes));
    }
    if (d2h_bytes) {
        CCK(cudaHostAlloc(&hout, d2h_bytes, cudaHostAllocDefault));
        memset(hout, 1, d2h_bytes);
        CCK(cudaMalloc(&dout, d2h_bytes));
    }
    CCK(cudaEventCreate(&e0));
    CCK(cudaEventCreate(&e1));
    CCK(cudaEventCreate(&f0));
    CCK(cudaEventCreate(&f1));
    // one untimed round, then the timed ones
    if (h2d_bytes) CCK(cudaMemcpyAsync(din, hin, h2d_bytes, cudaMemcpyHostToDevice, h->s_in));
    if (d2h_bytes) CCK(cudaMemcpyAsync(hout, dout, d2h_bytes, cudaMemcpyDeviceToHost, h->s_out));
    CCK(cudaStreamSynchronize(h->s_in));
    CCK(cudaStreamSynchronize(h->s_out));
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    CCK(cudaEventRecord(e0, h->s_in));
    CCK(cudaEventRecord(f0, h->s_out));
    for (int i = 0; i < iters; ++i) {
        if (h2d_bytes) CCK(cudaMemcpyAsync(din, hin, h2d_bytes, cudaMemcpyHostToDevice, h->s_in));
        if (d2h_bytes) CCK(cudaMemcpyAsync(hout, dout, d2h_bytes, cudaMemcpyDeviceToHost, h->s_out));
    }
    CCK(cudaEventRecord(e1, h->s_in));
    CCK(cudaEventRecord(f1, h->s_out));
    CCK(cudaStreamSynchronize(h->s_in));
    CCK(cudaStreamSynchronize(h->s_out));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    float ms_in = 0.f, ms_out = 0.f;
    CCK(cudaEventElapsedTime(&ms_in, e0, e1));
    CCK(cudaEventElapsedTime(&ms_out, f0, f1));
    if (h2d_gbs) *h2d_gbs = (h2d_bytes && ms_in > 0) ? (double)h2d_bytes * iters / (ms_in * 1e-3) / 1e9 : 0.0;
    if (d2h_gbs) *d2h_gbs = (d2h_bytes && ms_out > 0) ? (double)d2h_bytes * iters / (ms_out * 1e-3) / 1e9 : 0.0;
    if (wall_s) *wall_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    h->stats.h2d_bytes += h2d_bytes * (size_t)(iters + 1);
    h->stats.d2h_bytes += d2h_bytes * (size_t)(iters + 1);
#undef CCK
    cleanup();
    return RCB_OK;
}

// =================================================================================================
// K1  PFB + FM
// =================================================================================================
extern "C" int rcb_pfb_config(rcb_t* h, int nchans, const float* taps, int ntaps, int out_mask, float fm_gain) {
    if (!h || !taps || nchans < 1 || ntaps < 1) return RCB_EINVAL;
    if (!(out_mask & (RCB_OUT_IQ | RCB_OUT_FM)) || (out_mask & ~(RCB_OUT_IQ | RCB_OUT_FM))) return RCB_EINVAL;
    if (nchans > 65536) return RCB_EUNSUPPORTED;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    pfb_free(h);
    auto& s = h->pfb;
    s.N = nchans;
    s.L = ntaps;
    s.P = (ntaps + nchans - 1) / nchans;
    s.mode = out_mask;
    s.gain = fm_gain;
    s.R = (nchans == 64) ? 8 : (nchans == 256) ? 16 : (nchans == 1024) ? 32 : 0;
#ifdef RCB_EXPERIMENTS
    {
        const char* v = getenv("RCB_PFB_VARIANT");
        s.variant = v ? atoi(v) : 0;
    }
#else
    s.variant = 0;
#endif
    const int N = s.N, P = s.P, R = s.R;
    std::vector<float> hp((size_t)P * N, 0.f);
    for (int i = 0; i < ntaps; ++i) hp[i] = taps[i];
    std::vector<float> tperm((size_t)P * N);
    std::vector<float2> tw;
    if (R) {
        for (int k = 0; k < P; ++k)
            for (int jj = 0; jj < R; ++jj)
                for (int ll = 0; ll < R; ++ll)
                    tperm[(((size_t)k * (R / 4) + jj / 4) * R + ll) * 4 + (jj % 4)] =
                        hp[(size_t)R * (R - 1 - jj) + (R - 1 - ll) + (size_t)k * N];
        const int S = R + 2;
        tw.assign((size_t)R * S, make_float2(0.f, 0.f));
        for (int ll = 0; ll < R; ++ll)
            for (int m1 = 0; m1 < R; ++m1) {
                const int q = ((R - 1 - ll) * m1) % N;
                const double a = 2.0 * M_PI * (double)q / (double)N;
                tw[(size_t)ll * S + m1] = make_float2((float)cos(a), (float)sin(a));
            }
        s.taps_smem = ((size_t)P * N * sizeof(float) <= 16384);
        // FM-only, one tap per arm: the TMA-staged kernel (RCB_PFB_VARIANT=9 keeps the register-prefetch one)
        // 2..16 taps per arm: same kernel with the time-blocked arm FIR in front (P rounded up to 2, 4, 8, 16 with
        // zero taps; RCB_PFB_VARIANT=9 keeps the register-prefetch kernel for both)
        s.PT = 1;
        while (s.PT < P) s.PT *= 2;
        s.use_tma = (s.PT <= 16 && s.variant != 9);
        if (s.use_tma && s.PT > 1) {
            std::vector<float> kc((size_t)s.PT * N, 0.f);
            for (int k = 0; k < P; ++k)
                for (int c = 0; c < N; ++c) kc[(size_t)k * N + c] = hp[(size_t)(N - 1 - c) + (size_t)k * N];
            CK(cudaMalloc(&s.d_taps_kc, kc.size() * sizeof(float)));
            CK(cudaMemcpy(s.d_taps_kc, kc.data(), kc.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        // FM only, N = 256 / 1024: cluster / register-window kernel (RCB_PFB_VARIANT=20 in experiment builds: round-1 kernels)
        s.use_cl = (s.use_tma && R == 32 && s.mode == RCB_OUT_FM && s.variant != 20);  // (N = 256: the round-1 kernels are faster)
        if (s.use_cl) {
            if (!s.d_taps_kc) {  // one tap per arm: the [PT][N] column-major table is just the reversed prototype
                std::vector<float> kc((size_t)N, 0.f);
                for (int c = 0; c < N; ++c) kc[c] = hp[(size_t)(N - 1 - c)];
                CK(cudaMalloc(&s.d_taps_kc, kc.size() * sizeof(float)));
                CK(cudaMemcpy(s.d_taps_kc, kc.data(), kc.size() * sizeof(float), cudaMemcpyHostToDevice));
            }
            const int CS = R / 16;
            std::vector<float4> t4((size_t)CS * (R / 2) * 16);
            for (int rk = 0; rk < CS; ++rk)
                for (int c = 0; c < R / 2; ++c)
                    for (int l = 0; l < 16; ++l) {
                        const int ll = 16 * rk + l;
                        const int q0 = ((R - 1 - ll) * (2 * c)) % N, q1 = ((R - 1 - ll) * (2 * c + 1)) % N;
                        const double a0 = 2.0 * M_PI * (double)q0 / (double)N, a1 = 2.0 * M_PI * (double)q1 / (double)N;
                        t4[((size_t)rk * (R / 2) + c) * 16 + l] =
                            make_float4((float)cos(a0), (float)sin(a0), (float)cos(a1), (float)sin(a1));
                    }
            CK(cudaMalloc(&s.d_tw4, t4.size() * sizeof(float4)));
            CK(cudaMemcpy(s.d_tw4, t4.data(), t4.size() * sizeof(float4), cudaMemcpyHostToDevice));
        }
        // many taps per arm (FM only, 1024 channels): arm FIR kernel + the one-tap kernel with unit taps
        s.use_bigp = (R == 32 && P > 16 && P <= 1024 && s.mode == RCB_OUT_FM && s.variant != 20);
        if (s.use_bigp) {
            std::vector<float> kc((size_t)P * N, 0.f);
            for (int k = 0; k < P; ++k)
                for (int c = 0; c < N; ++c) kc[(size_t)k * N + c] = hp[(size_t)(N - 1 - c) + (size_t)k * N];
            cudaFree(s.d_taps_kc);
            s.d_taps_kc = nullptr;
            CK(cudaMalloc(&s.d_taps_kc, kc.size() * sizeof(float)));
            CK(cudaMemcpy(s.d_taps_kc, kc.data(), kc.size() * sizeof(float), cudaMemcpyHostToDevice));
            std::vector<float> ones((size_t)N, 1.f);
            CK(cudaMalloc(&s.d_taps_unit, ones.size() * sizeof(float)));
            CK(cudaMemcpy(s.d_taps_unit, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
            CK(cudaMalloc(&s.d_u_prev, (size_t)N * sizeof(float2)));
            CK(cudaMemset(s.d_u_prev, 0, (size_t)N * sizeof(float2)));
        }
        if (s.use_tma || s.use_bigp) {
            std::vector<float2> tt((size_t)N);
            for (int ll = 0; ll < R; ++ll) {
                const int sw = (R == 8) ? ((ll >> 1) & 3) : (ll & (R / 2 - 1));
                for (int m1 = 0; m1 < R; ++m1) {
                    const int q = ((R - 1 - ll) * m1) % N;
                    const double a = 2.0 * M_PI * (double)q / (double)N;
                    tt[(size_t)ll * R + ((((m1 >> 1) ^ sw) << 1) | (m1 & 1))] = make_float2((float)cos(a), (float)sin(a));
                }
            }
            CK(cudaMalloc(&s.d_tw_tma, tt.size() * sizeof(float2)));
            CK(cudaMemcpy(s.d_tw_tma, tt.data(), tt.size() * sizeof(float2), cudaMemcpyHostToDevice));
        }
    } else {
        tperm = hp;
        tw.resize(N);
        for (int q = 0; q < N; ++q) {
            const double a = 2.0 * M_PI * (double)q / (double)N;
            tw[q] = make_float2((float)cos(a), (float)sin(a));
        }
        if (2 * (size_t)N * sizeof(float2) > 48 * 1024) return RCB_EUNSUPPORTED;
    }
    if (R) {
        std::vector<float2> twg((size_t)N);
        for (int q = 0; q < N; ++q) {
            const double a = 2.0 * M_PI * (double)q / (double)N;
            twg[q] = make_float2((float)cos(a), (float)sin(a));
        }
        CK(cudaMalloc(&s.d_taps_gen, hp.size() * sizeof(float)));
        CK(cudaMalloc(&s.d_tw_gen, twg.size() * sizeof(float2)));
        CK(cudaMemcpy(s.d_taps_gen, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(s.d_tw_gen, twg.data(), twg.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&s.d_taps, tperm.size() * sizeof(float)));
    CK(cudaMalloc(&s.d_tw, tw.size() * sizeof(float2)));
    CK(cudaMemcpyAsync(s.d_taps, tperm.data(), tperm.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(s.d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&s.d_hist[b], (size_t)P * N * sizeof(float2)));
        CK(cudaMemsetAsync(s.d_hist[b], 0, (size_t)P * N * sizeof(float2), h->stream));
    }
    s.hist_cur = 0;
    CK(cudaMalloc(&s.d_counter, sizeof(int)));
    CK(cudaMalloc(&s.d_counter2, 2 * sizeof(int)));
    CK(cudaMemsetAsync(s.d_counter2, 0, 2 * sizeof(int), h->stream));
    CK(cudaMalloc(&s.d_zeros, (size_t)N * sizeof(float2)));
    CK(cudaMemsetAsync(s.d_zeros, 0, (size_t)N * sizeof(float2), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (R) {
        PfbParams p{};
        p.P = P;
        p.N = N;
        int rc = pfb_launch_fast(h, p, true);
        if (rc) return rc;
    }
    s.configured = true;
    return RCB_OK;
}

extern "C" int rcb_pfb_set_out_block(rcb_t* h, int frames) {
    if (!h) return RCB_EINVAL;
    if (!h->pfb.configured) return RCB_ESTATE;
    if (frames == 0) {
        if (h->pfb.oblock_log2) pfb_free_stages(h);
        h->pfb.oblock_log2 = 0;
        return RCB_OK;
    }
    if (frames < 8 || (frames & (frames - 1))) return RCB_EINVAL;  // power of two >= 8
    if (frames > (1 << 24)) return RCB_ERANGE;
    int k = 0;
    while ((1 << k) < frames) ++k;
    if (k != h->pfb.oblock_log2) pfb_free_stages(h);  // the host pipeline's chunk is a whole number of blocks
    h->pfb.oblock_log2 = k;
    return RCB_OK;
}

extern "C" int rcb_pfb_reset(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    auto& s = h->pfb;
    if (!s.configured) return RCB_ESTATE;
    CK(cudaSetDevice(h->device));
    for (int b = 0; b < 2; ++b) CK(cudaMemsetAsync(s.d_hist[b], 0, (size_t)s.P * s.N * sizeof(float2), h->stream));
    s.hist_valid = 0;
    if (s.d_u_prev) CK(cudaMemsetAsync(s.d_u_prev, 0, (size_t)s.N * sizeof(float2), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RCB_OK;
}

extern "C" int rcb_pfb_set_input_format(rcb_t* h, int fmt, float offset, float scale) {
    if (!h) return RCB_EINVAL;
    auto& s = h->pfb;
    if (!s.configured) return RCB_ESTATE;
    if (fmt != 0 && fmt != RCB_FMT_U8 && fmt != RCB_FMT_S8 && fmt != RCB_FMT_S16) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    pfb_free_stages(h);  // the staging buffers are sized for the wire format
    s.in_fmt = fmt;
    s.in_off = fmt ? offset : 0.f;
    s.in_scale = fmt ? scale : 1.f;
    s.hist_valid = 0;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMemsetAsync(s.d_hist[b], 0, (size_t)s.P * s.N * sizeof(float2), h->stream));
        if (fmt && !s.d_hist_raw[b]) CK(cudaMalloc(&s.d_hist_raw[b], (size_t)s.P * s.N * 4));
        if (s.d_hist_raw[b]) CK(cudaMemsetAsync(s.d_hist_raw[b], 0, (size_t)s.P * s.N * 4, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return RCB_OK;
}

extern "C" int rcb_pfb_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem, void* out_iq, void* out_fm,
                               size_t out_stride, int out_mem, size_t* nout) {
    if (!h) return RCB_EINVAL;
    auto& s = h->pfb;
    if (!s.configured) return RCB_ESTATE;
    if (nout) *nout = 0;
    if ((!iq && nsamples) || nsamples % (size_t)s.N) return RCB_EINVAL;
    const size_t frames = nsamples / (size_t)s.N;
    if (frames > 0x7fffff00u) return RCB_ERANGE;
    if ((s.mode & RCB_OUT_IQ) && !out_iq) return RCB_EINVAL;
    if ((s.mode & RCB_OUT_FM) && !out_fm) return RCB_EINVAL;
    if (out_stride < frames && !s.oblock_log2) return RCB_EINVAL;
    if ((in_mem != RCB_MEM_HOST && in_mem != RCB_MEM_DEVICE) || (out_mem != RCB_MEM_HOST && out_mem != RCB_MEM_DEVICE))
        return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    if (frames == 0) return RCB_OK;

    const size_t bps = pfb_in_bytes_per_sample(s.in_fmt);  // complex64, or the wire format set by rcb_pfb_set_input_format
    if (in_mem == RCB_MEM_DEVICE && ((uintptr_t)iq & 15)) return RCB_EINVAL;
    if (in_mem == RCB_MEM_DEVICE && out_mem == RCB_MEM_DEVICE) {
        int rc = pfb_run_device(h, iq, frames, (float2*)out_iq, (float*)out_fm, out_stride);
        if (rc) return rc;
        if (nout) *nout = frames;
        return RCB_OK;
    }

    // host-facing path: chunked 3-stage pipeline  H2D (s_in) | kernel (stream) | D2H (s_out)
    // Blocked layout (rcb_pfb_set_out_block) with host outputs: the pipeline chunk is a whole number of time blocks, a
    // chunk's staging buffer is then bit-for-bit the piece of the host buffer it belongs to, and the D2H is ONE
    // contiguous copy per chunk (instead of a 2-D copy of nchans row pieces).
    const size_t blk = s.oblock_log2 ? ((size_t)1 << s.oblock_log2) : 0;
    int rc = pfb_ensure_stages(h);
    if (rc) return rc;
    const size_t cf = s.chunk_frames;
    if (blk && (cf % blk)) return RCB_EUNSUPPORTED;
    size_t done = 0;
    int si = 0;
    while (done < frames) {
        const size_t f = std::min(cf, frames - done);
        Stage& st = s.st[si];
        const void* d_x;
        if (in_mem == RCB_MEM_HOST) {
            if (st.used) CK(cudaStreamWaitEvent(h->s_in, st.ev_k, 0));  // previous kernel finished reading d_in
            CK(cudaMemcpyAsync(st.d_in, (const char*)iq + done * s.N * bps, f * s.N * bps, cudaMemcpyHostToDevice, h->s_in));
            h->stats.h2d_bytes += f * s.N * bps;
            CK(cudaEventRecord(st.ev_in, h->s_in));
            CK(cudaStreamWaitEvent(h->stream, st.ev_in, 0));
            d_x = st.d_in;
        } else {
            d_x = (const char*)iq + done * s.N * bps;
        }
        float2* d_iq;
        float* d_fm;
        size_t dstride;
        if (out_mem == RCB_MEM_HOST) {
            if (st.used) CK(cudaStreamWaitEvent(h->stream, st.ev_out, 0));  // previous D2H of this stage done
            d_iq = st.d_iq;
            d_fm = st.d_fm;
            dstride = cf;
        } else if (blk) {  // device output, blocked: this chunk starts at block done / blk
            d_iq = out_iq ? (float2*)out_iq + (done / blk) * s.N * blk : nullptr;
            d_fm = out_fm ? (float*)out_fm + (done / blk) * s.N * blk : nullptr;
            dstride = out_stride;
        } else {
            d_iq = out_iq ? (float2*)out_iq + done : nullptr;
            d_fm = out_fm ? (float*)out_fm + done : nullptr;
            dstride = out_stride;
        }
        rc = pfb_run_device(h, d_x, f, d_iq, d_fm, dstride);
        if (rc) return rc;
        CK(cudaEventRecord(st.ev_k, h->stream));
        if (out_mem == RCB_MEM_HOST) {
            CK(cudaStreamWaitEvent(h->s_out, st.ev_k, 0));
            if (blk) {
                const size_t nb = (f + blk - 1) / blk, off = (done / blk) * s.N * blk, cnt = nb * s.N * blk;
                if (s.mode & RCB_OUT_FM) {
                    CK(cudaMemcpyAsync((float*)out_fm + off, st.d_fm, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += cnt * sizeof(float);
                }
                if (s.mode & RCB_OUT_IQ) {
                    CK(cudaMemcpyAsync((float2*)out_iq + off, st.d_iq, cnt * sizeof(float2), cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += cnt * sizeof(float2);
                }
            } else {
                if (s.mode & RCB_OUT_FM) {
                    CK(cudaMemcpy2DAsync((float*)out_fm + done, out_stride * sizeof(float), st.d_fm, cf * sizeof(float),
                                         f * sizeof(float), s.N, cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += f * s.N * sizeof(float);
                }
                if (s.mode & RCB_OUT_IQ) {
                    CK(cudaMemcpy2DAsync((float2*)out_iq + done, out_stride * sizeof(float2), st.d_iq, cf * sizeof(float2),
                                         f * sizeof(float2), s.N, cudaMemcpyDeviceToHost, h->s_out));
                    h->stats.d2h_bytes += f * s.N * sizeof(float2);
                }
            }
            CK(cudaEventRecord(st.ev_out, h->s_out));
        }
        st.used = true;
        done += f;
        si = (si + 1) % kStages;
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->s_out));
    if (nout) *nout = frames;
    return RCB_OK;
}

// ---- several independent streams of one shape in ONE launch --------------------------------------------------------
namespace {
template <int R, int PT>
int pfb_launch_multi_t(rcb_t* lead, const PfbParams* d_ps, int nstreams, int frames) {
    using G = PfbTmaGeom<R, 8, PFB_OUT_FM>;
    auto kern = pfb_fm_tma_multi_kernel<R, 8, true, PT, PFB_OUT_FM>;
    rcb_t* h = lead;
    static bool attr_dev[64] = {};
    static int per_sm[64] = {};
    const int di = h->device & 63;
    if (!attr_dev[di]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::THREADS, G::smem_bytes));
        per_sm[di] = std::max(nb, 1);
        attr_dev[di] = true;
    }
    const int NI = (frames + G::FPI - 1) / G::FPI;
    const int grid = std::max(1, std::min(NI, per_sm[di] * h->sm_count));
    kern<<<grid, G::THREADS, G::smem_bytes, h->stream>>>(d_ps, nstreams);
    CKL(h);
    return RCB_OK;
}
template <int R>
int pfb_launch_multi_r(rcb_t* lead, const PfbParams* d_ps, int nstreams, int frames) {
    switch (lead->pfb.PT) {
        case 1: return pfb_launch_multi_t<R, 1>(lead, d_ps, nstreams, frames);
        case 2: return pfb_launch_multi_t<R, 2>(lead, d_ps, nstreams, frames);
        case 4: return pfb_launch_multi_t<R, 4>(lead, d_ps, nstreams, frames);
        case 8: return pfb_launch_multi_t<R, 8>(lead, d_ps, nstreams, frames);
        case 16: return pfb_launch_multi_t<R, 16>(lead, d_ps, nstreams, frames);
    }
    return RCB_EUNSUPPORTED;
}
constexpr int kMultiSlots = 4;
}  // namespace

extern "C" int rcb_pfb_process_multi(rcb_t* const* hs, int nstreams, const void* const* iq, size_t nsamples,
                                     void* const* out_fm, size_t out_stride) {
    if (!hs || nstreams < 1 || !iq || !out_fm) return RCB_EINVAL;
    rcb_t* h = hs[0];
    if (!h) return RCB_EINVAL;
    auto& s0 = h->pfb;
    if (!s0.configured) return RCB_ESTATE;
    if (nsamples % (size_t)s0.N) return RCB_EINVAL;
    const size_t frames = nsamples / (size_t)s0.N;
    if (frames > 0x7fffff00u) return RCB_ERANGE;
    bool batched = (s0.R == 16 && s0.mode == RCB_OUT_FM && s0.use_tma && s0.PT <= 16 && s0.in_fmt == 0 && nstreams <= 1024);
    for (int i = 0; i < nstreams; ++i) {
        rcb_t* g = hs[i];
        if (!g || !iq[i] || !out_fm[i]) return RCB_EINVAL;
        auto& s = g->pfb;
        if (!s.configured) return RCB_ESTATE;
        if (g->device != h->device || s.N != s0.N || s.L != s0.L || s.mode != s0.mode || s.oblock_log2 != s0.oblock_log2 ||
            s.in_fmt != s0.in_fmt)
            return RCB_EINVAL;   // one shape, one device per batch
        if (out_stride < frames && !s.oblock_log2) return RCB_EINVAL;
        const bool aligned32 = (s.oblock_log2 > 0 || (out_stride % 8) == 0) && (((uintptr_t)out_fm[i] & 31) == 0) &&
                               (((uintptr_t)iq[i] & 15) == 0);
        batched = batched && aligned32;
    }
    CK(cudaSetDevice(h->device));
    if (frames == 0) return RCB_OK;
    if (!batched) {  // any other shape: the streams one after the other (same results, nstreams launches)
        for (int i = 0; i < nstreams; ++i) {
            int rc = rcb_pfb_process(hs[i], iq[i], nsamples, RCB_MEM_DEVICE, nullptr, out_fm[i], out_stride, RCB_MEM_DEVICE,
                                     nullptr);
            if (rc) return rc;
        }
        return RCB_OK;
    }
    // parameter blocks of all streams -> one pinned slot -> device
    const size_t per = sizeof(PfbParams) + sizeof(PfbHistJob);
    if (s0.multi_cap < (size_t)nstreams) {
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(s0.d_multi);
        if (s0.h_multi) cudaFreeHost(s0.h_multi);
        cudaFree(s0.d_mcounters);
        s0.d_multi = nullptr;
        s0.h_multi = nullptr;
        s0.d_mcounters = nullptr;
        s0.multi_cap = 0;
        const size_t cap = (size_t)nstreams;
        CK(cudaMalloc(&s0.d_multi, kMultiSlots * cap * per));
        CK(cudaHostAlloc(&s0.h_multi, kMultiSlots * cap * per, cudaHostAllocDefault));
        CK(cudaMalloc(&s0.d_mcounters, cap * sizeof(int)));
        s0.multi_cap = cap;
    }
    const int slot = (int)(s0.multi_no++ % kMultiSlots);
    if (!s0.ev_mslot[slot]) CK(cudaEventCreateWithFlags(&s0.ev_mslot[slot], cudaEventDisableTiming));
    else CK(cudaEventSynchronize(s0.ev_mslot[slot]));
    unsigned char* hb = s0.h_multi + (size_t)slot * s0.multi_cap * per;
    unsigned char* db = s0.d_multi + (size_t)slot * s0.multi_cap * per;
    PfbParams* hp = reinterpret_cast<PfbParams*>(hb);
    PfbHistJob* hj = reinterpret_cast<PfbHistJob*>(hb + (size_t)nstreams * sizeof(PfbParams));
    const PfbParams* dp = reinterpret_cast<const PfbParams*>(db);
    const PfbHistJob* dj = reinterpret_cast<const PfbHistJob*>(db + (size_t)nstreams * sizeof(PfbParams));
    for (int i = 0; i < nstreams; ++i) {
        rcb_t* g = hs[i];
        auto& s = g->pfb;
        // whatever is queued on the stream of handle i (its previous block, the producer of its input) comes first
        if (i > 0) {
            if (!s.ev_multi) CK(cudaEventCreateWithFlags(&s.ev_multi, cudaEventDisableTiming));
            CK(cudaEventRecord(s.ev_multi, g->stream));
            CK(cudaStreamWaitEvent(h->stream, s.ev_multi, 0));
        }
        PfbParams p{};
        p.x = (const float2*)iq[i];
        p.hist = s.d_hist[s.hist_cur];
        p.taps = s.d_taps;
        p.twiddle = s.d_tw_tma;
        p.taps_kc = s.d_taps_kc;
        p.zeros = s.d_zeros;
        p.work_counter = s0.d_mcounters + i;
        p.out_fm = (float*)out_fm[i];
        p.out_iq = nullptr;
        p.ostride = (long long)out_stride;
        p.oblock_log2 = s.oblock_log2;
        p.T = (int)frames;
        p.P = s.P;
        p.N = s.N;
        p.gain = s.gain;
        hp[i] = p;
        hj[i].old_hist = s.d_hist[s.hist_cur];
        hj[i].x = (const float2*)iq[i];
        hj[i].new_hist = s.d_hist[s.hist_cur ^ 1];
    }
    CK(cudaMemcpyAsync(db, hb, (size_t)nstreams * per, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(s0.ev_mslot[slot], h->stream));
    CK(cudaMemsetAsync(s0.d_mcounters, 0, (size_t)nstreams * sizeof(int), h->stream));
    int rc = pfb_launch_multi_r<16>(h, dp, nstreams, (int)frames);
    if (rc) return rc;
    const long long cap = (long long)s0.P * s0.N;
    dim3 hg((unsigned)((cap + 255) / 256), (unsigned)nstreams);
    pfb_hist_multi_kernel<<<hg, 256, 0, h->stream>>>(dj, (long long)nsamples, cap);
    CKL(h);
    // the other handles' streams continue after the batch
    if (!s0.ev_multi) CK(cudaEventCreateWithFlags(&s0.ev_multi, cudaEventDisableTiming));
    CK(cudaEventRecord(s0.ev_multi, h->stream));
    for (int i = 0; i < nstreams; ++i) {
        rcb_t* g = hs[i];
        if (i > 0) CK(cudaStreamWaitEvent(g->stream, s0.ev_multi, 0));
        g->pfb.hist_cur ^= 1;
        g->stats.samples_in += frames * (uint64_t)s0.N;
        g->stats.channel_samples += frames * (uint64_t)s0.N;
    }
    return RCB_OK;
}

// =================================================================================================
// K2  DDC bank
// =================================================================================================
namespace {
int ddc_upload_taps(rcb_t* h, DdcChan& c) {
    const double w = 2.0 * M_PI * c.center_freq / c.samp_rate;
    std::vector<float2> rev(c.ntaps);
    for (int r = 0; r < c.ntaps; ++r) {
        const int k = c.ntaps - 1 - r;
        const double a = fmod(w * (double)k, 2.0 * M_PI);
        rev[r] = make_float2((float)((double)c.taps[k] * cos(a)), (float)((double)c.taps[k] * sin(a)));
    }
    if (c.d_ctaps_rev) cudaFree(c.d_ctaps_rev);
    c.d_ctaps_rev = nullptr;
    CK(cudaMalloc(&c.d_ctaps_rev, sizeof(float2) * c.ntaps));
    CK(cudaMemcpyAsync(c.d_ctaps_rev, rev.data(), sizeof(float2) * c.ntaps, cudaMemcpyHostToDevice, h->stream));
    std::vector<float4> rev4(c.ntaps);
    for (int r = 0; r < c.ntaps; ++r) rev4[r] = make_float4(rev[r].x, rev[r].y, -rev[r].y, rev[r].x);
    if (c.d_ctaps4_rev) cudaFree(c.d_ctaps4_rev);
    c.d_ctaps4_rev = nullptr;
    CK(cudaMalloc(&c.d_ctaps4_rev, sizeof(float4) * c.ntaps));
    CK(cudaMemcpyAsync(c.d_ctaps4_rev, rev4.data(), sizeof(float4) * c.ntaps, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double cyc = c.center_freq * (double)c.decim / c.samp_rate;
    cyc -= floor(cyc);
    c.cyc = cyc;
    static uint64_t ver_counter = 0;
    c.taps_ver = ++ver_counter;
    return RCB_OK;
}
int ddc_ensure_hist(rcb_t* h) {
    auto& d = h->ddc;
    if (d.d_hist[0]) return RCB_OK;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&d.d_hist[b], sizeof(float2) * kDdcHistCap));
        CK(cudaMemsetAsync(d.d_hist[b], 0, sizeof(float2) * kDdcHistCap, h->stream));
    }
    return RCB_OK;
}
double ddc_phase_at(const DdcChan& c, uint64_t i) {
    double p = c.phase_base + frac_mul(c.cyc, i - c.i_base);
    p -= floor(p);
    return p;
}
}  // namespace

extern "C" int rcb_ddc_open(rcb_t* h, int decim, const float* taps, int ntaps, double center_freq, double samp_rate,
                            int out_mask, float fm_gain, int* chan_id) {
    if (!h || !taps || !chan_id || decim < 1 || ntaps < 1 || !(samp_rate > 0)) return RCB_EINVAL;
    if (!(out_mask & (RCB_OUT_IQ | RCB_OUT_FM)) || (out_mask & ~(RCB_OUT_IQ | RCB_OUT_FM))) return RCB_EINVAL;
    if (ntaps - 1 > kDdcHistCap) return RCB_EUNSUPPORTED;
    CK(cudaSetDevice(h->device));
    int rc = ddc_ensure_hist(h);
    if (rc) return rc;
    DdcChan c;
    c.id = h->ddc.next_id++;
    c.decim = decim;
    c.ntaps = ntaps;
    c.out_mask = out_mask;
    c.gain = fm_gain;
    c.center_freq = center_freq;
    c.samp_rate = samp_rate;
    c.taps.assign(taps, taps + ntaps);
    // all channels of a source with the same decimation share one decimation grid (outputs at stream
    // positions that are multiples of decim): a channel opened mid-stream starts at the next grid point.
    // (A GNU Radio channel block starts at whatever sample its ZMQ SUB happens to receive first.)
    c.start_sample = ((h->ddc.n_consumed + (uint64_t)decim - 1) / (uint64_t)decim) * (uint64_t)decim;
    rc = ddc_upload_taps(h, c);
    if (rc) return rc;
    CK(cudaMalloc(&c.d_prev, 2 * sizeof(float2)));
    CK(cudaMemsetAsync(c.d_prev, 0, 2 * sizeof(float2), h->stream));
    h->ddc.chans[c.id] = c;
    *chan_id = c.id;
    return RCB_OK;
}

extern "C" int rcb_ddc_retune(rcb_t* h, int chan_id, double center_freq) {
    if (!h) return RCB_EINVAL;
    auto it = h->ddc.chans.find(chan_id);
    if (it == h->ddc.chans.end()) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    DdcChan& c = it->second;
    // keep the rotator phase continuous, like set_center_freq -> rotator::set_phase_incr
    c.phase_base = ddc_phase_at(c, c.i_next);
    c.i_base = c.i_next;
    c.center_freq = center_freq;
    return ddc_upload_taps(h, c);
}

extern "C" int rcb_ddc_set_taps(rcb_t* h, int chan_id, const float* taps, int ntaps) {
    if (!h || !taps || ntaps < 1) return RCB_EINVAL;
    if (ntaps - 1 > kDdcHistCap) return RCB_EUNSUPPORTED;
    auto it = h->ddc.chans.find(chan_id);
    if (it == h->ddc.chans.end()) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    DdcChan& c = it->second;
    c.taps.assign(taps, taps + ntaps);
    c.ntaps = ntaps;
    return ddc_upload_taps(h, c);
}

extern "C" int rcb_ddc_set_tensor_cores(rcb_t* h, int mode, int seg) {
    if (!h || mode < 0 || mode > 2 || seg < 0) return RCB_EINVAL;
    if (mode == 2 && seg > 3) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->ddc.use_mma = (mode != 0);
    if (mode) h->ddc.mma_gen = (mode == 2) ? 1 : 2;
    if (mode == 1 && seg) h->ddc.mma_seg_len = seg;
    if (mode == 2 && seg) h->ddc.mma_nseg = seg;
    return RCB_OK;
}

extern "C" int rcb_ddc_tensor_core_launches(rcb_t* h, uint64_t* launches) {
    if (!h || !launches) return RCB_EINVAL;
    *launches = h->ddc.mma_launches;
    return RCB_OK;
}

extern "C" int rcb_ddc_close(rcb_t* h, int chan_id) {
    if (!h) return RCB_EINVAL;
    auto it = h->ddc.chans.find(chan_id);
    if (it == h->ddc.chans.end()) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(it->second.d_ctaps_rev);
    cudaFree(it->second.d_ctaps4_rev);
    cudaFree(it->second.d_out_iq);
    cudaFree(it->second.d_out_fm);
    cudaFree(it->second.d_prev);
    h->ddc.chans.erase(it);
    return RCB_OK;
}

namespace {

constexpr int kDdcSlots = 4;            // pinned parameter blocks in flight (one per chunk)
constexpr size_t kDdcChunk = 1u << 22;  // host input is staged in chunks of this many samples (H2D overlaps compute)

// one device-resident piece of the wideband stream through every open channel; outputs are APPENDED to what this
// rcb_ddc_process call has produced so far (c.nout_last)
int ddc_process_chunk(rcb_t* h, const float2* d_x, size_t nsamples) {
    auto& d = h->ddc;
    const size_t M = d.chans.size();
    if (M) {
        if (d.chans_cap < M) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_chans_ring);
            if (d.h_chans_ring) cudaFreeHost(d.h_chans_ring);
            d.d_chans_ring = nullptr;
            d.h_chans_ring = nullptr;
            d.chans_cap = 0;
            const size_t cap = std::max<size_t>(16, M * 2);
            CK(cudaMalloc(&d.d_chans_ring, kDdcSlots * cap * sizeof(DdcChanDev)));
            CK(cudaHostAlloc(&d.h_chans_ring, kDdcSlots * cap * sizeof(DdcChanDev), cudaHostAllocDefault));
            d.chans_cap = cap;
        }
        // parameter blocks live in a ring of pinned slots: no stream synchronisation per block, a slot is reused only
        // after the upload that read it has completed
        const int slot = (int)(d.slot_no++ % kDdcSlots);
        if (!d.ev_slot[slot]) CK(cudaEventCreateWithFlags(&d.ev_slot[slot], cudaEventDisableTiming));
        else CK(cudaEventSynchronize(d.ev_slot[slot]));
        d.h_chans = d.h_chans_ring + (size_t)slot * d.chans_cap;
        d.d_chans = d.d_chans_ring + (size_t)slot * d.chans_cap;
        const uint64_t n0 = d.n_consumed, n1 = d.n_consumed + nsamples;
        size_t max_nout = 0;
        size_t ci = 0;
        bool any_fm = false;
        std::vector<const DdcChan*> order;
        order.reserve(M);
        for (auto& kv : d.chans) {
            DdcChan& c = kv.second;
            order.push_back(&c);
            // outputs i with newest sample  start + i*D  in [n0, n1)
            const uint64_t s_next = c.start_sample + c.i_next * (uint64_t)c.decim;
            size_t nout = 0;
            if (s_next < n1) nout = (size_t)((n1 - 1 - s_next) / (uint64_t)c.decim) + 1;
            const size_t base = c.nout_last;  // appended after the earlier chunks of this call
            if (c.out_cap < base + nout) return RCB_ERANGE;  // (capacity is reserved by rcb_ddc_process)
            DdcChanDev& dv = d.h_chans[ci++];
            dv.ctaps_rev = c.d_ctaps_rev;
            dv.ctaps4_rev = c.d_ctaps4_rev;
            dv.out_iq = c.d_out_iq + base;
            dv.out_fm = (c.out_mask & RCB_OUT_FM) ? c.d_out_fm + base : nullptr;
            dv.prev = c.d_prev + (c.prev_cur & 1);          // FM carry in
            dv.prev_out = c.d_prev + ((c.prev_cur ^ 1) & 1);  // carry out (double buffered: ddc_post_kernel reads and
            if (nout > 0) c.prev_cur ^= 1;                    // writes it in the same launch)
            dv.cyc = c.cyc;
            dv.phase0 = ddc_phase_at(c, c.i_next);
            dv.s_first = (long long)s_next - (long long)n0;
            dv.s_open = (long long)c.start_sample - (long long)n0;
            dv.ntaps = c.ntaps;
            dv.decim = c.decim;
            dv.nout = (int)nout;
            dv.gain = c.gain;
            // fast path needs every window sample of this block to lie at / after the channel's first sample
            dv.fast = (nout > 0 && dv.s_open <= dv.s_first - (long long)(c.ntaps - 1) &&
                       (size_t)(7 * c.decim + c.ntaps) * sizeof(float2) <= 160 * 1024) ? 1 : 0;
            c.nout_last = base + nout;
            c.i_next += nout;
            max_nout = std::max(max_nout, nout);
            any_fm |= (dv.out_fm != nullptr);
            h->stats.channel_samples += nout;
        }
        // group the fast channels by (decim, ntaps, s_first, nout), 16 per work group
        std::map<std::vector<long long>, std::vector<int>> buckets;
        bool any_slow = false;
        for (size_t i = 0; i < M; ++i) {
            const DdcChanDev& dv = d.h_chans[i];
            if (dv.fast)
                buckets[{dv.decim, dv.ntaps, dv.s_first, dv.nout}].push_back((int)i);
            else if (dv.nout > 0)
                any_slow = true;
        }
        size_t ngroups = 0;
        for (auto& b : buckets) ngroups += (b.second.size() + 15) / 16;
        // buckets that go to the tensor cores: outputs [o_head, nout) by ddc_mma_kernel, the head (windows reaching into
        // the history buffer) stays on ddc_tile_kernel
        struct MmaPlan {
            int o_head, lead, kchunks, ng;
            size_t b_off;  // floats into d_mma_b
            const float2* base;
        };
        std::map<std::vector<long long>, MmaPlan> mma_plans;
        size_t mma_groups = 0, mma_b_floats = 0;
        if (d.use_mma && tmap_encoder()) {
            for (auto& b : buckets) {
                const long long decim = b.first[0], ntaps = b.first[1], s_first = b.first[2], nout = b.first[3];
                if ((int)b.second.size() < kDdcMmaMinChans || (decim & 1)) continue;
                const long long need = std::max<long long>(0, ntaps - s_first);   // window start >= 1
                const long long o_head = (need + decim - 1) / decim;
                if (nout - o_head < 64) continue;
                MmaPlan pl;
                pl.o_head = (int)o_head;
                const long long w0 = s_first + o_head * decim - (ntaps - 1);
                const float2* base = d_x + w0;
                pl.lead = (reinterpret_cast<uintptr_t>(base) & 15) ? 1 : 0;
                pl.base = base - pl.lead;
                if (reinterpret_cast<uintptr_t>(pl.base) & 15) continue;  // input not even 8-byte aligned
                pl.kchunks = (int)((2 * (ntaps + pl.lead) + 31) / 32);
                pl.ng = (int)((b.second.size() + 63) / 64);
                pl.b_off = mma_b_floats;
                mma_b_floats += (size_t)2 * pl.ng * 128 * pl.kchunks * 32;
                mma_groups += pl.ng;
                mma_plans[b.first] = pl;
            }
        }
        if (mma_groups > d.mgroups_cap) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_mgroups_ring);
            if (d.h_mgroups_ring) cudaFreeHost(d.h_mgroups_ring);
            d.d_mgroups_ring = nullptr;
            d.h_mgroups_ring = nullptr;
            d.mgroups_cap = 0;
            const size_t cap = std::max<size_t>(4, mma_groups * 2);
            CK(cudaMalloc(&d.d_mgroups_ring, kDdcSlots * cap * sizeof(DdcMmaGroupDev)));
            CK(cudaHostAlloc(&d.h_mgroups_ring, kDdcSlots * cap * sizeof(DdcMmaGroupDev), cudaHostAllocDefault));
            d.mgroups_cap = cap;
        }
        if (mma_b_floats > d.mma_b_cap) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_mma_b);
            d.d_mma_b = nullptr;
            d.mma_b_cap = 0;
            d.mma_b_sig.clear();
            CK(cudaMalloc(&d.d_mma_b, mma_b_floats * sizeof(float)));
            d.mma_b_cap = mma_b_floats;
        }
        bool mma_pack = true;
        if (mma_groups) {
            std::vector<long long> sig;
            for (auto& b : buckets) {
                auto mp = mma_plans.find(b.first);
                if (mp == mma_plans.end()) continue;
                sig.push_back(b.first[0]);
                sig.push_back(b.first[1]);
                sig.push_back(mp->second.lead);
                sig.push_back((long long)mp->second.b_off);
                sig.push_back((long long)b.second.size());
                for (int idx : b.second) {
                    sig.push_back(order[(size_t)idx]->id);
                    sig.push_back((long long)order[(size_t)idx]->taps_ver);
                }
            }
            mma_pack = (sig != d.mma_b_sig);
            if (mma_pack) d.mma_b_sig.swap(sig);
        }
        DdcMmaGroupDev* h_mgroups = d.h_mgroups_ring ? d.h_mgroups_ring + (size_t)slot * d.mgroups_cap : nullptr;
        DdcMmaGroupDev* d_mgroups = d.d_mgroups_ring ? d.d_mgroups_ring + (size_t)slot * d.mgroups_cap : nullptr;
        if (ngroups > d.groups_cap) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_groups_ring);
            if (d.h_groups_ring) cudaFreeHost(d.h_groups_ring);
            d.d_groups_ring = nullptr;
            d.h_groups_ring = nullptr;
            d.groups_cap = 0;
            const size_t cap = std::max<size_t>(8, ngroups * 2);
            CK(cudaMalloc(&d.d_groups_ring, kDdcSlots * cap * sizeof(DdcGroupDev)));
            CK(cudaHostAlloc(&d.h_groups_ring, kDdcSlots * cap * sizeof(DdcGroupDev), cudaHostAllocDefault));
            d.groups_cap = cap;
        }
        d.h_groups = d.h_groups_ring + (size_t)slot * d.groups_cap;
        d.d_groups = d.d_groups_ring + (size_t)slot * d.groups_cap;
        CK(cudaMemcpyAsync(d.d_chans, d.h_chans, M * sizeof(DdcChanDev), cudaMemcpyHostToDevice, h->stream));
        if (max_nout) {
            size_t gi = 0;
            for (auto& b : buckets) {
                const size_t first_group = gi;
                for (size_t off = 0; off < b.second.size(); off += 16) {
                    DdcGroupDev& g = d.h_groups[gi++];
                    g.nch = (int)std::min<size_t>(16, b.second.size() - off);
                    for (int u = 0; u < 16; ++u) g.ch[u] = (u < g.nch) ? b.second[off + u] : -1;
                    g.decim = (int)b.first[0];
                    g.ntaps = (int)b.first[1];
                    g.s_first = b.first[2];
                    g.nout = (int)b.first[3];
                    auto mp = mma_plans.find(b.first);
                    if (mp != mma_plans.end()) g.nout = 0;  // ddc_head_kernel + ddc_mma_kernel cover the bucket
                }
                (void)first_group;
            }
            if (mma_groups) {
                size_t mg = 0;
                for (auto& b : buckets) {
                    auto mp = mma_plans.find(b.first);
                    if (mp == mma_plans.end()) continue;
                    const MmaPlan& pl = mp->second;
                    for (int q = 0; q < pl.ng; ++q) {
                        DdcMmaGroupDev& g = h_mgroups[mg++];
                        g.nch = (int)std::min<size_t>(64, b.second.size() - (size_t)q * 64);
                        for (int u = 0; u < 64; ++u) g.ch[u] = (u < g.nch) ? b.second[(size_t)q * 64 + u] : -1;
                        g.ncols = ((2 * g.nch + 15) / 16) * 16;
                        g.ntaps = (int)b.first[1];
                        g.decim = (int)b.first[0];
                        g.lead = pl.lead;
                        g.kchunks = pl.kchunks;
                        g.o_head = pl.o_head;
                        g.nout = (int)b.first[3];
                        g.ldb = pl.kchunks * 32;
                        g.nseg = std::max(1, std::min(std::min(3, d.mma_nseg), pl.kchunks));
                        g.seg_len = std::max(2, d.mma_seg_len);
                        g.kq = std::max(1, (2 * g.decim + 16) / 32);
#ifdef RCB_EXPERIMENTS
                        if (const char* e = getenv("RCB_DDC_KQ")) g.kq = std::max(1, atoi(e));
#endif
                    }
                }
                CK(cudaMemcpyAsync(d_mgroups, h_mgroups, mma_groups * sizeof(DdcMmaGroupDev), cudaMemcpyHostToDevice, h->stream));
                static bool mma_attr_dev[64] = {};
                if (!mma_attr_dev[h->device & 63]) {
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kM2Smem));
#ifdef RCB_EXPERIMENTS
                    CK(cudaFuncSetAttribute(ddc_mma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kM2Smem));
#endif
#ifdef RCB_EXPERIMENTS
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<35>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
                    CK(cudaFuncSetAttribute(ddc_mma_kernel<27>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdcMmaSmem));
#endif
                    mma_attr_dev[h->device & 63] = true;
                }
                if (!d.s_aux) {
                    CK(cudaStreamCreateWithFlags(&d.s_aux, cudaStreamNonBlocking));
                    CK(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
                    CK(cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming));
                }
                // the few outputs per channel whose windows start in the history buffer: small CTAs that fit beside the
                // one-per-SM MMA CTAs, on their own stream
                CK(cudaEventRecord(d.ev_fork, h->stream));
                CK(cudaStreamWaitEvent(d.s_aux, d.ev_fork, 0));
                bool any_head = false;
                mg = 0;
                for (auto& b : buckets) {
                    auto mp = mma_plans.find(b.first);
                    if (mp == mma_plans.end()) continue;
                    const MmaPlan& pl = mp->second;
                    if (pl.o_head > 0) {
                        dim3 hg((unsigned)pl.o_head, 64, (unsigned)pl.ng);
                        ddc_head_kernel<<<hg, 128, 0, d.s_aux>>>(d.d_chans, d_mgroups + mg, d_x, (long long)nsamples,
                                                                 d.d_hist[d.hist_cur], kDdcHistCap);
                        CKL(h);
                        any_head = true;
                    }
                    mg += pl.ng;
                }
                if (any_head) CK(cudaEventRecord(d.ev_join, d.s_aux));
                mg = 0;
                for (auto& b : buckets) {
                    auto mp = mma_plans.find(b.first);
                    if (mp == mma_plans.end()) continue;
                    const MmaPlan& pl = mp->second;
                    const int decim = (int)b.first[0], ntaps = (int)b.first[1], nout = (int)b.first[3];
                    const uint64_t ldb = (uint64_t)pl.kchunks * 32;
                    float* bptr = d.d_mma_b + pl.b_off;
                    CUtensorMap tm_a, tm_b;
                    const uint64_t nrows = (uint64_t)(nout - pl.o_head);
                    const uint64_t ad[2] = {(uint64_t)2 * (uint64_t)(ntaps + pl.lead), nrows};
                    const uint64_t as[1] = {(uint64_t)decim * 8};
                    const uint64_t bd[2] = {ldb, (uint64_t)2 * pl.ng * 128};
                    const uint64_t bs[1] = {ldb * 4};
                    const uint32_t box[2] = {32, 128};
                    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
#ifdef RCB_EXPERIMENTS
                    if (const char* e = getenv("RCB_DDC_PROMO"))
                        if (atoi(e) == 256) promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
#endif
                    if (!tmap_encode_f32(&tm_a, 2, pl.base, ad, as, box, CU_TENSOR_MAP_SWIZZLE_128B, promo) ||
                        !tmap_encode_f32(&tm_b, 2, bptr, bd, bs, box, CU_TENSOR_MAP_SWIZZLE_128B))
                        {
                        snprintf(h->err, sizeof(h->err), "ddc: tensor map encode failed");
                        return RCB_ECUDA;
                    }
                    if (mma_pack) {
                        dim3 pg((unsigned)((ldb + 255) / 256), 128, (unsigned)pl.ng);
                        ddc_mma_pack_kernel<<<pg, 256, 0, h->stream>>>(d.d_chans, d_mgroups + mg, pl.ng, bptr);
                        CKL(h);
                    }

                    dim3 grid((unsigned)((nrows + 127) / 128), (unsigned)pl.ng);
                    int dbg = 0;
#ifdef RCB_EXPERIMENTS
                    if (const char* e = getenv("RCB_DDC_DBG")) dbg = atoi(e);
#endif
                    if (d.mma_gen == 2 && dbg == 0)
                        ddc_mma2_kernel<false><<<grid, kM2Threads, kM2Smem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
#ifdef RCB_EXPERIMENTS
                    else if (d.mma_gen == 2 && dbg == 100)
                        ddc_mma2_kernel<true><<<grid, kM2Threads, kM2Smem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
#endif
                    else if (dbg == 0)
                        ddc_mma_kernel<0><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
#ifdef RCB_EXPERIMENTS
                    else if (dbg == 1)
                        ddc_mma_kernel<1><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else if (dbg == 2)
                        ddc_mma_kernel<2><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else if (dbg == 3)
                        ddc_mma_kernel<3><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else if (dbg == 32)
                        ddc_mma_kernel<32><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else if (dbg == 35)
                        ddc_mma_kernel<35><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else if (dbg == 27)
                        ddc_mma_kernel<27><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
                    else
                        ddc_mma_kernel<4><<<grid, kDdcMmaThreads, kDdcMmaSmem, h->stream>>>(tm_a, tm_b, d.d_chans, d_mgroups + mg, pl.ng);
#endif
                    CKL(h);
                    d.mma_launches++;
                    mg += pl.ng;
                }
                if (any_head) CK(cudaStreamWaitEvent(h->stream, d.ev_join, 0));
            }
            if (ngroups) {
                CK(cudaMemcpyAsync(d.d_groups, d.h_groups, ngroups * sizeof(DdcGroupDev), cudaMemcpyHostToDevice, h->stream));
                gi = 0;
                for (auto& b : buckets) {
                    const size_t ng = (b.second.size() + 15) / 16;
                    const int decim = (int)b.first[0], ntaps = (int)b.first[1];
                    int nout = (int)b.first[3];
                    {
                        auto mp = mma_plans.find(b.first);
                        if (mp != mma_plans.end()) nout = 0;
                    }
                    if (nout == 0) {
                        gi += ng;
                        continue;
                    }
                    // output quads per CTA: as many as the group's channel count leaves warps for and smem allows
                    const size_t per_group = std::min<size_t>(16, b.second.size());
                    int oq = per_group <= 4 ? 8 : (per_group <= 8 ? 4 : 2);
                    // a lone channel: 16 outputs x 1 channel per warp (no idle accumulators) when the tile fits
                    const bool lone = (per_group == 1 && (size_t)(127 * decim + ntaps) * sizeof(float2) <= 160 * 1024);
                    const int opw = lone ? 16 : 4;
                    while (oq > 2 && (size_t)((opw * oq - 1) * decim + ntaps) * sizeof(float2) > 160 * 1024) oq >>= 1;
                    const size_t smem = (size_t)((opw * oq - 1) * decim + ntaps) * sizeof(float2);
                    {
                        // per function and per DEVICE (handles share devices): opt in once to the 160 KB cap the tile
                        // sizes above are clamped to - a per-handle "largest so far" would be lowered by another handle
                        static bool tile_attr_dev[64] = {};
                        bool& done = tile_attr_dev[h->device & 63];
                        if (!done) {
                            const int cap = 160 * 1024 + 1024;
                            CK(cudaFuncSetAttribute(ddc_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_tile_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            done = true;
                        }
                    }
                    dim3 grid((unsigned)((nout + opw * oq - 1) / (opw * oq)), (unsigned)ng);
                    // a lone channel with a decimation that is not tiny: frame-per-lane kernel (P MACs per sample load)
                    const int lone_p = (ntaps + decim - 1) / decim;
                    const size_t lone_smem = 16 + (size_t)decim * ((lone_p + 1) & ~1) * sizeof(float2) +
                                             (size_t)(kDdcLoneWarps * (33 - lone_p) + lone_p - 1) * ddc_lone_pitch(decim) * sizeof(float2);
                    bool use_lone = d.use_lone;
#ifdef RCB_EXPERIMENTS
                    if (const char* e = getenv("RCB_DDC_LONE")) use_lone = atoi(e) != 0;
#endif
                    if (b.second.size() == 1 && decim >= 8 && lone_p >= 1 && lone_p <= 8 && lone_smem <= 200 * 1024 &&
                        use_lone) {
                        static bool lone_attr_dev[64] = {};
                        if (!lone_attr_dev[h->device & 63]) {
                            const int cap = 200 * 1024;
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            CK(cudaFuncSetAttribute(ddc_lone_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                            lone_attr_dev[h->device & 63] = true;
                        }
                        const int per_cta = kDdcLoneWarps * (33 - lone_p);
                        const unsigned lg = (unsigned)((nout + per_cta - 1) / per_cta);
                        const int ci0 = b.second[0];
                        const float2* hp = d.d_hist[d.hist_cur];
#define RCB_LONE(PP)                                                                                                   \
    ddc_lone_kernel<PP><<<lg, kDdcLoneWarps * 32, lone_smem, h->stream>>>(d.d_chans, ci0, d_x, (long long)nsamples, hp, \
                                                                         kDdcHistCap)
                        switch (lone_p) {
                            case 1: RCB_LONE(1); break;
                            case 2: RCB_LONE(2); break;
                            case 3: RCB_LONE(3); break;
                            case 4: RCB_LONE(4); break;
                            case 5: RCB_LONE(5); break;
                            case 6: RCB_LONE(6); break;
                            case 7: RCB_LONE(7); break;
                            default: RCB_LONE(8); break;
                        }
#undef RCB_LONE
                    } else if (lone)
                        ddc_tile_kernel<8, 1><<<grid, 256, smem, h->stream>>>(d.d_chans, d.d_groups + gi, d_x, (long long)nsamples,
                                                                              d.d_hist[d.hist_cur], kDdcHistCap);
                    else if (oq == 8)
                        ddc_tile_kernel<8><<<grid, 256, smem, h->stream>>>(d.d_chans, d.d_groups + gi, d_x, (long long)nsamples,
                                                                           d.d_hist[d.hist_cur], kDdcHistCap);
                    else if (oq == 4)
                        ddc_tile_kernel<4><<<grid, 256, smem, h->stream>>>(d.d_chans, d.d_groups + gi, d_x, (long long)nsamples,
                                                                           d.d_hist[d.hist_cur], kDdcHistCap);
                    else
                        ddc_tile_kernel<2><<<grid, 256, smem, h->stream>>>(d.d_chans, d.d_groups + gi, d_x, (long long)nsamples,
                                                                           d.d_hist[d.hist_cur], kDdcHistCap);
                    CKL(h);
                    gi += ng;
                }
            }
            if (any_slow) {
                dim3 grid((unsigned)((max_nout + 8 * kDdcOutPerWarp - 1) / (8 * kDdcOutPerWarp)), (unsigned)M);
                ddc_bank_kernel<<<grid, 256, 0, h->stream>>>(d.d_chans, d_x, (long long)nsamples, d.d_hist[d.hist_cur],
                                                            kDdcHistCap);
                CKL(h);
            }
        }
        CK(cudaEventRecord(d.ev_slot[slot], h->stream));  // both parameter uploads of this slot are behind this point
        // FM demod of this chunk's outputs + FM carry + wideband history update: ONE launch (ddc_post_kernel)
        {
            const unsigned gx = (unsigned)std::max<size_t>((max_nout + 255) / 256, (kDdcHistCap + 255) / 256);
            dim3 g2(gx, (unsigned)M + 1);
            const int nxt = d.hist_cur ^ 1;
            ddc_post_kernel<<<g2, 256, 0, h->stream>>>(d.d_chans, (int)M, d.d_hist[d.hist_cur], d_x, (long long)nsamples,
                                                     d.d_hist[nxt], kDdcHistCap);
            CKL(h);
            d.hist_cur = nxt;
        }
        (void)any_fm;
    } else {
        const int nxt = d.hist_cur ^ 1;
        hist_update_kernel<<<(kDdcHistCap + 255) / 256, 256, 0, h->stream>>>(d.d_hist[d.hist_cur], d_x, (long long)nsamples,
                                                                            d.d_hist[nxt], kDdcHistCap);
        CKL(h);
        d.hist_cur = nxt;
    }
    d.n_consumed += nsamples;
    h->stats.samples_in += nsamples;
    return RCB_OK;
}

}  // namespace

extern "C" int rcb_ddc_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem) {
    if (!h || (!iq && nsamples)) return RCB_EINVAL;
    if (in_mem != RCB_MEM_HOST && in_mem != RCB_MEM_DEVICE) return RCB_EINVAL;
    if (nsamples > ((size_t)1 << 31)) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    auto& d = h->ddc;
    int rc = ddc_ensure_hist(h);
    if (rc) return rc;
    for (auto& kv : d.chans) kv.second.nout_last = 0;
    if (nsamples == 0) return RCB_OK;
    // reserve every channel's output capacity for the whole call (chunks append)
    {
        const uint64_t n1 = d.n_consumed + nsamples;
        bool synced = false;
        for (auto& kv : d.chans) {
            DdcChan& c = kv.second;
            const uint64_t s_next = c.start_sample + c.i_next * (uint64_t)c.decim;
            size_t nout = 0;
            if (s_next < n1) nout = (size_t)((n1 - 1 - s_next) / (uint64_t)c.decim) + 1;
            if (c.out_cap < nout) {
                if (!synced) {
                    CK(cudaStreamSynchronize(h->stream));
                    synced = true;
                }
                cudaFree(c.d_out_iq);
                cudaFree(c.d_out_fm);
                c.d_out_iq = nullptr;
                c.d_out_fm = nullptr;
                c.out_cap = 0;
                const size_t cap = nout + nout / 4 + 16;
                CK(cudaMalloc(&c.d_out_iq, cap * sizeof(float2)));
                CK(cudaMalloc(&c.d_out_fm, cap * sizeof(float)));
                c.out_cap = cap;
            }
        }
    }
    if (in_mem == RCB_MEM_DEVICE && d.in_fmt == 0) return ddc_process_chunk(h, (const float2*)iq, nsamples);
    // host input (and any wire-format input): chunks through two staging buffers, the H2D of chunk k + 1 (copy stream)
    // overlaps the kernels of chunk k.  Wire formats travel as they are (2-4x fewer PCIe bytes) and become complex64 in
    // the staging buffer on the device (convert_iq_kernel, the K5 arithmetic: bit-identical to rcb_convert_iq).
    const size_t csz = std::min(nsamples, kDdcChunk);
    const size_t bps = pfb_in_bytes_per_sample(d.in_fmt);
    for (int i = 0; i < 2; ++i) {
        if (d.in_cap2[i] < csz) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_in2[i]);
            d.d_in2[i] = nullptr;
            d.in_cap2[i] = 0;
            CK(cudaMalloc(&d.d_in2[i], csz * sizeof(float2)));
            d.in_cap2[i] = csz;
        }
        if (d.in_fmt && in_mem == RCB_MEM_HOST && d.raw_cap2[i] < csz * bps) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(d.d_raw2[i]);
            d.d_raw2[i] = nullptr;
            d.raw_cap2[i] = 0;
            CK(cudaMalloc(&d.d_raw2[i], csz * bps));
            d.raw_cap2[i] = csz * bps;
        }
        if (!d.ev_in[i]) CK(cudaEventCreateWithFlags(&d.ev_in[i], cudaEventDisableTiming));
        if (!d.ev_done[i]) CK(cudaEventCreateWithFlags(&d.ev_done[i], cudaEventDisableTiming));
    }
    size_t done = 0;
    int k = 0;
    while (done < nsamples) {
        const int sl = k & 1;
        const size_t n = std::min(csz, nsamples - done);
        const void* d_raw = nullptr;
        if (in_mem == RCB_MEM_HOST) {
            if (d.in_used[sl]) CK(cudaStreamWaitEvent(h->s_in, d.ev_done[sl], 0));  // the kernels that read this buffer are done
            void* dst = d.in_fmt ? d.d_raw2[sl] : (void*)d.d_in2[sl];
            CK(cudaMemcpyAsync(dst, (const char*)iq + done * bps, n * bps, cudaMemcpyHostToDevice, h->s_in));
            h->stats.h2d_bytes += n * bps;
            CK(cudaEventRecord(d.ev_in[sl], h->s_in));
            CK(cudaStreamWaitEvent(h->stream, d.ev_in[sl], 0));
            d_raw = dst;
        } else {
            d_raw = (const char*)iq + done * bps;
        }
        if (d.in_fmt) {
            const unsigned grid = (unsigned)((n + 1023) / 1024);
            if (d.in_fmt == RCB_FMT_U8)
                convert_iq_kernel<uint8_t><<<grid, 256, 0, h->stream>>>((const uint8_t*)d_raw, d.d_in2[sl], (long long)n, d.in_off, d.in_scale);
            else if (d.in_fmt == RCB_FMT_S8)
                convert_iq_kernel<int8_t><<<grid, 256, 0, h->stream>>>((const int8_t*)d_raw, d.d_in2[sl], (long long)n, d.in_off, d.in_scale);
            else
                convert_iq_kernel<int16_t><<<grid, 256, 0, h->stream>>>((const int16_t*)d_raw, d.d_in2[sl], (long long)n, d.in_off, d.in_scale);
            CKL(h);
        }
        rc = ddc_process_chunk(h, d.d_in2[sl], n);
        if (rc) return rc;
        CK(cudaEventRecord(d.ev_done[sl], h->stream));
        d.in_used[sl] = true;
        done += n;
        ++k;
    }
    return RCB_OK;
}

// Wire-format input for the DDC bank: see include/b200chan.h.  The channels' streaming state is kept
// (the history is complex64 whatever the wire format), so the format may change between blocks.
extern "C" int rcb_ddc_set_input_format(rcb_t* h, int fmt, float offset, float scale) {
    if (!h) return RCB_EINVAL;
    if (fmt != 0 && fmt != RCB_FMT_U8 && fmt != RCB_FMT_S8 && fmt != RCB_FMT_S16) return RCB_EINVAL;
    h->ddc.in_fmt = fmt;
    h->ddc.in_off = fmt ? offset : 0.f;
    h->ddc.in_scale = fmt ? scale : 1.f;
    return RCB_OK;
}

// every channel's outputs of the last rcb_ddc_process call in ONE transfer: row c (channels in ascending id order) of
// dst holds counts[c] items (`which` = RCB_OUT_IQ: complex64, RCB_OUT_FM: float32; channels without FM give 0).
extern "C" int rcb_ddc_pull_all(rcb_t* h, int which, void* dst, size_t row_stride_items, int dst_mem, int* ids,
                                size_t* counts, size_t cap_rows, size_t* nrows) {
    if (!h || !nrows) return RCB_EINVAL;
    if (which != RCB_OUT_IQ && which != RCB_OUT_FM) return RCB_EINVAL;
    auto& d = h->ddc;
    const size_t M = d.chans.size();
    *nrows = M;
    if (M == 0) return RCB_OK;
    if (!dst || !counts || cap_rows < M) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    const size_t isz = (which == RCB_OUT_IQ) ? sizeof(float2) : sizeof(float);
    size_t mx = 0, r = 0;
    std::vector<const void*> src(M);
    std::vector<int> cnt(M);
    for (auto& kv : d.chans) {
        DdcChan& c = kv.second;
        const bool has = (which == RCB_OUT_IQ) || (c.out_mask & RCB_OUT_FM);
        const size_t n = has ? c.nout_last : 0;
        if (ids) ids[r] = c.id;
        counts[r] = n;
        cnt[r] = (int)n;
        src[r] = (which == RCB_OUT_IQ) ? (const void*)c.d_out_iq : (const void*)c.d_out_fm;
        mx = std::max(mx, n);
        ++r;
    }
    if (mx == 0) return RCB_OK;
    if (mx > row_stride_items) return RCB_ERANGE;
    // gather into one dense [M][mx] block, then one copy
    int rc = ensure_tmp(h, 0, M * mx * isz + M * (sizeof(void*) + sizeof(int)) + 64);
    if (rc) return rc;
    char* base = (char*)h->d_tmp[0];
    const void** d_src = (const void**)(base + ((M * mx * isz + 15) / 16) * 16);
    int* d_cnt = (int*)(d_src + M);
    CK(cudaMemcpyAsync(d_src, src.data(), M * sizeof(void*), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_cnt, cnt.data(), M * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    dim3 grid((unsigned)((mx * (isz / 4) + 255) / 256), (unsigned)M);
    ddc_gather_kernel<<<grid, 256, 0, h->stream>>>(d_src, d_cnt, (int)(isz / 4), (unsigned*)base, (long long)(mx * (isz / 4)));
    CKL(h);
    CK(cudaMemcpy2DAsync(dst, row_stride_items * isz, base, mx * isz, mx * isz, M,
                         dst_mem == RCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // src / cnt are stack vectors
    if (dst_mem == RCB_MEM_HOST) h->stats.d2h_bytes += M * mx * isz;
    return RCB_OK;
}

extern "C" int rcb_ddc_pull(rcb_t* h, int chan_id, int which, void* dst, size_t cap_items, int dst_mem, size_t* nitems) {
    if (!h || !nitems) return RCB_EINVAL;
    auto it = h->ddc.chans.find(chan_id);
    if (it == h->ddc.chans.end()) return RCB_ERANGE;
    DdcChan& c = it->second;
    if (which != RCB_OUT_IQ && which != RCB_OUT_FM) return RCB_EINVAL;
    if (which == RCB_OUT_FM && !(c.out_mask & RCB_OUT_FM)) return RCB_ESTATE;
    *nitems = c.nout_last;
    if (c.nout_last == 0) return RCB_OK;
    if (!dst || cap_items < c.nout_last) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    const size_t bytes = c.nout_last * (which == RCB_OUT_IQ ? sizeof(float2) : sizeof(float));
    const void* src = (which == RCB_OUT_IQ) ? (const void*)c.d_out_iq : (const void*)c.d_out_fm;
    CK(cudaMemcpyAsync(dst, src, bytes, dst_mem == RCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (dst_mem == RCB_MEM_HOST) h->stats.d2h_bytes += bytes;
    return RCB_OK;
}

// =================================================================================================
// K4  stand-alone demod / probe
// =================================================================================================
extern "C" int rcb_quad_demod(rcb_t* h, const void* iq, size_t rows, size_t n, size_t in_stride, float gain, void* prev,
                              void* out_fm, size_t out_stride, int mem) {
    if (!h || !iq || !out_fm || rows == 0 || in_stride < n || out_stride < n) return RCB_EINVAL;
    if (rows > 65535) return RCB_ERANGE;
    CK(cudaSetDevice(h->device));
    if (n == 0) return RCB_OK;
    const float2* d_x = (const float2*)iq;
    float* d_o = (float*)out_fm;
    float2* d_prev = (float2*)prev;
    if (mem == RCB_MEM_HOST) {
        const size_t inb = rows * in_stride * sizeof(float2), outb = rows * out_stride * sizeof(float);
        int rc = ensure_tmp(h, 0, inb + rows * sizeof(float2));
        if (rc) return rc;
        rc = ensure_tmp(h, 1, outb);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->d_tmp[0], iq, inb, cudaMemcpyHostToDevice, h->stream));
        d_x = (const float2*)h->d_tmp[0];
        d_o = (float*)h->d_tmp[1];
        d_prev = (float2*)((char*)h->d_tmp[0] + inb);
        if (prev)
            CK(cudaMemcpyAsync(d_prev, prev, rows * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
        else
            CK(cudaMemsetAsync(d_prev, 0, rows * sizeof(float2), h->stream));
        h->stats.h2d_bytes += inb;
    } else if (mem != RCB_MEM_DEVICE) {
        return RCB_EINVAL;
    }
    dim3 grid((unsigned)((n + 1023) / 1024), (unsigned)rows);
    quad_demod_rows_kernel<<<grid, 256, 0, h->stream>>>(d_x, (long long)in_stride, d_prev, d_o, (long long)out_stride,
                                                       (long long)n, gain, 0);
    CKL(h);
    if (d_prev) {
        save_last_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(d_x, (long long)in_stride,
                                                                              (long long)n - 1, d_prev, (int)rows);
        CKL(h);
    }
    if (mem == RCB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_fm, d_o, rows * out_stride * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (prev) CK(cudaMemcpyAsync(prev, d_prev, rows * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += rows * out_stride * sizeof(float);
    }
    CK(cudaStreamSynchronize(h->stream));
    h->stats.channel_samples += rows * n;
    return RCB_OK;
}

extern "C" int rcb_probe_mean(rcb_t* h, const void* x, size_t rows, size_t n, size_t stride, size_t length, float scale,
                              void* out, int mem) {
    if (!h || !x || !out || rows == 0 || stride < n || length == 0) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    const float* d_x = (const float*)x;
    float* d_o = (float*)out;
    if (mem == RCB_MEM_HOST) {
        int rc = ensure_tmp(h, 0, rows * stride * sizeof(float));
        if (rc) return rc;
        rc = ensure_tmp(h, 1, rows * sizeof(float));
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->d_tmp[0], x, rows * stride * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        d_x = (const float*)h->d_tmp[0];
        d_o = (float*)h->d_tmp[1];
    } else if (mem != RCB_MEM_DEVICE) {
        return RCB_EINVAL;
    }
    window_sum_rows_kernel<<<(unsigned)rows, 256, 0, h->stream>>>(d_x, (long long)stride, (long long)n, (long long)length,
                                                                 scale, d_o);
    CKL(h);
    if (mem == RCB_MEM_HOST) CK(cudaMemcpyAsync(out, d_o, rows * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return RCB_OK;
}

// =================================================================================================
// K5  ingest conversion
// =================================================================================================
extern "C" int rcb_convert_iq(rcb_t* h, const void* src, int fmt, float offset, float scale, size_t nsamples,
                              int src_mem, void* dst, int dst_mem) {
    if (!h || (!src && nsamples) || (!dst && nsamples)) return RCB_EINVAL;
    if (fmt != RCB_FMT_U8 && fmt != RCB_FMT_S8 && fmt != RCB_FMT_S16) return RCB_EINVAL;
    if ((src_mem != RCB_MEM_HOST && src_mem != RCB_MEM_DEVICE) || (dst_mem != RCB_MEM_HOST && dst_mem != RCB_MEM_DEVICE))
        return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    if (nsamples == 0) return RCB_OK;
    const size_t esz = (fmt == RCB_FMT_S16) ? 2 : 1;
    const size_t inb = nsamples * 2 * esz, outb = nsamples * sizeof(float2);
    const void* d_src = src;
    float2* d_dst = (float2*)dst;
    if (src_mem == RCB_MEM_HOST) {
        int rc = ensure_tmp(h, 0, inb + 16);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->d_tmp[0], src, inb, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += inb;
        d_src = h->d_tmp[0];
    } else if (((uintptr_t)src) & 15) {
        return RCB_EINVAL;  // vector loads need 16-byte aligned device input
    }
    if (dst_mem == RCB_MEM_HOST) {
        int rc = ensure_tmp(h, 1, outb);
        if (rc) return rc;
        d_dst = (float2*)h->d_tmp[1];
    }
    const unsigned grid = (unsigned)((nsamples + 1023) / 1024);
    if (fmt == RCB_FMT_U8)
        convert_iq_kernel<uint8_t><<<grid, 256, 0, h->stream>>>((const uint8_t*)d_src, d_dst, (long long)nsamples, offset, scale);
    else if (fmt == RCB_FMT_S8)
        convert_iq_kernel<int8_t><<<grid, 256, 0, h->stream>>>((const int8_t*)d_src, d_dst, (long long)nsamples, offset, scale);
    else
        convert_iq_kernel<int16_t><<<grid, 256, 0, h->stream>>>((const int16_t*)d_src, d_dst, (long long)nsamples, offset, scale);
    CKL(h);
    if (dst_mem == RCB_MEM_HOST) {
        CK(cudaMemcpyAsync(dst, d_dst, outb, cudaMemcpyDeviceToHost, h->stream));
        h->stats.d2h_bytes += outb;
    }
    if (src_mem == RCB_MEM_HOST || dst_mem == RCB_MEM_HOST) CK(cudaStreamSynchronize(h->stream));
    return RCB_OK;
}

// =================================================================================================
// K3  streaming FFT + log power
// =================================================================================================
extern "C" int rcb_fft_config(rcb_t* h, int length, const float* window, int avg_frames) {
    if (!h || !window || avg_frames < 1) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    int rc = fft_config(h->fft, length, window, avg_frames, h->stream, h->sm_count);
    if (rc == RCB_ECUDA) return fail_cuda(h, cudaGetLastError(), "fft_config");
    return rc;
}
extern "C" int rcb_fft_reset(rcb_t* h) {
    if (!h) return RCB_EINVAL;
    if (!h->fft.configured) return RCB_ESTATE;
    CK(cudaSetDevice(h->device));
    int rc = fft_reset(h->fft, h->stream);
    if (rc == RCB_ECUDA) return fail_cuda(h, cudaGetLastError(), "fft_reset");
    return rc;
}
extern "C" int rcb_fft_set_pipeline(rcb_t* h, int persistent) {
    if (!h) return RCB_EINVAL;
    if (!h->fft.configured) return RCB_ESTATE;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (persistent < 0 || persistent > 2) return RCB_EINVAL;
    h->fft.use_scan = (persistent == 1);
    h->fft.frame_off = (persistent != 0);
    // the pipelines associate the block sum differently (frame order / group order): a switch restarts the block
    if (fft_reset(h->fft, h->stream)) return fail_cuda(h, cudaGetLastError(), "fft_reset");
    return RCB_OK;
}
extern "C" int rcb_fft_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem, void* out_sums, size_t cap_vectors,
                               int out_mem, size_t* nvec) {
    if (!h || (!iq && nsamples) || !nvec) return RCB_EINVAL;
    if (!h->fft.configured) return RCB_ESTATE;
    if (nsamples % (size_t)h->fft.L) return RCB_EINVAL;
    CK(cudaSetDevice(h->device));
    uint64_t launches = 0, h2d = 0, d2h = 0;
    // one persistent launch per call (fft_scan.cuh); the round-1 three-kernel pipeline remains as the fallback for inputs
    // a TMA descriptor cannot address
    int rc = 1;
    if (h->fft.use_frame && !h->fft.frame_off)
        rc = fft_process_frames(h->fft, (const float2*)iq, nsamples, in_mem, (float*)out_sums, cap_vectors, out_mem, nvec,
                                h->stream, &launches, &h2d, &d2h);
    else if (h->fft.use_scan)
        rc = fft_process_scan(h->fft, (const float2*)iq, nsamples, in_mem, (float*)out_sums, cap_vectors, out_mem, nvec,
                              h->stream, &launches, &h2d, &d2h, h->sm_count);
    if (rc == 1)
        rc = fft_process(h->fft, (const float2*)iq, nsamples, in_mem, (float*)out_sums, cap_vectors, out_mem, nvec,
                         h->stream, &launches, &h2d, &d2h, h->sm_count);
    h->stats.kernel_launches += launches;
    h->stats.h2d_bytes += h2d;
    h->stats.d2h_bytes += d2h;
    if (rc == RCB_ECUDA) return fail_cuda(h, cudaGetLastError(), "fft_process");
    if (rc == RCB_OK) h->stats.samples_in += nsamples;
    return rc;
}
