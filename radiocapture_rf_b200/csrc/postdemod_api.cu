// K6 host side: rcb_post_* (include/b200chan.h) - the data-parallel post-demod stages of the backend demods, batched
// over channel rows.  Separate translation unit (compiled in parallel with b200chan.cu).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <vector>

#include "internal.h"
#include "postdemod.cuh"

using namespace rcb;

namespace {

#define PCK(call)                                                          \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return rcb_internal_fail(h, e__, #call);   \
    } while (0)
#define PCKL()                                                                         \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) return rcb_internal_fail(h, e__, "post kernel launch"); \
        rcb_internal_count(h, 1, 0, 0);                                                \
    } while (0)

struct FirStage {          // one polyphase rational FIR stage over rows
    int NT = 0, KH = 0, I = 1, D = 1;
    float* d_taps = nullptr;
    float* d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    unsigned long long* d_ctr = nullptr;  // [rows][2]
    float* d_out = nullptr;               // [rows][cap]
    int* d_cnt = nullptr;                 // outputs of the last block per row
    size_t cap = 0;
};

struct PostChain {
    rcb_post_cfg cfg{};
    int rows = 0;
    // P25
    float* d_taps0 = nullptr;   // prefilter
    float* d_taps1 = nullptr;   // symbol filter
    float2* d_hist[2] = {nullptr, nullptr};
    int hist_cur = 0;
    int H = 0;
    float* d_fm = nullptr;      // demod output of the last block [rows][cap] (probe input)
    float* d_sym = nullptr;
    float* d_probe_ring[2] = {nullptr, nullptr};  // last probe_len demod samples per row
    int probe_cur = 0;
    float* d_probe = nullptr;   // [rows]
    size_t cap = 0;
    // analog
    double* d_sq_state = nullptr;   // [rows]
    double* d_de_state = nullptr;   // [rows][4]
    float2* d_gated = nullptr;      // [rows][cap]
    float* d_de = nullptr;          // [rows][cap]
    int* d_cnt0 = nullptr;          // samples surviving the squelch gate
    FirStage fir[3];                // audio low-pass, high-pass, resampler
    int* d_map = nullptr;           // row map (device copy)
    int* h_cnt = nullptr;           // pinned [rows]
    float* h_probe = nullptr;       // pinned [rows]
};

struct PostState {
    std::map<int, PostChain> chains;
    int next_id = 1;
};

PostState* state_of(rcb_t* h, bool create) {
    void** slot = rcb_internal_post_slot(h);
    if (!*slot && create) *slot = new (std::nothrow) PostState();
    return static_cast<PostState*>(*slot);
}

void free_fir(FirStage& f) {
    cudaFree(f.d_taps);
    cudaFree(f.d_hist[0]);
    cudaFree(f.d_hist[1]);
    cudaFree(f.d_ctr);
    cudaFree(f.d_out);
    cudaFree(f.d_cnt);
    f = FirStage{};
}

void free_chain(PostChain& c) {
    cudaFree(c.d_taps0);
    cudaFree(c.d_taps1);
    cudaFree(c.d_hist[0]);
    cudaFree(c.d_hist[1]);
    cudaFree(c.d_fm);
    cudaFree(c.d_sym);
    cudaFree(c.d_probe_ring[0]);
    cudaFree(c.d_probe_ring[1]);
    cudaFree(c.d_probe);
    cudaFree(c.d_sq_state);
    cudaFree(c.d_de_state);
    cudaFree(c.d_gated);
    cudaFree(c.d_de);
    cudaFree(c.d_cnt0);
    cudaFree(c.d_map);
    if (c.h_cnt) cudaFreeHost(c.h_cnt);
    if (c.h_probe) cudaFreeHost(c.h_probe);
    for (auto& f : c.fir) free_fir(f);
}

int init_fir(rcb_t* h, FirStage& f, int rows, const float* taps, int nt, int I, int D) {
    f.NT = nt;
    f.I = I;
    f.D = D;
    f.KH = (nt + I - 1) / I;
    PCK(cudaMalloc(&f.d_taps, (size_t)nt * sizeof(float)));
    PCK(cudaMemcpy(f.d_taps, taps, (size_t)nt * sizeof(float), cudaMemcpyHostToDevice));
    for (int b = 0; b < 2; ++b) {
        PCK(cudaMalloc(&f.d_hist[b], (size_t)rows * f.KH * sizeof(float)));
        PCK(cudaMemset(f.d_hist[b], 0, (size_t)rows * f.KH * sizeof(float)));
    }
    PCK(cudaMalloc(&f.d_ctr, (size_t)rows * 2 * sizeof(unsigned long long)));
    PCK(cudaMemset(f.d_ctr, 0, (size_t)rows * 2 * sizeof(unsigned long long)));
    PCK(cudaMalloc(&f.d_cnt, (size_t)rows * sizeof(int)));
    PCK(cudaMemset(f.d_cnt, 0, (size_t)rows * sizeof(int)));
    return RCB_OK;
}

int ensure_fir_out(rcb_t* h, FirStage& f, int rows, size_t max_in) {
    const size_t need = (max_in * (size_t)f.I) / (size_t)f.D + 2;
    if (f.cap >= need) return RCB_OK;
    cudaFree(f.d_out);
    f.d_out = nullptr;
    f.cap = 0;
    PCK(cudaMalloc(&f.d_out, (size_t)rows * need * sizeof(float)));
    f.cap = need;
    return RCB_OK;
}

// one FIR stage over the rows at d_in (row stride in_stride, per-row counts d_cnt_in); max_in bounds the counts
int run_fir(rcb_t* h, FirStage& f, int rows, const float* d_in, size_t in_stride, const int* d_cnt_in, size_t max_in) {
    cudaStream_t st = rcb_internal_stream(h);
    int rc = ensure_fir_out(h, f, rows, max_in);
    if (rc) return rc;
    PostFirParams p{};
    p.x = d_in;
    p.xs = (long long)in_stride;
    p.cnt_in = d_cnt_in;
    p.hist = f.d_hist[f.cur];
    p.taps = f.d_taps;
    p.ctr = f.d_ctr;
    p.y = f.d_out;
    p.ys = (long long)f.cap;
    p.NT = f.NT;
    p.KH = f.KH;
    p.I = f.I;
    p.D = f.D;
    const size_t max_out = (max_in * (size_t)f.I) / (size_t)f.D + 2;
    dim3 grid((unsigned)((max_out + 255) / 256), (unsigned)rows);
    post_fir_rat_kernel<<<grid, 256, 0, st>>>(p);
    PCKL();
    rows_hist_update_kernel<float><<<rows, 128, 0, st>>>(f.d_hist[f.cur], d_in, (long long)in_stride, nullptr, d_cnt_in,
                                                         0, f.KH, f.d_hist[f.cur ^ 1]);
    PCKL();
    f.cur ^= 1;
    post_fir_finish_kernel<<<(rows + 127) / 128, 128, 0, st>>>(f.d_ctr, d_cnt_in, f.d_cnt, rows, f.I, f.D);
    PCKL();
    return RCB_OK;
}

}  // namespace

void rcb_post_free_all(rcb_t* h) {
    PostState* s = state_of(h, false);
    if (!s) return;
    for (auto& kv : s->chains) free_chain(kv.second);
    delete s;
    *rcb_internal_post_slot(h) = nullptr;
}

extern "C" int rcb_post_open(rcb_t* h, const rcb_post_cfg* cfg, int* chain_id) {
    if (!h || !cfg || !chain_id) return RCB_EINVAL;
    if (cfg->rows < 1 || cfg->rows > 65535) return RCB_EINVAL;
    if (cfg->kind != RCB_POST_P25_C4FM && cfg->kind != RCB_POST_ANALOG_FM) return RCB_EINVAL;
    if (!cfg->taps0 || cfg->ntaps0 < 1 || !cfg->taps1 || cfg->ntaps1 < 1) return RCB_EINVAL;
    if (cfg->kind == RCB_POST_ANALOG_FM && (!cfg->taps2 || cfg->ntaps2 < 1 || cfg->interp < 1 || cfg->decim < 1))
        return RCB_EINVAL;
    if (cfg->kind == RCB_POST_P25_C4FM && (cfg->ntaps0 > 4096 || cfg->ntaps1 > 256)) return RCB_EUNSUPPORTED;
    PCK(cudaSetDevice(rcb_internal_device(h)));
    PostState* s = state_of(h, true);
    if (!s) return RCB_ENOMEM;
    PostChain c;
    c.cfg = *cfg;
    c.cfg.taps0 = c.cfg.taps1 = c.cfg.taps2 = nullptr;
    c.rows = cfg->rows;
    const int rows = c.rows;
    PCK(cudaHostAlloc(&c.h_cnt, (size_t)rows * sizeof(int), cudaHostAllocDefault));
    PCK(cudaHostAlloc(&c.h_probe, (size_t)rows * sizeof(float), cudaHostAllocDefault));
    PCK(cudaMalloc(&c.d_map, (size_t)rows * sizeof(int)));
    if (cfg->kind == RCB_POST_P25_C4FM) {
        c.H = cfg->ntaps0 - 1 + cfg->ntaps1;
        PCK(cudaMalloc(&c.d_taps0, (size_t)cfg->ntaps0 * sizeof(float)));
        PCK(cudaMalloc(&c.d_taps1, (size_t)cfg->ntaps1 * sizeof(float)));
        PCK(cudaMemcpy(c.d_taps0, cfg->taps0, (size_t)cfg->ntaps0 * sizeof(float), cudaMemcpyHostToDevice));
        PCK(cudaMemcpy(c.d_taps1, cfg->taps1, (size_t)cfg->ntaps1 * sizeof(float), cudaMemcpyHostToDevice));
        for (int b = 0; b < 2; ++b) {
            PCK(cudaMalloc(&c.d_hist[b], (size_t)rows * c.H * sizeof(float2)));
            PCK(cudaMemset(c.d_hist[b], 0, (size_t)rows * c.H * sizeof(float2)));
        }
        if (cfg->probe_len > 0) {
            for (int b = 0; b < 2; ++b) {
                PCK(cudaMalloc(&c.d_probe_ring[b], (size_t)rows * cfg->probe_len * sizeof(float)));
                PCK(cudaMemset(c.d_probe_ring[b], 0, (size_t)rows * cfg->probe_len * sizeof(float)));
            }
            PCK(cudaMalloc(&c.d_probe, (size_t)rows * sizeof(float)));
        }
    } else {
        PCK(cudaMalloc(&c.d_sq_state, (size_t)rows * sizeof(double)));
        PCK(cudaMemset(c.d_sq_state, 0, (size_t)rows * sizeof(double)));
        PCK(cudaMalloc(&c.d_de_state, (size_t)rows * 4 * sizeof(double)));
        PCK(cudaMemset(c.d_de_state, 0, (size_t)rows * 4 * sizeof(double)));
        PCK(cudaMalloc(&c.d_cnt0, (size_t)rows * sizeof(int)));
        int rc = init_fir(h, c.fir[0], rows, cfg->taps0, cfg->ntaps0, 1, 1);
        if (!rc) rc = init_fir(h, c.fir[1], rows, cfg->taps1, cfg->ntaps1, 1, 1);
        if (!rc) rc = init_fir(h, c.fir[2], rows, cfg->taps2, cfg->ntaps2, cfg->interp, cfg->decim);
        if (rc) {
            free_chain(c);
            return rc;
        }
    }
    const int id = s->next_id++;
    s->chains[id] = c;
    *chain_id = id;
    return RCB_OK;
}

extern "C" int rcb_post_close(rcb_t* h, int chain_id) {
    if (!h) return RCB_EINVAL;
    PostState* s = state_of(h, false);
    if (!s) return RCB_ERANGE;
    auto it = s->chains.find(chain_id);
    if (it == s->chains.end()) return RCB_ERANGE;
    PCK(cudaSetDevice(rcb_internal_device(h)));
    PCK(cudaStreamSynchronize(rcb_internal_stream(h)));
    free_chain(it->second);
    s->chains.erase(it);
    return RCB_OK;
}

extern "C" int rcb_post_process(rcb_t* h, int chain_id, const void* iq, size_t n, size_t in_stride, const int* row_map,
                                int in_mem, void* out, size_t out_stride, int out_mem, int* nout, float* probe) {
    if (!h || (!iq && n) || !out || !nout) return RCB_EINVAL;
    if ((in_mem != RCB_MEM_HOST && in_mem != RCB_MEM_DEVICE) || (out_mem != RCB_MEM_HOST && out_mem != RCB_MEM_DEVICE))
        return RCB_EINVAL;
    if (n > (size_t)1 << 28) return RCB_ERANGE;
    PostState* s = state_of(h, false);
    if (!s) return RCB_ERANGE;
    auto it = s->chains.find(chain_id);
    if (it == s->chains.end()) return RCB_ERANGE;
    PostChain& c = it->second;
    const int rows = c.rows;
    if (in_stride < n) return RCB_EINVAL;
    PCK(cudaSetDevice(rcb_internal_device(h)));
    cudaStream_t st = rcb_internal_stream(h);
    if (n == 0) {
        for (int r = 0; r < rows; ++r) nout[r] = 0;
        return RCB_OK;
    }
    // stage the input rows
    const float2* d_x = (const float2*)iq;
    size_t xs = in_stride;
    const int* d_map = nullptr;
    if (in_mem == RCB_MEM_HOST) {
        // host rows are gathered by the copy itself (row_map applied here), dense [rows][n] on the device
        float2* stage = nullptr;
        PCK(cudaMallocAsync((void**)&stage, (size_t)rows * n * sizeof(float2), st));
        for (int r = 0; r < rows; ++r) {
            const int sr = row_map ? row_map[r] : r;
            PCK(cudaMemcpyAsync(stage + (size_t)r * n, (const float2*)iq + (size_t)sr * in_stride, n * sizeof(float2),
                                cudaMemcpyHostToDevice, st));
        }
        rcb_internal_count(h, 0, (size_t)rows * n * sizeof(float2), 0);
        d_x = stage;
        xs = n;
    } else if (row_map) {
        PCK(cudaMemcpyAsync(c.d_map, row_map, (size_t)rows * sizeof(int), cudaMemcpyHostToDevice, st));
        d_map = c.d_map;
    }
    auto release_stage = [&]() {
        if (in_mem == RCB_MEM_HOST) cudaFreeAsync((void*)d_x, st);
    };
    // per-chain work buffers sized for n inputs per row
    if (c.cap < n) {
        cudaStreamSynchronize(st);
        cudaFree(c.d_fm);
        cudaFree(c.d_sym);
        cudaFree(c.d_gated);
        cudaFree(c.d_de);
        c.d_fm = c.d_sym = c.d_de = nullptr;
        c.d_gated = nullptr;
        c.cap = 0;
        const size_t cap = n + n / 4 + 64;
        if (c.cfg.kind == RCB_POST_P25_C4FM) {
            PCK(cudaMalloc(&c.d_fm, (size_t)rows * cap * sizeof(float)));
            PCK(cudaMalloc(&c.d_sym, (size_t)rows * cap * sizeof(float)));
        } else {
            PCK(cudaMalloc(&c.d_gated, (size_t)rows * cap * sizeof(float2)));
            PCK(cudaMalloc(&c.d_de, (size_t)rows * cap * sizeof(float)));
        }
        c.cap = cap;
    }
    const float* d_res = nullptr;   // result rows on the device
    size_t res_stride = 0;
    const int* d_res_cnt = nullptr;  // null: n for every row
    if (c.cfg.kind == RCB_POST_P25_C4FM) {
        PostP25Params p{};
        p.x = d_x;
        p.xs = (long long)xs;
        p.row_map = d_map;
        p.hist = c.d_hist[c.hist_cur];
        p.taps = c.d_taps0;
        p.sym = c.d_taps1;
        p.out_sym = c.d_sym;
        p.out_fm = c.d_fm;
        p.os = (long long)c.cap;
        p.n = (int)n;
        p.NT = c.cfg.ntaps0;
        p.SPS = c.cfg.ntaps1;
        p.gain = c.cfg.gain;
        const int H = c.H;
        const size_t smem = (size_t)(256 + H) * 8 + (size_t)(256 + p.SPS) * 8 + (size_t)(256 + p.SPS) * 4 +
                            (size_t)(p.NT + p.SPS) * 4;
        if (smem > 48 * 1024) {
            static bool attr_dev[64] = {};
            bool& done = attr_dev[rcb_internal_device(h) & 63];
            if (!done) {
                PCK(cudaFuncSetAttribute(post_p25_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                done = true;
            }
        }
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)rows);
        post_p25_kernel<<<grid, 256, smem, st>>>(p);
        PCKL();
        rows_hist_update_kernel<float2><<<rows, 128, 0, st>>>(c.d_hist[c.hist_cur], d_x, (long long)xs, d_map, nullptr,
                                                              (int)n, H, c.d_hist[c.hist_cur ^ 1]);
        PCKL();
        c.hist_cur ^= 1;
        if (c.cfg.probe_len > 0) {
            // moving_average_ff(probe_len, 1, ..) -> multiply_const(probe_scale) -> probe_signal_f: the value after
            // the block = scale * sum of the last probe_len demod samples (ring of them carried across blocks)
            rows_hist_update_kernel<float><<<rows, 128, 0, st>>>(c.d_probe_ring[c.probe_cur], c.d_fm, (long long)c.cap,
                                                                 nullptr, nullptr, (int)n, c.cfg.probe_len,
                                                                 c.d_probe_ring[c.probe_cur ^ 1]);
            PCKL();
            c.probe_cur ^= 1;
            post_row_sum_kernel<<<rows, 256, 0, st>>>(c.d_probe_ring[c.probe_cur], (long long)c.cfg.probe_len,
                                                      c.cfg.probe_len, c.cfg.probe_scale, c.d_probe);
            PCKL();
        }
        d_res = c.d_sym;
        res_stride = c.cap;
    } else {
        const double thr = pow(10.0, c.cfg.squelch_db / 10.0);
        post_squelch_kernel<<<(rows + 3) / 4, 128, 0, st>>>(d_x, (long long)xs, d_map, (int)n, rows, c.cfg.squelch_alpha,
                                                            thr, c.cfg.squelch_gate ? 1 : 0, c.d_sq_state, c.d_gated,
                                                            (long long)c.cap, c.d_cnt0);
        PCKL();
        post_fm_deemph_kernel<<<(rows + 3) / 4, 128, 0, st>>>(c.d_gated, (long long)c.cap, c.d_cnt0, rows, c.cfg.gain,
                                                              c.cfg.deemph_b0, c.cfg.deemph_b1, c.cfg.deemph_a1,
                                                              c.d_de_state, c.d_de, (long long)c.cap);
        PCKL();
        int rc = run_fir(h, c.fir[0], rows, c.d_de, c.cap, c.d_cnt0, n);
        if (!rc) rc = run_fir(h, c.fir[1], rows, c.fir[0].d_out, c.fir[0].cap, c.fir[0].d_cnt, n);
        if (!rc) rc = run_fir(h, c.fir[2], rows, c.fir[1].d_out, c.fir[1].cap, c.fir[1].d_cnt, n);
        if (rc) {
            release_stage();
            return rc;
        }
        d_res = c.fir[2].d_out;
        res_stride = c.fir[2].cap;
        d_res_cnt = c.fir[2].d_cnt;
    }
    // counts (and probe) to the host
    if (d_res_cnt) {
        PCK(cudaMemcpyAsync(c.h_cnt, d_res_cnt, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    if (probe && c.d_probe) PCK(cudaMemcpyAsync(c.h_probe, c.d_probe, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    size_t max_out = 0;
    for (int r = 0; r < rows; ++r) {
        nout[r] = d_res_cnt ? c.h_cnt[r] : (int)n;
        max_out = std::max(max_out, (size_t)nout[r]);
        if (probe) probe[r] = c.d_probe ? c.h_probe[r] : 0.f;
    }
    if (max_out > out_stride) {
        release_stage();
        return RCB_ERANGE;
    }
    if (max_out) {
        PCK(cudaMemcpy2DAsync(out, out_stride * sizeof(float), d_res, res_stride * sizeof(float), max_out * sizeof(float),
                              rows, out_mem == RCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
        if (out_mem == RCB_MEM_HOST) rcb_internal_count(h, 0, 0, max_out * sizeof(float) * rows);
    }
    release_stage();
    PCK(cudaStreamSynchronize(st));
    return RCB_OK;
}
