// K6  post-demod data-parallel stages of the backend demods, batched over channel rows.
//
// Everything between the frontend's narrowband complex64 stream and the first sample-serial timing loop
// (fsk4_demod_ff / clock recovery / vocoders - out of scope) is linear filtering plus atan2 and maps onto rows of
// channels:
//   P25 C4FM front (p25_control_demod.py:106-133, logging_receiver.py:229-245):
//     freq_xlating_fir_filter_ccc(1, low_pass_2(1, 25000, 6250, 500, 30, BLACKMAN), 0, 25000)   69-tap prefilter
//     -> analog.quadrature_demod_cf(rate / (2 pi 600)) -> fir_filter_fff(1, (1/5,)*5) symbol filter,
//     with moving_average_ff(10000, 1, 40000) * 1e-4 -> probe_signal_f (the AFC probe) on the demod output
//     => ONE fused kernel (post_p25_kernel): a CTA stages the input window of a 256-output tile in shared memory,
//     prefilters, demodulates and box-filters it there; the prefiltered and demodulated streams never touch HBM.
//   analog voice (logging_receiver.py:210-222):
//     pwr_squelch_cc(-100, 0.01, 0, True) -> fm_demod_cf (quadrature demod, fm_deemph(75 us) first-order IIR in
//     double, optfir audio low-pass) -> 300 Hz firdes.high_pass (2007 taps at 25 kHz) -> rational_resampler_fff(8000,
//     rate) (8/25, 821-tap Kaiser prototype)
//     => post_squelch_kernel (power IIR as a warp-wide affine scan + gate compaction), post_fm_deemph_kernel (demod +
//     IIR as an affine scan), post_fir_rat_kernel (polyphase rational FIR, one output per thread, double accumulation)
//     x 3, post_stage_finish_kernel (history / counters).
// First-order recurrences y[n] = a y[n-1] + b[n] are evaluated exactly like the sequential loop up to fp64 rounding:
// 32 consecutive samples per warp step, inclusive scan of the affine maps (a, b) by shuffles, carry in lane 31.
#pragma once
#include "common.cuh"

namespace rcb {

// ---------------------------------------------------------------------------------------------------------------
// P25 C4FM front, fused.  grid (ceil(n / 256), rows), 256 threads.
//   hist[r][H], H = NT - 1 + SPS: the input samples preceding this block (zeros at stream start - the boxcar and the
//   demod then see exactly GNU Radio's zero histories because the prefilter of zeros is zero and atan2(0,0) = 0)
// ---------------------------------------------------------------------------------------------------------------
struct PostP25Params {
    const float2* x;       // [rows][xs]
    long long xs;
    const int* row_map;    // optional: input row of output row r
    const float2* hist;    // [rows][H]
    const float* taps;     // prefilter taps [NT] (real)
    const float* sym;      // symbol filter taps [SPS]
    float* out_sym;        // [rows][os]  symbol-filtered FM
    float* out_fm;         // [rows][os]  demod output (AFC probe input) or null
    long long os;
    int n, NT, SPS;
    float gain;
};

__global__ void __launch_bounds__(256) post_p25_kernel(const PostP25Params p) {
    constexpr int TB = 256;
    extern __shared__ __align__(16) unsigned char post_smem[];
    const int NT = p.NT, SPS = p.SPS, H = NT - 1 + SPS;
    float2* xw = reinterpret_cast<float2*>(post_smem);           // TB + H input samples: index i <-> time t0 - H + i
    float2* zw = xw + (TB + H);                                   // TB + SPS prefiltered: index i <-> time t0 - SPS + i
    float* dw = reinterpret_cast<float*>(zw + (TB + SPS));        // TB + SPS - 1 demodulated: i <-> t0 - (SPS-1) + i
    float* tp = dw + (TB + SPS);                                  // NT + SPS taps
    const int r = blockIdx.y, tid = threadIdx.x;
    const int t0 = blockIdx.x * TB;
    const int src_row = p.row_map ? p.row_map[r] : r;
    const float2* xr = p.x + (long long)src_row * p.xs;
    const float2* hr = p.hist + (long long)r * H;
    for (int i = tid; i < TB + H; i += 256) {
        const int t = t0 - H + i;
        float2 v = make_float2(0.f, 0.f);
        if (t < 0) v = hr[H + t];
        else if (t < p.n) v = xr[t];
        xw[i] = v;
    }
    for (int i = tid; i < NT; i += 256) tp[i] = p.taps[i];
    for (int i = tid; i < SPS; i += 256) tp[NT + i] = p.sym[i];
    __syncthreads();
    for (int i = tid; i < TB + SPS; i += 256) {
        // z[t] = sum_k h[k] x[t - k], t = t0 - SPS + i  ->  x index (t - k) - (t0 - H) = i + NT - 1 - k
        float2 acc = make_float2(0.f, 0.f);
        const float2* xp = xw + i + NT - 1;
        for (int k = 0; k < NT; ++k) {
            const float h = tp[k];
            const float2 v = xp[-k];
            acc.x = fmaf(h, v.x, acc.x);
            acc.y = fmaf(h, v.y, acc.y);
        }
        zw[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < TB + SPS - 1; i += 256) {
        const float2 pr = cmul_conj(zw[i + 1], zw[i]);
        dw[i] = p.gain * atan2_fast(pr.y, pr.x);
    }
    __syncthreads();
    const int t = t0 + tid;
    if (t < p.n) {
        float acc = 0.f;
        for (int j = 0; j < SPS; ++j) acc = fmaf(tp[NT + j], dw[tid + SPS - 1 - j], acc);
        p.out_sym[(long long)r * p.os + t] = acc;
        if (p.out_fm) p.out_fm[(long long)r * p.os + t] = dw[tid + SPS - 1];
    }
}

// new_hist[r] = last H items of (hist[r] ++ x[r][0 .. n_r))  for item sizes of 4 / 8 bytes (T = float / float2).
// counts: per-row item counts (device) or null (= n for every row).  grid (rows), 128 threads.
template <typename T>
__global__ void rows_hist_update_kernel(const T* __restrict__ hist, const T* __restrict__ x, long long xs,
                                        const int* __restrict__ row_map, const int* __restrict__ counts, int n, int H,
                                        T* __restrict__ out) {
    const int r = blockIdx.x;
    const int nr = counts ? counts[r] : n;
    const T* xr = x + (long long)(row_map ? row_map[r] : r) * xs;
    const T* hr = hist + (long long)r * H;
    T* o = out + (long long)r * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        const int t = nr - H + i;  // position in the new block (negative: still history)
        o[i] = (t >= 0) ? xr[t] : hr[H + t];   // H + t = i + nr < H
    }
}

// ---------------------------------------------------------------------------------------------------------------
// warp-wide inclusive scan of affine maps f_l(v) = a_l v + b_l (composition in lane order), fp64
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void affine_scan_warp(double& a, double& b, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double ap = __shfl_up_sync(0xffffffffu, a, d);
        const double bp = __shfl_up_sync(0xffffffffu, b, d);
        if (lane >= d) {  // (a, b) o (ap, bp): first the earlier lanes' map, then mine
            b = fma(a, bp, b);
            a = a * ap;
        }
    }
}

// pwr_squelch_cc(db, alpha, ramp = 0, gate): one warp per row.
//   state[r] = running power (double); y rows receive the kept samples (gate) or the input with muted samples zeroed;
//   cnt[r] = samples written.
__global__ void __launch_bounds__(128) post_squelch_kernel(const float2* __restrict__ x, long long xs,
                                                           const int* __restrict__ row_map, int n, int rows,
                                                           double alpha, double thr, int gate,
                                                           double* __restrict__ state, float2* __restrict__ y,
                                                           long long ys, int* __restrict__ cnt) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float2* xr = x + (long long)(row_map ? row_map[r] : r) * xs;
    float2* yr = y + (long long)r * ys;
    double pw = state[r];
    int w = 0;
    for (int base = 0; base < n; base += 32) {
        const int t = base + lane;
        const bool in = t < n;
        const float2 v = in ? xr[t] : make_float2(0.f, 0.f);
        // p[t] = (1 - alpha) p[t-1] + alpha |x|^2 ; lanes past the end are the identity map
        double a = in ? (1.0 - alpha) : 1.0;
        double b = in ? alpha * ((double)v.x * (double)v.x + (double)v.y * (double)v.y) : 0.0;
        affine_scan_warp(a, b, lane);
        const double pt = fma(a, pw, b);
        const bool keep = in && !(pt < thr);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (gate) {
            if (keep) yr[w + __popc(m & ((1u << lane) - 1u))] = v;
            w += __popc(m);
        } else {
            if (in) yr[t] = keep ? v : make_float2(0.f, 0.f);
            w = min(n, base + 32);
        }
        pw = __shfl_sync(0xffffffffu, pt, 31);
    }
    if (lane == 0) {
        state[r] = pw;
        cnt[r] = w;
    }
}

// quadrature_demod_cf(gain) -> iir_filter_ffd([b0, b1], [1, a1]) (fm_deemph): one warp per row, per-row counts.
//   st[r] = {x_prev.re, x_prev.im, d_prev, y_prev} (double)
__global__ void __launch_bounds__(128) post_fm_deemph_kernel(const float2* __restrict__ x, long long xs,
                                                             const int* __restrict__ cnt, int rows, float gain,
                                                             double b0, double b1, double a1, double* __restrict__ st,
                                                             float* __restrict__ out, long long os) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int n = cnt[r];
    const float2* xr = x + (long long)r * xs;
    float* orow = out + (long long)r * os;
    float2 xprev = make_float2((float)st[4 * r + 0], (float)st[4 * r + 1]);
    double dprev = st[4 * r + 2], yprev = st[4 * r + 3];
    for (int base = 0; base < n; base += 32) {
        const int t = base + lane;
        const bool in = t < n;
        const float2 v = in ? xr[t] : make_float2(0.f, 0.f);
        float2 pv = make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
        if (lane == 0) pv = xprev;
        const float2 pr = cmul_conj(v, pv);
        const float d = gain * atan2_fast(pr.y, pr.x);
        float dp = __shfl_up_sync(0xffffffffu, d, 1);
        const double dpd = (lane == 0) ? dprev : (double)dp;
        // y[t] = -a1 y[t-1] + (b0 d[t] + b1 d[t-1])
        double a = in ? -a1 : 1.0;
        double b = in ? fma(b0, (double)d, b1 * dpd) : 0.0;
        affine_scan_warp(a, b, lane);
        const double yt = fma(a, yprev, b);
        if (in) orow[t] = (float)yt;
        // carries: the last valid lane of this step
        const int last = min(31, n - 1 - base);
        yprev = __shfl_sync(0xffffffffu, yt, last);
        dprev = (double)__shfl_sync(0xffffffffu, d, last);
        xprev.x = __shfl_sync(0xffffffffu, v.x, last);
        xprev.y = __shfl_sync(0xffffffffu, v.y, last);
    }
    if (lane == 0) {
        st[4 * r + 0] = (double)xprev.x;
        st[4 * r + 1] = (double)xprev.y;
        st[4 * r + 2] = dprev;
        st[4 * r + 3] = yprev;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// polyphase rational FIR on rows (fir_filter_fff: I = 1; rational_resampler_fff: I / D):
//   y[m] = sum_n x[n] h[m D - n I],  outputs m produced while floor(m D / I) < (inputs seen so far)
//   ctr[r] = {n0 = inputs consumed before this block, m0 = next output index} (64-bit)
//   hist[r][KH]: the KH = ceil(NT / I) inputs before this block.  One output per thread, double accumulation.
// grid (ceil(max_out / 256), rows)
// ---------------------------------------------------------------------------------------------------------------
struct PostFirParams {
    const float* x;        // [rows][xs]
    long long xs;
    const int* cnt_in;     // per-row input counts (device)
    const float* hist;     // [rows][KH]
    const float* taps;     // [NT]
    const unsigned long long* ctr;  // [rows][2]
    float* y;              // [rows][ys]
    long long ys;
    int NT, KH, I, D;
};

__device__ __forceinline__ int post_fir_count(unsigned long long n0, unsigned long long m0, int nin, int I, int D) {
    // outputs m >= m0 with floor(m D / I) < n0 + nin  <=>  m D < (n0 + nin) I  <=>  m <= ((n0 + nin) I - 1) / D
    const unsigned long long lim = (n0 + (unsigned long long)nin) * (unsigned long long)I;
    if (lim == 0) return 0;
    const unsigned long long mlast = (lim - 1) / (unsigned long long)D;
    return (mlast + 1 > m0) ? (int)(mlast + 1 - m0) : 0;
}

__global__ void __launch_bounds__(256) post_fir_rat_kernel(const PostFirParams p) {
    const int r = blockIdx.y;
    const unsigned long long n0 = p.ctr[2 * r], m0 = p.ctr[2 * r + 1];
    const int nin = p.cnt_in[r];
    const int nout = post_fir_count(n0, m0, nin, p.I, p.D);
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= nout) return;
    const unsigned long long m = m0 + (unsigned long long)j;
    const unsigned long long md = m * (unsigned long long)p.D;
    const long long nn = (long long)(md / (unsigned long long)p.I) - (long long)n0;  // newest input, block relative
    const int ph = (int)(md % (unsigned long long)p.I);
    const float* xr = p.x + (long long)r * p.xs;
    const float* hr = p.hist + (long long)r * p.KH;
    double acc = 0.0;
    int k = 0;
    for (int ti = ph; ti < p.NT; ti += p.I, ++k) {
        const long long t = nn - k;
        const float v = (t >= 0) ? xr[t] : ((t >= -(long long)p.KH) ? hr[p.KH + t] : 0.f);
        acc = fma((double)__ldg(p.taps + ti), (double)v, acc);
    }
    p.y[(long long)r * p.ys + j] = (float)acc;
}

// after a FIR stage: counters, per-row output count for the next stage.  (History is updated by
// rows_hist_update_kernel before this runs.)
__global__ void post_fir_finish_kernel(unsigned long long* __restrict__ ctr, const int* __restrict__ cnt_in,
                                       int* __restrict__ cnt_out, int rows, int I, int D) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const unsigned long long n0 = ctr[2 * r], m0 = ctr[2 * r + 1];
    const int nin = cnt_in[r];
    const int nout = post_fir_count(n0, m0, nin, I, D);
    ctr[2 * r] = n0 + (unsigned long long)nin;
    ctr[2 * r + 1] = m0 + (unsigned long long)nout;
    cnt_out[r] = nout;
}

// scale * sum of a row of `length` floats (the AFC probe ring), double accumulation.  One CTA per row.
__global__ void __launch_bounds__(256) post_row_sum_kernel(const float* __restrict__ x, long long xs, int length,
                                                           float scale, float* __restrict__ out) {
    const float* xr = x + (long long)blockIdx.x * xs;
    double acc = 0.0;
    for (int i = threadIdx.x; i < length; i += 256) acc += (double)xr[i];
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(sh[0] * (double)scale);
}

}  // namespace rcb
