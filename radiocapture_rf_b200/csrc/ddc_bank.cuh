// K2  ddc_bank : M arbitrary-offset digital down-converters over one wideband block.
//
// Each channel c is one  filter.freq_xlating_fir_filter_ccc(D, taps, f0, fs)  (GNU Radio gr-filter
// freq_xlating_fir_filter_impl.cc + fir_filter.cc + gr-blocks rotator.h), i.e. what one
// rc_frontend/channel.py:35 `channel` top_block computes (also the split-2 half-band pair
// rc_frontend/receiver.py:85-86 and the 1x prefilters p25_control_demod.py:108 /
// logging_receiver.py:231):
//
//   y[i] = e^{-j w D i} * sum_k h[k] e^{+j w k} x[iD - k],        w = 2 pi f0 / fs
//
// All M channels read the SAME staged wideband block (the reference re-copies the full-rate stream
// through ZMQ once per channel, rc_frontend/channel.py:29).  One warp produces 4 consecutive outputs
// of one channel: lanes stride the composite taps (coalesced tap and sample loads, the re-reads of
// overlapping windows are L1/L2 hits), shuffle-reduce, then lane 0 derotates with an EXACT phase
// (double sincospi of frac(phase0 + cyc*i), no recursive float rotator drift) and stores.
// The optional FM demod of the narrowband result runs in ddc_fm_kernel (output rates are fs/D).
#pragma once
#include "common.cuh"

namespace rcb {

struct DdcChanDev {
    const float2* ctaps_rev;  // [ntaps]  ct_rev[r] = h[K-1-r] e^{+j w (K-1-r)}
    float2* out_iq;           // [nout] this block's outputs
    float* out_fm;            // [nout] or null
    float2* prev;             // device scalar: last output of the previous block (FM carry)
    double cyc;               // frac(f0 * D / fs)  (cycles per output, reduced mod 1)
    double phase0;            // frac(cyc * i_first)
    long long s_first;        // block-relative index of the newest input sample of output 0 (may be < 0)
    long long s_open;         // block-relative index of the first sample the channel ever saw: older samples
                              // read as 0 (a new GNU Radio `channel` top_block starts with zero filter history)
    int ntaps;
    int decim;
    int nout;
    float gain;
};

__device__ __forceinline__ float2 ddc_x_at(const float2* __restrict__ x, const float2* __restrict__ hist,
                                           int hist_cap, long long idx) {
    // idx >= 0: new block; idx < 0: the hist_cap samples preceding the block
    if (idx >= 0) return __ldg(x + idx);
    return (idx >= -(long long)hist_cap) ? __ldg(hist + hist_cap + idx) : make_float2(0.f, 0.f);
}

constexpr int kDdcOutPerWarp = 4;

// grid: (ceil(max_nout / (8 warps * 4)), M)   block: 256
__global__ void __launch_bounds__(256) ddc_bank_kernel(const DdcChanDev* __restrict__ chans,
                                                       const float2* __restrict__ x, long long nsamp,
                                                       const float2* __restrict__ hist, int hist_cap) {
    const DdcChanDev ch = chans[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = (blockIdx.x * 8 + warp) * kDdcOutPerWarp;
    if (o0 >= ch.nout) return;
    float2 acc[kDdcOutPerWarp];
#pragma unroll
    for (int q = 0; q < kDdcOutPerWarp; ++q) acc[q] = make_float2(0.f, 0.f);
    // window of output o starts at sample  s_first + o*D - (ntaps-1)
    const long long w0 = ch.s_first + (long long)o0 * ch.decim - (ch.ntaps - 1);
    for (int r = lane; r < ch.ntaps; r += 32) {
        const float2 t = __ldg(ch.ctaps_rev + r);
#pragma unroll
        for (int q = 0; q < kDdcOutPerWarp; ++q) {
            const long long idx = w0 + (long long)q * ch.decim + r;
            float2 xv = make_float2(0.f, 0.f);
            if (idx < nsamp && idx >= ch.s_open) xv = ddc_x_at(x, hist, hist_cap, idx);
            acc[q].x = fmaf(t.x, xv.x, fmaf(-t.y, xv.y, acc[q].x));
            acc[q].y = fmaf(t.x, xv.y, fmaf(t.y, xv.x, acc[q].y));
        }
    }
#pragma unroll
    for (int q = 0; q < kDdcOutPerWarp; ++q) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, s);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, s);
        }
    }
    if (lane < kDdcOutPerWarp) {
        const int o = o0 + lane;
        if (o < ch.nout) {
            float2 a = acc[0];
#pragma unroll
            for (int q = 1; q < kDdcOutPerWarp; ++q)
                if (lane == q) a = acc[q];
            double ph = ch.phase0 + ch.cyc * (double)o;
            ph -= floor(ph);
            double s, c;
            sincospi(-2.0 * ph, &s, &c);
            const float cf = (float)c, sf = (float)s;
            ch.out_iq[o] = make_float2(fmaf(a.x, cf, -a.y * sf), fmaf(a.x, sf, a.y * cf));
        }
    }
}

// FM demod of each channel's narrowband block + carry of the last sample.  grid (ceil(max_nout/256), M)
__global__ void __launch_bounds__(256) ddc_fm_kernel(const DdcChanDev* __restrict__ chans) {
    const DdcChanDev ch = chans[blockIdx.y];
    if (!ch.out_fm) return;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ch.nout) return;
    const float2 c = ch.out_iq[o];
    const float2 pv = (o == 0) ? *ch.prev : ch.out_iq[o - 1];
    const float2 pr = cmul_conj(c, pv);
    ch.out_fm[o] = ch.gain * atan2_fast(pr.y, pr.x);
}
// must run after ddc_fm_kernel (same stream): prev <- last output
__global__ void ddc_carry_kernel(const DdcChanDev* __restrict__ chans, int M) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    const DdcChanDev ch = chans[c];
    if (ch.nout > 0) *ch.prev = ch.out_iq[ch.nout - 1];
}

// new_hist (cap samples) = last `cap` samples of (old_hist ++ x[0..n))
__global__ void hist_update_kernel(const float2* __restrict__ old_hist, const float2* __restrict__ x,
                                   long long n, float2* __restrict__ new_hist, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const long long src = i + n - cap;  // index into x; negative -> old hist
    new_hist[i] = (src >= 0) ? x[src] : ((src >= -cap) ? old_hist[cap + src] : make_float2(0.f, 0.f));
}

}  // namespace rcb
