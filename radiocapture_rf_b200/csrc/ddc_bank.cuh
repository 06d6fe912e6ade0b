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
#include "fft_packed.cuh"
#include "pfb_fm_tma.cuh"  // mbarrier / bulk-copy helpers

namespace rcb {

struct DdcChanDev {
    const float2* ctaps_rev;  // [ntaps]  ct_rev[r] = h[K-1-r] e^{+j w (K-1-r)}
    const float4* ctaps4_rev; // [ntaps]  (t.re, t.im, -t.im, t.re) of ct_rev[r]: the two f32x2 multiplicands of a
                              //          packed complex MAC  acc += x.re * (t.re, t.im) + x.im * (-t.im, t.re)
    float2* out_iq;           // [nout] this block's outputs
    float* out_fm;            // [nout] or null
    float2* prev;             // device scalar: last output of the previous block (FM carry in)
    float2* prev_out;         // where this block's last output goes (the other half of the double buffer)
    double cyc;               // frac(f0 * D / fs)  (cycles per output, reduced mod 1)
    double phase0;            // frac(cyc * i_first)
    long long s_first;        // block-relative index of the newest input sample of output 0 (may be < 0)
    long long s_open;         // block-relative index of the first sample the channel ever saw: older samples
                              // read as 0 (a new GNU Radio `channel` top_block starts with zero filter history)
    int ntaps;
    int decim;
    int nout;
    float gain;
    int fast;                 // 1: this block is computed by ddc_tile_kernel, ddc_bank_kernel skips it
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
    if (ch.fast || o0 >= ch.nout) return;
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

// ------------------------------------------------------------------------------------------------
// Tiled fast path.  Channels of one source that share (decim, ntaps) sit on a common decimation grid
// (outputs at stream positions that are multiples of decim), so they need the SAME input windows.
// A CTA stages the window of 8 consecutive outputs (7*D + K samples) in shared memory once and its 8
// warps form a 2 x 4 grid of (output quad, channel quad): every lane keeps 4 outputs x 4 channels = 16
// complex accumulators, so one tap step costs 4 LDS (samples) + 4 LDG (composite taps, coalesced, L1/L2
// resident) for 64 FMAs - versus 5 loads per 16 FMAs and per-element bounds checks in ddc_bank_kernel,
// which stays as the path for channels opened inside the current windows (zero-history gating).
// ------------------------------------------------------------------------------------------------
struct DdcGroupDev {
    int ch[16];        // indices into the DdcChanDev array, -1 = unused
    int nch;
    int decim, ntaps, nout;
    long long s_first;
};

// OQ = output groups per CTA (8 / 4 / 2 for groups of <= 4 / <= 8 / <= 16 channels), 8/OQ channel groups.
// CPW = channels per warp (4, or 1 for a lone channel), OPW = 16 / CPW consecutive outputs per warp: every lane
// keeps OPW x CPW = 16 complex accumulators either way (a lone channel would waste 3/4 of a 4 x 4 block).
// grid (ceil(nout / (OPW*OQ)), ngroups), block 256, dynamic smem ((OPW*OQ-1)*decim + ntaps) * 8 bytes
template <int OQ, int CPW = 4>
__global__ void __launch_bounds__(256) ddc_tile_kernel(const DdcChanDev* __restrict__ chans,
                                                       const DdcGroupDev* __restrict__ groups,
                                                       const float2* __restrict__ x, long long nsamp,
                                                       const float2* __restrict__ hist, int hist_cap) {
    constexpr int OPW = 16 / CPW;
    extern __shared__ __align__(16) float2 ddc_tile[];
    const DdcGroupDev& g = groups[blockIdx.y];
    const int o_base = blockIdx.x * (OPW * OQ);
    if (o_base >= g.nout) return;
    const int D = g.decim, K = g.ntaps;
    const long long w0 = g.s_first + (long long)o_base * D - (K - 1);
    const int tile_len = (OPW * OQ - 1) * D + K;
    for (int t = threadIdx.x; t < tile_len; t += 256) {
        const long long idx = w0 + t;
        ddc_tile[t] = (idx < nsamp) ? ddc_x_at(x, hist, hist_cap, idx) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int oq = warp % OQ, cq = warp / OQ;
    if (cq * CPW >= g.nch || o_base + oq * OPW >= g.nout) return;
    int cidx[CPW];
    const float4* tp[CPW];
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        cidx[i] = g.ch[cq * CPW + i];
        tp[i] = chans[cidx[i] >= 0 ? cidx[i] : g.ch[cq * CPW]].ctaps4_rev;
    }
    float2 acc[OPW][CPW];
#pragma unroll
    for (int q = 0; q < OPW; ++q)
#pragma unroll
        for (int i = 0; i < CPW; ++i) acc[q][i] = make_float2(0.f, 0.f);
    const float2* xt = ddc_tile + (oq * OPW) * D;
    for (int r = lane; r < K; r += 32) {
        float2 xv[OPW];
        float4 tv[CPW];
#pragma unroll
        for (int q = 0; q < OPW; ++q) xv[q] = xt[q * D + r];
#pragma unroll
        for (int i = 0; i < CPW; ++i) tv[i] = __ldg(tp[i] + r);
        // complex MAC = two packed FFMA2 (sample part broadcast, tap pair / rotated tap pair): 32 issue slots
        // per 16 MACs instead of 64 scalar FFMA
#pragma unroll
        for (int q = 0; q < OPW; ++q)
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                acc[q][i] = p2fmas(make_float2(tv[i].x, tv[i].y), xv[q].x, acc[q][i]);
                acc[q][i] = p2fmas(make_float2(tv[i].z, tv[i].w), xv[q].y, acc[q][i]);
            }
    }
    float2 mine = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < OPW; ++q)
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            float2 a = acc[q][i];
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, sft);
                a.y += __shfl_xor_sync(0xffffffffu, a.y, sft);
            }
            if (lane == q * CPW + i) mine = a;
        }
    if (lane < 16) {
        const int q = lane / CPW, i = lane % CPW;
        const int o = o_base + oq * OPW + q;
        int ci = cidx[0];
#pragma unroll
        for (int u = 1; u < CPW; ++u)
            if (i == u) ci = cidx[u];
        if (ci >= 0 && o < g.nout) {
            const DdcChanDev& ch = chans[ci];
            double ph = ch.phase0 + ch.cyc * (double)o;
            ph -= floor(ph);
            double sn, cs;
            sincospi(-2.0 * ph, &sn, &cs);
            const float cf = (float)cs, sf = (float)sn;
            ch.out_iq[o] = make_float2(fmaf(mine.x, cf, -mine.y * sf), fmaf(mine.x, sf, mine.y * cf));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A lone channel (BASELINE config 1: one rc_frontend/channel.py DDC on a 2.4 Msps source, D 96, 349 taps).
// ddc_tile_kernel gives a lone channel 16 outputs per warp with the LANES striding the taps: every MAC costs one LDS
// and the tile load is not overlapped (104 Gsps = 0.13 of the HBM roofline).  Here the window is cut into FRAMES of D
// samples, frame g = samples [B + gD, B + (g+1)D),  B = window start of output 0:
//
//     y[o] = sum_{p < P} sum_{q < D} ct_rev[pD + q] * X[o + p][q],      P = ceil(K / D)
//
// A lane owns one frame: it reads X[g][q] ONCE (frames sit ddc_lone_pitch(D) samples apart in shared memory: the 32
// lanes' 8-byte reads are 2-way conflicted at worst) and feeds P accumulators - the partial sums of outputs g, g-1, ..., g-P+1 - with taps that
// are the same for all lanes (broadcast LDS.128 = two taps (re, im)): P complex MACs per sample load instead of one.
// The P partials of an output meet by shuffle (lane o takes part[p] from lane o + p), so a warp of 32 frames completes
// 32 - (P-1) outputs and consecutive warps overlap by P - 1 frames.
// grid (ceil(nout / (W * (33 - P))),), block W * 32 (W = kDdcLoneWarps), dynamic smem 16 + taps D*PP*8 + frames
// (W*(33-P)+P-1) * ddc_lone_pitch(D) * 8 bytes
// ------------------------------------------------------------------------------------------------
constexpr int kDdcLoneWarps = 2;   // small CTAs (48 KB tile for D 96): four per SM, their fill and MAC phases interleave
// frame pitch in shared memory (samples): >= D + 2 (a bulk copy starts on a 16-byte boundary, i.e. up to one sample
// early, and moves an even number of samples) and = 2 (mod 4), so 16 lanes x 8 bytes hit 8 distinct bank pairs (2-way)
__host__ __device__ constexpr int ddc_lone_pitch(int D) { return D + 2 + ((2 - (D + 2)) & 3); }
template <int P>
__global__ void __launch_bounds__(kDdcLoneWarps * 32) ddc_lone_kernel(const DdcChanDev* __restrict__ chans, int ci,
                                                                      const float2* __restrict__ x, long long nsamp,
                                                                      const float2* __restrict__ hist, int hist_cap) {
    extern __shared__ __align__(16) unsigned char ddc_lone_smem[];
    const DdcChanDev ch = chans[ci];
    const int D = ch.decim, K = ch.ntaps;
    constexpr int S = 33 - P;                       // outputs a warp completes
    constexpr int F = kDdcLoneWarps * S + P - 1;    // frames of the CTA tile
    const int pitch = ddc_lone_pitch(D);
    constexpr int PP = (P + 1) & ~1;                // taps per q, padded to an even count (one LDS.128 = two taps)
    uint64_t* bar = reinterpret_cast<uint64_t*>(ddc_lone_smem);
    float2* s_taps = reinterpret_cast<float2*>(ddc_lone_smem + 16);       // [D][PP]: ct_rev[p * D + q] at [q][p], zero beyond K
    float2* s_x = s_taps + D * PP;                                        // [F][pitch]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o_base = blockIdx.x * (kDdcLoneWarps * S);
    if (o_base >= ch.nout) return;
    // frames o_base .. o_base + F - 1 = F * D consecutive samples starting at b0
    const long long b0 = ch.s_first - (K - 1) + (long long)o_base * D;
    // interior tile: every frame arrives by ONE bulk copy (cp.async.bulk, SASS UBLKCP) - asynchronous, no registers, the
    // whole 47 KB in flight at once.  (Version 1 used 8-byte cp.async: LDGSTS.64 costs one LSU wavefront PER LANE, the
    // fill took three times as long as the MACs; a register-staged loop kept only the 8 loads in flight the scheduler
    // chose to: 4 GB/s per CTA by Little's law, profiles/r02_ddc_lone_v3_summary.txt.)
    const bool interior = (b0 >= 1 && b0 + (long long)F * D + 2 <= nsamp);
    if (threadIdx.x == 0) {
        mbar_init(bar, kDdcLoneWarps * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float2* src0 = x + b0;
    if (interior) {
        // every thread starts the copies of its own frames (UBLKCP is issued lane by lane: spread over all warps), after
        // posting their byte count: the barrier expects one arrival per thread
        uint32_t bytes = 0;
        for (int g = threadIdx.x; g < F; g += kDdcLoneWarps * 32) {
            const uint32_t off = (uint32_t)((reinterpret_cast<uintptr_t>(src0 + (long long)g * D) >> 3) & 1u);
            bytes += (((uint32_t)D + off + 1u) & ~1u) * 8u;
        }
        mbar_expect_tx(bar, bytes);
        for (int g = threadIdx.x; g < F; g += kDdcLoneWarps * 32) {
            const float2* sg = src0 + (long long)g * D;
            const uint32_t off = (uint32_t)((reinterpret_cast<uintptr_t>(sg) >> 3) & 1u);
            tma_bulk_g2s(s_x + g * pitch, sg - off, (((uint32_t)D + off + 1u) & ~1u) * 8u, bar);
        }
    }
    for (int i = threadIdx.x; i < D * PP; i += kDdcLoneWarps * 32) {
        const int q = i / PP, p = i - q * PP;
        const int r = p * D + q;
        s_taps[i] = (p < P && r < K) ? __ldg(ch.ctaps_rev + r) : make_float2(0.f, 0.f);
    }
    if (!interior) {
        // block edges: history before the block, zeros after it
        for (int g = warp; g < F; g += kDdcLoneWarps) {
            const long long fb = b0 + (long long)g * D;
            for (int q = lane; q < D; q += 32) {
                const long long idx = fb + q;
                s_x[g * pitch + q] = (idx < nsamp) ? ddc_x_at(x, hist, hist_cap, idx) : make_float2(0.f, 0.f);
            }
        }
    }
    __syncthreads();
    if (interior) mbar_wait(bar, 0);
    const int g = warp * S + lane;  // this lane's frame within the tile
    // complex MAC as two real-scalar MACs on the (re, im) pair of the tap:  A += t * x.re,  B += t * x.im,
    // acc = (A.x - B.y, A.y + B.x) - only (re, im) of a tap is loaded, two taps per LDS.128
    float2 pa[P], pb[P];
#pragma unroll
    for (int p = 0; p < P; ++p) pa[p] = pb[p] = make_float2(0.f, 0.f);
    if (g < F) {
        // (a bulk copy started one sample early when the frame's first sample sits on an odd 8-byte address)
        const float2* xr = s_x + g * pitch +
                           (interior ? (int)((reinterpret_cast<uintptr_t>(src0 + (long long)g * D) >> 3) & 1u) : 0);
#pragma unroll 4
        for (int q = 0; q < D; ++q) {
            const float2 xv = xr[q];
            const float4* tq = reinterpret_cast<const float4*>(s_taps + q * PP);
#pragma unroll
            for (int pp = 0; pp < PP / 2; ++pp) {
                const float4 tv = tq[pp];
                pa[2 * pp] = p2fmas(make_float2(tv.x, tv.y), xv.x, pa[2 * pp]);
                pb[2 * pp] = p2fmas(make_float2(tv.x, tv.y), xv.y, pb[2 * pp]);
                if (2 * pp + 1 < P) {
                    pa[2 * pp + 1] = p2fmas(make_float2(tv.z, tv.w), xv.x, pa[2 * pp + 1]);
                    pb[2 * pp + 1] = p2fmas(make_float2(tv.z, tv.w), xv.y, pb[2 * pp + 1]);
                }
            }
        }
    }
    float2 part[P];
#pragma unroll
    for (int p = 0; p < P; ++p) part[p] = make_float2(pa[p].x - pb[p].y, pa[p].y + pb[p].x);
    // output o = frame index of its first frame; frame o + p holds its p-th partial
    float2 acc = part[0];
#pragma unroll
    for (int p = 1; p < P; ++p) {
        acc.x += __shfl_down_sync(0xffffffffu, part[p].x, p);
        acc.y += __shfl_down_sync(0xffffffffu, part[p].y, p);
    }
    const int o = o_base + g;
    if (lane < S && o < ch.nout) {
        double ph = ch.phase0 + ch.cyc * (double)o;
        ph -= floor(ph);
        double sn, cs;
        sincospi(-2.0 * ph, &sn, &cs);
        const float cf = (float)cs, sf = (float)sn;
        ch.out_iq[o] = make_float2(fmaf(acc.x, cf, -acc.y * sf), fmaf(acc.x, sf, acc.y * cf));
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

// FM demod of every channel's block + FM carry + wideband history update in ONE launch (was three).
// grid (max(ceil(max_nout / 256), ceil(cap / 256)), M + 1): rows 0..M-1 = channels, row M = history copy.
__global__ void __launch_bounds__(256) ddc_post_kernel(const DdcChanDev* __restrict__ chans, int M,
                                                       const float2* __restrict__ old_hist, const float2* __restrict__ x,
                                                       long long n, float2* __restrict__ new_hist, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)blockIdx.y == M) {
        if (i >= cap) return;
        const long long src = i + n - cap;  // index into x; negative -> old hist
        new_hist[i] = (src >= 0) ? x[src] : ((src >= -cap) ? old_hist[cap + src] : make_float2(0.f, 0.f));
        return;
    }
    const DdcChanDev ch = chans[blockIdx.y];
    if (i >= ch.nout) return;
    const int o = (int)i;
    const float2 c = ch.out_iq[o];
    if (ch.out_fm) {
        const float2 pv = (o == 0) ? *ch.prev : ch.out_iq[o - 1];
        const float2 pr = cmul_conj(c, pv);
        ch.out_fm[o] = ch.gain * atan2_fast(pr.y, pr.x);
    }
    if (o == ch.nout - 1) *ch.prev_out = c;
}

// rows of 32-bit words from M separately allocated channel buffers into one dense [M][stride] block (rcb_ddc_pull_all)
__global__ void __launch_bounds__(256) ddc_gather_kernel(const void* const* __restrict__ src, const int* __restrict__ cnt,
                                                         int words_per_item, unsigned* __restrict__ dst, long long stride) {
    const int r = blockIdx.y;
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)cnt[r] * words_per_item) return;
    dst[(long long)r * stride + w] = reinterpret_cast<const unsigned*>(src[r])[w];
}

// new_hist (cap samples) = last `cap` samples of (old_hist ++ x[0..n))
__global__ void hist_update_kernel(const float2* __restrict__ old_hist, const float2* __restrict__ x,
                                   long long n, float2* __restrict__ new_hist, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const long long src = i + n - cap;  // index into x; negative -> old hist
    new_hist[i] = (src >= 0) ? x[src] : ((src >= -cap) ? old_hist[cap + src] : make_float2(0.f, 0.f));
}

// same for raw integer I/Q kept as 16-bit units (u8 / s8: one unit per complex sample, s16: two): byte-exact copy
__global__ void hist_update_raw_kernel(const unsigned short* __restrict__ old_hist, const unsigned short* __restrict__ x,
                                       long long n, unsigned short* __restrict__ new_hist, long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const long long src = i + n - cap;
    new_hist[i] = (src >= 0) ? x[src] : ((src >= -cap) ? old_hist[cap + src] : (unsigned short)0);
}

// complex64 history from raw integer input: new_hist = last `cap` samples of (old_hist ++ convert(x[0..n)))
// fmt: 1 u8, 2 s8, 3 s16 (include/b200chan.h RCB_FMT_*); out = (v + offset) * scale like convert_iq_kernel
__global__ void hist_update_convert_kernel(const float2* __restrict__ old_hist, const void* __restrict__ x, int fmt,
                                           float offset, float scale, long long n, float2* __restrict__ new_hist,
                                           long long cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const long long src = i + n - cap;
    float2 v = make_float2(0.f, 0.f);
    if (src >= 0) {
        float a, b;
        if (fmt == 3) {
            const short* q = reinterpret_cast<const short*>(x) + 2 * src;
            a = (float)q[0];
            b = (float)q[1];
        } else if (fmt == 2) {
            const signed char* q = reinterpret_cast<const signed char*>(x) + 2 * src;
            a = (float)q[0];
            b = (float)q[1];
        } else {
            const unsigned char* q = reinterpret_cast<const unsigned char*>(x) + 2 * src;
            a = (float)q[0];
            b = (float)q[1];
        }
        v = make_float2((a + offset) * scale, (b + offset) * scale);
    } else if (src >= -cap) {
        v = old_hist[cap + src];
    }
    new_hist[i] = v;
}

}  // namespace rcb
