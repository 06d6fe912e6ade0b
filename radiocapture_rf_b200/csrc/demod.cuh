// K4  quadrature FM demod over already-channelised complex streams, and the AFC mean probe.
//
// quad_demod_rows: out[r][n] = gain * atan2(Im p, Re p), p = x[r][n] conj(x[r][n-1]), x[r][-1] = prev[r]
//   = analog.quadrature_demod_cf(gain)  (gr-analog quadrature_demod_cf_impl.cc; reference call sites
//   moto_control_demod.py:105, edacs_control_demod.py:82-84, p25_control_demod.py:120-121,
//   logging_receiver.py:214,234,336,346), one row per channel.  12 algorithmic bytes / sample.
// window_sum_rows: scale * sum of the last `length` samples of each row = what
//   moving_average_ff(length, 1, ...) -> multiply_const(scale) -> probe_signal_f holds after the block
//   (p25_control_demod.py:123-127, moto_control_demod.py:121-125, edacs_control_demod.py:98-101).
#pragma once
#include "common.cuh"

namespace rcb {

// grid: (ceil(n / (256*4)), rows).  x row stride xs (complex), out row stride os (float).
__global__ void __launch_bounds__(256) quad_demod_rows_kernel(const float2* __restrict__ x, long long xs,
                                                              const float2* __restrict__ prev,
                                                              float* __restrict__ out, long long os,
                                                              long long n, float gain, long long x_col0) {
    const int r = blockIdx.y;
    const float2* xr = x + (long long)r * xs + x_col0;
    float* orow = out + (long long)r * os;
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    float2 p = (i0 == 0) ? (prev ? prev[r] : make_float2(0.f, 0.f)) : xr[i0 - 1];
    float o[4];
    int cnt = (int)min((long long)4, n - i0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < cnt) {
            const float2 c = xr[i0 + j];
            const float2 pr = cmul_conj(c, p);
            o[j] = gain * atan2_fast(pr.y, pr.x);
            p = c;
        }
    }
    if (cnt == 4 && ((os & 3) == 0) && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
        *reinterpret_cast<float4*>(orow + i0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        for (int j = 0; j < cnt; ++j) orow[i0 + j] = o[j];
    }
}

// last sample of every row -> prev[r]  (streaming carry for the next block)
__global__ void save_last_kernel(const float2* __restrict__ x, long long xs, long long last_col,
                                 float2* __restrict__ prev, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) prev[r] = x[(long long)r * xs + last_col];
}

// generic-PFB emit: ys is [N][T+1] (column 0 = frame -1).  Writes IQ and / or FM channel-major.
__global__ void __launch_bounds__(256) pfb_generic_emit_kernel(const float2* __restrict__ ys, long long ystride,
                                                               float2* __restrict__ out_iq,
                                                               float* __restrict__ out_fm, long long ostride,
                                                               int T, float gain, int oblock_log2, int N) {
    const int m = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float2 c = ys[(long long)m * ystride + t + 1];
    const long long oi = (oblock_log2 > 0)
                             ? (((((t >> oblock_log2) * N) + m) << oblock_log2) | (t & ((1LL << oblock_log2) - 1)))
                             : ((long long)m * ostride + t);
    if (out_iq) out_iq[oi] = c;
    if (out_fm) {
        const float2 pv = ys[(long long)m * ystride + t];
        const float2 pr = cmul_conj(c, pv);
        out_fm[oi] = gain * atan2_fast(pr.y, pr.x);
    }
}

// K5  ingest conversion: interleaved integer I/Q as the SDR delivers it over the wire -> complex64.
//   fmt 1: u8  offset-binary (RTL-SDR; gr-osmosdr rtl_source_c: (v - 127.4) / 128)
//   fmt 2: s8  (USRP otw_format sc8, configs/config_denver_usrp.py:20)        fmt 3: s16 (sc16)
// out = (v + offset) * scale, 4 samples per thread (8..16 B in, 32 B out).  SURVEY 8(f) row 4.
template <typename T>
__global__ void __launch_bounds__(256) convert_iq_kernel(const T* __restrict__ in, float2* __restrict__ out,
                                                         long long nsamp, float offset, float scale) {
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= nsamp) return;
    if (i0 + 4 <= nsamp) {
        T v[8];
        if constexpr (sizeof(T) == 1) {
            const uint2 w = *reinterpret_cast<const uint2*>(in + 2 * i0);
            *reinterpret_cast<uint2*>(v) = w;
        } else {
            const uint4 w = *reinterpret_cast<const uint4*>(in + 2 * i0);
            *reinterpret_cast<uint4*>(v) = w;
        }
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = ((float)v[j] + offset) * scale;
        float4* o = reinterpret_cast<float4*>(out + i0);
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    } else {
        for (long long i = i0; i < nsamp; ++i)
            out[i] = make_float2(((float)in[2 * i] + offset) * scale, ((float)in[2 * i + 1] + offset) * scale);
    }
}

// one CTA per row: scale * sum of x[r][n-length .. n-1] (clamped at 0), double accumulate.
__global__ void __launch_bounds__(256) window_sum_rows_kernel(const float* __restrict__ x, long long xs,
                                                              long long n, long long length, float scale,
                                                              float* __restrict__ out) {
    const int r = blockIdx.x;
    const float* xr = x + (long long)r * xs;
    const long long lo = (n > length) ? n - length : 0;
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < n; i += blockDim.x) acc += (double)xr[i];
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[r] = (float)(sh[0] * (double)scale);
}

}  // namespace rcb
