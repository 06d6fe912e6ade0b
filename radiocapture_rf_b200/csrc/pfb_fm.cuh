// K1  pfb_fm : critically-sampled polyphase filterbank channelizer with fused quadrature FM demod.
//
// Replaces, in one launch, the reference's  pfb.channelizer_ccf(N, taps, 1.0)  (rc_frontend/receiver.py:
// 249-261; GNU Radio gr-filter pfb_channelizer_ccf_impl.cc / polyphase_filterbank.cc) plus one
// analog.quadrature_demod_cf(gain) per output bin (moto_control_demod.py:105, edacs_control_demod.py:
// 82-84, p25_control_demod.py:120-121, logging_receiver.py:234):
//
//   u_i[n] = sum_k h[i + kN] x[(n-k)N + N-1-i]            (arm FIR, P = ceil(L/N) taps per arm)
//   Y_m[n] = sum_i u_i[n] e^{+j 2 pi m i / N}             (backward, unnormalised DFT over arms)
//   fm_m[n] = gain * atan2(Im p, Re p),  p = Y_m[n] conj(Y_m[n-1])
//
// Fast path (N = R*R, R in {8,16,32}):  HBM-bound streaming kernel, no tensor cores.
//   * one warp owns F = 32/R whole frames per iteration; lane (fr, ll) holds the R points
//     r = R*jj + ll of its frame  ->  every global load instruction is F fully used R*8-byte segments
//     (256 B / warp for N = 1024); each input byte is read from HBM exactly once (k = 0 row);
//     the k >= 1 rows of the arm FIR are L1/L2 hits (they are the k = 0 rows of the frames the other
//     warps of the CTA own) and the next frame's k = 0 row is prefetched into registers
//     before the CTA barrier, so the loads fly while the demod phase runs;
//   * N-point DFT = two in-register radix-R passes (fft_inreg.cuh) with the W_N^{l m1} twiddle and a
//     bank-conflict-free (row stride R+2) 32x32 shared-memory transpose in between - warp-private,
//     only __syncwarp();
//   * the CTA's 8 warps fill FPI = 8*F frame slots of a shared ring (+1 slot = previous frame), then
//     every thread demodulates 8 consecutive frames of one channel (conj-multiply + atan2_fast) and
//     emits them with ONE 256-bit store = one full 32 B sector of the channel-major [N][T] output.
//   * persistent grid (2 CTAs/SM), each CTA walks a contiguous run of iterations; the frame before the
//     run is recomputed (1 extra iteration per ~100) instead of carrying Y state between CTAs/launches,
//     so the only streaming state is the last P input rows ("hist").
// Algorithmic HBM bytes per input sample: 8 (read) + 4 (FM out) and/or 8 (IQ out).
#pragma once
#include "common.cuh"
#include "fft_inreg.cuh"

namespace rcb {

enum { PFB_OUT_IQ = 1, PFB_OUT_FM = 2 };

struct PfbParams {
    const float2* x;        // T*N new samples (row n = frame n)
    const float2* hist;     // P*N samples preceding x (rows -P .. -1)
    const float* taps;      // permuted: taps[(k*R + jj)*R + ll] = h[R*(R-1-jj) + (R-1-ll) + k*N]  (fast path)
                            // generic path: taps[k*N + i] = h[i + k*N]
    const float2* twiddle;  // fast: [R][R+2]: tw[ll*(R+2) + m1] = W_N^{+(R-1-ll) m1};  generic: [N] W_N^{+q}
    float* out_fm;          // [N][ostride] floats (or null)
    float2* out_iq;         // [N][ostride] complex (or null)
    long long ostride;      // elements between consecutive channels
    int T;                  // frames in this launch
    int P;                  // taps per arm
    int N;                  // channels
    float gain;
};

template <int R>
struct PfbGeom {
    static constexpr int N = R * R;
    static constexpr int F = 32 / R;
    static constexpr int WARPS = 8;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int FPI = WARPS * F;
    static constexpr int S = R + 2;
    static constexpr int FS = (R == 8) ? 88 : R * S;
    static constexpr int NSLOT = FPI + 1;
    static constexpr size_t ring_bytes = (size_t)NSLOT * FS * sizeof(float2);
    static constexpr size_t tw_bytes = (size_t)R * S * sizeof(float2);
    static size_t smem_bytes(int P, bool taps_smem) {
        return ring_bytes + tw_bytes + (taps_smem ? (size_t)P * N * sizeof(float) : 0);
    }
};

template <int R>
__device__ __forceinline__ const float2* pfb_row_ptr(const PfbParams& p, long long r) {
    constexpr int N = R * R;
    if (r >= 0) return (r < p.T) ? p.x + r * N : nullptr;
    return (r >= -(long long)p.P) ? p.hist + (r + p.P) * N : nullptr;
}

template <int R, int MODE, bool TAPS_SMEM, int PT /* compile-time P, 0 = runtime */>
__global__ void __launch_bounds__(256, 2) pfb_fm_kernel(const PfbParams p) {
    using G = PfbGeom<R>;
    constexpr int N = G::N, F = G::F, FPI = G::FPI, S = G::S, FS = G::FS, NSLOT = G::NSLOT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float2* tws = ring + NSLOT * FS;
    float* taps_s = reinterpret_cast<float*>(tws + R * S);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane / R, ll = lane % R;
    const int P = (PT > 0) ? PT : p.P;

    for (int i = tid; i < R * S; i += G::THREADS) tws[i] = p.twiddle[i];
    if (TAPS_SMEM)
        for (int i = tid; i < P * N; i += G::THREADS) taps_s[i] = p.taps[i];
    __syncthreads();
    const float* tap_base = TAPS_SMEM ? taps_s : p.taps;

    // contiguous run of iterations for this CTA
    const int NI = (p.T + FPI - 1) / FPI;
    const int per = NI / gridDim.x, rem = NI % gridDim.x;
    const int it0 = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int cnt = per + ((int)blockIdx.x < rem ? 1 : 0);
    if (cnt == 0) return;
    const int it1 = it0 + cnt;

    float2 xr[R];
    long long frame = (long long)(it0 - 1) * FPI + warp * F + fr;
    {
        const float2* rp = pfb_row_ptr<R>(p, frame);
#pragma unroll
        for (int jj = 0; jj < R; ++jj)
            xr[jj] = rp ? (PT == 1 ? ld_stream_f2(rp + jj * R + ll) : __ldg(rp + jj * R + ll)) : make_float2(0.f, 0.f);
    }

    int c = 0;
    for (int it = it0 - 1; it < it1; ++it, ++c) {
        // ---------------- phase 1: arm FIR + N-point backward DFT for this warp's F frames -------------
        float2 v[R];  // v[j], j = R-1-jj  (DFT input index i = R*j + (R-1-ll))
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
            const float h = tap_base[jj * R + ll];
            v[R - 1 - jj] = make_float2(h * xr[jj].x, h * xr[jj].y);
        }
        for (int k = 1; k < P; ++k) {
            const float2* rp = pfb_row_ptr<R>(p, frame - k);
            if (rp) {
                const float* tk = tap_base + k * N;
#pragma unroll
                for (int jj = 0; jj < R; ++jj) {
                    const float2 xv = __ldg(rp + jj * R + ll);
                    const float h = tk[jj * R + ll];
                    v[R - 1 - jj].x = fmaf(h, xv.x, v[R - 1 - jj].x);
                    v[R - 1 - jj].y = fmaf(h, xv.y, v[R - 1 - jj].y);
                }
            }
        }
        const int slot = (c * FPI + warp * F + fr + 1) % NSLOT;
        float2* buf = ring + slot * FS;
        warp_fft_2pass<R, +1, true>(v, buf, tws, ll);  // v[m2] = Y[ll + R*m2]
#pragma unroll
        for (int m2 = 0; m2 < R; ++m2) buf[m2 * S + ll] = v[m2];

        // prefetch the k = 0 row of this warp's next frame; the loads fly during phase 2
        frame += FPI;
        if (it + 1 < it1) {
            const float2* rp = pfb_row_ptr<R>(p, frame);
#pragma unroll
            for (int jj = 0; jj < R; ++jj)
                xr[jj] = rp ? (PT == 1 ? ld_stream_f2(rp + jj * R + ll) : __ldg(rp + jj * R + ll)) : make_float2(0.f, 0.f);
        }
        __syncthreads();

        // ---------------- phase 2: demod 8 consecutive frames of one channel, 32 B store ---------------
        if (it >= it0) {
            constexpr int ITEMS = N * (FPI / 8) / G::THREADS;
#pragma unroll
            for (int q = 0; q < ITEMS; ++q) {
                const int item = q * G::THREADS + tid;
                const int m = item % N, g = item / N;
                const int pos = (m / R) * S + (m % R);
                const int base = c * FPI + 8 * g;  // slot of frame f is (base + j + 1) % NSLOT, j = f - 8g
                const long long t0 = (long long)it * FPI + 8 * g;
                float2 y[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) y[j] = ring[((base + j) % NSLOT) * FS + pos];
                const bool full = (t0 + 8 <= p.T);
                if (MODE & PFB_OUT_FM) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 pr = cmul_conj(y[j + 1], y[j]);
                        o[j] = p.gain * atan2_fast(pr.y, pr.x);
                    }
                    float* dst = p.out_fm + (long long)m * p.ostride + t0;
                    if (full) {
                        st_global_v8(dst, o);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (t0 + j < p.T) dst[j] = o[j];
                    }
                }
                if (MODE & PFB_OUT_IQ) {
                    float2* dst = p.out_iq + (long long)m * p.ostride + t0;
                    if (full) {
                        float o[8];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                o[2 * j] = y[1 + 4 * hh + j].x;
                                o[2 * j + 1] = y[1 + 4 * hh + j].y;
                            }
                            st_global_v8(reinterpret_cast<float*>(dst + 4 * hh), o);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (t0 + j < p.T) dst[j] = y[j + 1];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Generic path: any N >= 1 (the reference's own PFB shapes are N = fs/400 kHz = 5..40, not powers of
// two, rc_frontend/receiver.py:244-249).  Direct O(N^2) DFT per frame out of shared memory; writes
// channel-major IQ; FM is produced by quad_demod_rows (demod.cuh) over the IQ rows.  Correctness
// path, not the benchmarked one.
// ------------------------------------------------------------------------------------------------
__global__ void pfb_generic_kernel(const PfbParams p, float2* __restrict__ y_out, long long ystride) {
    // column c of y_out holds frame c-1 (column 0 = the frame before the block, for the FM carry)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* u = reinterpret_cast<float2*>(smem_raw);  // [N]
    float2* tw = u + p.N;                             // [N]
    const int N = p.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = p.twiddle[i];
    for (long long col = blockIdx.x; col <= p.T; col += gridDim.x) {
        const long long n = col - 1;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float2 acc = make_float2(0.f, 0.f);
            for (int k = 0; k < p.P; ++k) {
                const long long r = n - k;
                const float2* rp = (r >= 0) ? p.x + r * N : ((r >= -(long long)p.P) ? p.hist + (r + p.P) * N : nullptr);
                if (rp) {
                    const float2 xv = rp[N - 1 - i];
                    const float h = p.taps[k * N + i];
                    acc.x = fmaf(h, xv.x, acc.x);
                    acc.y = fmaf(h, xv.y, acc.y);
                }
            }
            u[i] = acc;
        }
        __syncthreads();
        for (int m = threadIdx.x; m < N; m += blockDim.x) {
            float2 acc = make_float2(0.f, 0.f);
            int q = 0;
            for (int i = 0; i < N; ++i) {
                const float2 w = tw[q];
                const float2 a = u[i];
                acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
                acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
                q += m;
                if (q >= N) q -= N;
            }
            y_out[(long long)m * ystride + col] = acc;
        }
    }
}

}  // namespace rcb
