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
#include "fft_packed.cuh"

namespace rcb {

enum { PFB_OUT_IQ = 1, PFB_OUT_FM = 2, PFB_LOGPOW = 4 /* K3 row pass on the K1 pipeline (pfb_fm_tma.cuh) */ };

struct PfbParams {
    const float2* x;        // T*N new samples (row n = frame n)
    const float2* hist;     // P*N samples preceding x (rows -P .. -1)
    const float* taps;      // fast path, float4 groups: taps[((k*(R/4) + jj/4)*R + ll)*4 + jj%4] =
                            //   h[R*(R-1-jj) + (R-1-ll) + k*N];  generic path: taps[k*N + i] = h[i + k*N]
    const float* taps_kc;   // [P][N]: taps_kc[k*N + c] = h[(N-1-c) + k*N] (tap of column c; pfb_fm_tma_kernel, P > 1)
    const float2* zeros;    // N complex zeros (rows outside the stream)
    int* work_counter;      // zeroed before each launch: dynamic tail-chunk counter (pfb_fm_tma_kernel)
    int* next_counter;      // pfb_fm1_kernel: the counter of the NEXT launch, zeroed by this one (no memset per launch)
    float2* hist_out;       // pfb_fm1_kernel<0>: receives the last input row of this block (no history kernel per launch)
    const float2* twiddle;  // fast: [R][R+2]: tw[ll*(R+2) + m1] = W_N^{+(R-1-ll) m1};  generic: [N] W_N^{+q}
    float* out_fm;          // [N][ostride] floats (or null)
    float2* out_iq;         // [N][ostride] complex (or null)
    long long ostride;      // elements between consecutive channels (plain channel-major layout)
    int oblock_log2;        // > 0: channel-major inside time blocks of 2^k frames: (m, t) at
                            //   ((t >> k) * N + m) << k | (t & (2^k - 1)); keeps the 1024 rows a CTA writes
                            //   per iteration within a few MB (TLB / DRAM-page locality, +15 % on cfg3)
    // fused integer ingest (pfb_fm1_kernel<FMT != 0>, pfb_cl_kernel): x / hist_raw hold raw interleaved integer I/Q
    // (RCB_FMT_U8 / S8 / S16), sample value = (v + in_off) * in_scale; rows before -hist_valid are zero samples
    const void* hist_raw;   // P rows of raw samples preceding x
    int in_fmt;             // 0 = complex64
    int hist_valid;         // rows of hist_raw that hold real samples (0 at stream start)
    float in_off, in_scale;
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
        return ring_bytes + tw_bytes + 16 /* mbarrier */ + (taps_smem ? (size_t)P * N * sizeof(float) : 0);
    }
};

__device__ __forceinline__ long long pfb_out_index(const PfbParams& p, int m, long long t) {
    if (p.oblock_log2 > 0) {
        const int k = p.oblock_log2;
        return ((((t >> k) * p.N) + m) << k) | (t & ((1LL << k) - 1));
    }
    return (long long)m * p.ostride + t;
}

template <int R>
__device__ __forceinline__ const float2* pfb_row_ptr(const PfbParams& p, long long r) {
    // rows outside [-P, T) read an all-zero row (p.zeros) so the loads stay unconditional
    constexpr int N = R * R;
    if (r >= 0) return (r < p.T) ? p.x + r * N : p.zeros;
    return (r >= -(long long)p.P) ? p.hist + (r + p.P) * N : p.zeros;
}

// ---- mbarrier (arrive early / wait late) used to overlap the next frame's FIR + first FFT pass with the
// ---- tail of the other warps' demod phase -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// atan2 variant for the phase ring: (0,0) -> NaN (0/0), which the demod turns into the reference's 0
__device__ __forceinline__ float atan2_nan(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float z = __fdividef(mn, mx);
    const float s = z * z;
    float r = -0.004370174370706081f;
    r = fmaf(r, s, 0.023092154413461685f);
    r = fmaf(r, s, -0.05784549191594124f);
    r = fmaf(r, s, 0.0979914739727974f);
    r = fmaf(r, s, -0.13978290557861328f);
    r = fmaf(r, s, 0.1996297985315323f);
    r = fmaf(r, s, -0.33331674337387085f);
    r = r * s;
    r = fmaf(r, z, z);
    r = (ay > ax) ? (1.57079632679489661923f - r) : r;
    r = (x < 0.0f) ? (3.14159265358979323846f - r) : r;
    return copysignf(r, y);
}

template <int R, int MODE, bool TAPS_SMEM, int PT /* compile-time P, 0 = runtime */, int PF = 0 /* 0: register prefetch, 1: L2 prefetch */>
__global__ void __launch_bounds__(256, 2) pfb_fm_kernel(const PfbParams p) {
    using G = PfbGeom<R>;
    constexpr int N = G::N, F = G::F, FPI = G::FPI, S = G::S, FS = G::FS, NSLOT = G::NSLOT;
    constexpr bool PHI = (MODE == PFB_OUT_FM);  // FM only: the ring holds angle(Y) (4 B) instead of Y (8 B)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float2* tws = ring + NSLOT * FS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tws + R * S);
    float* taps_s = reinterpret_cast<float*>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane / R, ll = lane % R;
    const int P = (PT > 0) ? PT : p.P;

    for (int i = tid; i < R * S; i += G::THREADS) tws[i] = p.twiddle[i];
    if (TAPS_SMEM)
        for (int i = tid; i < P * N; i += G::THREADS) taps_s[i] = p.taps[i];
    if (tid == 0) mbar_init(bar, G::THREADS);
    __syncthreads();
    const float4* tap4 = reinterpret_cast<const float4*>(TAPS_SMEM ? taps_s : p.taps);

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
        const float2* rp = pfb_row_ptr<R>(p, frame) + ll;
#pragma unroll
        for (int jj = 0; jj < R; ++jj) xr[jj] = (PT == 1) ? ld_stream_f2(rp + jj * R) : __ldg(rp + jj * R);
    }

    int base_slot = 0;  // slot of the frame before this iteration's first frame; frame f -> base_slot + f + 1
    uint32_t parity = 0;
#pragma unroll 1
    for (int it = it0 - 1; it < it1; ++it) {
        // ---------------- phase 1: arm FIR + N-point backward DFT for this warp's F frames -------------
        if constexpr (PF == 1) {
            if (it > it0 - 1) {  // the row was pulled into L2 before the demod phase: load it now
                const float2* rp = pfb_row_ptr<R>(p, frame) + ll;
#pragma unroll
                for (int jj = 0; jj < R; ++jj) xr[jj] = (PT == 1) ? ld_stream_f2(rp + jj * R) : __ldg(rp + jj * R);
            }
        }
        float2 v[R];  // v[j], j = R-1-jj  (DFT input index i = R*j + (R-1-ll))
#pragma unroll
        for (int jq = 0; jq < R / 4; ++jq) {
            const float4 h = tap4[jq * R + ll];
            v[R - 1 - (4 * jq + 0)] = make_float2(h.x * xr[4 * jq + 0].x, h.x * xr[4 * jq + 0].y);
            v[R - 1 - (4 * jq + 1)] = make_float2(h.y * xr[4 * jq + 1].x, h.y * xr[4 * jq + 1].y);
            v[R - 1 - (4 * jq + 2)] = make_float2(h.z * xr[4 * jq + 2].x, h.z * xr[4 * jq + 2].y);
            v[R - 1 - (4 * jq + 3)] = make_float2(h.w * xr[4 * jq + 3].x, h.w * xr[4 * jq + 3].y);
        }
#pragma unroll 1
        for (int k = 1; k < P; ++k) {
            const float2* rp = pfb_row_ptr<R>(p, frame - k) + ll;
            const float4* tk = tap4 + k * (N / 4);
#pragma unroll
            for (int jq = 0; jq < R / 4; ++jq) {
                const float4 h = tk[jq * R + ll];
                const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 xv = __ldg(rp + (4 * jq + u) * R);
                    v[R - 1 - (4 * jq + u)].x = fmaf(hh[u], xv.x, v[R - 1 - (4 * jq + u)].x);
                    v[R - 1 - (4 * jq + u)].y = fmaf(hh[u], xv.y, v[R - 1 - (4 * jq + u)].y);
                }
            }
        }
        fft_inreg<R, +1>(v);  // pass 1 needs no shared memory: overlaps the other warps' demod tail

        int slot = base_slot + warp * F + fr + 1;
        slot = (slot >= NSLOT) ? slot - NSLOT : slot;
        float2* buf = ring + slot * FS;
        if (it > it0 - 1) {  // everyone finished reading the ring in the previous demod phase?
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        warp_fft_xpose_pass2<R, +1, true>(v, buf, tws, ll);  // v[m2] = Y[ll + R*m2]
        if constexpr (PHI) {
            float* fb = reinterpret_cast<float*>(buf);
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) {
                fb[m2 * R + ll] = atan2_nan(v[m2].y, v[m2].x);
                if ((m2 & 3) == 3) asm volatile("" ::: "memory");  // bound the atan2 ILP (register pressure)
            }
        } else {
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) buf[m2 * S + ll] = v[m2];
        }

        // prefetch the k = 0 row of this warp's next frame; the loads fly during the demod phase
        frame += FPI;
        if (it + 1 < it1) {
            if constexpr (PF == 1) {
                // one prefetch per 128 B line of the F*N*8-byte chunk this warp reads next
                const char* base = reinterpret_cast<const char*>(pfb_row_ptr<R>(p, frame - fr)) ;
                if (frame - fr >= 0 && frame - fr + F <= p.T) {
#pragma unroll
                    for (int q = 0; q < (F * N * 8) / (128 * 32); ++q)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (q * 32 + lane) * 128));
                }
            } else {
                const float2* rp = pfb_row_ptr<R>(p, frame) + ll;
#pragma unroll
                for (int jj = 0; jj < R; ++jj) xr[jj] = (PT == 1) ? ld_stream_f2(rp + jj * R) : __ldg(rp + jj * R);
            }
        }
        __syncthreads();

        // ---------------- phase 2: 8 consecutive frames of CPT channels per thread, 32 B stores --------
        if (it >= it0) {
            if constexpr (PHI) {
                constexpr int CPT = N * (FPI / 8) / G::THREADS;  // 4, 2, 1 for R = 32, 16, 8
                const int g = tid / (N / CPT);
                const int m0 = (tid % (N / CPT)) * CPT;
                const long long t0 = (long long)it * FPI + 8 * g;
                int s = base_slot + 8 * g;
                s = (s >= NSLOT) ? s - NSLOT : s;
                const bool full = (t0 + 8 <= p.T);
                constexpr int W = (CPT >= 2) ? 2 : 1;  // channels handled together (register footprint)
#pragma unroll
                for (int hq = 0; hq < CPT / W; ++hq) {
                    float ph[9][W];
                    int sj = s;
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        const float* src = reinterpret_cast<const float*>(ring + sj * FS) + m0 + hq * W;
                        if constexpr (W == 2) {
                            const float2 t = *reinterpret_cast<const float2*>(src);
                            ph[j][0] = t.x; ph[j][1] = t.y;
                        } else {
                            ph[j][0] = *src;
                        }
                        sj = (sj + 1 == NSLOT) ? 0 : sj + 1;
                    }
#pragma unroll
                    for (int q = 0; q < W; ++q) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float d = ph[j + 1][q] - ph[j][q];                       // in (-2pi, 2pi)
                            const float k = (d * 0.15915494309189535f + 12582912.0f) - 12582912.0f;  // rint(d/2pi)
                            d = fmaf(k, -6.283185307179586f, d);
                            d *= p.gain;
                            o[j] = (d != d) ? 0.0f : d;                             // Y == 0 -> 0 like fast_atan2f
                        }
                        float* dst = p.out_fm + pfb_out_index(p, m0 + hq * W + q, t0);
                        if (full) {
                            st_global_v8(dst, o);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (t0 + j < p.T) dst[j] = o[j];
                        }
                    }
                }
            } else {
                constexpr int ITEMS = N * (FPI / 8) / G::THREADS;
#pragma unroll 1
                for (int q = 0; q < ITEMS; ++q) {
                    const int item = q * G::THREADS + tid;
                    const int m = item % N, g = item / N;
                    const int pos = (m / R) * S + (m % R);
                    const long long t0 = (long long)it * FPI + 8 * g;
                    int s = base_slot + 8 * g;
                    s = (s >= NSLOT) ? s - NSLOT : s;
                    float2 y[9];
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        y[j] = ring[s * FS + pos];
                        s = (s + 1 == NSLOT) ? 0 : s + 1;
                    }
                    const bool full = (t0 + 8 <= p.T);
                    if (MODE & PFB_OUT_FM) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 pr = cmul_conj(y[j + 1], y[j]);
                            o[j] = p.gain * atan2_fast(pr.y, pr.x);
                        }
                        float* dst = p.out_fm + pfb_out_index(p, m, t0);
                        if (full) {
                            st_global_v8(dst, o);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (t0 + j < p.T) dst[j] = o[j];
                        }
                    }
                    if (MODE & PFB_OUT_IQ) {
                        float2* dst = p.out_iq + pfb_out_index(p, m, t0);
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
        }
        mbar_arrive(bar);  // this thread is done reading the ring
        base_slot = (base_slot == 0) ? NSLOT - 1 : base_slot - 1;  // FPI == -1 (mod NSLOT)
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

namespace rcb {

// ---------------------------------------------------------------------------------------------------------------------
// Arm FIR for MANY taps per arm (P > 16): u[t][c] = sum_k h_c[k] x[t - k][c], the polyphase filter alone.  The FFT + FM
// demod of u then run as a one-tap channelizer with unit taps (pfb_fm1_kernel / pfb_fm_tma_kernel), so the third reading of
// BASELINE's "256-tap / 1024-channel" (256 taps PER ARM, compute bound: 512 fp32 lane-operations per sample) runs at
// FP32-pipe speed instead of re-reading 256 rows per frame through L2 as pfb_fm_kernel does.
// CTA = 16 columns x 128 frames: the (128 + P - 1) x 16 input tile and the P x 16 taps are staged in shared memory once
// (L2 read amplification (127 + P) / 128), a thread owns one column for 8 consecutive frames and slides an 8-sample
// register window over the taps (unrolled by 8 so the ring is statically indexed): per tap one LDS.64 (new sample), one
// LDS.32 (tap, broadcast across the half-warps) and 8 packed FFMA2.
// grid (N / 16, ceil(count / 128)), 256 threads, dynamic smem ((128 + P - 1) * 16 * 8 + P * 16 * 4) bytes.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pfb_arm_fir_kernel(const float2* __restrict__ x, const float2* __restrict__ hist,
                                                          const float* __restrict__ taps_kc, float2* __restrict__ u,
                                                          int N, int P, long long T, long long t_begin, long long count) {
    constexpr int CB = 16, TB = 128, OPT = 8;
    extern __shared__ __align__(16) unsigned char fir_smem[];
    const int rows = TB + P - 1;
    float2* xt = reinterpret_cast<float2*>(fir_smem);              // [rows][CB]: row i <-> frame t0 - (P - 1) + i
    float* ht = reinterpret_cast<float*>(xt + (size_t)rows * CB);  // [P][CB]
    const int tid = threadIdx.x, c = tid & 15, g = tid >> 4;
    const int c0 = blockIdx.x * CB;
    const long long t0 = t_begin + (long long)blockIdx.y * TB;
    for (int i = tid; i < rows * CB; i += 256) {
        const int r = i / CB, cc = i % CB;
        const long long f = t0 - (P - 1) + r;
        float2 v = make_float2(0.f, 0.f);
        if (f >= 0) {
            if (f < T) v = __ldg(x + f * N + c0 + cc);
        } else if (f >= -(long long)P) {
            v = __ldg(hist + (f + P) * N + c0 + cc);
        }
        xt[i] = v;
    }
    for (int i = tid; i < P * CB; i += 256) ht[i] = __ldg(taps_kc + (size_t)(i / CB) * N + c0 + (i % CB));
    __syncthreads();
    // outputs t0 + 8 g + j, j = 0..7; ring[p & 7] = x[p] for the 8 samples the current tap needs
    const int rb = (P - 1) + OPT * g;      // tile row of frame t0 + 8 g
    float2 ring[OPT], acc[OPT];
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        ring[j] = xt[(rb + j) * CB + c];
        acc[j] = make_float2(0.f, 0.f);
    }
    for (int k0 = 0; k0 < P; k0 += OPT) {
#pragma unroll
        for (int kk = 0; kk < OPT; ++kk) {
            const int k = k0 + kk;
            if (k < P) {
                const float h = ht[k * CB + c];
#pragma unroll
                for (int j = 0; j < OPT; ++j) acc[j] = p2fmas(ring[(j - kk + 8 * OPT) % OPT], h, acc[j]);
                // next tap needs x[t - k - 1]: it replaces the newest sample of the window (same residue mod 8)
                if (k + 1 < P) ring[(OPT - 1 - kk) % OPT] = xt[(rb - k - 1) * CB + c];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        const long long t = t0 + OPT * g + j;
        if (t < t_begin + count && t < T) u[(t - t_begin) * N + c0 + c] = acc[j];
    }
}

}  // namespace rcb
