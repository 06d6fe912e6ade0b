// K1 (multi-tap, FM-only) pfb_fm_ws : warp-specialised producer / consumer version of the polyphase channelizer
// for 2..16 taps per arm.
//
// The phase-serial kernel (pfb_fm_tma_kernel, PT > 1) runs arm FIR -> CTA barrier -> FFT -> CTA barrier -> demod
// with every warp in the same phase: the FIR phase waits on its global loads with nothing else to issue
// (profiles/r01_pfb_fm_p16_v2_w16_summary.txt: issue slots 35 % busy, long_sb + lg = 28 % of the stall samples).
// Here one 16-warp CTA per SM is split into
//   * 8 PRODUCER warps: nothing but the time-blocked arm FIR, one buffer set ahead (one column x 8 frames per task,
//     PT packed FFMA2 per output, written to one of two sets of frame buffers);
//   * 8 CONSUMER warps: one frame each per iteration - both packed radix-R passes (in-place swizzled transpose in
//     the frame buffer), packed atan2, angle ring - then, after a named barrier of the consumer warps only, the
//     demod + sector stores of that iteration (with the demod on the producer side the producers were the critical
//     path and the consumers spun 35 % of the time);
// handing the frame buffers back and forth through full / empty mbarriers per buffer set (one elected lane per warp
// arriving after __syncwarp()).  The producers' load latency now overlaps the consumers' arithmetic instead of
// stalling the whole SM.
// Arithmetic, ring layout, output layout and the warm-up-iteration scheme (no state carried between CTAs or
// launches) are those of pfb_fm_tma_kernel; shared memory: 2 x 8 x 8 KB frame buffers + 36 KB ring + 8 KB twiddles.
#pragma once
#include "pfb_fm_tma.cuh"

namespace rcb {

template <int R, int PW = 8, int MODE = PFB_OUT_FM>
struct PfbWsGeom {
    static constexpr int N = R * R;
    static constexpr int F = 32 / R;
    static constexpr int CW = 8;  // consumer warps; PW producer warps (8, or 12 with setmaxnreg register rebalancing)
    static constexpr int THREADS = (CW + PW) * 32;
    static constexpr int FPI = CW * F;
    static constexpr int NSLOT = FPI + 1;
    static constexpr int FSW = N + (R == 8 ? 8 : 0);
    static constexpr int WORK = F * FSW;  // complex per consumer warp and buffer set
    static constexpr size_t set_bytes = (size_t)CW * WORK * 8;
    static constexpr size_t ring_bytes = (size_t)NSLOT * N * (MODE == PFB_OUT_FM ? 4 : 8);  // angle(Y), or Y when IQ is an output
    static constexpr size_t tw_bytes = (size_t)N * 8;
    static constexpr size_t smem_bytes(int /*PT*/) { return 2 * set_bytes + ring_bytes + tw_bytes + 128; }
};

template <int R, int PT, int PW = 8, int MODE = PFB_OUT_FM>
__global__ void __launch_bounds__((8 + PW) * 32, 1) pfb_fm_ws_kernel(const PfbParams p) {
    using G = PfbWsGeom<R, PW, MODE>;
    constexpr int N = G::N, F = G::F, FPI = G::FPI, NSLOT = G::NSLOT, FSW = G::FSW, CW = G::CW;
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    float2* work_all = reinterpret_cast<float2*>(smem_raw128);  // [2 sets][CW][WORK]
    float* ring = reinterpret_cast<float*>(smem_raw128 + 2 * G::set_bytes);
    float2* tws = reinterpret_cast<float2*>(smem_raw128 + 2 * G::set_bytes + G::ring_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw128 + 2 * G::set_bytes + G::ring_bytes + G::tw_bytes);
    uint64_t* u_full = bars;       // [2] producers -> consumers: frame buffers of set s hold filtered frames
    uint64_t* u_empty = bars + 2;  // [2] consumers -> producers: set s has been transformed
    uint64_t* ring_free = bars + 4;  // consumer threads: arrive after the demod reads / wait before the next ring write

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // contiguous run of iterations per CTA (+ one warm-up iteration that recomputes the frame before the run)
    const int NI = (p.T + FPI - 1) / FPI;
    const int per = (NI + (int)gridDim.x - 1) / (int)gridDim.x;
    const int it0 = (int)blockIdx.x * per;
    const int it1 = min(it0 + per, NI);
    if (it0 >= NI) return;

    for (int i = tid; i < N; i += G::THREADS) tws[i] = p.twiddle[i];
    if (tid == 0) {
        mbar_init(u_full + 0, PW);
        mbar_init(u_full + 1, PW);
        mbar_init(u_empty + 0, CW);
        mbar_init(u_empty + 1, CW);
        mbar_init(ring_free, CW * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= CW) {
        // =========================== producers: arm FIR, one buffer set ahead ===========================
        // 12 producer warps: 640 threads start with 96 registers each; the producers give registers back so that the
        // consumer warpgroups can grow to 120 (640 x 96 = 384 x 80 + 256 x 120: the pool is what the CTA was launched with)
        if constexpr (PW > 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        const int ptid = tid - CW * 32;
        constexpr int PTHREADS = PW * 32;
        constexpr int TBK = 8, NX = TBK + PT - 1;
        constexpr int NTASK = N * (FPI / TBK);
        // (Staging the task windows one task ahead with cp.async was measured slower: 184 vs 200 Gsps on cfg3_p16 -
        // 8-byte LDGSTS cost an LSU wavefront each and 7x the shared-memory bank conflicts.)
        int k = 0;
        for (int it = it0 - 1; it < it1; ++it, ++k) {
            {
                const int s = k & 1;
                if (k >= 2) mbar_wait(u_empty + s, (uint32_t)(((k >> 1) - 1) & 1));
                float2* wset = work_all + (size_t)s * CW * G::WORK;
#pragma unroll 1
                for (int task = ptid; task < NTASK; task += PTHREADS) {
                    const int c = task % N, g = task / N;
                    const long long f0 = (long long)it * FPI + TBK * g;
                    const long long r0 = f0 - (PT - 1);
                    float2 xs[NX];
                    if (r0 >= 0 && f0 + TBK <= p.T) {
                        const float2* b = p.x + r0 * N + c;
#pragma unroll
                        for (int j = 0; j < NX; ++j) xs[j] = __ldg(b + (long long)j * N);
                    } else {
#pragma unroll
                        for (int j = 0; j < NX; ++j) xs[j] = __ldg(pfb_row_ptr<R>(p, r0 + j) + c);
                    }
                    float hk[PT];
#pragma unroll
                    for (int kk = 0; kk < PT; ++kk) hk[kk] = __ldg(p.taps_kc + kk * N + c);
#pragma unroll
                    for (int t = 0; t < TBK; ++t) {
                        float2 acc = p2muls(xs[t + PT - 1], hk[0]);
#pragma unroll
                        for (int kk = 1; kk < PT; ++kk) acc = p2fmas(xs[t + PT - 1 - kk], hk[kk], acc);
                        const int fi = TBK * g + t;
                        wset[(fi / F) * G::WORK + (fi % F) * FSW + c] = acc;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(u_full + s);
            }
        }
    } else {
        // =========================== consumers: one frame set per warp and iteration ===========================
        if constexpr (PW > 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        const int fr = lane / R, ll = lane % R;
        const uint64_t pol_stream = l2_policy_evict_first();
        int base_slot = 0;
        int k = 0;
        for (int it = it0 - 1; it < it1; ++it, ++k) {
            const int s = k & 1;
            float2* wf = work_all + (size_t)s * CW * G::WORK + warp * G::WORK + fr * FSW;
            if (PT == 16) {  // (measured: +7 % at 16 taps per arm, -4 % at 8)
                // the consumers have slack: pull the new rows the producers will need two iterations from now into L2,
                // so the FIR loads (the critical path) pay an L2 instead of an HBM latency
                const long long nf = (long long)(it + 2) * FPI;
                if (nf >= 0 && nf + FPI <= p.T && it + 2 < it1) {
                    const char* b = reinterpret_cast<const char*>(p.x + nf * N);
#pragma unroll
                    for (int u = 0; u < (FPI * N * 8) / (128 * CW * 32); ++u) prefetch_l2(b + (size_t)(u * CW * 32 + tid) * 128);
                }
            }
            mbar_wait(u_full + s, (uint32_t)((k >> 1) & 1));
            float2 pr[R / 2], pi[R / 2];
            {
                auto get = [&](auto j) { return wf[(R - 1 - decltype(j)::value) * R + ll]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);
            }
            __syncwarp();  // every lane has read its samples: the buffer becomes the transpose scratch
            {
                const int sw = pfb_swz<R>(ll);
                const float4* twp = reinterpret_cast<const float4*>(tws + ll * R);
                float4* bp = reinterpret_cast<float4*>(wf + ll * R);
#pragma unroll
                for (int c = 0; c < R / 2; ++c) {
                    const float4 t = twp[c ^ sw];
                    const float2 b0 = make_float2(fmaf(pr[c].x, t.x, -pi[c].x * t.y), fmaf(pr[c].x, t.y, pi[c].x * t.x));
                    const float2 b1 = make_float2(fmaf(pr[c].y, t.z, -pi[c].y * t.w), fmaf(pr[c].y, t.w, pi[c].y * t.z));
                    bp[c ^ sw] = make_float4(b0.x, b0.y, b1.x, b1.y);
                }
            }
            __syncwarp();
            {
                const int ch = ll >> 1, wi = ll & 1;
                float2 u[R];
#pragma unroll
                for (int l2 = 0; l2 < R; ++l2) u[R - 1 - l2] = wf[l2 * R + (((ch ^ pfb_swz<R>(l2)) << 1) | wi)];
                __syncwarp();  // frame buffer fully consumed: hand it back to the producers
                if (lane == 0) mbar_arrive(u_empty + s);
                auto get = [&](auto j) { return u[decltype(j)::value]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = Y[ll + R*(2q)], Y[ll + R*(2q+1)]
            }
            float ph[MODE == PFB_OUT_FM ? R : 1];
            if constexpr (MODE == PFB_OUT_FM) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const float2 a = atan2_nan_p2(pi[q], pr[q]);
                    ph[2 * q] = a.x;
                    ph[2 * q + 1] = a.y;
                }
            }
            int slot = base_slot + warp * F + fr + 1;
            slot = (slot >= NSLOT) ? slot - NSLOT : slot;
            if (k >= 1) mbar_wait(ring_free, (uint32_t)((k - 1) & 1));  // every consumer thread has read the ring
            if constexpr (MODE == PFB_OUT_FM) {
                float* fb = ring + slot * N;
#pragma unroll
                for (int m2 = 0; m2 < R; ++m2) fb[m2 * R + ll] = ph[m2];
            } else {  // IQ is an output: the ring holds Y
                float2* fb = reinterpret_cast<float2*>(ring) + slot * N;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    fb[(2 * q) * R + ll] = make_float2(pr[q].x, pi[q].x);
                    fb[(2 * q + 1) * R + ll] = make_float2(pr[q].y, pi[q].y);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");  // consumer warps only: this iteration's angles are in the ring
            // ---- demod: CPT channels x 8 consecutive frames per consumer thread ----
            if (it >= it0) {
                constexpr int CTHREADS = CW * 32;
                constexpr int CPT = N * (FPI / 8) / CTHREADS;  // 4, 2, 1 for R = 32, 16, 8
                constexpr int GRP = FPI / 8;
                constexpr bool kPairLanes = (GRP == 2 && CPT >= 2);
                const int g = kPairLanes ? (lane >> 4) : tid / (N / CPT);
                const int m0 = kPairLanes ? (warp * 16 + (lane & 15)) * CPT : (tid % (N / CPT)) * CPT;
                const long long t0 = (long long)it * FPI + 8 * g;
                int sl = base_slot + 8 * g;
                sl = (sl >= NSLOT) ? sl - NSLOT : sl;
                const bool full = (t0 + 8 <= p.T);
                const long long rowstride = (p.oblock_log2 > 0) ? (1LL << p.oblock_log2) : p.ostride;
                if constexpr (MODE != PFB_OUT_FM) {
                    pfb_demod_from_y<N, NSLOT, CPT, MODE>(p, ring, sl, m0, t0, full, rowstride);
                } else {
                float* dst0 = p.out_fm + pfb_out_index(p, m0, t0);
                float pw[9][CPT];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const float* src = ring + sl * N + m0;
                    if constexpr (CPT == 4) {
                        const float4 t = *reinterpret_cast<const float4*>(src);
                        pw[j][0] = t.x; pw[j][1] = t.y; pw[j][2] = t.z; pw[j][3] = t.w;
                    } else if constexpr (CPT == 2) {
                        const float2 t = *reinterpret_cast<const float2*>(src);
                        pw[j][0] = t.x; pw[j][1] = t.y;
                    } else {
                        pw[j][0] = *src;
                    }
                    sl = (sl + 1 == NSLOT) ? 0 : sl + 1;
                }
                float o[CPT][8];
                if constexpr (CPT >= 2) {
#pragma unroll
                    for (int q = 0; q < CPT; q += 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float2 d = p2sub(make_float2(pw[j + 1][q], pw[j + 1][q + 1]), make_float2(pw[j][q], pw[j][q + 1]));
                            const float2 kk = p2add(p2fmas(d, 0.15915494309189535f, make_float2(12582912.0f, 12582912.0f)),
                                                    make_float2(-12582912.0f, -12582912.0f));
                            d = p2fmas(kk, -6.283185307179586f, d);
                            d = p2muls(d, p.gain);
                            o[q][j] = (d.x != d.x) ? 0.0f : d.x;
                            o[q + 1][j] = (d.y != d.y) ? 0.0f : d.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float d = pw[j + 1][0] - pw[j][0];
                        const float kk = (d * 0.15915494309189535f + 12582912.0f) - 12582912.0f;
                        d = fmaf(kk, -6.283185307179586f, d);
                        d *= p.gain;
                        o[0][j] = (d != d) ? 0.0f : d;
                    }
                }
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
                    float* dst = dst0 + q * rowstride;
                    if (full) {
                        st_global_v8_hint(dst, o[q], pol_stream);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (t0 + j < p.T) dst[j] = o[q][j];
                    }
                }
                }
            }
            mbar_arrive(ring_free);
            base_slot = (base_slot == 0) ? NSLOT - 1 : base_slot - 1;
        }
    }
}

}  // namespace rcb
