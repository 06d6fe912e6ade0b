// K1 (headline variant)  pfb_fm_tma : FM-only polyphase channelizer, 1 tap per arm (P = 1), N = R*R.
//
// Same arithmetic and launch geometry as pfb_fm_kernel (pfb_fm.cuh) but with the input staged by TMA:
//   * each warp owns an 8 KB shared-memory work buffer.  One elected lane issues cp.async.bulk (1-D TMA,
//     SASS UBLKCP) copies of the warp's next F frames (8 KB of contiguous wideband samples) into it and the
//     warp waits on its own mbarrier - no registers are tied up by in-flight loads (the register-prefetch
//     variant spilled half of its 64 prefetch registers, profiles/r01_pfb_fm_v2_*), and the copy for frame
//     n+1 is issued as soon as the second FFT pass of frame n has been read out of the buffer, i.e. a full
//     atan2 + demod phase ahead of its use;
//   * the same buffer is then reused as the 32x32 transpose scratch between the two in-register radix-R
//     passes, dense (no padding) with a 16-byte-chunk XOR swizzle so STS.128 / LDS.64 stay conflict free;
//   * the ring shared by the CTA holds angle(Y) (4 B) instead of Y: atan2 runs once per sample in the FFT
//     warp, the demod phase is subtract + wrap (magic-number rint) + gain, 32 B sector stores.
//   * CTA-level ordering uses an mbarrier (arrive after the demod reads, wait just before the next ring
//     write) so the next frame's FIR + both FFT passes + atan2 overlap the other warps' demod tail.
// Shared memory: 8 x 8 KB work + 9 x 4 KB phase ring + 8 KB twiddles + 4 KB taps = 112.1 KB -> 2 CTAs/SM.
#pragma once
#include "pfb_fm.cuh"
#include "fft_packed.cuh"

namespace rcb {

// the ring holds one float per (frame, bin) - angle(Y) or log power - unless IQ is an output
__host__ __device__ constexpr bool pfb_ring_f32(int mode) { return mode == PFB_OUT_FM || mode == PFB_LOGPOW; }

template <int R, int W = 8, int MODE = PFB_OUT_FM>
struct PfbTmaGeom {
    static constexpr int N = R * R;
    static constexpr int F = 32 / R;
    static constexpr int WARPS = W;
    static constexpr int THREADS = 32 * W;
    static constexpr int FPI = WARPS * F;
    static constexpr int NSLOT = FPI + 1;
    static constexpr int FSW = N + (R == 8 ? 8 : 0);       // frame stride inside a warp's work buffer (complex)
    static constexpr int WORK = F * FSW;                    // complex per warp
    static constexpr size_t work_bytes = (size_t)WARPS * WORK * 8;
    // ring element: angle(Y) (4 B) when only FM is produced, Y itself (8 B) when IQ is an output
    static constexpr size_t ring_bytes = (size_t)NSLOT * N * (pfb_ring_f32(MODE) ? 4 : 8);
    static constexpr size_t tw_bytes = (size_t)N * 8;
    static constexpr size_t taps_bytes = (size_t)N * 4;
    static constexpr size_t smem_bytes = work_bytes + ring_bytes + tw_bytes + taps_bytes + 256;
    // resident CTAs per SM the register allocation is tuned for (3-4 CTAs/SM for N = 64 / 256 were measured: no gain)
    static constexpr int MIN_CTAS = (2 * smem_bytes <= 227 * 1024 && W <= 8) ? 2 : 1;
};

template <int R>
__device__ __forceinline__ int pfb_swz(int l) {
    return (R == 8) ? ((l >> 1) & 3) : (l & (R / 2 - 1));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// demod phase when the ring holds Y (IQ / IQ+FM outputs): thread = CPT adjacent channels x 8 consecutive frames
// starting at ring slot s (the slot of the frame before them); shared by the phase-serial and the warp-specialised
// kernels.
template <int N, int NSLOT, int CPT, int MODE>
__device__ __forceinline__ void pfb_demod_from_y(const PfbParams& p, const float* ring, int s, int m0, long long t0,
                                                 bool full, long long rowstride) {
                const float2* ring2 = reinterpret_cast<const float2*>(ring);
                float2* dq0 = (MODE & PFB_OUT_IQ) ? p.out_iq + pfb_out_index(p, m0, t0) : nullptr;
                float* dst0 = (MODE & PFB_OUT_FM) ? p.out_fm + pfb_out_index(p, m0, t0) : nullptr;
                constexpr int CW = (CPT >= 2) ? 2 : 1;  // channels handled together
#pragma unroll
                for (int q0 = 0; q0 < CPT; q0 += CW) {
                    float2 y[9][CW];
                    int sj = s;
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        const float2* src = ring2 + sj * N + m0 + q0;
                        if constexpr (CW == 2) {
                            const float4 t = *reinterpret_cast<const float4*>(src);
                            y[j][0] = make_float2(t.x, t.y);
                            y[j][1] = make_float2(t.z, t.w);
                        } else {
                            y[j][0] = *src;
                        }
                        sj = (sj + 1 == NSLOT) ? 0 : sj + 1;
                    }
                    if constexpr ((MODE & PFB_OUT_IQ) != 0) {
#pragma unroll
                        for (int c = 0; c < CW; ++c) {
                            float2* dst = dq0 + (q0 + c) * rowstride;
                            if (full) {
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    float o[8];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        o[2 * j] = y[1 + 4 * hh + j][c].x;
                                        o[2 * j + 1] = y[1 + 4 * hh + j][c].y;
                                    }
                                    st_global_v8(reinterpret_cast<float*>(dst + 4 * hh), o);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (t0 + j < p.T) dst[j] = y[j + 1][c];
                            }
                        }
                    }
                    if constexpr ((MODE & PFB_OUT_FM) != 0) {
                        float o[CW][8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if constexpr (CW == 2) {  // p = y[j+1] conj(y[j]) for both channels, packed atan2
                                const float2 a0 = cmul_conj(y[j + 1][0], y[j][0]), a1 = cmul_conj(y[j + 1][1], y[j][1]);
                                const float2 ang = atan2_zero_p2(make_float2(a0.y, a1.y), make_float2(a0.x, a1.x));
                                o[0][j] = p.gain * ang.x;
                                o[1][j] = p.gain * ang.y;
                            } else {
                                const float2 a0 = cmul_conj(y[j + 1][0], y[j][0]);
                                o[0][j] = p.gain * atan2_fast(a0.y, a0.x);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < CW; ++c) {
                            float* dst = dst0 + (q0 + c) * rowstride;
                            if (full) {
                                st_global_v8(dst, o[c]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (t0 + j < p.T) dst[j] = o[c][j];
                            }
                        }
                    }
                }
}

// twiddle table layout expected in p.twiddle for this kernel: dense [R][R] complex, 16-byte chunks swizzled:
//   tw[ll*R + (((m1>>1) ^ swz(ll))<<1 | (m1&1))] = W_N^{+(R-1-ll) m1}
// taps layout: float4 groups as in pfb_fm.cuh (P = 1).
// PK = true: packed f32x2 arithmetic (fft_packed.cuh) for both radix-R passes, atan2 and the demod.
// PT  > 1  : PT taps per arm.  The TMA row copy is replaced by a time-blocked arm FIR: a thread owns one
//             column (arm) for 8 consecutive frames, loads the 8 + PT - 1 samples of that column once
//             (coalesced 256 B per warp and row; the PT - 1 history rows are L1/L2 hits) and its PT taps,
//             and produces the 8 filtered samples with PT packed FFMA2 each (re/im pair x broadcast tap),
//             written to the frame buffers the FFT warps then transform exactly like TMA-landed rows.
// MODE    : PFB_OUT_FM (angle ring, headline), PFB_OUT_IQ or PFB_OUT_IQ | PFB_OUT_FM (the ring holds Y; the demod
//             phase emits 8 consecutive frames of a channel as two 32 B sector stores of complex64 and, for
//             IQ + FM, conj-multiplies neighbouring frames and runs the packed atan2 itself).  PK only.
// OB8     : the device output layout is channel-major inside time blocks of exactly 8 frames (oblock_log2 == 3),
//             i.e. one CTA iteration writes one contiguous 32 KB piece.  The demod threads then take channels
//             lane + 32 q, so a store instruction covers 1 KB of contiguous memory (8 full lines instead of 32
//             scattered sectors: a quarter of the LSU wavefronts of the store path).
template <int R, int W = 8, bool PK = false, int PT = 1, int MODE = PFB_OUT_FM, bool OB8 = false>
__device__ __forceinline__ void pfb_fm_tma_body(const PfbParams& p) {
    static_assert(!OB8 || (R == 32 && W == 8 && MODE == PFB_OUT_FM), "OB8 is an FM-only 1024-channel layout variant");
    using G = PfbTmaGeom<R, W, MODE>;
    static_assert(PK || MODE == PFB_OUT_FM, "IQ outputs are implemented on the packed path only");
    // MODE = PFB_LOGPOW: the K3 row pass (fft_vector.py: fft_vcc + mag^2 + nlog10_ff) on the same pipeline - forward
    // transform of contiguous rows in natural order, no taps, log10(|X|^2) + 1 into the ring, the "demod" phase
    // copies 8 consecutive rows of a bin as one sector into vals[f][fftshift(k2) * L1 + k1] (ostride = L1).
    static_assert(MODE != PFB_LOGPOW || (PK && PT == 1 && !OB8), "log-power mode: packed, no taps");
    constexpr int SG = (MODE == PFB_LOGPOW) ? -1 : +1;
    constexpr bool REVI = (MODE != PFB_LOGPOW);  // PFB arms are commutated in reverse order
    constexpr int THREADS = G::THREADS;
    constexpr int N = G::N, F = G::F, FPI = G::FPI, NSLOT = G::NSLOT, FSW = G::FSW;
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    float2* work_all = reinterpret_cast<float2*>(smem_raw128);
    float* ring = reinterpret_cast<float*>(smem_raw128 + G::work_bytes);
    float2* tws = reinterpret_cast<float2*>(smem_raw128 + G::work_bytes + G::ring_bytes);
    float* taps_s = reinterpret_cast<float*>(smem_raw128 + G::work_bytes + G::ring_bytes + G::tw_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw128 + G::work_bytes + G::ring_bytes + G::tw_bytes + G::taps_bytes);
    uint64_t* cta_bar = bars + W;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane / R, ll = lane % R;
    float2* work = work_all + warp * G::WORK;   // this warp's buffer
    float2* wf = work + fr * FSW;               // this lane's frame inside it
    uint64_t* row_bar = bars + warp;

    for (int i = tid; i < N; i += THREADS) {
        tws[i] = p.twiddle[i];
        if constexpr (PT == 1 && MODE != PFB_LOGPOW) taps_s[i] = p.taps[i];
    }
    if (tid < W) mbar_init(bars + tid, 1);
    if (tid == W) mbar_init(cta_bar, THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const float4* tap4 = reinterpret_cast<const float4*>(taps_s);
    const uint64_t pol_stream = l2_policy_evict_first();  // (evict_last on the row loads and L2 prefetches: no gain)

    // ---- work distribution: a static contiguous run (7/8 of the even share) per CTA, then the tail of the
    // ---- launch is handed out dynamically in chunks of kTailChunk iterations (atomic counter).  SMs differ
    // ---- by up to +-10 % in sustained speed with TMA staging (profiles/r01_pfb_fm_v3_*): a purely static split
    // ---- leaves the fast SMs idle for 11 % of the launch.  Every range starts with one warm-up iteration
    // ---- that recomputes the frame before it (no Y / phase state is carried between ranges or launches).
    // log-power rows are independent: no warm-up iteration, single-iteration dynamic chunks
    constexpr int WARM = (MODE == PFB_LOGPOW) ? 0 : 1;
    // (chunk lengths 2..8 and static shares 12/16..15/16 were measured: all within 1 %)
    constexpr int kTailChunk = (MODE == PFB_LOGPOW) ? 2 : 4;
    const int NI = (p.T + FPI - 1) / FPI;
    const int stat = (int)(((long long)(NI / (int)gridDim.x) * 7) / 8);
    const int tail0 = stat * (int)gridDim.x;
    __shared__ int s_next;
    int cur0, cur1;
    if (stat > 0) {
        cur0 = blockIdx.x * stat;
        cur1 = cur0 + stat;
    } else {
        if (tid == 0) s_next = atomicAdd(p.work_counter, 1);
        __syncthreads();
        cur0 = tail0 + s_next * kTailChunk;
        cur1 = min(cur0 + kTailChunk, NI);
        __syncthreads();
        if (cur0 >= NI) return;
    }
    int nxt0 = NI, nxt1 = NI;  // following range (nxt0 >= NI: none)

    long long frame0 = (long long)(cur0 - WARM) * FPI + warp * F;  // first of this warp's F frames
    auto issue_rows = [&](long long f0) {
        if (PT == 1 && lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(row_bar, (uint32_t)(F * N * 8));
#pragma unroll
            for (int q = 0; q < F; ++q)
                tma_bulk_g2s(work + q * FSW, pfb_row_ptr<R>(p, f0 + q), (uint32_t)(N * 8), row_bar);
        }
    };
    issue_rows(frame0);

    int base_slot = 0;
    uint32_t row_par = 0, cta_par = 0;
    bool first_ever = true;
#pragma unroll 1
    for (int it = cur0 - WARM;; ++it) {
        const bool range_first = (it == cur0 - WARM);
        if (range_first && tid == 0) {  // reserve the range that follows the current one
            const int c = atomicAdd(p.work_counter, 1);
            s_next = tail0 + c * kTailChunk;
        }
        if constexpr (PT == 1) {
            // ---- wait for the TMA copy of this warp's frames ----
            mbar_wait(row_bar, row_par);
            row_par ^= 1u;
        } else {
            // ---- time-blocked arm FIR of this iteration's FPI frames into the warps' frame buffers ----
            // task = (column c, group of TBK consecutive frames); 16-frame groups where the iteration has them
            constexpr int TBK = (FPI % 16 == 0 && N * (FPI / 16) >= THREADS) ? 16 : 8;
            constexpr int NX = TBK + PT - 1;
            constexpr int TASKS = N * (FPI / TBK) / THREADS;
#pragma unroll 1
            for (int q = 0; q < TASKS; ++q) {
                const int task = q * THREADS + tid;
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
                for (int k = 0; k < PT; ++k) hk[k] = __ldg(p.taps_kc + k * N + c);
#pragma unroll
                for (int t = 0; t < TBK; ++t) {
                    float2 acc = p2muls(xs[t + PT - 1], hk[0]);
#pragma unroll
                    for (int k = 1; k < PT; ++k) acc = p2fmas(xs[t + PT - 1 - k], hk[k], acc);
                    const int fi = TBK * g + t;  // frame within the iteration -> owning warp's buffer
                    work_all[(fi / F) * G::WORK + (fi % F) * FSW + c] = acc;
                }
            }
            __syncthreads();
        }
        float ph[R];
        float yre[pfb_ring_f32(MODE) ? 1 : R], yim[pfb_ring_f32(MODE) ? 1 : R];
        const bool range_last = (it + 1 == cur1);
        // issue the TMA copy of the frames this warp transforms next (its buffer has been fully consumed)
        auto issue_next = [&]() {
            if (!range_last) {
                frame0 += FPI;
                issue_rows(frame0);
            } else if (nxt0 < NI) {  // seamless hand-over: prefetch the warm-up frames of the next range
                frame0 = (long long)(nxt0 - WARM) * FPI + warp * F;
                issue_rows(frame0);
            }
        };
        if constexpr (PK) {
            // ---- packed path: scalar DIF stage (taps folded in) + R/2-point transform on f32x2 pairs ----
            float2 pr[R / 2], pi[R / 2];
            {
                float hreg[R];
#pragma unroll
                for (int jq = 0; PT == 1 && MODE != PFB_LOGPOW && jq < R / 4; ++jq) {
                    const float4 h = tap4[jq * R + ll];
                    hreg[4 * jq + 0] = h.x; hreg[4 * jq + 1] = h.y; hreg[4 * jq + 2] = h.z; hreg[4 * jq + 3] = h.w;
                }
                // DFT input j is sample row jj = R-1-j of this lane's column
                auto get = [&](auto j) { return wf[(REVI ? R - 1 - decltype(j)::value : decltype(j)::value) * R + ll]; };
                auto tap = [&](auto j) { return hreg[R - 1 - decltype(j)::value]; };
                fft_packed<R, SG, PT == 1 && MODE != PFB_LOGPOW>(pr, pi, get, tap);
            }
            __syncwarp();  // every lane has read its samples: the buffer becomes the transpose scratch
            {
                const int sw = pfb_swz<R>(ll);
                const float4* twp = reinterpret_cast<const float4*>(tws + ll * R);
                float4* bp = reinterpret_cast<float4*>(wf + ll * R);
#pragma unroll
                for (int c = 0; c < R / 2; ++c) {  // pair c = (X[2c], X[2c+1]) -> one 16-byte chunk
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
                for (int l2 = 0; l2 < R; ++l2) u[REVI ? R - 1 - l2 : l2] = wf[l2 * R + (((ch ^ pfb_swz<R>(l2)) << 1) | wi)];
                __syncwarp();  // scratch fully consumed: start the copy of the next frames right away
                issue_next();
                auto get = [&](auto j) { return u[decltype(j)::value]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, SG, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = Y[ll + R*(2q)], Y[ll + R*(2q+1)]
            }
            if constexpr (MODE == PFB_OUT_FM) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const float2 a = atan2_nan_p2(pi[q], pr[q]);
                    ph[2 * q] = a.x;
                    ph[2 * q + 1] = a.y;
                }
            } else if constexpr (MODE == PFB_LOGPOW) {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {  // nlog10_ff(1, L, 1): log10(max(|X|^2, 1e-18)) + 1
                    const float2 pw2 = p2fma(pr[q], pr[q], p2mul(pi[q], pi[q]));
                    ph[2 * q] = fmaf(log2f(fmaxf(pw2.x, 1e-18f)), 0.30102999566398120f, 1.0f);
                    ph[2 * q + 1] = fmaf(log2f(fmaxf(pw2.y, 1e-18f)), 0.30102999566398120f, 1.0f);
                }
            } else {  // keep Y: written to the ring below (needs pr/pi in scope -> stash in yre/yim)
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    yre[2 * q] = pr[q].x; yre[2 * q + 1] = pr[q].y;
                    yim[2 * q] = pi[q].x; yim[2 * q + 1] = pi[q].y;
                }
            }
        } else {
        float2 v[R];
#pragma unroll
        for (int jq = 0; jq < R / 4; ++jq) {
            const float4 h = tap4[jq * R + ll];
            const float2 x0 = wf[(4 * jq + 0) * R + ll], x1 = wf[(4 * jq + 1) * R + ll];
            const float2 x2 = wf[(4 * jq + 2) * R + ll], x3 = wf[(4 * jq + 3) * R + ll];
            v[R - 1 - (4 * jq + 0)] = make_float2(h.x * x0.x, h.x * x0.y);
            v[R - 1 - (4 * jq + 1)] = make_float2(h.y * x1.x, h.y * x1.y);
            v[R - 1 - (4 * jq + 2)] = make_float2(h.z * x2.x, h.z * x2.y);
            v[R - 1 - (4 * jq + 3)] = make_float2(h.w * x3.x, h.w * x3.y);
        }
        fft_inreg<R, +1>(v);
        __syncwarp();  // every lane has read its samples: the buffer becomes the transpose scratch
        // ---- twiddle + swizzled transpose + second pass ----
        {
            const int sw = pfb_swz<R>(ll);
            const float4* twp = reinterpret_cast<const float4*>(tws + ll * R);
            float4* bp = reinterpret_cast<float4*>(wf + ll * R);
#pragma unroll
            for (int c = 0; c < R / 2; ++c) {
                const float4 t = twp[c ^ sw];
                const int m1 = 2 * c;
                const float2 b0 = make_float2(fmaf(v[m1].x, t.x, -v[m1].y * t.y), fmaf(v[m1].x, t.y, v[m1].y * t.x));
                const float2 b1 = make_float2(fmaf(v[m1 + 1].x, t.z, -v[m1 + 1].y * t.w),
                                              fmaf(v[m1 + 1].x, t.w, v[m1 + 1].y * t.z));
                bp[c ^ sw] = make_float4(b0.x, b0.y, b1.x, b1.y);
            }
        }
        __syncwarp();
        {
            const int ch = ll >> 1, wi = ll & 1;
#pragma unroll
            for (int l2 = 0; l2 < R; ++l2) v[R - 1 - l2] = wf[l2 * R + (((ch ^ pfb_swz<R>(l2)) << 1) | wi)];
        }
        __syncwarp();  // scratch fully consumed: start the copy of the next frames right away
        issue_next();
        fft_inreg<R, +1>(v);  // v[m2] = Y[ll + R*m2]
#pragma unroll
        for (int m2 = 0; m2 < R; ++m2) ph[m2] = atan2_nan(v[m2].y, v[m2].x);
        }

        int slot = base_slot + warp * F + fr + 1;
        slot = (slot >= NSLOT) ? slot - NSLOT : slot;
        if (!first_ever) {  // previous demod phase finished reading the ring?
            mbar_wait(cta_bar, cta_par);
            cta_par ^= 1u;
        }
        first_ever = false;
        if constexpr (pfb_ring_f32(MODE)) {
            float* fb = ring + slot * N;
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) fb[m2 * R + ll] = ph[m2];
        } else {
            float2* fb = reinterpret_cast<float2*>(ring) + slot * N;
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) fb[m2 * R + ll] = make_float2(yre[m2], yim[m2]);
        }
        __syncthreads();
        if (range_first) {
            nxt0 = s_next;
            nxt1 = min(nxt0 + kTailChunk, NI);
            if (range_last && nxt0 < NI) {  // single-iteration range: the follow-up range was not known in time
                frame0 = (long long)(nxt0 - WARM) * FPI + warp * F;
                issue_rows(frame0);
            }
        }

        // ---- demod: 8 consecutive frames of CPT channels per thread ----
        if (it >= cur0) {
            constexpr int CPT = N * (FPI / 8) / THREADS;  // 4, 2, 1 for R = 32, 16, 8
            // thread -> (8-frame group g, CPT adjacent channels).  With more than one group per iteration the
            // groups of one channel set sit in the two halves of the same warp, so one store instruction
            // covers 64 contiguous bytes of each channel row (half the L2 write requests of 32 B sectors)
            constexpr int GRP = FPI / 8;
            constexpr bool kPairLanes = (GRP == 2 && CPT >= 2);
            constexpr int MSTR = OB8 ? 32 : 1;  // channel stride between the CPT channels of a thread
            const int g = kPairLanes ? (lane >> 4) : tid / (N / CPT);
            const int m0 = OB8 ? (warp * (32 * CPT) + lane)
                               : (kPairLanes ? (warp * 16 + (lane & 15)) * CPT : (tid % (N / CPT)) * CPT);
            const long long t0 = (long long)it * FPI + 8 * g;
            int s = base_slot + 8 * g;
            s = (s >= NSLOT) ? s - NSLOT : s;
            const bool full = (t0 + 8 <= p.T);
            // (m0, t0) -> element index; consecutive channels are `rowstride` apart in both layouts
            float* dst0 = (MODE & PFB_OUT_FM) ? p.out_fm + pfb_out_index(p, m0, t0) : nullptr;
            if constexpr (MODE == PFB_LOGPOW) {  // t = f * L1 + k1 (ostride = L1), bin m = k2 -> vals[f][fftshift(k2)][k1]
                const long long fidx = t0 / p.ostride;
                dst0 = p.out_fm + fidx * (long long)N * p.ostride + (long long)((m0 + N / 2) % N) * p.ostride + (t0 - fidx * p.ostride);
            }
            const long long rowstride = (p.oblock_log2 > 0) ? (1LL << p.oblock_log2) : p.ostride;
            if constexpr (pfb_ring_f32(MODE)) {
            float pw[9][CPT];  // all ring loads first (9 vector LDS), then CPT*8 independent wrap chains
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float* src = ring + s * N + m0;
                if constexpr (OB8) {
#pragma unroll
                    for (int q = 0; q < CPT; ++q) pw[j][q] = src[32 * q];
                } else if constexpr (CPT == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(src);
                    pw[j][0] = t.x; pw[j][1] = t.y; pw[j][2] = t.z; pw[j][3] = t.w;
                } else if constexpr (CPT == 2) {
                    const float2 t = *reinterpret_cast<const float2*>(src);
                    pw[j][0] = t.x; pw[j][1] = t.y;
                } else {
                    pw[j][0] = *src;
                }
                s = (s + 1 == NSLOT) ? 0 : s + 1;
            }
            float o[CPT][8];
            if constexpr (MODE == PFB_LOGPOW) {
#pragma unroll
                for (int q = 0; q < CPT; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[q][j] = pw[j + 1][q];
            } else if constexpr (PK && CPT >= 2) {  // packed over channel pairs; the NaN select unpacks for free
#pragma unroll
                for (int q = 0; q < CPT; q += 2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float2 d = p2sub(make_float2(pw[j + 1][q], pw[j + 1][q + 1]), make_float2(pw[j][q], pw[j][q + 1]));
                        const float2 k = p2add(p2fmas(d, 0.15915494309189535f, make_float2(12582912.0f, 12582912.0f)),
                                               make_float2(-12582912.0f, -12582912.0f));
                        d = p2fmas(k, -6.283185307179586f, d);
                        d = p2muls(d, p.gain);
                        o[q][j] = (d.x != d.x) ? 0.0f : d.x;
                        o[q + 1][j] = (d.y != d.y) ? 0.0f : d.y;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < CPT; ++q) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float d = pw[j + 1][q] - pw[j][q];
                        const float k = (d * 0.15915494309189535f + 12582912.0f) - 12582912.0f;
                        d = fmaf(k, -6.283185307179586f, d);
                        d *= p.gain;
                        o[q][j] = (d != d) ? 0.0f : d;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CPT; ++q) {
                float* dst = dst0 + q * MSTR * rowstride;
                if (full) {
                    if (PT > 1) st_global_v8_hint(dst, o[q], pol_stream);  // keeps the input rows in L2 (+6 %)
                    else st_global_v8(dst, o[q]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (t0 + j < p.T) dst[j] = o[q][j];
                }
            }
                    } else {
                pfb_demod_from_y<N, NSLOT, CPT, MODE>(p, ring, s, m0, t0, full, rowstride);
            }
        }
        mbar_arrive(cta_bar);
        base_slot = (base_slot == 0) ? NSLOT - 1 : base_slot - 1;
        if (range_last) {
            if (nxt0 >= NI) break;
            cur0 = nxt0;
            cur1 = nxt1;
            it = cur0 - 1 - WARM;  // ++it -> warm-up iteration of the new range
        }
    }
}

template <int R, int W = 8, bool PK = false, int PT = 1, int MODE = PFB_OUT_FM, bool OB8 = false>
__global__ void __launch_bounds__(32 * W, (PfbTmaGeom<R, W, MODE>::MIN_CTAS)) pfb_fm_tma_kernel(const PfbParams p) {
    pfb_fm_tma_body<R, W, PK, PT, MODE, OB8>(p);
}

// Several independent streams of the same shape in ONE launch (one stream per SDR source,
// rc_frontend/receiver.py:67-70; BASELINE config 5 = 8 x 256-channel streams per GPU).  The persistent grid walks the
// streams one after the other - every CTA takes its share of stream 0, then of stream 1, ... - so the launch gaps and
// wave tails of per-stream launches disappear (a CTA that finishes its part of stream s starts on s + 1 at once) while
// the L2 working set stays that of ONE stream.  (blockIdx.y = stream, all streams at once, was measured 8 % SLOWER than
// separate launches: the history rows the time-blocked FIR re-reads fall out of L2 with eight streams interleaved.)
template <int R, int W = 8, bool PK = false, int PT = 1, int MODE = PFB_OUT_FM>
__global__ void __launch_bounds__(32 * W, (PfbTmaGeom<R, W, MODE>::MIN_CTAS))
    pfb_fm_tma_multi_kernel(const PfbParams* __restrict__ ps, int nstreams) {
    for (int s = 0; s < nstreams; ++s) {
        const PfbParams p = ps[s];
        pfb_fm_tma_body<R, W, PK, PT, MODE, false>(p);
        __syncthreads();  // the body's shared memory (mbarriers included) is re-initialised by the next stream
    }
}

// history update of every stream of a multi-stream launch: new_hist = last cap samples of (old_hist ++ x)
struct PfbHistJob {
    const float2* old_hist;
    const float2* x;
    float2* new_hist;
};
__global__ void pfb_hist_multi_kernel(const PfbHistJob* __restrict__ jobs, long long n, long long cap) {
    const PfbHistJob j = jobs[blockIdx.y];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const long long src = i + n - cap;
    j.new_hist[i] = (src >= 0) ? j.x[src] : ((src >= -cap) ? j.old_hist[cap + src] : make_float2(0.f, 0.f));
}

}  // namespace rcb
