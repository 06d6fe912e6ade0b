// K1 (FM-only, 1..16 taps per arm)  pfb_cl : polyphase channelizer + fused FM demod with a REGISTER-RESIDENT arm
// FIR window, a TMA-fed raw-row ring, a dense output tile emitted by TMA tensor stores and - for N = 1024 - the
// frame split over a 2-CTA thread-block cluster with the transpose between the two radix-32 passes exchanged through
// distributed shared memory.
//
// Replaces pfb.channelizer_ccf(N, taps, 1.0) + one analog.quadrature_demod_cf(gain) per bin
// (rc_frontend/receiver.py:249-261; moto_control_demod.py:105, p25_control_demod.py:120-121, ...).  Arithmetic is that
// of pfb_fm.cuh:
//   u_c[n]  = sum_k h_c[k] x[(n-k) N + c]            (arm FIR of column c, taps h_c[k] = h[(N-1-c) + kN])
//   Y_m[n]  = sum_c u_c[n] e^{+j 2 pi m (N-1-c) / N}  (two packed radix-R passes, fft_packed.cuh)
//   fm_m[n] = gain * wrap(angle Y_m[n] - angle Y_m[n-1])
//
// Why: the round-1 multi-tap kernels re-read the P-1 history rows of every 8-frame FIR task from L2 with __ldg
// (31 rows per 16 frames) and stalled on those loads (long_sb 42 %, profiles/r01_pfb_fm_ws_p16_v1_summary.txt); the
// history of a 1024-channel / 16-tap filter (15 rows x 8 KB) does not fit beside the frame buffers in one SM's shared
// memory.  Here a column's window never leaves the register file:
//   * an SM owns NC = 16 R columns: all 256 for N = 256 (R = 16), the 512 columns {c : (c mod 32) / 16 == rank} for
//     N = 1024 (R = 32) where rank is the CTA's rank in a cluster of two;
//   * 8 FIR warps: a thread owns NC/256 adjacent columns for the whole run of its CTA, keeps their last PT samples
//     (register ring, statically indexed: the 16-frame iteration is fully unrolled) and their PT taps in registers;
//     per frame it reads its NEW sample(s) from the raw-row ring (one LDS.64/128), runs PT packed FFMA2 per column
//     and writes the filtered sample(s) to the frame buffers: every input byte crosses L2 -> SM exactly once;
//   * the raw-row ring (8 rows) is filled by cp.async.bulk.tensor (SASS UTMALDG) issued by one lane six rows ahead;
//     rows before the block (streaming history) come from the hist tensor, rows outside the stream are zero-filled
//     by the TMA unit (out-of-bounds coordinates);
//   * 8 FFT warps, two frames each (half-warp per frame, lane = ll mod 16): first packed radix-R pass over the
//     columns ll + R jj, twiddle, transpose IN PLACE in the frame buffer - for N = 1024 the half of the first-pass
//     outputs that belongs to the other CTA's bins (m1 = m mod 32 in its half) is written straight into the peer's
//     frame buffer with st.shared::cluster and handed over with remote mbarrier arrives - then the second pass over
//     ll for the CTA's own 16 values of m1, packed atan2 and the angle ring (16 slots);
//   * demod: thread = (NC/256 channels) x 16 consecutive frames, previous angle carried in registers; the 16 x NC
//     float results overwrite the ring region as a dense [m2][m1][16] tile (64 B per channel, 64B-swizzled so the
//     STS.128 stay conflict free) and ONE cp.async.bulk.tensor store (SASS UTMASTG) per iteration and CTA writes
//     it to the channel-major output - no per-thread sector stores, no LSU store wavefronts.
// One iteration = 16 frames; a run starts with one warm-up iteration that refills the register windows and the
// previous angles (no state is carried between CTAs or launches, so any split of a stream is bit exact).
// Shared memory (N = 1024): 32 KB raw ring + 2 x 64 KB frame sets + 32 KB ring/tile + 4 KB twiddles = 196 KB.
#pragma once
#include "pfb_fm_tma.cuh"
#include "tma_utils.cuh"

namespace rcb {

struct PfbClParams {
    const float* taps_kc;  // [PT][N]: taps_kc[k*N + c] = h[(N-1-c) + k*N]
    const float4* tw4;     // [CS][R/2][16]: (W_N^{(R-1-ll) 2c}, W_N^{(R-1-ll)(2c+1)}), ll = 16 rank + l
    float* out_fm;
    long long ostride;
    int oblock_log2;
    int out_rank;  // rank of the output tensor map: 3 (plain [N][ostride]) or 4 (time blocks of 2^k frames)
    int T, P, N;
    float gain;
    int* work_counter;  // zeroed before each launch: dynamic tail chunks (one atomic per cluster and chunk)
};

__device__ __forceinline__ long long pfb_cl_out_index(const PfbClParams& p, int m, long long t) {
    if (p.oblock_log2 > 0) {
        const int k = p.oblock_log2;
        return ((((t >> k) * p.N) + m) << k) | (t & ((1LL << k) - 1));
    }
    return (long long)m * p.ostride + t;
}

template <int R>
struct PfbClGeom {
    static constexpr int CS = R / 16;  // CTAs per cluster
    static constexpr int N = R * R;
    static constexpr int NC = N / CS;  // columns and channels per CTA (16 R)
    static constexpr int FPI = 16;     // frames per iteration
    static constexpr int THREADS = 512;
    static constexpr int CPF = NC / 256;  // columns per FIR thread = channels per demod thread
    static constexpr int RS = 8;          // raw-row ring slots
    static constexpr int LAG = 2;         // a slot is refilled LAG rows after its last use (prefetch distance RS - LAG)
    static constexpr size_t row_bytes = (size_t)NC * 8;
    static constexpr size_t tile_bytes = (size_t)FPI * NC * 4;  // angle ring == output tile
    static constexpr size_t raw_bytes = RS * row_bytes;
    static constexpr size_t set_bytes = FPI * row_bytes;
    static constexpr size_t tw_bytes = (size_t)(R / 2) * 16 * 16;
    static constexpr size_t off_raw = tile_bytes;
    static constexpr size_t off_sets = off_raw + raw_bytes;
    static constexpr size_t off_tw = off_sets + 2 * set_bytes;
    static constexpr size_t off_bar = off_tw + tw_bytes;
    static constexpr size_t smem_bytes = off_bar + 512 + 1024;  // + barriers + alignment slack
};

enum {  // mbarrier slots (8 B each) behind off_bar
    CLB_RAW_FULL = 0,    // [8]  TMA -> FIR warps
    CLB_RAW_EMPTY = 8,   // [8]  FIR warps -> TMA issuer
    CLB_SET_FULL = 16,   // [2]  FIR -> FFT
    CLB_SET_EMPTY = 18,  // [2]  FFT -> FIR
    CLB_T_FREE = 20,     // [8]  peer FFT warp w has read its frames: its buffer may receive my first-pass outputs
    CLB_T_FULL = 28,     // [8]  peer FFT warp w has written its half of my transpose buffer
    CLB_RING_FREE = 36,  // [1]  output tile read by the TMA store: the ring may be rewritten
    CLB_COUNT = 37
};

template <int R, int PT>
__global__ void __launch_bounds__(512, 1)
    pfb_cl_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hist,
                  const __grid_constant__ CUtensorMap tm_out, const PfbClParams p) {
    using G = PfbClGeom<R>;
    constexpr int CS = G::CS, N = G::N, NC = G::NC, CPF = G::CPF, RS = G::RS, LAG = G::LAG;
    static_assert(R == 16 || R == 32, "N = 256 (one CTA) or N = 1024 (cluster of two)");
    static_assert(PT >= 1 && PT <= 16 && (PT & (PT - 1)) == 0, "taps per arm rounded up to a power of two");
    extern __shared__ unsigned char smem_cl_raw[];
    const uint32_t base = (smem_addr_u32(smem_cl_raw) + 1023u) & ~1023u;  // swizzled TMA tiles want their pattern aligned
    const uint32_t a_tile = base, a_raw = base + (uint32_t)G::off_raw, a_sets = base + (uint32_t)G::off_sets;
    const uint32_t a_tw = base + (uint32_t)G::off_tw, a_bar = base + (uint32_t)G::off_bar;
    unsigned char* gbase = smem_cl_raw + (base - smem_addr_u32(smem_cl_raw));
    auto bar = [&](int i) { return a_bar + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CS == 2) ? cluster_ctarank() : 0u;
    const uint32_t peer = rank ^ 1u;
    const int ncl = (int)gridDim.x / CS, cl = (int)blockIdx.x / CS;

    // contiguous run of 16-frame iterations per cluster (+ one warm-up iteration in front)
    const int NI = (p.T + 15) / 16;
    const int per = (NI + ncl - 1) / ncl;
    const int it0 = cl * per;
    const int it1 = min(it0 + per, NI);
    if (it0 >= NI) return;  // the whole cluster leaves
    const int nit = it1 - it0 + 1;

    {
        const float4* src = p.tw4 + (size_t)rank * (R / 2) * 16;
        float4* dst = reinterpret_cast<float4*>(gbase + G::off_tw);
        for (int i = tid; i < (R / 2) * 16; i += G::THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int i = 0; i < RS; ++i) {
            mbar_init_a(bar(CLB_RAW_FULL + i), 1);
            mbar_init_a(bar(CLB_RAW_EMPTY + i), 8);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(bar(CLB_SET_FULL + i), 8);
            mbar_init_a(bar(CLB_SET_EMPTY + i), 8);
        }
        for (int i = 0; i < 8; ++i) {
            mbar_init_a(bar(CLB_T_FREE + i), 1);
            mbar_init_a(bar(CLB_T_FULL + i), 32);
        }
        mbar_init_a(bar(CLB_RING_FREE), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&tm_x);
        prefetch_tmap(&tm_hist);
        prefetch_tmap(&tm_out);
    }
    if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();

    const long long fbase = (long long)(it0 - 1) * 16;  // stream frame of row 0 of this run
    const int nrows = nit * 16;

    if (warp < 8) {
        // =============================== FIR warps ===============================
        const int ft = tid;  // 0..255
        float hk[CPF][PT];
        float2 win[CPF][PT];
#pragma unroll
        for (int q = 0; q < CPF; ++q) {
            const int lc = CPF * ft + q, jj = lc >> 4, l = lc & 15;
            const int c = R * jj + 16 * (int)rank + l;
#pragma unroll
            for (int k = 0; k < PT; ++k) {
                hk[q][k] = __ldg(p.taps_kc + (size_t)k * N + c);
                win[q][k] = make_float2(0.f, 0.f);
            }
        }
        const bool issuer = (tid == 0);
        auto issue_row = [&](int g) {  // row g of the run -> slot g % RS
            const int slot = g & (RS - 1);
            const long long f = fbase + g;
            const uint32_t dst = a_raw + (uint32_t)slot * (uint32_t)G::row_bytes;
            mbar_expect_tx_a(bar(CLB_RAW_FULL + slot), (uint32_t)G::row_bytes);
            if (f >= 0 || f < -(long long)p.P)
                tma_load_3d(dst, &tm_x, 32 * (int)rank, 0, (f >= 0) ? (int)f : -1, bar(CLB_RAW_FULL + slot));
            else
                tma_load_3d(dst, &tm_hist, 32 * (int)rank, 0, (int)(f + p.P), bar(CLB_RAW_FULL + slot));
        };
        if (issuer) {
            for (int g = 0; g < RS && g < nrows; ++g) issue_row(g);
        }
#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
            const int s = k & 1;
            if (k >= 2) mbar_wait_a(bar(CLB_SET_EMPTY + s), (uint32_t)(((k >> 1) - 1) & 1));
            const uint32_t a_set = a_sets + (uint32_t)s * (uint32_t)G::set_bytes;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int slot = t & (RS - 1);
                const int use = k * (16 / RS) + t / RS;  // how often this slot has been used before
                mbar_wait_a(bar(CLB_RAW_FULL + slot), (uint32_t)(use & 1));
                const uint32_t src = a_raw + (uint32_t)slot * (uint32_t)G::row_bytes + (uint32_t)ft * (8u * CPF);
                float2 x[CPF];
                if constexpr (CPF == 2) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
                    x[0] = make_float2(v.x, v.y);
                    x[1] = make_float2(v.z, v.w);
                } else {
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[0].x), "=f"(x[0].y) : "r"(src));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(bar(CLB_RAW_EMPTY + slot));
                if (issuer) {  // refill the slot that was last used LAG rows ago
                    const int g = k * 16 + t;
                    const int gd = g - LAG;          // row whose slot is recycled
                    const int gn = gd + RS;          // row that goes into it
                    if (gd >= 0 && gn < nrows) {
                        mbar_wait_a(bar(CLB_RAW_EMPTY + (gd & (RS - 1))), (uint32_t)((gd / RS) & 1));
                        fence_async_smem();
                        issue_row(gn);
                    }
                }
                float2 y[CPF];
#pragma unroll
                for (int q = 0; q < CPF; ++q) {
                    win[q][t % PT] = x[q];
                    if constexpr (PT == 1) {
                        y[q] = p2muls(x[q], hk[q][0]);
                    } else {
                        // two interleaved partial sums (even / odd taps) keep four FFMA2 chains per thread in flight
                        float2 a0 = p2muls(win[q][t % PT], hk[q][0]);
                        float2 a1 = p2muls(win[q][(t + PT - 1) % PT], hk[q][1]);
#pragma unroll
                        for (int kk = 2; kk < PT; kk += 2) {
                            a0 = p2fmas(win[q][(t + PT - kk) % PT], hk[q][kk], a0);
                            a1 = p2fmas(win[q][(t + PT - kk - 1) % PT], hk[q][kk + 1], a1);
                        }
                        y[q] = p2add(a0, a1);
                    }
                }
                const uint32_t dst = a_set + (uint32_t)t * (uint32_t)G::row_bytes + (uint32_t)ft * (8u * CPF);
                if constexpr (CPF == 2) {
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(y[0].x), "f"(y[0].y), "f"(y[1].x), "f"(y[1].y) : "memory");
                } else {
                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(dst), "f"(y[0].x), "f"(y[0].y) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar(CLB_SET_FULL + s));
        }
    } else {
        // =============================== FFT / demod warps ===============================
        const int w = warp - 8, dt = tid - 256;
        const int fr = lane >> 4, l = lane & 15;
        const int ll = 16 * (int)rank + l;  // this lane's column residue (pass 1) and bin residue m1 (pass 2)
        const float4* tws = reinterpret_cast<const float4*>(gbase + G::off_tw);
        float* ring = reinterpret_cast<float*>(gbase);
        const uint32_t peer_tfree = (CS == 2) ? dsmem_map(bar(CLB_T_FREE + w), peer) : 0u;
        const uint32_t peer_tfull = (CS == 2) ? dsmem_map(bar(CLB_T_FULL + w), peer) : 0u;
        // demod ownership: R = 32: bins m2 = 2 kk + {0,1}, kk = dt >> 4;  R = 16: bin m2 = dt >> 4;  m1 = 16 rank + (dt & 15)
        const int dm1 = dt & 15, dmh = dt >> 4;
        float prev[CPF];
#pragma unroll
        for (int q = 0; q < CPF; ++q) prev[q] = 0.f;

#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
            const int it = it0 - 1 + k;
            const int s = k & 1;
            const uint32_t a_fr = a_sets + (uint32_t)s * (uint32_t)G::set_bytes + (uint32_t)(2 * w + fr) * (uint32_t)G::row_bytes;
            const float2* wf = reinterpret_cast<const float2*>(gbase + G::off_sets + (size_t)s * G::set_bytes +
                                                               (size_t)(2 * w + fr) * G::row_bytes);
            if (dt == 0 && k >= 1) {  // the previous iteration's tile has been read by the TMA unit: release the ring
                tma_store_wait_read();
                mbar_arrive_a(bar(CLB_RING_FREE));
            }
            mbar_wait_a(bar(CLB_SET_FULL + s), (uint32_t)((k >> 1) & 1));
            float2 pr[R / 2], pi[R / 2];
            {
                auto get = [&](auto j) { return wf[(R - 1 - decltype(j)::value) * 16 + l]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);
            }
            __syncwarp();  // both frames of this warp are in registers: the buffer becomes the transpose scratch
            if constexpr (CS == 2) {
                if (lane == 0) mbar_arrive_remote(peer_tfree);          // the peer may write its half into my buffer
                mbar_wait_cluster_a(bar(CLB_T_FREE + w), (uint32_t)(k & 1));  // and I may write mine into the peer's
            }
            {
                // row ll of the destination's [R rows][8 chunks of 16 B] buffer, 16-byte chunks XOR-swizzled by row
                const uint32_t rowoff = (uint32_t)ll * 128u;
                const uint32_t peer_fr = (CS == 2) ? dsmem_map(a_fr, peer) : 0u;
#pragma unroll
                for (int c = 0; c < R / 2; ++c) {
                    const float4 t = tws[c * 16 + l];
                    const float2 b0 = make_float2(fmaf(pr[c].x, t.x, -pi[c].x * t.y), fmaf(pr[c].x, t.y, pi[c].x * t.x));
                    const float2 b1 = make_float2(fmaf(pr[c].y, t.z, -pi[c].y * t.w), fmaf(pr[c].y, t.w, pi[c].y * t.z));
                    const uint32_t off = rowoff + (uint32_t)(((c & 7) ^ (ll & 7)) << 4);
                    if (CS == 1 || (uint32_t)(c >> 3) == rank) {
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_fr + off), "f"(b0.x), "f"(b0.y), "f"(b1.x), "f"(b1.y) : "memory");
                    } else {
                        dsmem_st_v4(peer_fr + off, b0.x, b0.y, b1.x, b1.y);
                    }
                }
            }
            if constexpr (CS == 2) {
                mbar_arrive_remote(peer_tfull);  // every lane: its remote stores are released to the peer
                __syncwarp();
                mbar_wait_cluster_a(bar(CLB_T_FULL + w), (uint32_t)(k & 1));
            } else {
                __syncwarp();
            }
            {
                float2 u[R];
                const uint32_t coff = (uint32_t)(l & 1) * 8u;
#pragma unroll
                for (int l2 = 0; l2 < R; ++l2) {
                    const uint32_t a = a_fr + (uint32_t)l2 * 128u + (uint32_t)((((l >> 1) ^ (l2 & 7))) << 4) + coff;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(u[R - 1 - l2].x), "=f"(u[R - 1 - l2].y) : "r"(a));
                }
                __syncwarp();  // frame buffers of this warp fully consumed: hand the set back to the FIR warps
                if (lane == 0) mbar_arrive_a(bar(CLB_SET_EMPTY + s));
                auto get = [&](auto j) { return u[decltype(j)::value]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = Y[ll + R*(2q)], Y[ll + R*(2q+1)]
            }
            float2 ph[R / 2];
#pragma unroll
            for (int q = 0; q < R / 2; ++q) ph[q] = atan2_nan_p2(pi[q], pr[q]);
            if (k >= 1) mbar_wait_a(bar(CLB_RING_FREE), (uint32_t)((k - 1) & 1));
            {
                // ring[slot][kk = m2 / 2][m1 (16)][2]
                float2* rb = reinterpret_cast<float2*>(ring) + (size_t)(2 * w + fr) * (NC / 2) + l;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) rb[q * 16] = ph[q];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");  // FFT warps only: the 16 angle rows of this iteration are in the ring
            float o[CPF][16];
            {
                float a[16][CPF];
                if constexpr (CPF == 2) {
                    const float2* rp = reinterpret_cast<const float2*>(ring) + dmh * 16 + dm1;
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const float2 v = rp[(size_t)t * (NC / 2)];
                        a[t][0] = v.x;
                        a[t][1] = v.y;
                    }
                } else {
                    const float* rp = ring + (dmh >> 1) * 32 + dm1 * 2 + (dmh & 1);
#pragma unroll
                    for (int t = 0; t < 16; ++t) a[t][0] = rp[(size_t)t * NC];
                }
#pragma unroll
                for (int q = 0; q < CPF; ++q) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        float d = a[t][q] - (t == 0 ? prev[q] : a[t - 1][q]);
                        const float kk = (d * 0.15915494309189535f + 12582912.0f) - 12582912.0f;
                        d = fmaf(kk, -6.283185307179586f, d);
                        d *= p.gain;
                        o[q][t] = (d != d) ? 0.0f : d;
                    }
                    prev[q] = a[15][q];
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");  // every ring read is done: the region becomes the output tile
            if (it >= it0) {
                const long long t0 = (long long)it * 16;
                const bool full = (t0 + 16 <= p.T);  // (a 16-frame group never straddles a time block of 2^k >= 16 frames)
                if (full) {
                    // tile [m2][m1 (16)][t (16)] floats, CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk j of a 64-byte row
                    // lands at j ^ (address bits 8:7) = j ^ ((m1 >> 1) & 3)
#pragma unroll
                    for (int q = 0; q < CPF; ++q) {
                        const int m2 = (CPF == 2) ? 2 * dmh + q : dmh;
                        const uint32_t rowa = a_tile + (uint32_t)(m2 * 16 + dm1) * 64u;
                        const uint32_t sw = (uint32_t)((dm1 >> 1) & 3);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowa + ((uint32_t)(j ^ sw) << 4)), "f"(o[q][4 * j]),
                                         "f"(o[q][4 * j + 1]), "f"(o[q][4 * j + 2]), "f"(o[q][4 * j + 3])
                                         : "memory");
                    }
                    fence_async_smem();
                } else {
                    // ragged tail of the block: guarded scalar stores straight from the registers
#pragma unroll
                    for (int q = 0; q < CPF; ++q) {
                        const int m2 = (CPF == 2) ? 2 * dmh + q : dmh;
                        const int m = 16 * (int)rank + dm1 + R * m2;
#pragma unroll
                        for (int t = 0; t < 16; ++t)
                            if (t0 + t < p.T) p.out_fm[pfb_cl_out_index(p, m, t0 + t)] = o[q][t];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (dt == 0 && full) {
                    if (p.out_rank == 3) {
                        tma_store_3d(&tm_out, (int)t0, 16 * (int)rank, 0, a_tile);
                    } else {
                        const int kb = p.oblock_log2;
                        tma_store_4d(&tm_out, (int)(t0 & ((1LL << kb) - 1)), 16 * (int)rank, 0, (int)(t0 >> kb), a_tile);
                    }
                    tma_store_commit();
                }
            }
        }
        if (dt == 0) tma_store_wait_all();
    }
    if constexpr (CS == 2) cluster_sync_all();
}

}  // namespace rcb
