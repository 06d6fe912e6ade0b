// K1 (FM-only, 1..16 taps per arm)  pfb_cl : polyphase channelizer + fused FM demod with a REGISTER-RESIDENT arm
// FIR window, a TMA-fed raw-row ring, dense output tiles emitted by TMA tensor stores and - for N = 1024 - the
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
//     and writes the filtered sample(s) to a ring of three 8-frame buffer sets: every input byte crosses L2 -> SM once;
//   * the raw-row ring (8 rows) is filled by cp.async.bulk.tensor (SASS UTMALDG) issued by one lane six rows ahead;
//     rows before the block (streaming history) come from the hist tensor, rows outside the stream are zero-filled
//     by the TMA unit (out-of-bounds coordinates);
//   * 8 FFT warps in two groups of four that take alternate 8-frame sets (so the groups run half an iteration apart
//     and hide each other's latencies), two frames per warp (half-warp per frame, lane = ll mod 16): first packed
//     radix-R pass over the columns ll + R jj, twiddle, transpose IN PLACE in the frame buffer - for N = 1024 the half
//     of the first-pass outputs that belongs to the other CTA's bins (m1 = m mod 32 in its half) is written straight
//     into the peer's frame buffer with st.shared::cluster and handed over with remote mbarrier arrives - then the
//     second pass over ll for the CTA's own 16 values of m1, packed atan2, and the angles go into one of two
//     [m2][m1][16 frames] tiles (64 B per channel, CU_TENSOR_MAP_SWIZZLE_64B layout);
//   * demod + store run on the FIR warps (they have the issue slots to spare): a thread owns NC/256 channel rows of
//     the tile, replaces the 16 angles of a row by gain * wrap(difference) IN PLACE (previous angle carried in a
//     register - read set == write set, so no barrier), and each warp emits its 4 KB slice of the tile with one
//     cp.async.bulk.tensor store (SASS UTMASTG) - no per-thread sector stores, no LSU store wavefronts.
// There is no CTA-wide barrier in the steady state: all hand-overs are mbarriers (full / empty per buffer).
// One iteration = 16 frames; a run starts with one warm-up iteration that refills the register windows and the
// previous angles (no state is carried between CTAs or launches, so any split of a stream is bit exact).
// Shared memory (N = 1024): 2 x 32 KB tiles + 32 KB raw ring + 3 x 32 KB frame sets + 4 KB twiddles = 196 KB.
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
    static constexpr int CPF = NC / 256;  // columns per FIR thread = channel rows per demod thread
    static constexpr int RS = 8;          // raw-row ring: 4 slots of a row PAIR (one TMA box = 2 rows of the CTA's columns)
    static constexpr int M2W = R / 8;     // m2 rows of the tile owned (demodulated and stored) by one FIR warp = jj
                                          //   rows of a frame owned by one FIR warp
    static constexpr size_t row_bytes = (size_t)NC * 8;
    static constexpr size_t tile_bytes = (size_t)FPI * NC * 4;  // [R m2][16 m1][16 t] floats
    static constexpr size_t raw_bytes = RS * row_bytes;
    static constexpr size_t hset_bytes = 8 * row_bytes;         // one 8-frame buffer set
    static constexpr size_t tw_bytes = (size_t)(R / 2) * 16 * 16;
    static constexpr size_t off_raw = 2 * tile_bytes;
    static constexpr size_t off_sets = off_raw + raw_bytes;
    static constexpr size_t off_tw = off_sets + 3 * hset_bytes;
    static constexpr size_t off_bar = off_tw + tw_bytes;
    static constexpr size_t smem_bytes = off_bar + 1024 + 1024;  // + barriers + alignment slack
};

enum {  // mbarrier slots (8 B each) behind off_bar
    CLB_SET_FULL = 0,     // [3]  FIR -> FFT group: 8 filtered frames
    CLB_SET_EMPTY = 3,    // [3]  FFT group (4 warps) -> FIR
    CLB_T_FREE = 6,       // [8]  peer FFT warp w has read its frames: its buffer may receive my first-pass outputs
    CLB_T_FULL = 14,      // [8]  my half of the transpose buffer of warp w is complete (local arrive + peer's st.async bytes)
    CLB_RING_FULL = 22,   // [2]  FFT warps (8) -> demod: the 16 angle columns of an iteration are in the tile
    CLB_RING_FREE = 24,   // [2]  demod warps (8) -> FFT: the tile has been read by the TMA stores
    CLB_RAW_FULL = 26,    // [4]  TMA -> FIR warps: a row pair has landed
    CLB_COUNT = 30        // (+ 4 int consumer counters of the row-pair slots behind the barriers)
};

template <int R, int PT>
__global__ void __launch_bounds__(512, 1)
    pfb_cl_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hist,
                  const __grid_constant__ CUtensorMap tm_out, const PfbClParams p) {
    using G = PfbClGeom<R>;
    constexpr int CS = G::CS, N = G::N, CPF = G::CPF, RS = G::RS, M2W = G::M2W;
    static_assert(CLB_COUNT * 8 <= 1024, "barrier area");
    static_assert(R == 16 || R == 32, "N = 256 (one CTA) or N = 1024 (cluster of two)");
    static_assert(PT >= 1 && PT <= 16 && (PT & (PT - 1)) == 0, "taps per arm rounded up to a power of two");
    extern __shared__ unsigned char smem_cl_raw[];
    const uint32_t base = (smem_addr_u32(smem_cl_raw) + 1023u) & ~1023u;  // swizzled TMA tiles want their pattern aligned
    const uint32_t a_tile = base, a_raw = base + (uint32_t)G::off_raw, a_sets = base + (uint32_t)G::off_sets;
    const uint32_t a_bar = base + (uint32_t)G::off_bar;
    unsigned char* gbase = smem_cl_raw + (base - smem_addr_u32(smem_cl_raw));
    auto bar = [&](int i) { return a_bar + 8u * (uint32_t)i; };
    int* raw_cnt = reinterpret_cast<int*>(gbase + G::off_bar + 8 * CLB_COUNT);  // consumers (FIR warps) done with a slot

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
        for (int i = 0; i < RS / 2; ++i) {
            mbar_init_a(bar(CLB_RAW_FULL + i), 1);
            raw_cnt[i] = 0;
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init_a(bar(CLB_SET_FULL + i), 8);
            mbar_init_a(bar(CLB_SET_EMPTY + i), 4);
        }
        for (int i = 0; i < 8; ++i) {
            mbar_init_a(bar(CLB_T_FREE + i), 1);
            mbar_init_a(bar(CLB_T_FULL + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init_a(bar(CLB_RING_FULL + i), 8);
            mbar_init_a(bar(CLB_RING_FREE + i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&tm_x);
        prefetch_tmap(&tm_hist);
        prefetch_tmap(&tm_out);
    }
    if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();

    const long long fbase = (long long)(it0 - 1) * 16;  // stream frame of row 0 of this run
    const int nrows = nit * 16;

    if (warp < 8) {
        // =============================== FIR + demod warps ===============================
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
        // demod ownership: R = 32: tile rows m2 = 2 (ft >> 4) + {0, 1};  R = 16: m2 = ft >> 4;  m1 = ft & 15
        const int dm1 = ft & 15, dmh = ft >> 4;
        float prev[CPF];
#pragma unroll
        for (int q = 0; q < CPF; ++q) prev[q] = 0.f;

        // shared ring of 4 row pairs.  Each FIR warp counts itself off a slot after reading it (shared-memory atomic);
        // the LAST warp to do so refills the slot with the pair 8 rows ahead - one TMA per two rows and CTA, issued by
        // whichever warp is slowest, so no warp ever waits for another one before it can go on.
        auto issue_pair = [&](int g, int slot) {  // rows g, g + 1 of the run (g even)
            const long long f = fbase + g;
            const uint32_t dst = a_raw + (uint32_t)slot * (uint32_t)(2 * G::row_bytes);
            mbar_expect_tx_a(bar(CLB_RAW_FULL + slot), (uint32_t)(2 * G::row_bytes));
            if (f >= 0)
                tma_load_3d(dst, &tm_x, 32 * (int)rank, 0, (int)f, bar(CLB_RAW_FULL + slot));  // rows >= T: zero filled
            else  // streaming history; rows before it are out of bounds of the hist tensor = zero filled
                tma_load_3d(dst, &tm_hist, 32 * (int)rank, 0, (int)(f + p.P), bar(CLB_RAW_FULL + slot));
        };
        if (tid == 0) {
            for (int g = 0; g < RS && g < nrows; g += 2) issue_pair(g, g >> 1);
        }
        // demod of iteration j (runs one iteration behind the FIR): in place in tile j & 1, then this warp's slice of
        // the tile goes out with one TMA store
        auto demod = [&](int j) {
            const int it = it0 - 1 + j;
            const int buf = j & 1;
            if (lane == 0 && j >= 1) {  // the slice stored an iteration ago has been read: its tile may be rewritten
                tma_store_wait_read();
                mbar_arrive_a(bar(CLB_RING_FREE + ((j - 1) & 1)));
            }
            mbar_wait_a(bar(CLB_RING_FULL + buf), (uint32_t)((j >> 1) & 1));
            const uint32_t tile = a_tile + (uint32_t)buf * (uint32_t)G::tile_bytes;
            const uint32_t sw = (uint32_t)((dm1 >> 1) & 3);
            const long long t0 = (long long)it * 16;
            const bool live = (it >= it0);
            const bool full = live && (t0 + 16 <= p.T);
#pragma unroll
            for (int q = 0; q < CPF; ++q) {
                const int m2 = (CPF == 2) ? 2 * dmh + q : dmh;
                const uint32_t rowa = tile + (uint32_t)(m2 * 16 + dm1) * 64u;
                float a[16];
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(a[4 * j4]), "=f"(a[4 * j4 + 1]), "=f"(a[4 * j4 + 2]), "=f"(a[4 * j4 + 3])
                                 : "r"(rowa + ((uint32_t)(j4 ^ sw) << 4)));
                float o[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    float d = a[t] - (t == 0 ? prev[q] : a[t - 1]);
                    const float kk = (d * 0.15915494309189535f + 12582912.0f) - 12582912.0f;
                    d = fmaf(kk, -6.283185307179586f, d);
                    d *= p.gain;
                    o[t] = (d != d) ? 0.0f : d;
                }
                prev[q] = a[15];
                if (full) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowa + ((uint32_t)(j4 ^ sw) << 4)), "f"(o[4 * j4]),
                                     "f"(o[4 * j4 + 1]), "f"(o[4 * j4 + 2]), "f"(o[4 * j4 + 3])
                                     : "memory");
                } else if (live) {  // ragged tail of the block: guarded scalar stores straight from the registers
                    const int m = 16 * (int)rank + dm1 + R * m2;
#pragma unroll
                    for (int t = 0; t < 16; ++t)
                        if (t0 + t < p.T) p.out_fm[pfb_cl_out_index(p, m, t0 + t)] = o[t];
                }
            }
            if (full) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const uint32_t src = tile + (uint32_t)(M2W * warp) * 1024u;
                    if (p.out_rank == 3) {
                        tma_store_3d(&tm_out, (int)t0, 16 * (int)rank, M2W * warp, src);
                    } else {
                        const int kb = p.oblock_log2;
                        tma_store_4d(&tm_out, (int)(t0 & ((1LL << kb) - 1)), 16 * (int)rank, M2W * warp, (int)(t0 >> kb), src);
                    }
                    tma_store_commit();
                }
            }
        };
#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
            uint32_t a_set = 0;
            int sb = 0;
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
                if ((t & 7) == 0) {  // next 8-frame set of the ring of three
                    const int h = 2 * k + (t >> 3);
                    sb = h % 3;
                    if (h >= 3) mbar_wait_a(bar(CLB_SET_EMPTY + sb), (uint32_t)((h / 3 - 1) & 1));
                    a_set = a_sets + (uint32_t)sb * (uint32_t)G::hset_bytes;
                }
                const int ps = (t >> 1) & 3;                                 // static: the iteration is unrolled
                mbar_wait_a(bar(CLB_RAW_FULL + ps), (uint32_t)((2 * k + (t >> 3)) & 1));
                const uint32_t src = a_raw + (uint32_t)ps * (uint32_t)(2 * G::row_bytes) + (uint32_t)ft * (8u * CPF);
                float2 x[2][CPF];
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {
                    if constexpr (CPF == 2) {
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                     : "r"(src + (uint32_t)r2 * (uint32_t)G::row_bytes));
                        x[r2][0] = make_float2(v.x, v.y);
                        x[r2][1] = make_float2(v.z, v.w);
                    } else {
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];"
                                     : "=f"(x[r2][0].x), "=f"(x[r2][0].y)
                                     : "r"(src + (uint32_t)r2 * (uint32_t)G::row_bytes));
                    }
                }
                __syncwarp();
                // count this warp off the slot now, look at the answer after the arithmetic below (a shared-memory atomic
                // with a used result is a ~60-cycle stall: profiles/r02_pfb_cl_v4_p16_summary.txt)
                int seen = 0;
                if (lane == 0) seen = atomicAdd(&raw_cnt[ps], 1);
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2) {
                    const int tt = t + r2;
                    float2 y[CPF];
#pragma unroll
                    for (int q = 0; q < CPF; ++q) {
                        win[q][tt % PT] = x[r2][q];
                        if constexpr (PT == 1) {
                            y[q] = p2muls(x[r2][q], hk[q][0]);
                        } else {
                            // two interleaved partial sums (even / odd taps) keep four FFMA2 chains per thread in flight
                            float2 a0 = p2muls(win[q][tt % PT], hk[q][0]);
                            float2 a1 = p2muls(win[q][(tt + PT - 1) % PT], hk[q][1]);
#pragma unroll
                            for (int kk = 2; kk < PT; kk += 2) {
                                a0 = p2fmas(win[q][(tt + PT - kk) % PT], hk[q][kk], a0);
                                a1 = p2fmas(win[q][(tt + PT - kk - 1) % PT], hk[q][kk + 1], a1);
                            }
                            y[q] = p2add(a0, a1);
                        }
                    }
                    const uint32_t dst = a_set + (uint32_t)(tt & 7) * (uint32_t)G::row_bytes + (uint32_t)ft * (8u * CPF);
                    if constexpr (CPF == 2) {
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(y[0].x), "f"(y[0].y), "f"(y[1].x), "f"(y[1].y) : "memory");
                    } else {
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(dst), "f"(y[0].x), "f"(y[0].y) : "memory");
                    }
                }
                if (lane == 0 && seen == 7) {  // every FIR warp has read the pair: this warp refills the slot
                    raw_cnt[ps] = 0;
                    const int gn = k * 16 + t + RS;
                    if (gn < nrows) {
                        fence_async_smem();
                        issue_pair(gn, ps);
                    }
                }
                if ((t & 7) == 6) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(bar(CLB_SET_FULL + sb));
                }
            }
            if (k >= 1) demod(k - 1);
        }
        demod(nit - 1);
        if (lane == 0) tma_store_wait_all();
    } else {
        // =============================== FFT warps ===============================
        const int w = warp - 8;
        const int g = w >> 2, wq = w & 3;  // group g takes the 8-frame sets 2k + g; frames 2 wq + {0, 1} of a set
        const int fr = lane >> 4, l = lane & 15;
        const int ll = 16 * (int)rank + l;  // this lane's column residue (pass 1) and bin residue m1 (pass 2)
        const float4* tws = reinterpret_cast<const float4*>(gbase + G::off_tw);
        const uint32_t peer_tfree = (CS == 2) ? dsmem_map(bar(CLB_T_FREE + w), peer) : 0u;
        const uint32_t peer_tfull = (CS == 2) ? dsmem_map(bar(CLB_T_FULL + w), peer) : 0u;
        const int tcol = 8 * g + 2 * wq + fr;  // this lane's frame inside the iteration = tile column
        const uint32_t tile_lane = (uint32_t)l * 64u + ((uint32_t)((tcol >> 2) ^ ((l >> 1) & 3)) << 4) + (uint32_t)(tcol & 3) * 4u;

#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
            const int h = 2 * k + g;
            const int sb = h % 3;
            const size_t fr_off = G::off_sets + (size_t)sb * G::hset_bytes + (size_t)(2 * wq + fr) * G::row_bytes;
            const uint32_t a_fr = base + (uint32_t)fr_off;
            const float2* wf = reinterpret_cast<const float2*>(gbase + fr_off);
            mbar_wait_a(bar(CLB_SET_FULL + sb), (uint32_t)((h / 3) & 1));
            float2 pr[R / 2], pi[R / 2];
            {
                auto get = [&](auto j) { return wf[(R - 1 - decltype(j)::value) * 16 + l]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);
            }
            __syncwarp();  // both frames of this warp are in registers: the buffer becomes the transpose scratch
            if constexpr (CS == 2) {
                // (relaxed: the only accesses to order are this warp's loads above, already consumed by the transform)
                if (lane == 0) mbar_arrive_remote_relaxed(peer_tfree);  // the peer may write its half into my buffer
                mbar_wait_a(bar(CLB_T_FREE + w), (uint32_t)(k & 1));     // and I may write mine into the peer's
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
                        dsmem_st_async_v4(peer_fr + off, b0.x, b0.y, b1.x, b1.y, peer_tfull);  // + 16 bytes of complete_tx
                    }
                }
            }
            if constexpr (CS == 2) {
                __syncwarp();  // local half written
                // my buffer is complete when the peer's R/2 x 256 bytes of st.async have landed as well
                if (lane == 0) mbar_expect_tx_a(bar(CLB_T_FULL + w), (uint32_t)(32 * (R / 4) * 16));
                mbar_wait_a(bar(CLB_T_FULL + w), (uint32_t)(k & 1));
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
                if (lane == 0) mbar_arrive_a(bar(CLB_SET_EMPTY + sb));
                auto get = [&](auto j) { return u[decltype(j)::value]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = Y[ll + R*(2q)], Y[ll + R*(2q+1)]
            }
            float2 ph[R / 2];
#pragma unroll
            for (int q = 0; q < R / 2; ++q) ph[q] = atan2_nan_p2(pi[q], pr[q]);
            const int buf = k & 1;
            if (k >= 2) mbar_wait_a(bar(CLB_RING_FREE + buf), (uint32_t)(((k >> 1) - 1) & 1));
            {
                // angle of (bin m2, m1 = l) at frame tcol -> tile[m2][l][tcol] (64B-swizzled rows)
                const uint32_t ta = a_tile + (uint32_t)buf * (uint32_t)G::tile_bytes + tile_lane;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ta + (uint32_t)(2 * q) * 1024u), "f"(ph[q].x) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ta + (uint32_t)(2 * q + 1) * 1024u), "f"(ph[q].y) : "memory");
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar(CLB_RING_FULL + buf));
        }
    }
    if constexpr (CS == 2) cluster_sync_all();
}

}  // namespace rcb
