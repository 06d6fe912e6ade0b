// K1 (headline, 1 tap per arm, FM only, N = 1024)  pfb_fm1 : pfb_fm_tma_kernel<32, 8, PK, PT = 1> with the output path
// moved off the LSU: the FM samples of an 8-frame iteration are assembled as a dense [1024 channels][8 frames] tile in
// shared memory (in place of the angle ring, 32B-swizzled) and written by cp.async.bulk.tensor stores (SASS UTMASTG).
//
// Round-1 profile of the predecessor (profiles/r01_pfb_fm_tma_v6_packed_summary.txt): LSU data pipe 69 % busy, of which
// the per-thread 32-byte sector stores (one L1 wavefront per lane and store: 32 scattered lines per warp instruction)
// were 0.19 of 0.58 wavefronts per sample; with the stores suppressed the kernel ran 16 % faster.  Here
//   * input staging, both packed radix-32 passes (taps folded into the first DIF stage), packed atan2 and the
//     per-warp TMA row prefetch are unchanged (pfb_fm_tma.cuh);
//   * the angle ring shrinks to the 8 frames of the iteration ([slot][channel], conflict-free 4-byte stores by the
//     FFT lanes); the previous frame's angle of a channel is carried in a register of the thread that demodulates
//     that channel (thread t owns channels t, t + 256, t + 512, t + 768);
//   * demod: 8 conflict-free LDS.32 per channel, wrap / gain packed over channel pairs, then - after a CTA barrier,
//     because every thread's tile rows overlap ring entries of other threads - the 32-byte rows are written back
//     with two STS.128 per channel (chunk order XORed with address bit 7 = CU_TENSOR_MAP_SWIZZLE_32B, which makes
//     the 32-byte-stride stores of a warp conflict free), and one elected thread issues four [256 x 8] tensor
//     stores once every thread has arrived on an mbarrier; the tile is released for the next ring write when
//     cp.async.bulk.wait_group.read reports the stores have read it.
// Shared memory: 8 x 8 KB frame buffers + 32 KB ring / tile + 8 KB twiddles + 4 KB taps = 108 KB -> 2 CTAs / SM.
#pragma once
#include "pfb_fm_tma.cuh"
#include "tma_utils.cuh"

// static share of a CTA's even part of the work (the rest is handed out in chunks of RCB_FM1_TAIL_CHUNK iterations, each
// preceded by a warm-up iteration) - compile-time knobs; measured on the bench shape (2^28 samples, 296 CTAs): static
// share 5/8: 0.6735 of the roofline, 3/4: 0.7017, 7/8: 0.6987-0.6997, 15/16: 0.6893, 31/32: 0.6632; tail
// chunks of 8 instead of 4 iterations (at 7/8): 0.6975
#ifndef RCB_FM1_STAT_NUM
#define RCB_FM1_STAT_NUM 3
#define RCB_FM1_STAT_DEN 4
#endif
#ifndef RCB_FM1_TAIL_CHUNK
#define RCB_FM1_TAIL_CHUNK 4
#endif

namespace rcb {

struct PfbFm1Geom {
    static constexpr int R = 32, N = 1024, WARPS = 8, THREADS = 256, FPI = 8;
    static constexpr size_t work_bytes = (size_t)WARPS * N * 8;
    static constexpr size_t tile_bytes = (size_t)FPI * N * 4;
    static constexpr size_t tw_bytes = (size_t)N * 8;
    static constexpr size_t taps_bytes = (size_t)N * 4;
    static constexpr size_t off_work = tile_bytes;                 // the tile sits first (1 KB aligned for the swizzle)
    static constexpr size_t off_tw = off_work + work_bytes;
    static constexpr size_t off_taps = off_tw + tw_bytes;
    static constexpr size_t off_bar = off_taps + taps_bytes;
    static constexpr size_t smem_bytes = off_bar + 128 + 1024;
};

// Fused ingest: FMT = RCB_FMT_U8 / S8 / S16 reads the SDR's wire format directly - the TMA row copy
// moves 2 / 2 / 4 bytes per sample instead of 8 (HBM read traffic and, on the host path, PCIe bytes drop 4x / 4x / 2x),
// the first radix pass converts while it loads ((v + offset), the scale is folded into the taps), nothing else
// changes.  configs/config_denver_usrp.py:20 (otw_format sc8), every rtlsdr config (u8), logging_receiver.py:107-109.
template <int FMT>
__device__ __forceinline__ float2 fm1_load_sample(const float2* wf, int idx, float off) {
    if constexpr (FMT == 0) {
        return wf[idx];
    } else if constexpr (FMT == 3) {  // s16 pairs
        const int v = reinterpret_cast<const int*>(wf)[idx];
        return make_float2((float)(short)(v & 0xffff) + off, (float)(short)(v >> 16) + off);
    } else if constexpr (FMT == 2) {  // s8 pairs
        const unsigned short v = reinterpret_cast<const unsigned short*>(wf)[idx];
        return make_float2((float)(signed char)(v & 0xff) + off, (float)(signed char)(v >> 8) + off);
    } else {                          // u8 pairs
        const unsigned short v = reinterpret_cast<const unsigned short*>(wf)[idx];
        return make_float2((float)(v & 0xff) + off, (float)(v >> 8) + off);
    }
}

// p.twiddle / p.taps layouts as for pfb_fm_tma_kernel (dense swizzled table, float4 tap groups).
// out_rank: 2 = plain [N][ostride] (tensor (t, m)), 3 = time blocks of 2^k >= 8 frames (tensor (t_lo, m, t_hi)).
template <int FMT = 0>
__global__ void __launch_bounds__(256, 2) pfb_fm1_kernel(const __grid_constant__ CUtensorMap tm_out, const PfbParams p,
                                                         const int out_rank) {
    constexpr uint32_t ROWB = 1024u * (FMT == 0 ? 8u : (FMT == 3 ? 4u : 2u));  // bytes of one input row
    using G = PfbFm1Geom;
    constexpr int R = 32, N = 1024, W = 8, FPI = 8, THREADS = 256;
    extern __shared__ unsigned char smem_fm1_raw[];
    const uint32_t base = (smem_addr_u32(smem_fm1_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_fm1_raw + (base - smem_addr_u32(smem_fm1_raw));
    float* ring = reinterpret_cast<float*>(gbase);  // [8 slots][1024] angles, then [1024][8] output rows
    float2* work_all = reinterpret_cast<float2*>(gbase + G::off_work);
    float2* tws = reinterpret_cast<float2*>(gbase + G::off_tw);
    float* taps_s = reinterpret_cast<float*>(gbase + G::off_taps);
    uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + G::off_bar);
    uint64_t* tile_done = bars + W;      // all 256 threads have written their tile rows (thread 0 waits, then stores)
    uint64_t* ring_free = bars + W + 1;  // the stores have read the tile: the ring may be rewritten
    const uint32_t a_tile = base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ll = lane;
    float2* work = work_all + warp * N;
    float2* wf = work;
    uint64_t* row_bar = bars + warp;

    for (int i = tid; i < N; i += THREADS) {
        tws[i] = p.twiddle[i];
        taps_s[i] = (FMT == 0) ? p.taps[i] : p.taps[i] * p.in_scale;
    }
    if (tid < W) mbar_init(bars + tid, 1);
    if (tid == W) mbar_init(tile_done, THREADS);
    if (tid == W + 1) mbar_init(ring_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tid == 0) prefetch_tmap(&tm_out);
    __syncthreads();
    const float4* tap4 = reinterpret_cast<const float4*>(taps_s);
    // housekeeping that used to be two more stream operations per block: the last CTA zeroes the work counter of the NEXT
    // launch (ping-pong pair) and, for complex64 input, saves the block's last row as the next block's "frame -1"
    if (blockIdx.x == gridDim.x - 1) {
        if (tid == 0 && p.next_counter) *p.next_counter = 0;
        if constexpr (FMT == 0) {
            if (p.hist_out && p.T >= 1) {
                const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)(p.T - 1) * N);
                float4* dst = reinterpret_cast<float4*>(p.hist_out);
                for (int i = tid; i < N / 2; i += THREADS) dst[i] = __ldg(src + i);
            }
        }
    }

    // work distribution as in pfb_fm_tma_kernel: static run (RCB_FM1_STAT_NUM / DEN of the even share) + dynamic tail chunks, each range
    // preceded by a warm-up iteration (recomputes the 8 frames before it: the carried angles are rebuilt, nothing
    // is stored)
    constexpr int kTailChunk = RCB_FM1_TAIL_CHUNK;
    const int NI = (p.T + FPI - 1) / FPI;
    const int stat = (int)(((long long)(NI / (int)gridDim.x) * RCB_FM1_STAT_NUM) / RCB_FM1_STAT_DEN);
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
    int nxt0 = NI, nxt1 = NI;

    long long frame0 = (long long)(cur0 - 1) * FPI + warp;
    auto row_src = [&](long long f) -> const void* {
        if constexpr (FMT == 0) {
            return pfb_row_ptr<R>(p, f);
        } else {  // raw rows; rows that do not exist (zero samples) are flagged by the caller, any valid row is loaded
            if (f >= p.T) f = p.T - 1;
            if (f >= 0) return reinterpret_cast<const char*>(p.x) + (size_t)f * ROWB;
            if (f >= -(long long)p.hist_valid) return reinterpret_cast<const char*>(p.hist_raw) + (size_t)(f + p.P) * ROWB;
            return reinterpret_cast<const char*>(p.x);
        }
    };
    auto issue_rows = [&](long long f0) {
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(row_bar, ROWB);
            tma_bulk_g2s(work, row_src(f0), ROWB, row_bar);
        }
    };
    // A warm-up iteration exists only to rebuild the angles of the frame before the range: the warp that owns the LAST
    // frame of the iteration does its normal work, the other seven skip theirs (their issue slots go to the second
    // CTA of the SM) and just start the copy of their first real frame.
    const bool warm_owner = (warp == W - 1);
    if (warm_owner) issue_rows(frame0);

    uint32_t row_par = 0, done_par = 0, free_par = 0;
    bool first_ever = true;
    float prev[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int it = cur0 - 1;; ++it) {
        const bool range_first = (it == cur0 - 1);
        if (range_first && tid == 0) {
            const int c = atomicAdd(p.work_counter, 1);
            s_next = tail0 + c * kTailChunk;
        }
        const bool works = !range_first || warm_owner;
        if (works) {
            mbar_wait(row_bar, row_par);
            row_par ^= 1u;
        }
        const bool range_last = (it + 1 == cur1);
        auto issue_next = [&]() {
            if (!range_last) {
                frame0 += FPI;
                issue_rows(frame0);
            } else if (nxt0 < NI) {  // warm-up frames of the next range
                frame0 = (long long)(nxt0 - 1) * FPI + warp;
                if (warm_owner) issue_rows(frame0);
            }
        };
        float ph[R];
        if (!works) {
            issue_next();
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) ph[m2] = 0.f;
        } else {
            float2 pr[R / 2], pi[R / 2];
            {
                float hreg[R];
#pragma unroll
                for (int jq = 0; jq < R / 4; ++jq) {
                    const float4 h = tap4[jq * R + ll];
                    hreg[4 * jq + 0] = h.x; hreg[4 * jq + 1] = h.y; hreg[4 * jq + 2] = h.z; hreg[4 * jq + 3] = h.w;
                }
                const float off = p.in_off;
                auto get = [&](auto j) { return fm1_load_sample<FMT>(wf, (R - 1 - decltype(j)::value) * R + ll, off); };
                auto tap = [&](auto j) { return hreg[R - 1 - decltype(j)::value]; };
                fft_packed<R, +1, true>(pr, pi, get, tap);
            }
            __syncwarp();
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
                __syncwarp();
                issue_next();
                auto get = [&](auto j) { return u[decltype(j)::value]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, +1, false>(pr, pi, get, tap);
            }
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                const float2 a = atan2_nan_p2(pi[q], pr[q]);
                ph[2 * q] = a.x;
                ph[2 * q + 1] = a.y;
            }
            if constexpr (FMT != 0) {
                // a frame before the first real sample is all zeros: its angle is the (0,0) sentinel.  (Frames past the
                // end of the block are never stored.)  The raw zero row cannot be represented for u8 (offset 127.4).
                const long long fcur = (long long)it * FPI + warp;
                if (fcur < -(long long)p.hist_valid) {
#pragma unroll
                    for (int m2 = 0; m2 < R; ++m2) ph[m2] = __int_as_float(0x7fc00000);
                }
            }
        }
        if (!first_ever) {  // the previous iteration's tile has been read by its TMA stores
            if (tid == 0) {   // (checked here, a whole FFT later, so the wait never stalls the issuing warp)
                tma_store_wait_read();
                mbar_arrive(ring_free);
            }
            mbar_wait(ring_free, free_par);
            free_par ^= 1u;
        }
        first_ever = false;
        {
            float* fb = ring + warp * N;  // slot = frame within the iteration = warp
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) fb[m2 * R + ll] = ph[m2];
        }
        __syncthreads();  // (A) the 8 angle rows of this iteration are in the ring
        if (range_first) {
            nxt0 = s_next;
            nxt1 = min(nxt0 + kTailChunk, NI);
            if (range_last && nxt0 < NI) {
                frame0 = (long long)(nxt0 - 1) * FPI + warp;
                if (warm_owner) issue_rows(frame0);
            }
        }
        // ---- demod: thread t owns channels t + 256 q ----
        float o[4][8];
        {
            float a[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) a[j][q] = ring[j * N + q * 256 + tid];
#pragma unroll
            for (int q = 0; q < 4; q += 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 cur = make_float2(a[j][q], a[j][q + 1]);
                    const float2 prv = (j == 0) ? make_float2(prev[q], prev[q + 1]) : make_float2(a[j - 1][q], a[j - 1][q + 1]);
                    float2 d = p2sub(cur, prv);
                    const float2 k = p2add(p2fmas(d, 0.15915494309189535f, make_float2(12582912.0f, 12582912.0f)),
                                           make_float2(-12582912.0f, -12582912.0f));
                    d = p2fmas(k, -6.283185307179586f, d);
                    d = p2muls(d, p.gain);
                    o[q][j] = (d.x != d.x) ? 0.0f : d.x;
                    o[q + 1][j] = (d.y != d.y) ? 0.0f : d.y;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) prev[q] = a[7][q];
        }
        __syncthreads();  // (B) every ring entry has been read: the region becomes the output tile
        const bool live = (it >= cur0);
        const long long t0 = (long long)it * FPI;
        const bool full = live && (t0 + 8 <= p.T);
        if (full) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = q * 256 + tid;
                const uint32_t rowa = a_tile + (uint32_t)m * 32u;
                const uint32_t sw = (uint32_t)((m >> 2) & 1);  // CU_TENSOR_MAP_SWIZZLE_32B: chunk ^= address bit 7
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowa + ((0u ^ sw) << 4)), "f"(o[q][0]), "f"(o[q][1]),
                             "f"(o[q][2]), "f"(o[q][3])
                             : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(rowa + ((1u ^ sw) << 4)), "f"(o[q][4]), "f"(o[q][5]),
                             "f"(o[q][6]), "f"(o[q][7])
                             : "memory");
            }
            fence_async_smem();
        } else if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = q * 256 + tid;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (t0 + j < p.T) p.out_fm[pfb_out_index(p, m, t0 + j)] = o[q][j];
            }
        }
        mbar_arrive(tile_done);
        if (tid == 0) {
            mbar_wait(tile_done, done_par);
            if (full) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (out_rank == 2) {
                        tma_store_2d(&tm_out, (int)t0, 256 * b, a_tile + (uint32_t)b * 8192u);
                    } else {
                        const int kb = p.oblock_log2;
                        tma_store_3d(&tm_out, (int)(t0 & ((1LL << kb) - 1)), 256 * b, (int)(t0 >> kb), a_tile + (uint32_t)b * 8192u);
                    }
                }
                tma_store_commit();
            }
        }
        done_par ^= 1u;
        if (range_last) {
            if (nxt0 >= NI) break;
            cur0 = nxt0;
            cur1 = nxt1;
            it = cur0 - 2;  // ++it -> warm-up iteration of the new range
        }
    }
    if (tid == 0) tma_store_wait_all();
}

}  // namespace rcb
