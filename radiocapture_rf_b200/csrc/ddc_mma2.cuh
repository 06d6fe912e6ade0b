// K2 on the tensor cores, second generation:  ddc_mma2_kernel  - the A operand lives in TENSOR MEMORY.
//
// ddc_mma_kernel (ddc_mma.cuh) is bound by shared-memory bandwidth, not by the tensor pipe: per k-chunk it moves
// 48 KB in by TMA, 48 KB through the hi / lo split (read 16, write 32) and 96 KB out to the 12 MMAs (every tf32
// MMA of 128 x 128 x 8 re-reads 4 KB of A and 4 KB of B) = 192 KB at 128 B/clk = 1500 cycles, against 800 cycles of
// tensor-pipe time (ncu: tensor pipe 55 % of the active cycles).  Here the split warps write A hi / A lo straight into
// TMEM (tcgen05.st: one accumulator-lane = one A row, 32 columns per k-chunk) and the MMAs take A from TMEM
// (tcgen05.mma [d], [a], b-desc): shared memory carries 48 KB in, 16 KB to the split and 48 KB of B to the MMAs
// = 112 KB per chunk.
//
// TMEM (512 columns): main accumulator at column 0, cross-term accumulator at 128, four A stages at 256 + 64 s
// (32 columns hi + 32 columns lo each).  Two A stages were not enough: tcgen05.commit -> mbarrier -> tcgen05.st ->
// mbarrier -> MMA issue is about 1300 cycles, 1.7 chunks of tensor-pipe time (ncu: the split warps waited for a free
// stage in 26 % of all samples, tensor pipe 59 % busy).
//
// Accumulation accuracy: the tensor core adds into the fp32 accumulator with truncation - measured 2.8e-8 relative per
// k-step, systematic, so a 2327-tap window (588 k-steps) in ONE accumulator is 1.6e-5 off.  The main accumulation
// therefore runs in segments of `seg_len` chunks; after each the split warps drain the main accumulator into registers
// (fp32 round-to-nearest adds; the MMA warp waits, about 6 % of a 16-chunk segment, while the A stages fill up):
// the error no longer grows with the tap count.
#pragma once
#include "ddc_mma.cuh"

namespace rcb {

constexpr int kM2SA = 4;  // raw A tiles (TMA -> split warps)
constexpr int kM2SB = 5;  // B hi + B lo tile pairs (TMA -> MMA)
constexpr int kM2OffB = kM2SA * kDdcMmaTile;
constexpr int kM2OffBar = kM2OffB + kM2SB * 2 * kDdcMmaTile;
constexpr int kM2OffPar = kM2OffBar + 256;
constexpr int kM2Smem = kM2OffPar + 64 * 24 + 1024 /*alignment*/;
constexpr int kM2Threads = 224;
constexpr int kM2ST = 4;   // A stages in TMEM
constexpr uint32_t kM2ColCross = 128, kM2ColA = 256;

// D[tmem] (+)= A[tmem] * B[smem], 128 x N x 8 tf32
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 registers of this thread -> 32 consecutive columns of its TMEM lane
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// FREERUN (experiments build, timing study, garbage results): the MMA warp issues its 12 MMAs per chunk without waiting
// for operands, nobody else does anything - the tensor pipe's own time for the tile
template <bool FREERUN = false>
__global__ void __launch_bounds__(kM2Threads, 1)
ddc_mma2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const DdcChanDev* __restrict__ chans, const DdcMmaGroupDev* __restrict__ groups, int ngroups) {
    extern __shared__ unsigned char ddc_mma2_raw[];
    const uint32_t raw_addr = smem_addr_u32(ddc_mma2_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    unsigned char* base_p = ddc_mma2_raw + (base - raw_addr);
    const uint32_t bars = base + kM2OffBar;
    // barriers (8 bytes each): full_a[4] @0, empty_a[4] @32, full_b[5] @64, empty_b[5] @104, conv[4] @144, empty_t[4] @176,
    // seg_full @208, drained @216, accum @224; tmem slot @232
    auto bar_full_a = [&](int s) { return bars + 8u * s; };
    auto bar_empty_a = [&](int s) { return bars + 32u + 8u * s; };
    auto bar_full_b = [&](int s) { return bars + 64u + 8u * s; };
    auto bar_empty_b = [&](int s) { return bars + 104u + 8u * s; };
    auto bar_conv = [&](int s) { return bars + 144u + 8u * s; };
    auto bar_empty_t = [&](int s) { return bars + 176u + 8u * s; };
    const uint32_t bar_seg_full = bars + 208u, bar_drained = bars + 216u;
    const uint32_t bar_accum = bars + 224u;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_p + kM2OffBar + 232);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DdcMmaGroupDev& g = groups[blockIdx.y];
    const int kchunks = g.kchunks;
    const int kq = g.kq;
    const int seg_len = g.seg_len;
    const int tile = blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < kM2SA; ++s) {
            mbar_init_a(bar_full_a(s), 1);
            mbar_init_a(bar_empty_a(s), 4);
        }
        for (int s = 0; s < kM2SB; ++s) {
            mbar_init_a(bar_full_b(s), 1);
            mbar_init_a(bar_empty_b(s), 1);
        }
        for (int s = 0; s < kM2ST; ++s) {
            mbar_init_a(bar_conv(s), 4);
            mbar_init_a(bar_empty_t(s), 1);
        }
        mbar_init_a(bar_seg_full, 1);
        mbar_init_a(bar_drained, 4);
        mbar_init_a(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
    }
    double* s_ph0 = reinterpret_cast<double*>(base_p + kM2OffPar);
    double* s_cyc = s_ph0 + 64;
    float2** s_out = reinterpret_cast<float2**>(s_cyc + 64);
    if (tid >= 64 && tid < 128) {
        const int cslot = tid - 64;
        const int ci = (cslot < g.nch) ? g.ch[cslot] : -1;
        s_ph0[cslot] = (ci >= 0) ? chans[ci].phase0 : 0.0;
        s_cyc[cslot] = (ci >= 0) ? chans[ci].cyc : 0.0;
        s_out[cslot] = (ci >= 0) ? chans[ci].out_iq : nullptr;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(bars + 232u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && !FREERUN) {
            DdcChunkOrder ord(kq, kchunks);
            for (int it = 0; it < kchunks; ++it, ord.next()) {
                const int s = it % kM2SA;
                mbar_wait_a(bar_empty_a(s), ((uint32_t)(it / kM2SA) & 1u) ^ 1u);
                mbar_expect_tx_a(bar_full_a(s), (uint32_t)kDdcMmaTile);
                tma_load_2d(base + (uint32_t)s * kDdcMmaTile, &tm_a, ord.kc * 32, tile * 128, bar_full_a(s));
            }
        }
    } else if (warp == 6) {
        if (lane == 0 && !FREERUN) {
            DdcChunkOrder ord(kq, kchunks);
            for (int it = 0; it < kchunks; ++it, ord.next()) {
                const int s = it % kM2SB, kc = ord.kc;
                mbar_wait_a(bar_empty_b(s), ((uint32_t)(it / kM2SB) & 1u) ^ 1u);
                const uint32_t st = base + (uint32_t)kM2OffB + (uint32_t)s * 2u * kDdcMmaTile;
                mbar_expect_tx_a(bar_full_b(s), 2u * kDdcMmaTile);
                tma_load_2d(st, &tm_b, kc * 32, (int)blockIdx.y * 128, bar_full_b(s));
                tma_load_2d(st + (uint32_t)kDdcMmaTile, &tm_b, kc * 32, (ngroups + (int)blockIdx.y) * 128, bar_full_b(s));
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D fp32, A / B tf32, both K-major, M 128, N ncols
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.ncols >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t t_cross = tmem + kM2ColCross;
        int seg = 0, seg_end = seg_len;
        bool seg_first = true;
        for (int it = 0; it < kchunks; ++it) {
            const int st = it % kM2ST, sb = it % kM2SB;
            if (it == seg_end && !FREERUN) {
                // the segment's sum goes to the split warps' registers; the next segment starts the accumulator afresh
                if (lane == 0) tc_commit(bar_seg_full);
                __syncwarp();
                mbar_wait_a(bar_drained, (uint32_t)seg & 1u);
                tc_fence_after();
                ++seg;
                seg_end += seg_len;
                seg_first = true;
            }
            if (!FREERUN) {
                mbar_wait_a(bar_conv(st), (uint32_t)(it / kM2ST) & 1u);
                mbar_wait_a(bar_full_b(sb), (uint32_t)(it / kM2SB) & 1u);
            }
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sb_addr = base + (uint32_t)kM2OffB + (uint32_t)sb * 2u * kDdcMmaTile;
                const uint64_t b_hi = tc_smem_desc(sb_addr), b_lo = tc_smem_desc(sb_addr + (uint32_t)kDdcMmaTile);
                const uint32_t a_hi = tmem + kM2ColA + 64u * (uint32_t)st, a_lo = a_hi + 32u;
                const uint32_t t_main = tmem;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t adv = (uint64_t)(2 * j);  // 8 floats = 32 bytes along K inside the 128-byte swizzle row
                    const uint32_t ac = 8u * (uint32_t)j;    // 8 TMEM columns along K
                    tc_mma_tf32_ts(t_main, a_hi + ac, b_hi + adv, idesc, (seg_first && j == 0) ? 0u : 1u);
                    tc_mma_tf32_ts(t_cross, a_hi + ac, b_lo + adv, idesc, (it == 0 && j == 0) ? 0u : 1u);
                    tc_mma_tf32_ts(t_cross, a_lo + ac, b_hi + adv, idesc, 1u);
                }
                tc_commit(bar_empty_t(st));
                tc_commit(bar_empty_b(sb));
            }
            seg_first = false;
            __syncwarp();
        }
        if (lane == 0) tc_commit(bar_accum);
        __syncwarp();
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const int ncols = g.ncols;
        const int nsegs = (kchunks + seg_len - 1) / seg_len;
        float total[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) total[i] = 0.f;
        int next_drain = 0;
        // segment j's main accumulator -> registers
        auto drain = [&](int j) {
            mbar_wait_a(bar_seg_full, (uint32_t)j & 1u);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 16) {
                if (c0 < ncols) {
                    float t[16];
                    tc_ld16(trow + (uint32_t)c0, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) total[c0 + i] += t[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar_drained);
        };
        for (int it = 0; it < (FREERUN ? 0 : kchunks); ++it) {
            const int sa = it % kM2SA, st = it % kM2ST;
            mbar_wait_a(bar_full_a(sa), (uint32_t)(it / kM2SA) & 1u);
            // this thread's row of the raw tile: 8 x 16 bytes at the swizzled positions
            const unsigned char* rp = base_p + (size_t)sa * kDdcMmaTile + (size_t)row * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t h = __float_as_uint(vv[e]) & 0xffffe000u;
                    hi[4 * c + e] = h;
                    lo[4 * c + e] = __float_as_uint(vv[e] - __uint_as_float(h));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar_empty_a(sa));  // the raw tile is in registers
            mbar_wait_a(bar_empty_t(st), ((uint32_t)(it / kM2ST) & 1u) ^ 1u);
            tc_fence_after();
            tc_st32(trow + kM2ColA + 64u * (uint32_t)st, hi);
            tc_st32(trow + kM2ColA + 64u * (uint32_t)st + 32u, lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(bar_conv(st));
            // drain once the A stages are refilled up to 3 chunks past the segment end (a 4th would wait for an MMA that
            // itself waits for this drain)
            if (next_drain < nsegs - 1 && it + 1 >= seg_len * (next_drain + 1) + 3) {
                drain(next_drain);
                ++next_drain;
            }
        }
        while (next_drain < nsegs - 1 && !FREERUN) {
            drain(next_drain);
            ++next_drain;
        }
        // ---- epilogue: last segment + cross terms, derotate, store ----
        mbar_wait_a(bar_accum, 0u);
        tc_fence_after();
        const int o = g.o_head + tile * 128 + row;
        const uint32_t last = trow;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
            if (c0 < ncols) {
                float t[16], u[16];
                tc_ld16(last + (uint32_t)c0, t);
                tc_ld16(trow + kM2ColCross + (uint32_t)c0, u);
                if (o < g.nout) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int cslot = (c0 >> 1) + i;
                        float2* op = s_out[cslot];
                        if (op) {
                            // phase bookkeeping in double (cyc * o reaches 1e4 cycles), the rotation itself in float:
                            // the reduced phase |ph| <= 0.5 rounds to float with 3e-8 cycles = 2e-7 rad
                            double phs = fma(s_cyc[cslot], (double)o, s_ph0[cslot]);
                            phs -= floor(phs);
                            if (phs >= 0.5) phs -= 1.0;
                            float sf, cf;
                            sincospif(-2.0f * (float)phs, &sf, &cf);
                            const float ax = (total[c0 + 2 * i] + t[2 * i]) + u[2 * i];
                            const float ay = (total[c0 + 2 * i + 1] + t[2 * i + 1]) + u[2 * i + 1];
                            op[o] = make_float2(fmaf(ax, cf, -ay * sf), fmaf(ax, sf, ay * cf));
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace rcb
