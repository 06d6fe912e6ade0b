// K3 (persistent)  fft_scan : the whole scan chain of fft_vector.py:37-60 - window, four-step FFT, |.|^2, log10 + 1,
// running block sums - for ALL frames of a call in ONE persistent launch.
//
// Round 1 ran three kernels per 4-frame sub-batch (column pass, row pass, fold) on two streams: 22 us launches with 1.7
// waves each, 960 launches per 20 steps, a `vals` round trip (4 B / sample written and re-read) and DRAM traffic for
// the intermediate whenever a sub-batch fell out of L2 (profiles/r01_fft_*: warps_active 22 %, long_sb 31-37 %,
// issue 34-36 %, 0.14 of the HBM roofline).  Here a persistent grid (2 CTAs / SM) pulls tasks from one queue:
//   cols(f, tile) : the TMA-tile column pass of fft_cols_tma_kernel (first two radix passes of frame f over n1, window
//                   folded into the first DIF stage, W_L^{-n2 k1} twiddle) -> scratch ring slot f mod RING;
//   rows(f, tile) : the row pass of fft_rows_kernel (two packed radix passes over n2, |.|^2, log10 + 1, fftshift) whose
//                   results are ADDED STRAIGHT INTO the block sum acc[L] - no vals array, no fold kernel.
// The queue is ordered cols(0..LA-1), then cols(f), rows(f - LA) interleaved, so column tiles (HBM-bound) and row tiles
// (compute-bound) of neighbouring frames overlap on every SM and only RING frames of the 8 B / sample intermediate are
// ever live: it stays in L2 (32 MB for 2^20-point frames), HBM sees the 8 B / sample of input and nothing else.
// Dependencies are global counters, all pointing at strictly EARLIER tasks of the queue (a persistent grid executes
// every dequeued task, so waiting cannot deadlock):
//   rows(f, .)    waits until all column tiles of frame f are done          (cols_done[f mod RING])
//   cols(f, .)    waits until all row tiles of frame f - RING are done      (rows_done[f mod RING])
//   rows(f, tile) adds into acc after rows(f - 1, tile) has                 (tile_seq[tile]) -> the sums are accumulated
//                 in frame order: results are bit-identical for any split of a stream into calls, like round 1.
#pragma once
#include "fft_logpow.cuh"

namespace rcb {

struct FftScanParams {
    FftParams fp;          // x = all frames of this launch, scratch = ring base (RING frames), vals unused
    int nfr;               // frames in this launch
    int ring;              // scratch ring depth in frames
    int la;                // column tiles run this many frames ahead of the row tiles
    int avg;               // averaging block length
    int in_block0;         // frames already summed into the current block before this launch
    int* task_counter;     // zeroed before the launch
    unsigned* cols_done;   // [ring]   zeroed before the launch
    unsigned* rows_done;   // [ring]   zeroed before the launch
    unsigned* tile_seq;    // [row tiles per frame]  zeroed before the launch
    float* acc;            // [L] running block sum (persists across launches)
    float* emit;           // completed vectors of this launch, vector v at emit + v * L
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 polls, the CTA barrier publishes the acquired state to the other threads
__device__ __forceinline__ void cta_wait_counter(const unsigned* p, unsigned target, int tid) {
    if (tid == 0) {
        while (ld_acquire_u32(p) < target) __nanosleep(20);
    }
    __syncthreads();
}

template <int R1, int R2>
struct FftScanGeom {
    using GC = FftColsGeom<R1>;
    using GR = FftGeom<R2>;
    static constexpr size_t cols_smem = GC::smem;
    static constexpr size_t rows_smem = GR::b_smem();
    static constexpr size_t smem = (cols_smem > rows_smem ? cols_smem : rows_smem) + 64;
    static constexpr int NC_DIV = GC::CB;   // columns per column tile
    static constexpr int NR_DIV = GR::CB;   // rows per row tile
};

template <int R1, int R2>
__global__ void __launch_bounds__(256, 2) fft_scan_kernel(const __grid_constant__ CUtensorMap tm, const FftScanParams sp) {
    using SG = FftScanGeom<R1, R2>;
    using GC = FftColsGeom<R1>;
    using GR = FftGeom<R2>;
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    __shared__ int s_task;
    const FftParams& p = sp.fp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L1 = p.L1, L2 = p.L2;
    const int NC = L2 / SG::NC_DIV, NR = L1 / SG::NR_DIV;
    const int la = min(sp.la, sp.nfr);
    const int total = sp.nfr * (NC + NR);
    uint64_t* tbar = reinterpret_cast<uint64_t*>(smem_raw128 + SG::smem - 16);  // TMA arrival barrier of the column pass
    uint32_t tma_par = 0;
    if (tid == 0) {
        mbar_init(tbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&tm);
    }
    __syncthreads();

    // the id of the next task is fetched while the current one runs (the atomic's round trip is off the critical path)
    if (tid == 0) s_task = atomicAdd(sp.task_counter, 1);
    __syncthreads();
    for (;;) {
        const int t = s_task;
        __syncthreads();   // everyone has read s_task (and is done with the previous task's shared memory)
        if (t >= total) break;
        if (tid == 0) s_task = atomicAdd(sp.task_counter, 1);
        // ---- decode: cols(0..la-1) | [cols(s), rows(s - la)] for s = la..nfr-1 | rows(nfr-la..nfr-1) ----
        bool is_cols;
        int f, tile;
        if (t < la * NC) {
            is_cols = true;
            f = t / NC;
            tile = t % NC;
        } else {
            const int t1 = t - la * NC;
            const int mid = (sp.nfr - la) * (NC + NR);
            if (t1 < mid) {
                const int s = la + t1 / (NC + NR), r = t1 % (NC + NR);
                if (r < NC) {
                    is_cols = true;
                    f = s;
                    tile = r;
                } else {
                    is_cols = false;
                    f = s - la;
                    tile = r - NC;
                }
            } else {
                const int t2 = t1 - mid;
                is_cols = false;
                f = sp.nfr - la + t2 / NR;
                tile = t2 % NR;
            }
        }
        const int slot = f % sp.ring;
        const unsigned gen = (unsigned)(f / sp.ring);

        if (is_cols) {
            // ================= column tile: body of fft_cols_tma_kernel<R1> =================
            constexpr int R = R1;
            constexpr int CB = GC::CB, BL = GC::BL, RP = GC::RP, S = GC::S, CL1 = GC::L1;
            float2* tile_s = reinterpret_cast<float2*>(smem_raw128);
            float2* tws = reinterpret_cast<float2*>(smem_raw128 + GC::tile_bytes);
            float2* rho = tws + R * S;
            const int sub = lane / CB, c = lane % CB;
            const int b = warp * BL + sub;
            const int c0 = tile * CB;
            const int n2 = c0 + c;
            if (tid == 0) {
                fence_proxy_async();  // the tile region was read / written through the generic proxy by the previous task
                mbar_expect_tx(tbar, (uint32_t)(CL1 * CB * 8));
#pragma unroll
                for (int r0 = 0; r0 < CL1; r0 += GC::BOX_ROWS) tma_load_2d(tile_s + r0 * CB, &tm, c0 * 2, f * CL1 + r0, tbar);
            }
            float wv[R];
#pragma unroll
            for (int a = 0; a < R; ++a) wv[a] = __ldg(p.window + (size_t)(R * a + b) * L2 + n2);
            for (int i = tid; i < R * S; i += 256) tws[i] = p.tw1[i];
            {
                const int kb = tid / CB, cc = tid % CB;
                const unsigned e = ((unsigned)(c0 + cc) * (unsigned)(R * kb)) & (unsigned)(p.L - 1);
                float sn, cs;
                sincospif(-2.0f * (float)e / (float)p.L, &sn, &cs);
                rho[kb * CB + cc] = make_float2(cs, sn);
            }
            // the ring slot may be overwritten once every row tile of its previous frame has read it
            cta_wait_counter(sp.rows_done + slot, (unsigned)NR * gen, tid);
            mbar_wait(tbar, tma_par);
            tma_par ^= 1u;

            float2 pr[R / 2], pi[R / 2];
            {
                auto get = [&](auto j) { return tile_s[(R * decltype(j)::value + b) * CB + c]; };
                auto tap = [&](auto j) { return wv[decltype(j)::value]; };
                fft_packed<R, -1, true>(pr, pi, get, tap);
            }
            __syncthreads();
            {
                const float4* twp = reinterpret_cast<const float4*>(tws + b * S);
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const float4 tq = twp[q];
                    tile_s[((2 * q) * RP + b) * CB + c] =
                        make_float2(fmaf(pr[q].x, tq.x, -pi[q].x * tq.y), fmaf(pr[q].x, tq.y, pi[q].x * tq.x));
                    tile_s[((2 * q + 1) * RP + b) * CB + c] =
                        make_float2(fmaf(pr[q].y, tq.z, -pi[q].y * tq.w), fmaf(pr[q].y, tq.w, pi[q].y * tq.z));
                }
            }
            __syncthreads();
            const int ka = b;
            {
                auto get = [&](auto j) { return tile_s[(ka * RP + decltype(j)::value) * CB + c]; };
                auto tap = [&](auto) { return 1.0f; };
                fft_packed<R, -1, false>(pr, pi, get, tap);
            }
            float2 b0;
            {
                const unsigned e = ((unsigned)n2 * (unsigned)ka) & (unsigned)(p.L - 1);
                float sn, cs;
                sincospif(-2.0f * (float)e / (float)p.L, &sn, &cs);
                b0 = make_float2(cs, sn);
            }
            float2* out = p.scratch + ((size_t)slot * CL1 + ka) * L2 + n2;
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
                const float2 w0 = cmul(b0, rho[(2 * q) * CB + c]), w1 = cmul(b0, rho[(2 * q + 1) * CB + c]);
                __stcg(out + (size_t)(R * (2 * q)) * L2, cmul(make_float2(pr[q].x, pi[q].x), w0));
                __stcg(out + (size_t)(R * (2 * q + 1)) * L2, cmul(make_float2(pr[q].y, pi[q].y), w1));
            }
            __syncthreads();
            if (tid == 0) red_release_add_u32(sp.cols_done + slot, 1u);   // release at gpu scope, cumulative over the barrier
        } else {
            // ================= row tile: body of fft_rows_kernel<R2> + accumulation into the block sum =================
            constexpr int R = R2;
            constexpr int N = GR::N, F = GR::F, CB = GR::CB, S = GR::S, FS = GR::FS;
            float2* bufs = reinterpret_cast<float2*>(smem_raw128);
            float2* tws = bufs + CB * FS;
            const int fr = lane / R, ll = lane % R;
            const int r0 = tile * CB;
            for (int i = tid; i < R * S; i += 256) tws[i] = p.tw2[i];
            cta_wait_counter(sp.cols_done + slot, (unsigned)NC * (gen + 1u), tid);   // (also orders the tws stores)
            const int row = warp * F + fr;
            const float2* src = p.scratch + (size_t)slot * p.L + (size_t)(r0 + row) * N;
            float2 v[R];
#pragma unroll
            for (int jj = 0; jj < R; ++jj) v[jj] = __ldcg(src + jj * R + ll);
            float2* buf = bufs + row * FS;
            warp_fft_2pass_packed<R, -1, false>(v, buf, tws, ll);
            float* fb = reinterpret_cast<float*>(buf);
#pragma unroll
            for (int m2 = 0; m2 < R; ++m2) {
                const float pw = fmaf(v[m2].x, v[m2].x, v[m2].y * v[m2].y);
                fb[m2 * S + ll] = fmaf(log2f(fmaxf(pw, 1e-18f)), 0.30102999566398120f, 1.0f);
            }
            __syncthreads();   // this CTA's scratch reads are done (consumed by the transforms) before the slot is handed back
            if (tid == 0) red_release_add_u32(sp.rows_done + slot, 1u);
            // sums in frame order: wait for this tile's turn
            cta_wait_counter(sp.tile_seq + tile, (unsigned)f, tid);
            const int pos_in_block = sp.in_block0 + f;                   // frames summed before this one
            const bool completes = ((pos_in_block + 1) % sp.avg) == 0;
            float* emit = completes ? sp.emit + (size_t)((pos_in_block + 1) / sp.avg - 1) * p.L : nullptr;
            constexpr int ITEMS = N * (CB / 8) / 256;
#pragma unroll
            for (int q = 0; q < ITEMS; ++q) {
                const int item = q * 256 + tid;
                const int k2 = item % N, g = item / N;
                const int pos = (k2 / R) * S + (k2 % R);
                const int k2s = (k2 + N / 2) % N;  // fftshift: only k2 moves since L/2 = L1*(L2/2)
                float* ap = sp.acc + (size_t)k2s * L1 + r0 + 8 * g;
                float4 a0 = __ldcg(reinterpret_cast<const float4*>(ap));
                float4 a1 = __ldcg(reinterpret_cast<const float4*>(ap) + 1);
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = reinterpret_cast<const float*>(bufs + (8 * g + j) * FS)[pos];
                a0.x += o[0]; a0.y += o[1]; a0.z += o[2]; a0.w += o[3];
                a1.x += o[4]; a1.y += o[5]; a1.z += o[6]; a1.w += o[7];
                if (emit) {
                    float* ep = emit + (size_t)k2s * L1 + r0 + 8 * g;
                    __stcg(reinterpret_cast<float4*>(ep), a0);
                    __stcg(reinterpret_cast<float4*>(ep) + 1, a1);
                    a0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    a1 = a0;
                }
                __stcg(reinterpret_cast<float4*>(ap), a0);
                __stcg(reinterpret_cast<float4*>(ap) + 1, a1);
            }
            __syncthreads();
            if (tid == 0) st_release_u32(sp.tile_seq + tile, (unsigned)(f + 1));   // release: cumulative over the barrier
        }
        __syncthreads();   // publishes the prefetched s_task
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define SCK(call)                              \
    do {                                       \
        if ((call) != cudaSuccess) return -3;  \
    } while (0)

// one persistent launch over nfr device-resident frames at d_x; completed vectors go to d_emit (device).
// Returns 1 when the TMA descriptor cannot be built (caller falls back to the three-kernel pipeline).
template <int R1, int R2>
inline int fft_scan_launch(FftState& s, const float2* d_x, int nfr, float* d_emit, cudaStream_t st, int sm_count) {
    using SG = FftScanGeom<R1, R2>;
    auto kern = fft_scan_kernel<R1, R2>;
    rcb_tmap_encode_fn enc = tmap_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(d_x) & 15)) return 1;
    const int NC = s.L2 / SG::NC_DIV, NR = s.L1 / SG::NR_DIV;
    if (!s.scan_grid) {
        static bool attr_dev[64] = {};
        bool& attr = attr_dev[fft_cur_device()];
        if (!attr) {
            SCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG::smem));
            attr = true;
        }
        int nb = 0;
        SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, SG::smem));
        s.scan_grid = std::max(1, nb) * sm_count;
        // the column tiles run far enough ahead of the row tiles that a row task rarely finds its frame unfinished
        s.la = std::max(2, (2 * s.scan_grid + NC + NR - 1) / (NC + NR) + 1);
        s.ring = s.la + 2;
        SCK(cudaMalloc(&s.d_ring, (size_t)s.ring * s.L * sizeof(float2)));
        s.ctl_words = 1 + 2 * (size_t)s.ring + (size_t)NR;
        SCK(cudaMalloc(&s.d_ctl, s.ctl_words * sizeof(unsigned)));
    }
    CUtensorMap tm;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)s.L2 * 2, (cuuint64_t)nfr * (cuuint64_t)s.L1};
        const cuuint64_t gstr[1] = {(cuuint64_t)s.L2 * 8};
        const cuuint32_t box[2] = {(cuuint32_t)(FftColsGeom<R1>::CB * 2), (cuuint32_t)FftColsGeom<R1>::BOX_ROWS};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(d_x), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 1;
    }
    SCK(cudaMemsetAsync(s.d_ctl, 0, s.ctl_words * sizeof(unsigned), st));
    FftScanParams sp{};
    sp.fp.x = d_x;
    sp.fp.window = s.d_window;
    sp.fp.scratch = s.d_ring;
    sp.fp.vals = nullptr;
    sp.fp.tw1 = s.d_tw1;
    sp.fp.tw2 = s.d_tw2;
    sp.fp.t_lo = s.d_tlo;
    sp.fp.t_hi = s.d_thi;
    sp.fp.L = s.L;
    sp.fp.L1 = s.L1;
    sp.fp.L2 = s.L2;
    sp.nfr = nfr;
    sp.ring = s.ring;
    sp.la = s.la;
    sp.avg = s.avg;
    sp.in_block0 = s.in_block;
    sp.task_counter = reinterpret_cast<int*>(s.d_ctl);
    sp.cols_done = s.d_ctl + 1;
    sp.rows_done = s.d_ctl + 1 + s.ring;
    sp.tile_seq = s.d_ctl + 1 + 2 * s.ring;
    sp.acc = s.d_acc;
    sp.emit = d_emit;
    const int total = nfr * (NC + NR);
    kern<<<std::min(total, s.scan_grid), 256, SG::smem, st>>>(tm, sp);
    SCK(cudaGetLastError());
    return 0;
}

inline int fft_scan_dispatch(FftState& s, const float2* d_x, int nfr, float* d_emit, cudaStream_t st, int sm_count) {
    if (s.R1 == 8 && s.R2 == 8) return fft_scan_launch<8, 8>(s, d_x, nfr, d_emit, st, sm_count);
    if (s.R1 == 8 && s.R2 == 16) return fft_scan_launch<8, 16>(s, d_x, nfr, d_emit, st, sm_count);
    if (s.R1 == 16 && s.R2 == 16) return fft_scan_launch<16, 16>(s, d_x, nfr, d_emit, st, sm_count);
    if (s.R1 == 16 && s.R2 == 32) return fft_scan_launch<16, 32>(s, d_x, nfr, d_emit, st, sm_count);
    if (s.R1 == 32 && s.R2 == 32) return fft_scan_launch<32, 32>(s, d_x, nfr, d_emit, st, sm_count);
    return 1;
}

// rcb_fft_process on the persistent kernel.  Device input: ONE launch for the whole call.  Host input: the block is
// staged in chunks of ~32 MiB through two buffers (copy stream ws[i] -> compute stream), one launch per chunk.
// Returns 1 when the scan kernel cannot serve the call (the caller then runs fft_process).
inline int fft_process_scan(FftState& s, const float2* iq, size_t nsamples, int in_mem, float* out, size_t cap_vec,
                            int out_mem, size_t* nvec, cudaStream_t st, uint64_t* launches, uint64_t* h2d, uint64_t* d2h,
                            int sm_count) {
    *nvec = 0;
    const size_t L = (size_t)s.L;
    const size_t nframes = nsamples / L;
    const size_t will_emit = (s.in_block + nframes) / (size_t)s.avg;
    if (will_emit > cap_vec || (will_emit && !out)) return -6;  // RCB_ERANGE
    if (nframes == 0) return 0;
    if (in_mem != 0 && (reinterpret_cast<uintptr_t>(iq) & 15)) return 1;
    if (out_mem == 0 && will_emit && s.emit_cap < will_emit * L) {
        cudaFree(s.d_emit);
        s.d_emit = nullptr;
        s.emit_cap = 0;
        SCK(cudaMalloc(&s.d_emit, will_emit * L * sizeof(float)));
        s.emit_cap = will_emit * L;
    }
    float* d_out = (out_mem == 0) ? s.d_emit : out;
    size_t emitted = 0;
    if (in_mem != 0) {
        int rc = fft_scan_dispatch(s, iq, (int)nframes, d_out, st, sm_count);
        if (rc) return rc;
        *launches += 1;
        emitted = will_emit;
        s.in_block = (int)((s.in_block + nframes) % (size_t)s.avg);
    } else {
        const size_t chunk = std::max<size_t>(1, ((size_t)32 << 20) / (L * sizeof(float2)));  // frames per staged chunk
        for (int i = 0; i < 2; ++i) {
            if (s.in_cap2[i] < chunk * L) {
                cudaFree(s.d_in2[i]);
                s.d_in2[i] = nullptr;
                s.in_cap2[i] = 0;
                SCK(cudaMalloc(&s.d_in2[i], chunk * L * sizeof(float2)));
                s.in_cap2[i] = chunk * L;
            }
            if (!s.ev_copy[i]) SCK(cudaEventCreateWithFlags(&s.ev_copy[i], cudaEventDisableTiming));
            if (!s.ev_used[i]) SCK(cudaEventCreateWithFlags(&s.ev_used[i], cudaEventDisableTiming));
        }
        SCK(cudaEventRecord(s.ev_start, st));
        SCK(cudaStreamWaitEvent(s.ws[0], s.ev_start, 0));
        SCK(cudaStreamWaitEvent(s.ws[1], s.ev_start, 0));
        size_t done = 0;
        int k = 0;
        while (done < nframes) {
            const int sl = k & 1;
            const size_t nfr = std::min(chunk, nframes - done);
            if (k >= 2) SCK(cudaStreamWaitEvent(s.ws[sl], s.ev_used[sl], 0));  // the launch that read this buffer is done
            SCK(cudaMemcpyAsync(s.d_in2[sl], iq + done * L, nfr * L * sizeof(float2), cudaMemcpyHostToDevice, s.ws[sl]));
            *h2d += nfr * L * sizeof(float2);
            SCK(cudaEventRecord(s.ev_copy[sl], s.ws[sl]));
            SCK(cudaStreamWaitEvent(st, s.ev_copy[sl], 0));
            const size_t em = (s.in_block + nfr) / (size_t)s.avg;
            int rc = fft_scan_dispatch(s, s.d_in2[sl], (int)nfr, d_out + emitted * L, st, sm_count);
            if (rc) return (rc == 1 && done == 0) ? 1 : (rc == 1 ? -3 : rc);
            SCK(cudaEventRecord(s.ev_used[sl], st));
            *launches += 1;
            emitted += em;
            s.in_block = (int)((s.in_block + nfr) % (size_t)s.avg);
            done += nfr;
            ++k;
        }
    }
    if (out_mem == 0 && emitted) {
        SCK(cudaMemcpyAsync(out, s.d_emit, emitted * L * sizeof(float), cudaMemcpyDeviceToHost, st));
        *d2h += emitted * L * sizeof(float);
    }
    if (in_mem == 0 || out_mem == 0) SCK(cudaStreamSynchronize(st));
    *nvec = emitted;
    return 0;
}
#undef SCK

}  // namespace rcb
