// K3  fft_logpow : streaming windowed FFT + |.|^2 + log10 + frame accumulation (the scan path).
//
// Replaces the GNU Radio chain of fft_vector.py:37-60:
//   stream_to_vector(L) -> fft.fft_vcc(L, forward, window, shift=True) -> complex_to_mag_squared
//   -> nlog10_ff(1, L, 1) -> moving_average_ff(avg, 1, ...)
// (gr-fft fft_vcc_fftw.cc, gr-blocks complex_to_mag_squared / nlog10_ff / moving_average impl).
//
// L = L1*L2 (L1 <= L2, each R*R with R in {8,16,32}: L = 2^12, 2^14, 2^16, 2^18, 2^20) is done as a
// four-step FFT in two kernels whose 8 MB-per-frame intermediate stays L2-resident (126 MB L2), so HBM
// sees ~8 B per input sample:
//   A  fft_cols : CTA = 8*F adjacent columns n2 of one frame.  Coalesced (CB*8-byte row segments)
//                 window-multiplied load -> smem transpose -> per-warp L1-point FFT (two in-register
//                 radix-R passes, warp_fft_2pass) -> W_L^{-n2 k1} twiddle from a two-level smem table
//                 -> smem transpose -> coalesced store of B[k1][n2] to the L2-resident scratch.
//   B  fft_rows : CTA = 8*F adjacent rows k1.  Per-warp L2-point FFT of a contiguous row -> power ->
//                 log10 + 1 -> smem transpose -> 32 B-sector stores of vals[frame][fftshift(k)],
//                 k = k1 + L1*k2 (8 consecutive k1 per store).
//   C  fold     : acc[k] += sum over the frames of the sub-batch (fixed order => deterministic);
//                 when an averaging block of `avg` frames completes the vector is emitted.
#pragma once
#include <stdlib.h>

#include <vector>

#include "tma_utils.cuh"

#include "common.cuh"
#include "fft_inreg.cuh"
#include "fft_packed.cuh"
#include "pfb_fm_tma.cuh"  // mbarrier / TMA helpers

namespace rcb {

struct FftParams {
    const float2* x;       // frames of this sub-batch, frame f at x + f*L
    const float* window;   // [L]
    float2* scratch;       // [SB][L1][L2]
    float* vals;           // [SB][L]
    const float2* tw1;     // [R1][R1+2]  W_{L1}^{-ll m1}
    const float2* tw2;     // [R2][R2+2]  W_{L2}^{-ll m1}
    const float2* t_lo;    // [min(L,1024)] W_L^{-q}
    const float2* t_hi;    // [max(1,L/1024)] W_L^{-1024 q}
    int L, L1, L2;
};

template <int R>
struct FftGeom {
    static constexpr int N = R * R;
    static constexpr int F = 32 / R;
    static constexpr int CB = 8 * F;                        // columns (A) / rows (B) per CTA
    static constexpr int S = R + 2;
    static constexpr int FS = (R == 8) ? 88 : R * S;        // per-frame FFT scratch (complex)
    static constexpr int CS = (R == 32) ? (N + 66) : (N + 1);  // column stride of the input tile (A)
    static constexpr int OS = CB + 1;                       // row stride of the output tile (A)
    // A: region0 = max(input tile CB*CS, out tile N*OS, per-frame scratch CB*FS)
    static constexpr int A_REGION = (CB * CS > N * OS ? (CB * CS > CB * FS ? CB * CS : CB * FS)
                                                     : (N * OS > CB * FS ? N * OS : CB * FS));
    static constexpr size_t a_smem(int L) {
        return (size_t)A_REGION * 8 + (size_t)R * S * 8 + (size_t)(L < 1024 ? L : 1024) * 8 +
               (size_t)(L / 1024 > 0 ? L / 1024 : 1) * 8 + (size_t)CB * R * 8 /* rho */;
    }
    static constexpr size_t b_smem() { return (size_t)CB * FS * 8 + (size_t)R * S * 8; }
};

// ---------------------------------------------------------------------------------------------
// A: column FFTs.  grid (L2/CB, SB), block 256.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 2) fft_cols_kernel(const FftParams p) {
    using G = FftGeom<R>;
    constexpr int N = G::N, F = G::F, CB = G::CB, S = G::S, FS = G::FS, CS = G::CS, OS = G::OS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* region = reinterpret_cast<float2*>(smem_raw);
    float2* tws = region + G::A_REGION;
    float2* tlo = tws + R * S;
    const int nlo = p.L < 1024 ? p.L : 1024;
    float2* thi = tlo + nlo;
    const int nhi = p.L / 1024 > 0 ? p.L / 1024 : 1;
    float2* rho = thi + nhi;  // [CB][R]: W_L^{-(n2 * R * m2)} per column

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane / R, ll = lane % R;
    const int c0 = blockIdx.x * CB;
    const int f = blockIdx.y;
    const float2* xf = p.x + (size_t)f * p.L;

    for (int i = tid; i < R * S; i += 256) tws[i] = p.tw1[i];
    for (int i = tid; i < nlo; i += 256) tlo[i] = p.t_lo[i];
    for (int i = tid; i < nhi; i += 256) thi[i] = p.t_hi[i];
    // coalesced, window-multiplied load of the [N rows][CB cols] tile, transposed into region[c*CS + n1].
    // All loads of a batch are issued before the first use (memory-level parallelism: 16 x 8 B + 16 x 4 B per
    // thread in flight) - the first version had 4 in flight and was latency bound (29 % issue utilisation).
    {
        constexpr int PER = N * CB / 256;       // 32 elements per thread
        constexpr int BATCH = PER < 16 ? PER : 16;
#pragma unroll
        for (int b0 = 0; b0 < PER; b0 += BATCH) {
            float2 xv[BATCH];
            float wv[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int idx = (b0 + u) * 256 + tid;
                const size_t g = (size_t)(idx / CB) * p.L2 + c0 + (idx % CB);
                xv[u] = ld_stream_f2(xf + g);
                wv[u] = __ldg(p.window + g);
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int idx = (b0 + u) * 256 + tid;
                region[(idx % CB) * CS + (idx / CB)] = make_float2(xv[u].x * wv[u], xv[u].y * wv[u]);
            }
        }
    }
    __syncthreads();
    const int col = warp * F + fr;  // this lane's column within the CTA
    // W_L^{-n2 k1}, k1 = ll + R m2, factors as W^{-n2 ll} (per lane) * W^{-n2 R m2} (per column): the
    // per-column powers go to shared memory once (two-level table, exact), so the per-point twiddle is one
    // broadcast LDS + one complex multiply instead of two scattered (bank-conflicting) table reads.
    {
        const int e = ((c0 + col) * R * ll) & (p.L - 1);
        rho[col * R + ll] = cmul(tlo[e & 1023], thi[e >> 10]);
    }
    float2 v[R];
#pragma unroll
    for (int jj = 0; jj < R; ++jj) v[jj] = region[col * CS + jj * R + ll];
    __syncthreads();  // everyone holds its column in registers: the region can be reused
    float2* buf = region + col * FS;
    warp_fft_2pass_packed<R, -1, false>(v, buf, tws, ll);  // v[m2] = A[k1 = ll + R*m2] for column c0+col
    __syncthreads();  // all per-frame scratch dead: region becomes the [k1][CB] output tile
    {
        const int n2 = c0 + col;
        const int eb = (n2 * ll) & (p.L - 1);
        const float2 b = cmul(tlo[eb & 1023], thi[eb >> 10]);
        const float2* rc = rho + col * R;
#pragma unroll
        for (int m2 = 0; m2 < R; ++m2) {
            const int k1 = ll + R * m2;
            const float2 w = cmul(b, rc[m2]);
            region[k1 * OS + col] = cmul(v[m2], w);
        }
    }
    __syncthreads();
    float2* out = p.scratch + (size_t)f * p.L;
#pragma unroll 8
    for (int idx = tid; idx < N * CB; idx += 256) {
        const int k1 = idx / CB, c = idx % CB;
        out[(size_t)k1 * p.L2 + c0 + c] = region[k1 * OS + c];
    }
}

// ---------------------------------------------------------------------------------------------
// A (TMA version): column FFTs.  grid (L2/CB, SB), block 256, 2 CTAs/SM.
//   * the [L1 rows][CB columns] tile of one frame (CB = 256/R columns = 64..256 B per row) is fetched by ONE
//     thread as a 2-D TMA tensor copy (UTMALDG) straight into shared memory - no registers are tied up by loads in
//     flight and no transposing store pass; the window column (32 B sectors, L2 resident) and the tables are loaded
//     while the copy flies;
//   * L1 = R*R point DFT of a column, n1 = R a + b: lane (b, c) of warp b/BL runs the first packed radix-R pass
//     over a (rows R a + b: a warp reads BL full rows = 256 contiguous bytes per step, conflict free) with the
//     window multiply folded into the scalar DIF stage, applies W_L1^{-b ka}, and writes V[ka][b][c] back into the
//     tile (row pitch R+1 rows per ka for R = 32 so the second pass reads conflict free);
//   * second packed radix-R pass over b by lane (ka, c), W_L^{-n2 k1} = W_L^{-n2 ka} (one sincospi per lane)
//     x W_L^{-n2 R kb} (256-entry table per CTA), 64..256 B row segments of B[k1][n2] to the L2-resident scratch.
// ---------------------------------------------------------------------------------------------
template <int R>
struct FftColsGeom {
    static constexpr int L1 = R * R;
    static constexpr int CB = 256 / R;  // columns per CTA
    static constexpr int BL = R / 8;    // b (pass 1) / ka (pass 2) values per warp
    static constexpr int RP = R + (R == 32 ? 1 : 0);
    static constexpr int S = R + 2;     // row stride of the inner twiddle table (tw1)
    static constexpr int BOX_ROWS = L1 < 256 ? L1 : 256;
    static constexpr size_t tile_bytes = (size_t)R * RP * CB * 8;
    static constexpr size_t smem = tile_bytes + (size_t)R * S * 8 + (size_t)R * CB * 8 + 64;
};

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

template <int R>
__global__ void __launch_bounds__(256, 2) fft_cols_tma_kernel(const __grid_constant__ CUtensorMap tm, const FftParams p) {
    using G = FftColsGeom<R>;
    constexpr int L1 = G::L1, CB = G::CB, BL = G::BL, RP = G::RP, S = G::S;
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    float2* tile = reinterpret_cast<float2*>(smem_raw128);
    float2* tws = reinterpret_cast<float2*>(smem_raw128 + G::tile_bytes);
    float2* rho = tws + R * S;  // [kb][c] = W_L^{-(c0 + c) R kb}
    uint64_t* bar = reinterpret_cast<uint64_t*>(rho + R * CB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane / CB, c = lane % CB;
    const int b = warp * BL + sub;  // pass 1: residue b of n1;  pass 2: ka
    const int c0 = blockIdx.x * CB, f = blockIdx.y;
    const int n2 = c0 + c;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(L1 * CB * 8));
#pragma unroll
        for (int r0 = 0; r0 < L1; r0 += G::BOX_ROWS) tma_load_2d(tile + r0 * CB, &tm, c0 * 2, f * L1 + r0, bar);
    }
    float wv[R];  // window column of this lane: w[(R a + b) L2 + n2]
#pragma unroll
    for (int a = 0; a < R; ++a) wv[a] = __ldg(p.window + (size_t)(R * a + b) * p.L2 + n2);
    for (int i = tid; i < R * S; i += 256) tws[i] = p.tw1[i];
    {
        const int kb = tid / CB, cc = tid % CB;  // R * CB = 256 entries, one per thread
        const unsigned e = ((unsigned)(c0 + cc) * (unsigned)(R * kb)) & (unsigned)(p.L - 1);
        float sn, cs;
        sincospif(-2.0f * (float)e / (float)p.L, &sn, &cs);
        rho[kb * CB + cc] = make_float2(cs, sn);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    float2 pr[R / 2], pi[R / 2];
    {
        auto get = [&](auto j) { return tile[(R * decltype(j)::value + b) * CB + c]; };
        auto tap = [&](auto j) { return wv[decltype(j)::value]; };
        fft_packed<R, -1, true>(pr, pi, get, tap);  // (pr[q], pi[q]) = V[ka = 2q], V[2q + 1]
    }
    __syncthreads();  // every lane holds its column: the tile becomes the exchange buffer
    {
        const float4* twp = reinterpret_cast<const float4*>(tws + b * S);
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            const float4 t = twp[q];
            tile[((2 * q) * RP + b) * CB + c] =
                make_float2(fmaf(pr[q].x, t.x, -pi[q].x * t.y), fmaf(pr[q].x, t.y, pi[q].x * t.x));
            tile[((2 * q + 1) * RP + b) * CB + c] =
                make_float2(fmaf(pr[q].y, t.z, -pi[q].y * t.w), fmaf(pr[q].y, t.w, pi[q].y * t.z));
        }
    }
    __syncthreads();
    const int ka = b;
    {
        auto get = [&](auto j) { return tile[(ka * RP + decltype(j)::value) * CB + c]; };
        auto tap = [&](auto) { return 1.0f; };
        fft_packed<R, -1, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = A[ka + R kb], kb = 2q, 2q + 1
    }
    float2 b0;
    {
        const unsigned e = ((unsigned)n2 * (unsigned)ka) & (unsigned)(p.L - 1);
        float sn, cs;
        sincospif(-2.0f * (float)e / (float)p.L, &sn, &cs);
        b0 = make_float2(cs, sn);
    }
    float2* out = p.scratch + ((size_t)f * L1 + ka) * p.L2 + n2;
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
        const float2 w0 = cmul(b0, rho[(2 * q) * CB + c]), w1 = cmul(b0, rho[(2 * q + 1) * CB + c]);
        out[(size_t)(R * (2 * q)) * p.L2] = cmul(make_float2(pr[q].x, pi[q].x), w0);
        out[(size_t)(R * (2 * q + 1)) * p.L2] = cmul(make_float2(pr[q].y, pi[q].y), w1);
    }
}

// ---------------------------------------------------------------------------------------------
// B: row FFTs + log power.  grid (L1/CB, SB), block 256.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 2) fft_rows_kernel(const FftParams p) {
    using G = FftGeom<R>;
    constexpr int N = G::N, F = G::F, CB = G::CB, S = G::S, FS = G::FS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);
    float2* tws = bufs + CB * FS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane / R, ll = lane % R;
    const int r0 = blockIdx.x * CB;
    const int f = blockIdx.y;
    for (int i = tid; i < R * S; i += 256) tws[i] = p.tw2[i];
    __syncthreads();
    const int row = warp * F + fr;
    const float2* src = p.scratch + (size_t)f * p.L + (size_t)(r0 + row) * N;
    float2 v[R];
#pragma unroll
    for (int jj = 0; jj < R; ++jj) v[jj] = __ldcg(src + jj * R + ll);
    float2* buf = bufs + row * FS;
    warp_fft_2pass_packed<R, -1, false>(v, buf, tws, ll);  // v[m2] = X[k1 + L1*k2], k2 = ll + R*m2
    float* fb = reinterpret_cast<float*>(buf);
#pragma unroll
    for (int m2 = 0; m2 < R; ++m2) {
        const float pw = fmaf(v[m2].x, v[m2].x, v[m2].y * v[m2].y);
        // nlog10_ff(1, L, 1): log10(max(p, 1e-18)) + 1
        fb[m2 * S + ll] = fmaf(log2f(fmaxf(pw, 1e-18f)), 0.30102999566398120f, 1.0f);
    }
    __syncthreads();
    // thread item = (k2, group of 8 consecutive rows): one 32 B store into vals[f][fftshift(k)]
    constexpr int ITEMS = N * (CB / 8) / 256;
    float* vf = p.vals + (size_t)f * p.L;
#pragma unroll
    for (int q = 0; q < ITEMS; ++q) {
        const int item = q * 256 + tid;
        const int k2 = item % N, g = item / N;
        const int pos = (k2 / R) * S + (k2 % R);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = reinterpret_cast<const float*>(bufs + (8 * g + j) * FS)[pos];
        const int k2s = (k2 + N / 2) % N;  // fftshift: only k2 moves since L/2 = L1*(L2/2)
        st_global_v8(vf + (size_t)k2s * p.L1 + r0 + 8 * g, o);
    }
}

// C: acc[k] += sum_f vals[f][k] (f ascending).  If emit != null the block is complete: emit and clear.
__global__ void __launch_bounds__(256) fft_fold_kernel(const float* __restrict__ vals, int nframes, int L,
                                                       float* __restrict__ acc, float* __restrict__ emit) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (k >= L) return;
    float4 a = *reinterpret_cast<const float4*>(acc + k);
    for (int f = 0; f < nframes; ++f) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(vals + (size_t)f * L + k));
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
    }
    if (emit) {
        *reinterpret_cast<float4*>(emit + k) = a;
        a = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(acc + k) = a;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct FftState {
    bool configured = false;
    int L = 0, L1 = 0, L2 = 0, R1 = 0, R2 = 0, avg = 0;
    int sb = 0;               // frames per sub-batch (scratch kept L2 resident)
    bool cols_tma = true;     // column pass: TMA tile kernel (RCB_FFT_VARIANT=1 selects the register-staged one)
    bool rows_k1 = false;     // RCB_FFT_VARIANT=3: row pass on the persistent K1 pipeline (pfb_fm_tma_kernel, MODE = PFB_LOGPOW).
                              // Same 21 us per 4-frame sub-batch as fft_rows_kernel but it fills every SM, which stops the
                              // other work slot's kernels from overlapping (cfg4 101 vs 116 Gsps): not the default.
    float2* d_tw_rows = nullptr;  // dense swizzled W_L2^{-ll m1} table for it
    float2* d_zeros = nullptr;    // one all-zero row
    int* d_counter[2] = {nullptr, nullptr};
    int rows_blocks_per_sm = 0;
    int in_block = 0;         // frames already folded into the current averaging block
    float* d_window = nullptr;
    float2 *d_tw1 = nullptr, *d_tw2 = nullptr, *d_tlo = nullptr, *d_thi = nullptr;
    // two work slots: sub-batch b runs on slot b & 1 (own stream, scratch, vals, host staging) so the column
    // pass of one sub-batch overlaps the row pass / fold of the previous one (fills wave tails and
    // overlaps the memory-bound and compute-bound phases of these short kernels)
    float2* d_scratch2[2] = {nullptr, nullptr};
    float* d_vals2[2] = {nullptr, nullptr};
    float2* d_in2[2] = {nullptr, nullptr};
    size_t in_cap2[2] = {0, 0};
    cudaStream_t ws[2] = {nullptr, nullptr};
    cudaEvent_t ev_fold[2] = {nullptr, nullptr};
    cudaEvent_t ev_start = nullptr;
    unsigned long long batch_no = 0;
    float2* d_scratch = nullptr;
    float* d_vals = nullptr;
    float* d_acc = nullptr;
    float* d_emit = nullptr;  // staging for vectors when the caller's buffer is host memory
    size_t emit_cap = 0;
    float2* d_in = nullptr;   // staging for host input
    size_t in_cap = 0;
    // persistent scan kernel (fft_scan.cuh)
    bool use_scan = false;
    float2* d_ring = nullptr;     // scratch ring, `ring` frames
    int ring = 0, la = 0, scan_grid = 0;
    unsigned* d_ctl = nullptr;    // [1 task counter | ring cols_done | ring rows_done | row tiles tile_seq]
    size_t ctl_words = 0;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    // frame-resident kernel (fft_frame.cuh): L = 16384, the reference's own scan length
    bool use_frame = false;
    bool frame_off = false;         // rcb_fft_set_pipeline(1 / 2): run the tiled pipelines instead
    int fr_G = 0, fr_gpb = 0;       // frames per group, groups per averaging block
    float2* d_tw_frame = nullptr;   // [1024] dense swizzled W_1024^{-ll m1}
    float* d_carry = nullptr;       // [L] sums of the group a call ended in
    float* d_acc_alt = nullptr;     // second block accumulator (fft_fold_groups_kernel reads one, writes the other)
    float* d_partial = nullptr;     // [partial_rows][L] group sums of one launch
    size_t partial_rows = 0;
};

inline void fft_free(FftState& s) {
    cudaFree(s.d_window);
    cudaFree(s.d_tw1);
    cudaFree(s.d_tw2);
    cudaFree(s.d_tlo);
    cudaFree(s.d_thi);
    cudaFree(s.d_tw_rows);
    cudaFree(s.d_zeros);
    cudaFree(s.d_counter[0]);
    cudaFree(s.d_counter[1]);
    for (int i = 0; i < 2; ++i) {
        cudaFree(s.d_scratch2[i]);
        cudaFree(s.d_vals2[i]);
        cudaFree(s.d_in2[i]);
        if (s.ws[i]) cudaStreamDestroy(s.ws[i]);
        if (s.ev_fold[i]) cudaEventDestroy(s.ev_fold[i]);
    }
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    cudaFree(s.d_scratch);
    cudaFree(s.d_vals);
    cudaFree(s.d_acc);
    cudaFree(s.d_emit);
    cudaFree(s.d_in);
    cudaFree(s.d_ring);
    cudaFree(s.d_ctl);
    cudaFree(s.d_tw_frame);
    cudaFree(s.d_carry);
    cudaFree(s.d_acc_alt);
    cudaFree(s.d_partial);
    for (int i = 0; i < 2; ++i) {
        if (s.ev_copy[i]) cudaEventDestroy(s.ev_copy[i]);
        if (s.ev_used[i]) cudaEventDestroy(s.ev_used[i]);
    }
    s = FftState{};
}

inline std::vector<float2> fft_tw_table(int R, int N) {
    const int S = R + 2;
    std::vector<float2> t((size_t)R * S, make_float2(0.f, 0.f));
    for (int ll = 0; ll < R; ++ll)
        for (int m1 = 0; m1 < R; ++m1) {
            const double a = -2.0 * M_PI * (double)((ll * m1) % N) / (double)N;
            t[(size_t)ll * S + m1] = make_float2((float)cos(a), (float)sin(a));
        }
    return t;
}

// function attributes are per device: one flag / high-water mark per device ordinal (a receiver may place its sources
// on several GPUs of one process)
inline int fft_cur_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
}

#define FCK(call)                              \
    do {                                       \
        if ((call) != cudaSuccess) return -3;  \
    } while (0)

inline int fft_config(FftState& s, int L, const float* window, int avg, cudaStream_t st, int sm_count) {
    (void)sm_count;
    fft_free(s);
    int r1 = 0, r2 = 0;
    switch (L) {
        case 1 << 12: r1 = 8; r2 = 8; break;
        case 1 << 14: r1 = 8; r2 = 16; break;
        case 1 << 16: r1 = 16; r2 = 16; break;
        case 1 << 18: r1 = 16; r2 = 32; break;
        case 1 << 20: r1 = 32; r2 = 32; break;
        default: return -7;  // RCB_EUNSUPPORTED
    }
    s.L = L;
    s.R1 = r1;
    s.R2 = r2;
    s.L1 = r1 * r1;
    s.L2 = r2 * r2;
    s.avg = avg;
    // default: the three-kernel pipeline.  The persistent scan kernel (fft_scan.cuh, rcb_fft_set_pipeline) is parity-green
    // but measured slower (2^20 points: 91 vs 116 Gsps; its per-task barriers and in-order accumulation expose the
    // latencies the hardware CTA scheduler hides, profiles/r02_fft_scan_v2_summary.txt; short frames serialise on
    // the per-tile counters), so it is opt-in
    s.use_scan = false;
    // sub-batch: keep scratch (8 B) + vals (4 B) per sample under ~48 MB so they stay in L2
    size_t budget_mb = 48;
#ifdef RCB_EXPERIMENTS  // tuning builds only (radiocapture_rf_b200.build.build_experiments): never in the shipped library
    if (const char* e = getenv("RCB_FFT_VARIANT")) {
        s.cols_tma = (atoi(e) != 1);
        s.rows_k1 = (atoi(e) == 3);
        s.use_scan = (atoi(e) == 4);   // 4: persistent scan kernel
    }
    if (const char* e = getenv("RCB_FFT_SCRATCH_MB")) budget_mb = (size_t)std::max(1, atoi(e));
#endif
    s.sb = (int)std::max<size_t>(1, std::min<size_t>((budget_mb << 20) / ((size_t)L * 12), 4096));
    FCK(cudaMalloc(&s.d_window, sizeof(float) * L));
    FCK(cudaMemcpyAsync(s.d_window, window, sizeof(float) * L, cudaMemcpyHostToDevice, st));
    auto t1 = fft_tw_table(r1, s.L1), t2 = fft_tw_table(r2, s.L2);
    const int nlo = L < 1024 ? L : 1024, nhi = L / 1024 > 0 ? L / 1024 : 1;
    std::vector<float2> lo(nlo), hi(nhi);
    for (int q = 0; q < nlo; ++q) {
        const double a = -2.0 * M_PI * (double)q / (double)L;
        lo[q] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int q = 0; q < nhi; ++q) {
        const double a = -2.0 * M_PI * (double)q * 1024.0 / (double)L;
        hi[q] = make_float2((float)cos(a), (float)sin(a));
    }
    FCK(cudaMalloc(&s.d_tw1, t1.size() * sizeof(float2)));
    FCK(cudaMalloc(&s.d_tw2, t2.size() * sizeof(float2)));
    FCK(cudaMalloc(&s.d_tlo, lo.size() * sizeof(float2)));
    FCK(cudaMalloc(&s.d_thi, hi.size() * sizeof(float2)));
    FCK(cudaMemcpyAsync(s.d_tw1, t1.data(), t1.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    FCK(cudaMemcpyAsync(s.d_tw2, t2.data(), t2.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    FCK(cudaMemcpyAsync(s.d_tlo, lo.data(), lo.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    FCK(cudaMemcpyAsync(s.d_thi, hi.data(), hi.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < 2; ++i) {
        FCK(cudaMalloc(&s.d_scratch2[i], (size_t)s.sb * L * sizeof(float2)));
        FCK(cudaMalloc(&s.d_vals2[i], (size_t)s.sb * L * sizeof(float)));
        FCK(cudaStreamCreateWithFlags(&s.ws[i], cudaStreamNonBlocking));
        FCK(cudaEventCreateWithFlags(&s.ev_fold[i], cudaEventDisableTiming));
    }
    {   // row pass on the K1 pipeline: dense twiddle table, 16-byte chunks XOR-swizzled like pfb_fm_tma_kernel expects
        const int R = r2, N = s.L2;
        std::vector<float2> tt((size_t)N);
        for (int ll = 0; ll < R; ++ll) {
            const int sw = (R == 8) ? ((ll >> 1) & 3) : (ll & (R / 2 - 1));
            for (int m1 = 0; m1 < R; ++m1) {
                const double a = -2.0 * M_PI * (double)((ll * m1) % N) / (double)N;
                tt[(size_t)ll * R + ((((m1 >> 1) ^ sw) << 1) | (m1 & 1))] = make_float2((float)cos(a), (float)sin(a));
            }
        }
        FCK(cudaMalloc(&s.d_tw_rows, tt.size() * sizeof(float2)));
        FCK(cudaMemcpyAsync(s.d_tw_rows, tt.data(), tt.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
        FCK(cudaMalloc(&s.d_zeros, (size_t)N * sizeof(float2)));
        FCK(cudaMemsetAsync(s.d_zeros, 0, (size_t)N * sizeof(float2), st));
        FCK(cudaMalloc(&s.d_counter[0], sizeof(int)));
        FCK(cudaMalloc(&s.d_counter[1], sizeof(int)));
        FCK(cudaStreamSynchronize(st));
    }
    if (L == (1 << 14)) {
        // fft_frame_kernel: whole frame in one SM's shared memory.  Groups of about a dozen frames: short enough that a
        // call of a few averaging blocks fills the GPU, long enough that the group's flush and the one exposed frame
        // load are a few per cent of its work.
        const int R = 32, N = 1024;
        std::vector<float2> tt((size_t)N);
        for (int ll = 0; ll < R; ++ll) {
            const int sw = ll & (R / 2 - 1);
            for (int m1 = 0; m1 < R; ++m1) {
                const double a = -2.0 * M_PI * (double)((ll * m1) % N) / (double)N;
                tt[(size_t)ll * R + ((((m1 >> 1) ^ sw) << 1) | (m1 & 1))] = make_float2((float)cos(a), (float)sin(a));
            }
        }
        FCK(cudaMalloc(&s.d_tw_frame, tt.size() * sizeof(float2)));
        FCK(cudaMemcpyAsync(s.d_tw_frame, tt.data(), tt.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
        FCK(cudaStreamSynchronize(st));
        FCK(cudaMalloc(&s.d_carry, (size_t)L * sizeof(float)));
        FCK(cudaMemsetAsync(s.d_carry, 0, (size_t)L * sizeof(float), st));
        FCK(cudaMalloc(&s.d_acc_alt, (size_t)L * sizeof(float)));
        FCK(cudaMemsetAsync(s.d_acc_alt, 0, (size_t)L * sizeof(float), st));
        const int ng = (avg + 11) / 12;
        s.fr_G = (avg + ng - 1) / ng;
        s.fr_gpb = (avg + s.fr_G - 1) / s.fr_G;
        s.use_frame = true;
    }
    FCK(cudaEventCreateWithFlags(&s.ev_start, cudaEventDisableTiming));
    FCK(cudaMalloc(&s.d_acc, (size_t)L * sizeof(float)));
    FCK(cudaMemsetAsync(s.d_acc, 0, (size_t)L * sizeof(float), st));
    FCK(cudaStreamSynchronize(st));
    s.in_block = 0;
    s.configured = true;
    return 0;
}

inline int fft_reset(FftState& s, cudaStream_t st) {
    FCK(cudaMemsetAsync(s.d_acc, 0, (size_t)s.L * sizeof(float), st));
    FCK(cudaStreamSynchronize(st));
    s.in_block = 0;
    return 0;
}

template <int R>
inline int fft_launch_cols(const FftParams& p, int nfr, cudaStream_t st) {
    using G = FftGeom<R>;
    const size_t smem = G::a_smem(p.L);
    static size_t attr_dev[64] = {};  // the opt-in limit must cover the largest L used with this R
    size_t& attr = attr_dev[fft_cur_device()];
    if (smem > attr) {
        FCK(cudaFuncSetAttribute(fft_cols_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid(p.L2 / G::CB, nfr);
    fft_cols_kernel<R><<<grid, 256, smem, st>>>(p);
    FCK(cudaGetLastError());
    return 0;
}
// TMA column pass over the nfr frames at p.x; returns 1 when it cannot run (no encoder / misaligned input)
template <int R>
inline int fft_launch_cols_tma(const FftParams& p, int nfr, cudaStream_t st) {
    using G = FftColsGeom<R>;
    rcb_tmap_encode_fn enc = tmap_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(p.x) & 15)) return 1;
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)p.L2 * 2, (cuuint64_t)nfr * (cuuint64_t)p.L1};
    const cuuint64_t gstr[1] = {(cuuint64_t)p.L2 * 8};
    const cuuint32_t box[2] = {(cuuint32_t)(G::CB * 2), (cuuint32_t)G::BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(p.x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    static bool attr_dev[64] = {};
    bool& attr = attr_dev[fft_cur_device()];
    if (!attr) {
        FCK(cudaFuncSetAttribute(fft_cols_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem));
        attr = true;
    }
    dim3 grid(p.L2 / G::CB, nfr);
    fft_cols_tma_kernel<R><<<grid, 256, G::smem, st>>>(tm, p);
    FCK(cudaGetLastError());
    return 0;
}

// row pass on the K1 pipeline: rows of the scratch are the "frames", bins k2 the "channels"
template <int R>
inline int fft_launch_rows_k1(FftState& s, const FftParams& fp, int nfr, int slot, cudaStream_t st, int sm_count) {
    using G = PfbTmaGeom<R, 8, PFB_LOGPOW>;
    auto kern = pfb_fm_tma_kernel<R, 8, true, 1, PFB_LOGPOW>;
    if (!s.rows_blocks_per_sm) {
        FCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes));
        int nb = 0;
        FCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::THREADS, G::smem_bytes));
        s.rows_blocks_per_sm = nb > 0 ? nb : 1;
    }
    PfbParams p{};
    p.x = fp.scratch;
    p.hist = s.d_zeros;
    p.zeros = s.d_zeros;
    p.taps = nullptr;
    p.twiddle = s.d_tw_rows;
    p.work_counter = s.d_counter[slot];
    p.out_fm = fp.vals;
    p.ostride = fp.L1;
    p.T = nfr * fp.L1;
    p.P = 0;
    p.N = fp.L2;
    p.gain = 1.0f;
    const int NI = (p.T + G::FPI - 1) / G::FPI;
    const int grid = std::max(1, std::min(NI, s.rows_blocks_per_sm * sm_count));
    FCK(cudaMemsetAsync(s.d_counter[slot], 0, sizeof(int), st));
    kern<<<grid, G::THREADS, G::smem_bytes, st>>>(p);
    FCK(cudaGetLastError());
    return 0;
}

template <int R>
inline int fft_launch_rows(const FftParams& p, int nfr, cudaStream_t st) {
    using G = FftGeom<R>;
    const size_t smem = G::b_smem();
    static size_t attr_dev[64] = {};
    size_t& attr = attr_dev[fft_cur_device()];
    if (smem > attr) {
        FCK(cudaFuncSetAttribute(fft_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid(p.L1 / G::CB, nfr);
    fft_rows_kernel<R><<<grid, 256, smem, st>>>(p);
    FCK(cudaGetLastError());
    return 0;
}

inline int fft_process(FftState& s, const float2* iq, size_t nsamples, int in_mem, float* out, size_t cap_vec,
                       int out_mem, size_t* nvec, cudaStream_t st, uint64_t* launches, uint64_t* h2d, uint64_t* d2h,
                       int sm_count = kNumSMsB200) {
    *nvec = 0;
    const size_t L = (size_t)s.L;
    size_t nframes = nsamples / L;
    // vectors this call will complete
    const size_t will_emit = (s.in_block + nframes) / (size_t)s.avg;
    if (will_emit > cap_vec || (will_emit && !out)) return -6;  // RCB_ERANGE
    const float2* d_x = iq;
    if (in_mem == 0) {  // host: stage one sub-batch at a time (one staging buffer per work slot)
        const size_t need = (size_t)s.sb * L;
        for (int i = 0; i < 2; ++i)
            if (s.in_cap2[i] < need) {
                cudaFree(s.d_in2[i]);
                s.d_in2[i] = nullptr;
                s.in_cap2[i] = 0;
                FCK(cudaMalloc(&s.d_in2[i], need * sizeof(float2)));
                s.in_cap2[i] = need;
            }
    }
    if (out_mem == 0 && will_emit) {
        if (s.emit_cap < will_emit * L) {
            cudaFree(s.d_emit);
            s.d_emit = nullptr;
            s.emit_cap = 0;
            FCK(cudaMalloc(&s.d_emit, will_emit * L * sizeof(float)));
            s.emit_cap = will_emit * L;
        }
    }
    float* d_out = (out_mem == 0) ? s.d_emit : out;
    size_t done = 0, emitted = 0;
    // everything queued on the caller's stream so far (input production, previous outputs consumed) first
    FCK(cudaEventRecord(s.ev_start, st));
    FCK(cudaStreamWaitEvent(s.ws[0], s.ev_start, 0));
    FCK(cudaStreamWaitEvent(s.ws[1], s.ev_start, 0));
    int last_slot = -1;
    while (done < nframes) {
        const int sl = (int)(s.batch_no++ & 1);
        cudaStream_t wst = s.ws[sl];
        const int room = s.avg - s.in_block;
        const int nfr = (int)std::min<size_t>(std::min<size_t>(s.sb, nframes - done), (size_t)room);
        if (in_mem == 0) {
            FCK(cudaMemcpyAsync(s.d_in2[sl], iq + done * L, (size_t)nfr * L * sizeof(float2), cudaMemcpyHostToDevice, wst));
            *h2d += (size_t)nfr * L * sizeof(float2);
            d_x = s.d_in2[sl];
        } else {
            d_x = iq + done * L;
        }
        FftParams p{};
        p.x = d_x;
        p.window = s.d_window;
        p.scratch = s.d_scratch2[sl];
        p.vals = s.d_vals2[sl];
        p.tw1 = s.d_tw1;
        p.tw2 = s.d_tw2;
        p.t_lo = s.d_tlo;
        p.t_hi = s.d_thi;
        p.L = s.L;
        p.L1 = s.L1;
        p.L2 = s.L2;
        int rc = 1;
        if (s.cols_tma)
            rc = (s.R1 == 8) ? fft_launch_cols_tma<8>(p, nfr, wst) : (s.R1 == 16) ? fft_launch_cols_tma<16>(p, nfr, wst)
                                                                                : fft_launch_cols_tma<32>(p, nfr, wst);
        if (rc == 1)  // register-staged predecessor (RCB_FFT_VARIANT=1, or no tensor-map encoder / misaligned input)
            rc = (s.R1 == 8) ? fft_launch_cols<8>(p, nfr, wst) : (s.R1 == 16) ? fft_launch_cols<16>(p, nfr, wst)
                                                                            : fft_launch_cols<32>(p, nfr, wst);
        if (rc) return rc;
        if (s.rows_k1)
            rc = (s.R2 == 8) ? fft_launch_rows_k1<8>(s, p, nfr, sl, wst, sm_count)
                             : (s.R2 == 16) ? fft_launch_rows_k1<16>(s, p, nfr, sl, wst, sm_count)
                                            : fft_launch_rows_k1<32>(s, p, nfr, sl, wst, sm_count);
        else
            rc = (s.R2 == 8) ? fft_launch_rows<8>(p, nfr, wst) : (s.R2 == 16) ? fft_launch_rows<16>(p, nfr, wst)
                                                                            : fft_launch_rows<32>(p, nfr, wst);
        if (rc) return rc;
        const bool complete = (s.in_block + nfr == s.avg);
        float* emit = complete ? d_out + emitted * L : nullptr;
        // the block sum is accumulated in stream order of the sub-batches: wait for the previous fold
        if (last_slot >= 0 && last_slot != sl) FCK(cudaStreamWaitEvent(wst, s.ev_fold[last_slot], 0));
        fft_fold_kernel<<<(unsigned)((L / 4 + 255) / 256), 256, 0, wst>>>(s.d_vals2[sl], nfr, s.L, s.d_acc, emit);
        FCK(cudaGetLastError());
        FCK(cudaEventRecord(s.ev_fold[sl], wst));
        last_slot = sl;
        *launches += 3;
        s.in_block += nfr;
        if (complete) {
            s.in_block = 0;
            ++emitted;
        }
        done += nfr;
    }
    // rejoin the caller's stream
    if (last_slot >= 0) {
        FCK(cudaStreamWaitEvent(st, s.ev_fold[0], 0));
        FCK(cudaStreamWaitEvent(st, s.ev_fold[1], 0));
    }
    if (out_mem == 0 && emitted) {
        FCK(cudaMemcpyAsync(out, s.d_emit, emitted * L * sizeof(float), cudaMemcpyDeviceToHost, st));
        *d2h += emitted * L * sizeof(float);
    }
    if (in_mem == 0 || out_mem == 0) FCK(cudaStreamSynchronize(st));
    *nvec = emitted;
    return 0;
}
#undef FCK

}  // namespace rcb
