// K3f  fft_frame : the reference's own scan shape - fft_vector.py:32 `length = 1024*16`, Blackman-Harris window,
// |.|^2 -> nlog10 -> moving sum of 100 frames (fft_vector.py:37-45) - with the WHOLE 16384-point frame resident in one
// SM's shared memory (128 KB of the 227 KB), one launch per call.
//
// The four-step pipeline of fft_logpow.cuh moves every frame through an L2-resident scratch between two kernels and
// folds a `vals` array in a third: for 16384-point frames that is 3 short launches per sub-batch (2 487 launches per
// 20 bench steps) at 0.08 of the HBM roofline.  Here a CTA (16 warps) owns a GROUP of consecutive frames:
//   load    16 bulk copies of 8 KB (cp.async.bulk, one per warp and row n1) fill the frame x[n1][n2], n = 1024 n1 + n2;
//   pass 1  thread t owns columns n2 = t, t + 512: 16-point FFT over n1 in registers (packed f32x2, the window multiply
//           folded into the scalar DIF stage), times W_16384^{-n2 k1} (powers of one table value, built pairwise by
//           a packed recurrence), written back IN PLACE as row k1;
//   pass 2  warp w owns row k1 = w: the 1024-point FFT of pfb_fm1_kernel (two packed radix-32 passes, in-place
//           swizzled transpose inside the row).  Once a warp has read its row for the second pass the row is dead and
//           the warp starts the bulk copy of the NEXT frame's row into it: the load of frame f + 1 runs under the second
//           radix pass, log-power and accumulation of frame f; one mbarrier (16 arrivals + 128 KB) is both the data
//           barrier and the CTA barrier between pass 2 of frame f and pass 1 of frame f + 1;
//   power   |X|^2 on pairs, log10(max(p, 1e-18)) + 1 (gr::blocks::nlog10_ff(1, L, 1)), added to the group's sums
//           (64 KB of shared memory, conflict free).
// When the group ends the sums leave coalesced, fftshifted (fft_vcc shift = True).
//
// Determinism / streaming: groups are cut relative to the position inside the averaging block (G frames each, the last
// one shorter), so the association of the sum - frame order inside a group, group order inside a block
// (fft_fold_groups_kernel) - does not depend on how the stream is split into calls; a group cut by the end of a call
// parks its sums in a carry row and the next call's first group starts from it.
#pragma once
#include "fft_logpow.cuh"  // FftState, fft_packed, mbarrier / bulk-copy helpers, pfb_swz

namespace rcb {

struct FftFrameParams {
    const float2* x;        // frames of this call, frame f at x + f * 16384 (16-byte aligned)
    const float* window;    // [16384]
    const float2* tw;       // [1024] dense swizzled W_1024^{-ll m1} (pass-2 inter-radix twiddles)
    const float2* t_lo;     // [1024] W_16384^{-q}
    float* partial;         // [ngroups][16384] sums of the groups that complete in this call (output bin order)
    float* carry;           // [16384] sums of a group cut by the end of a call
    int nframes;            // frames in this call
    int pos0;               // frames of the current averaging block consumed before this call
    int avg;                // frames per block
    int G;                  // frames per group
    int gpb;                // groups per block = ceil(avg / G)
};

struct FftFrameGeom {
    static constexpr int L = 16384, L1 = 16, L2 = 1024, R = 32, THREADS = 512, WARPS = 16;
    static constexpr size_t frame_bytes = (size_t)L * 8;
    static constexpr size_t acc_bytes = (size_t)L * 4;
    static constexpr size_t tw_bytes = (size_t)L2 * 8;
    static constexpr size_t off_acc = frame_bytes;
    static constexpr size_t off_tw = off_acc + acc_bytes;
    static constexpr size_t off_bar = off_tw + tw_bytes;
    static constexpr size_t smem_bytes = off_bar + 64;
};

// call-relative frame range [f0, f1) of group gi of the call, whether it continues a parked group / ends complete
__host__ __device__ inline void fft_frame_group(int gi, int nframes, int pos0, int avg, int G, int gpb, int* f0, int* f1,
                                                bool* from_carry, bool* complete) {
    const int v = pos0 / G + gi;          // virtual group index, counted from the start of the call's first block
    const int b = v / gpb, j = v - b * gpb;
    const long long lo = (long long)b * avg + (long long)j * G - pos0;
    const int gend = (j + 1) * G < avg ? (j + 1) * G : avg;
    const long long hi = (long long)b * avg + gend - pos0;
    *f0 = lo < 0 ? 0 : (int)lo;
    *f1 = hi > nframes ? nframes : (int)hi;
    *from_carry = (gi == 0 && pos0 % G != 0 && pos0 % avg != 0);
    *complete = (hi <= nframes);
}
// groups a call of nframes touches
inline int fft_frame_ngroups(int nframes, int pos0, int avg, int G, int gpb) {
    if (nframes <= 0) return 0;
    const long long last = (long long)pos0 + nframes - 1;  // block-0-relative index of the last frame
    const long long b = last / avg, r = last - b * avg;
    const long long vlast = b * gpb + r / G;
    return (int)(vlast - pos0 / G + 1);
}

__global__ void __launch_bounds__(512, 1) fft_frame_kernel(const FftFrameParams p) {
    using Gm = FftFrameGeom;
    constexpr int L2 = Gm::L2, R = Gm::R, W = Gm::WARPS;
    extern __shared__ __align__(128) unsigned char smem_ff[];
    float2* frame = reinterpret_cast<float2*>(smem_ff);                 // [16][1024]
    float* acc = reinterpret_cast<float*>(smem_ff + Gm::off_acc);       // [k1][k2]
    float2* tws = reinterpret_cast<float2*>(smem_ff + Gm::off_tw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_ff + Gm::off_bar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ll = lane;

    int f0, f1;
    bool from_carry, complete;
    fft_frame_group((int)blockIdx.x, p.nframes, p.pos0, p.avg, p.G, p.gpb, &f0, &f1, &from_carry, &complete);
    if (f0 >= f1) return;

    if (tid == 0) {
        mbar_init(full, W);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float2* wf = frame + warp * L2;   // pass 2: this warp's row k1 = warp; load: row n1 = warp
    auto issue_row = [&](int f) {
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(full, (uint32_t)(L2 * 8));
            tma_bulk_g2s(wf, p.x + (size_t)f * Gm::L + (size_t)warp * L2, (uint32_t)(L2 * 8), full);
        }
    };
    issue_row(f0);
    for (int i = tid; i < L2; i += Gm::THREADS) tws[i] = p.tw[i];
    // the group's sums: zero, or the parked sums of the group's earlier frames (carry is in output bin order)
    if (from_carry) {
        for (int i = tid; i < Gm::L; i += Gm::THREADS) {
            const int k1 = i >> 10, k2 = i & 1023;
            acc[i] = __ldcg(p.carry + k1 + 16 * ((k2 + 512) & 1023));
        }
    } else {
        for (int i = tid; i < Gm::L; i += Gm::THREADS) acc[i] = 0.f;
    }
    __syncthreads();

    uint32_t par = 0;
#pragma unroll 1
    for (int f = f0; f < f1; ++f) {
        mbar_wait(full, par);
        par ^= 1u;
        // ---- pass 1: 16-point FFTs over n1 (stride 1024), window folded in, twiddle, in place ----
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int n2 = tid + 512 * c;
            float wv[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) wv[a] = __ldg(p.window + a * L2 + n2);
            float2 pr[8], pi[8];
            {
                auto get = [&](auto j) { return frame[decltype(j)::value * L2 + n2]; };
                auto tap = [&](auto j) { return wv[decltype(j)::value]; };
                fft_packed<16, -1, true>(pr, pi, get, tap);  // (pr[q], pi[q]) = A[k1 = 2q], A[2q + 1]
            }
            // W^{-n2 k1} for the pair (2q, 2q + 1): (W^0, W^1) advanced by W^2 each step; W^1 = W_L^{-n2} from the table
            const float2 w1 = __ldg(p.t_lo + n2);
            const float c1 = w1.x, s1 = w1.y;
            const float c2 = fmaf(c1, c1, -s1 * s1), s2 = 2.0f * c1 * s1;
            float2 twr = make_float2(1.0f, c1), twi = make_float2(0.0f, s1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 orr = p2fma(pr[q], twr, p2neg(p2mul(pi[q], twi)));
                const float2 oii = p2fma(pr[q], twi, p2mul(pi[q], twr));
                frame[(2 * q) * L2 + n2] = make_float2(orr.x, oii.x);
                frame[(2 * q + 1) * L2 + n2] = make_float2(orr.y, oii.y);
                if (q < 7) {
                    const float2 nr = p2fmas(twr, c2, p2muls(twi, -s2));
                    const float2 ni = p2fmas(twr, s2, p2muls(twi, c2));
                    twr = nr;
                    twi = ni;
                }
            }
        }
        __syncthreads();
        // ---- pass 2: row k1 = warp, 1024-point FFT in place (pfb_fm1_kernel's two radix-32 passes, forward sign) ----
        float2 pr[R / 2], pi[R / 2];
        {
            auto get = [&](auto j) { return wf[decltype(j)::value * R + ll]; };
            auto tap = [&](auto) { return 1.0f; };
            fft_packed<R, -1, false>(pr, pi, get, tap);
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
            for (int l2 = 0; l2 < R; ++l2) u[l2] = wf[l2 * R + (((ch ^ pfb_swz<R>(l2)) << 1) | wi)];
            __syncwarp();  // the row is dead: the next frame's row n1 = warp streams in under the rest of this frame
            if (f + 1 < f1) issue_row(f + 1);
            auto get = [&](auto j) { return u[decltype(j)::value]; };
            auto tap = [&](auto) { return 1.0f; };
            fft_packed<R, -1, false>(pr, pi, get, tap);  // (pr[q], pi[q]) = X[k1 + 16 k2], k2 = ll + 32 (2q), ll + 32 (2q + 1)
        }
        // ---- |X|^2 -> log10 + 1 -> group sums (lg2.approx: the argument is clamped to 1e-18, far from denormals;
        //      absolute error 2^-22 in log2, 1e-7 in the value) ----
        float* arow = acc + warp * L2 + ll;
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            const float2 pw = p2fma(pr[q], pr[q], p2mul(pi[q], pi[q]));
            const float v0 = fmaf(__log2f(fmaxf(pw.x, 1e-18f)), 0.30102999566398120f, 1.0f);
            const float v1 = fmaf(__log2f(fmaxf(pw.y, 1e-18f)), 0.30102999566398120f, 1.0f);
            arow[(2 * q) * R] += v0;
            arow[(2 * q + 1) * R] += v1;
        }
    }
    __syncthreads();
    // ---- the group's sums leave in output order: bin ks = k1 + 16 ((k2 + 512) mod 1024)  (fftshift) ----
    float* dst = complete ? p.partial + (size_t)blockIdx.x * Gm::L : p.carry;
#pragma unroll
    for (int i = 0; i < Gm::L / 4 / Gm::THREADS; ++i) {
        const int i4 = i * Gm::THREADS + tid;
        const int ks0 = 4 * i4;
        const int k1 = ks0 & 15, k2 = ((ks0 >> 4) + 512) & 1023;
        float4 o;
        o.x = acc[(k1 + 0) * L2 + k2];
        o.y = acc[(k1 + 1) * L2 + k2];
        o.z = acc[(k1 + 2) * L2 + k2];
        o.w = acc[(k1 + 3) * L2 + k2];
        __stcg(reinterpret_cast<float4*>(dst) + i4, o);
    }
}

// Block sums from the group sums of one call.  Block b of the call (b = 0 is the block the call starts in) owns the
// virtual groups [b gpb, (b + 1) gpb); the complete ones among them are rows of `partial`.  acc_in carries the block the
// previous call left open (b = 0 starts from it), acc_out receives the block this call leaves open (two buffers: the
// rows of the grid run concurrently); complete blocks go to emit[b].  grid (L / 1024, nblocks), block 256
__global__ void __launch_bounds__(256) fft_fold_groups_kernel(const float* __restrict__ partial,
                                                              const float* __restrict__ acc_in, float* __restrict__ acc_out,
                                                              float* __restrict__ emit, int L, int nframes, int pos0,
                                                              int avg, int G, int gpb, int ngroups) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (k >= L) return;
    const int b = blockIdx.y;
    const int v0 = pos0 / G;
    // groups of the call inside block b: gi = v - v0 for v in [b gpb, (b + 1) gpb), gi in [0, ngroups)
    int ga = b * gpb - v0, gb = (b + 1) * gpb - v0;
    if (ga < 0) ga = 0;
    if (gb > ngroups) gb = ngroups;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b == 0 && pos0 >= G) a = *reinterpret_cast<const float4*>(acc_in + k);  // complete groups folded by earlier calls
    bool block_done = false;
    for (int gi = ga; gi < gb; ++gi) {
        int f0, f1;
        bool fc, complete;
        fft_frame_group(gi, nframes, pos0, avg, G, gpb, &f0, &f1, &fc, &complete);
        if (!complete) break;  // parked in the carry row: folded by the call that completes it
        const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)gi * L + k));
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
        if (gi + v0 == (b + 1) * gpb - 1) block_done = true;
    }
    if (block_done) {
        *reinterpret_cast<float4*>(emit + (size_t)b * L + k) = a;
    } else {
        *reinterpret_cast<float4*>(acc_out + k) = a;  // the call's last, open block (another CTA row may still read acc_in)
    }
}

#define FCK(call)                              \
    do {                                       \
        if ((call) != cudaSuccess) return -3;  \
    } while (0)

// One launch of fft_frame_kernel + one of fft_fold_groups_kernel per chunk of frames.  Host input - and device input
// the bulk copies cannot address (not 16-byte aligned) - is staged through two buffers, the copy of chunk c + 1 running
// under the kernels of chunk c.
inline int fft_process_frames(FftState& s, const float2* iq, size_t nsamples, int in_mem, float* out, size_t cap_vec,
                              int out_mem, size_t* nvec, cudaStream_t st, uint64_t* launches, uint64_t* h2d, uint64_t* d2h) {
    using Gm = FftFrameGeom;
    *nvec = 0;
    const size_t L = (size_t)s.L;
    const size_t nframes = nsamples / L;
    const bool staged = (in_mem == 0) || (reinterpret_cast<uintptr_t>(iq) & 15);
    const size_t will_emit = (s.in_block + nframes) / (size_t)s.avg;
    if (will_emit > cap_vec || (will_emit && !out)) return -6;  // RCB_ERANGE
    static bool attr_dev[64] = {};
    bool& attr = attr_dev[fft_cur_device()];
    if (!attr) {
        FCK(cudaFuncSetAttribute(fft_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::smem_bytes));
        attr = true;
    }
    // chunks: device input is cut only to bound the group-sum buffer
    const size_t chunk = staged ? (size_t)1024 : (size_t)8192;
    if (staged) {
        const size_t need = std::min(chunk, std::max<size_t>(nframes, 1)) * L;
        for (int i = 0; i < 2; ++i) {
            if (s.in_cap2[i] < need) {
                FCK(cudaStreamSynchronize(st));
                cudaFree(s.d_in2[i]);
                s.d_in2[i] = nullptr;
                s.in_cap2[i] = 0;
                FCK(cudaMalloc(&s.d_in2[i], need * sizeof(float2)));
                s.in_cap2[i] = need;
            }
            if (!s.ev_copy[i]) FCK(cudaEventCreateWithFlags(&s.ev_copy[i], cudaEventDisableTiming));
            if (!s.ev_used[i]) FCK(cudaEventCreateWithFlags(&s.ev_used[i], cudaEventDisableTiming));
        }
        // the copies run on ws[0]: ordered after everything queued on the caller's stream so far
        FCK(cudaEventRecord(s.ev_start, st));
        FCK(cudaStreamWaitEvent(s.ws[0], s.ev_start, 0));
    }
    if (out_mem == 0 && will_emit && s.emit_cap < will_emit * L) {
        FCK(cudaStreamSynchronize(st));
        cudaFree(s.d_emit);
        s.d_emit = nullptr;
        s.emit_cap = 0;
        FCK(cudaMalloc(&s.d_emit, will_emit * L * sizeof(float)));
        s.emit_cap = will_emit * L;
    }
    float* d_out = (out_mem == 0) ? s.d_emit : out;
    size_t done = 0, emitted = 0;
    int k = 0;
    while (done < nframes) {
        const int nfr = (int)std::min(chunk, nframes - done);
        const float2* d_x = iq + done * L;
        if (staged) {
            const int sl = k & 1;
            if (k >= 2) FCK(cudaStreamWaitEvent(s.ws[0], s.ev_used[sl], 0));
            FCK(cudaMemcpyAsync(s.d_in2[sl], iq + done * L, (size_t)nfr * L * sizeof(float2),
                                in_mem == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s.ws[0]));
            if (in_mem == 0) *h2d += (size_t)nfr * L * sizeof(float2);
            FCK(cudaEventRecord(s.ev_copy[sl], s.ws[0]));
            FCK(cudaStreamWaitEvent(st, s.ev_copy[sl], 0));
            d_x = s.d_in2[sl];
        }
        const int pos0 = s.in_block;
        const int ngroups = fft_frame_ngroups(nfr, pos0, s.avg, s.fr_G, s.fr_gpb);
        if (s.partial_rows < (size_t)ngroups) {
            FCK(cudaStreamSynchronize(st));
            cudaFree(s.d_partial);
            s.d_partial = nullptr;
            s.partial_rows = 0;
            const size_t rows = (size_t)ngroups + 16;
            FCK(cudaMalloc(&s.d_partial, rows * L * sizeof(float)));
            s.partial_rows = rows;
        }
        FftFrameParams p{};
        p.x = d_x;
        p.window = s.d_window;
        p.tw = s.d_tw_frame;
        p.t_lo = s.d_tlo;
        p.partial = s.d_partial;
        p.carry = s.d_carry;
        p.nframes = nfr;
        p.pos0 = pos0;
        p.avg = s.avg;
        p.G = s.fr_G;
        p.gpb = s.fr_gpb;
        fft_frame_kernel<<<(unsigned)ngroups, Gm::THREADS, Gm::smem_bytes, st>>>(p);
        FCK(cudaGetLastError());
        const int nblocks = (int)((pos0 + (size_t)nfr + s.avg - 1) / (size_t)s.avg);   // blocks the chunk touches
        dim3 fgrid((unsigned)(L / 1024), (unsigned)nblocks);
        fft_fold_groups_kernel<<<fgrid, 256, 0, st>>>(s.d_partial, s.d_acc, s.d_acc_alt, d_out + emitted * L, s.L, nfr, pos0,
                                                     s.avg, s.fr_G, s.fr_gpb, ngroups);
        FCK(cudaGetLastError());
        std::swap(s.d_acc, s.d_acc_alt);
        if (staged) FCK(cudaEventRecord(s.ev_used[k & 1], st));
        *launches += 2;
        emitted += (pos0 + (size_t)nfr) / (size_t)s.avg;
        s.in_block = (int)((pos0 + (size_t)nfr) % (size_t)s.avg);
        done += nfr;
        ++k;
    }
    if (out_mem == 0 && emitted) {
        FCK(cudaMemcpyAsync(out, s.d_emit, emitted * L * sizeof(float), cudaMemcpyDeviceToHost, st));
        *d2h += emitted * L * sizeof(float);
    }
    if (staged || out_mem == 0) FCK(cudaStreamSynchronize(st));
    *nvec = emitted;
    return 0;
}
#undef FCK

}  // namespace rcb
