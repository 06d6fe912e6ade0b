// In-register radix-R FFT on packed f32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2).
//
// The K1 kernel is instruction-issue bound (profiles/r01_pfb_fm_tma_v5_summary.txt: 73 thread
// instructions per sample, issue slots 67 % busy, FP pipes ~45 %).  Blackwell's packed fp32x2
// instructions do two independent fp32 operations on an aligned register pair per issue slot, with
// an immediate / scalar multiplier broadcast to both halves.  A radix-2 FFT packs perfectly when the
// pair holds the SAME element of TWO sub-transforms that share every twiddle:
//
//   X[2q]   = sum_i  a_i W_{R/2}^{iq},   a_i =  v[i] + v[i+R/2]
//   X[2q+1] = sum_i  b_i W_{R/2}^{iq},   b_i = (v[i] - v[i+R/2]) W_R^i          (decimation in frequency)
//
// so one thread runs ONE scalar DIF stage (which also converts the interleaved re/im samples it reads
// from shared memory into the pair layout for free, and in the first pass absorbs the polyphase tap
// multiply into its FMAs) followed by a complete R/2-point transform on pairs
//   pr[i] = (Re a_i, Re b_i),  pi[i] = (Im a_i, Im b_i)
// in which every butterfly - trivial, +-j, (1+-j)/sqrt2 or general - costs what ONE scalar butterfly
// cost before.  R = 32: 16*4 + 14*4 scalar + 148 packed = 268 instructions instead of 388
// (first pass incl. taps: 300 instead of 452).  Outputs stay packed: (X[2q], X[2q+1]).
#pragma once
#include "fft_inreg.cuh"

namespace rcb {

__device__ __forceinline__ float2 p2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 p2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 p2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 p2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// a * s + c with the scalar s broadcast to both halves (immediate form when s is a constant)
__device__ __forceinline__ float2 p2fmas(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 p2muls(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 p2neg(float2 a) { return make_float2(-a.x, -a.y); }

// one DIT butterfly on pairs (ar, ai) +- w (br, bi), w = e^{SIGN*j*2*pi*K/LEN} - same cases as butterfly<>
template <int K, int LEN, int SIGN>
__device__ __forceinline__ void butterfly_p2(float2& ar, float2& ai, float2& br, float2& bi) {
    if constexpr (K == 0) {
        const float2 tr = br, ti = bi;
        br = p2sub(ar, tr);
        bi = p2sub(ai, ti);
        ar = p2add(ar, tr);
        ai = p2add(ai, ti);
    } else if constexpr (4 * K == LEN) {  // w = SIGN*j : w*b = SIGN*(-b.im, b.re)
        const float2 tr = br, ti = bi;
        if constexpr (SIGN > 0) {
            br = p2add(ar, ti);
            bi = p2sub(ai, tr);
            ar = p2sub(ar, ti);
            ai = p2add(ai, tr);
        } else {
            br = p2sub(ar, ti);
            bi = p2add(ai, tr);
            ar = p2add(ar, ti);
            ai = p2sub(ai, tr);
        }
    } else if constexpr (8 * K == LEN || 8 * K == 3 * LEN) {
        constexpr float h = 0.70710678118654752440f;
        constexpr bool cpos = Tw<K, LEN, SIGN>::c > 0, spos = Tw<K, LEN, SIGN>::s > 0;
        // d0 = cs*b.re - ss*b.im, d1 = ss*b.re + cs*b.im   (cs, ss = +-1: negations fold into FADD2)
        const float2 cbr = cpos ? br : p2neg(br), cbi = cpos ? bi : p2neg(bi);
        const float2 sbr = spos ? br : p2neg(br), sbi = spos ? bi : p2neg(bi);
        const float2 d0 = p2sub(cbr, sbi);
        const float2 d1 = p2add(sbr, cbi);
        br = p2fmas(d0, -h, ar);
        bi = p2fmas(d1, -h, ai);
        ar = p2fmas(d0, h, ar);
        ai = p2fmas(d1, h, ai);
    } else {
        constexpr float c = Tw<K, LEN, SIGN>::c;
        constexpr float s = Tw<K, LEN, SIGN>::s;
        const float2 ox = p2fmas(br, c, p2fmas(bi, -s, ar));
        const float2 oy = p2fmas(br, s, p2fmas(bi, c, ai));
        br = p2fmas(ar, 2.0f, p2neg(ox));
        bi = p2fmas(ai, 2.0f, p2neg(oy));
        ar = ox;
        ai = oy;
    }
}

template <int H, int LEN, int SIGN, int IDX>
__device__ __forceinline__ void stage_one_p2(float2 (&wr)[H], float2 (&wi)[H]) {
    constexpr int half = LEN / 2;
    constexpr int blk = IDX / half;
    constexpr int k = IDX % half;
    constexpr int i0 = blk * LEN + k;
    butterfly_p2<k, LEN, SIGN>(wr[i0], wi[i0], wr[i0 + half], wi[i0 + half]);
}
template <int H, int LEN, int SIGN, int... I>
__device__ __forceinline__ void stage_all_p2(float2 (&wr)[H], float2 (&wi)[H], std::integer_sequence<int, I...>) {
    (stage_one_p2<H, LEN, SIGN, I>(wr, wi), ...);
}
template <int H, int LEN, int SIGN>
__device__ __forceinline__ void stages_p2(float2 (&wr)[H], float2 (&wi)[H]) {
    if constexpr (LEN <= H) {
        stage_all_p2<H, LEN, SIGN>(wr, wi, std::make_integer_sequence<int, H / 2>{});
        stages_p2<H, LEN * 2, SIGN>(wr, wi);
    }
}
template <int H, int... I>
__device__ __forceinline__ void brev_copy_p2(float2 (&dr)[H], float2 (&di)[H], const float2 (&sr)[H],
                                             const float2 (&si)[H], std::integer_sequence<int, I...>) {
    ((dr[I] = sr[BrevIdx<I, H>::value], di[I] = si[BrevIdx<I, H>::value]), ...);
}

// H-point transform over i of the pairs (pr[i], pi[i]); natural order in and out
template <int H, int SIGN>
__device__ __forceinline__ void fft_pairs(float2 (&pr)[H], float2 (&pi)[H]) {
    float2 wr[H], wi[H];
    brev_copy_p2<H>(wr, wi, pr, pi, std::make_integer_sequence<int, H>{});
    stages_p2<H, 2, SIGN>(wr, wi);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        pr[i] = wr[i];
        pi[i] = wi[i];
    }
}

// b = d * W_R^{SIGN*I} for the scalar DIF stage (special cases like butterfly<>)
template <int I, int R, int SIGN>
__device__ __forceinline__ float2 dif_twiddle(float2 d) {
    if constexpr (I == 0) {
        return d;
    } else if constexpr (4 * I == R) {
        return (SIGN > 0) ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    } else if constexpr (8 * I == R || 8 * I == 3 * R) {
        constexpr float h = 0.70710678118654752440f;
        constexpr float cs = Tw<I, R, SIGN>::c > 0 ? h : -h;
        constexpr float ss = Tw<I, R, SIGN>::s > 0 ? h : -h;
        // (c + j s)(x + j y) = (c x - s y) + j (s x + c y),  |c| = |s| = h
        return make_float2(fmaf(cs, d.x, -ss * d.y), fmaf(ss, d.x, cs * d.y));
    } else {
        constexpr float c = Tw<I, R, SIGN>::c;
        constexpr float s = Tw<I, R, SIGN>::s;
        return make_float2(fmaf(c, d.x, -s * d.y), fmaf(s, d.x, c * d.y));
    }
}

// Scalar DIF stage.  GET(j) returns the j-th input (float2); TAP(j) its real tap (FIR = true) -
// both are called with compile-time j.  Fills pr/pi.
template <int R, int SIGN, bool FIR, int I, class GET, class TAP>
__device__ __forceinline__ void dif_one(float2 (&pr)[R / 2], float2 (&pi)[R / 2], GET& get, TAP& tap) {
    constexpr int H = R / 2;
    const float2 x0 = get(std::integral_constant<int, I>{});
    const float2 x1 = get(std::integral_constant<int, I + H>{});
    float2 a, d;
    if constexpr (FIR) {
        const float h0 = tap(std::integral_constant<int, I>{});
        const float h1 = tap(std::integral_constant<int, I + H>{});
        const float tr = h0 * x0.x, ti = h0 * x0.y;
        a = make_float2(fmaf(h1, x1.x, tr), fmaf(h1, x1.y, ti));
        d = make_float2(fmaf(-h1, x1.x, tr), fmaf(-h1, x1.y, ti));
    } else {
        a = make_float2(x0.x + x1.x, x0.y + x1.y);
        d = make_float2(x0.x - x1.x, x0.y - x1.y);
    }
    const float2 b = dif_twiddle<I, R, SIGN>(d);
    pr[I] = make_float2(a.x, b.x);
    pi[I] = make_float2(a.y, b.y);
}
template <int R, int SIGN, bool FIR, class GET, class TAP, int... I>
__device__ __forceinline__ void dif_all(float2 (&pr)[R / 2], float2 (&pi)[R / 2], GET& get, TAP& tap,
                                        std::integer_sequence<int, I...>) {
    (dif_one<R, SIGN, FIR, I>(pr, pi, get, tap), ...);
}

// X[k] = sum_j in(j) [* tap(j)] e^{SIGN j 2 pi j k / R};  out: (pr[q], pi[q]) = (X[2q], X[2q+1]) as pairs
template <int R, int SIGN, bool FIR, class GET, class TAP>
__device__ __forceinline__ void fft_packed(float2 (&pr)[R / 2], float2 (&pi)[R / 2], GET get, TAP tap) {
    dif_all<R, SIGN, FIR>(pr, pi, get, tap, std::make_integer_sequence<int, R / 2>{});
    fft_pairs<R / 2, SIGN>(pr, pi);
}

// atan2 for two points at once: (y, x) = (im.x, re.x) and (im.y, re.y).  Same octant fold and degree-6
// minimax polynomial as atan2_nan (pfb_fm.cuh); the polynomial runs on the pair.  (0,0) -> NaN.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 atan2_nan_p2(float2 im, float2 re) {
    const float ax0 = fabsf(re.x), ay0 = fabsf(im.x), ax1 = fabsf(re.y), ay1 = fabsf(im.y);
    const float2 mn = make_float2(fminf(ax0, ay0), fminf(ax1, ay1));
    const float2 rc = make_float2(rcp_approx(fmaxf(ax0, ay0)), rcp_approx(fmaxf(ax1, ay1)));
    const float2 z = p2mul(mn, rc);
    const float2 s = p2mul(z, z);
    float2 r = p2fmas(s, -0.004370174370706081f, make_float2(0.023092154413461685f, 0.023092154413461685f));
    r = p2fma(r, s, make_float2(-0.05784549191594124f, -0.05784549191594124f));
    r = p2fma(r, s, make_float2(0.0979914739727974f, 0.0979914739727974f));
    r = p2fma(r, s, make_float2(-0.13978290557861328f, -0.13978290557861328f));
    r = p2fma(r, s, make_float2(0.1996297985315323f, 0.1996297985315323f));
    r = p2fma(r, s, make_float2(-0.33331674337387085f, -0.33331674337387085f));
    r = p2mul(r, s);
    r = p2fma(r, z, z);
    float r0 = r.x, r1 = r.y;
    r0 = (ay0 > ax0) ? (1.57079632679489661923f - r0) : r0;
    r1 = (ay1 > ax1) ? (1.57079632679489661923f - r1) : r1;
    r0 = (re.x < 0.0f) ? (3.14159265358979323846f - r0) : r0;
    r1 = (re.y < 0.0f) ? (3.14159265358979323846f - r1) : r1;
    return make_float2(copysignf(r0, im.x), copysignf(r1, im.y));
}

// same, with the reference's fast_atan2f(0, 0) = 0 (used on conj-products, where no NaN sentinel is needed)
__device__ __forceinline__ float2 atan2_zero_p2(float2 im, float2 re) {
    const float ax0 = fabsf(re.x), ay0 = fabsf(im.x), ax1 = fabsf(re.y), ay1 = fabsf(im.y);
    const float mx0 = fmaxf(ax0, ay0), mx1 = fmaxf(ax1, ay1);
    const float2 mn = make_float2(fminf(ax0, ay0), fminf(ax1, ay1));
    const float2 rc = make_float2(mx0 == 0.0f ? 0.0f : rcp_approx(mx0), mx1 == 0.0f ? 0.0f : rcp_approx(mx1));
    const float2 z = p2mul(mn, rc);
    const float2 s = p2mul(z, z);
    float2 r = p2fmas(s, -0.004370174370706081f, make_float2(0.023092154413461685f, 0.023092154413461685f));
    r = p2fma(r, s, make_float2(-0.05784549191594124f, -0.05784549191594124f));
    r = p2fma(r, s, make_float2(0.0979914739727974f, 0.0979914739727974f));
    r = p2fma(r, s, make_float2(-0.13978290557861328f, -0.13978290557861328f));
    r = p2fma(r, s, make_float2(0.1996297985315323f, 0.1996297985315323f));
    r = p2fma(r, s, make_float2(-0.33331674337387085f, -0.33331674337387085f));
    r = p2mul(r, s);
    r = p2fma(r, z, z);
    float r0 = r.x, r1 = r.y;
    r0 = (ay0 > ax0) ? (1.57079632679489661923f - r0) : r0;
    r1 = (ay1 > ax1) ? (1.57079632679489661923f - r1) : r1;
    r0 = (re.x < 0.0f) ? (3.14159265358979323846f - r0) : r0;
    r1 = (re.y < 0.0f) ? (3.14159265358979323846f - r1) : r1;
    return make_float2(copysignf(r0, im.x), copysignf(r1, im.y));
}

// Drop-in packed replacement of warp_fft_2pass (fft_inreg.cuh): N = R*R point DFT of one frame spread over
// R lanes, both radix-R passes on f32x2 pairs, same scratch / twiddle table layout (row stride R+2).
template <int R, int SIGN, bool REV>
__device__ __forceinline__ void warp_fft_2pass_packed(float2 (&v)[R], float2* __restrict__ buf,
                                                      const float2* __restrict__ tws, const int ll) {
    constexpr int S = R + 2;
    float2 pr[R / 2], pi[R / 2];
    {
        auto get = [&](auto j) { return v[decltype(j)::value]; };
        auto tap = [&](auto) { return 1.0f; };
        fft_packed<R, SIGN, false>(pr, pi, get, tap);  // (pr[c], pi[c]) = first-pass outputs m1 = 2c, 2c+1
    }
    {
        const float4* twp = reinterpret_cast<const float4*>(tws + ll * S);
        float4* bp = reinterpret_cast<float4*>(buf + ll * S);
#pragma unroll
        for (int c = 0; c < R / 2; ++c) {
            const float4 t = twp[c];
            const float2 b0 = make_float2(fmaf(pr[c].x, t.x, -pi[c].x * t.y), fmaf(pr[c].x, t.y, pi[c].x * t.x));
            const float2 b1 = make_float2(fmaf(pr[c].y, t.z, -pi[c].y * t.w), fmaf(pr[c].y, t.w, pi[c].y * t.z));
            bp[c] = make_float4(b0.x, b0.y, b1.x, b1.y);
        }
    }
    __syncwarp();
    {
        float2 u[R];
#pragma unroll
        for (int l2 = 0; l2 < R; ++l2) u[REV ? (R - 1 - l2) : l2] = buf[l2 * S + ll];
        auto get = [&](auto j) { return u[decltype(j)::value]; };
        auto tap = [&](auto) { return 1.0f; };
        fft_packed<R, SIGN, false>(pr, pi, get, tap);
    }
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {  // v[m2] = X[ll + R*m2]
        v[2 * q] = make_float2(pr[q].x, pi[q].x);
        v[2 * q + 1] = make_float2(pr[q].y, pi[q].y);
    }
    __syncwarp();
}

}  // namespace rcb
