// In-register radix-2 DIT FFT of R = 2..64 complex points held by ONE thread.
// Everything is unrolled at compile time: twiddles are immediates, the bit-reversal is a register
// renaming, trivial (1, +-j) and semi-trivial ((+-1+-j)/sqrt2) twiddles are special-cased, and a
// general butterfly costs 6 FMA-class instructions (out1 = 2a - out0).
//   R = 32: 46 trivial (4 op) + 14 semi (6 op) + 20 general (6 op) = 388 instr = 12.1 / point.
#pragma once
#include <utility>
#include <cuda_runtime.h>

namespace rcb {

// ---- compile-time trig (double Taylor, evaluated only in constant expressions) ----
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double ct_sin(double x) {
    double term = x, sum = x;
    for (int n = 1; n < 40; ++n) {
        term *= -x * x / ((2.0 * n) * (2.0 * n + 1.0));
        sum += term;
    }
    return sum;
}
constexpr double ct_cos(double x) {
    double term = 1.0, sum = 1.0;
    for (int n = 1; n < 40; ++n) {
        term *= -x * x / ((2.0 * n - 1.0) * (2.0 * n));
        sum += term;
    }
    return sum;
}
// reduce k/n turn to (-1/2, 1/2] before evaluating to keep the series short and accurate
constexpr double ct_turn(int k, int n) {
    int kk = k % n;
    if (kk < 0) kk += n;
    if (2 * kk > n) kk -= n;
    return 2.0 * kPi * (double)kk / (double)n;
}

template <int K, int N, int SIGN>
struct Tw {
    static constexpr float c = (float)ct_cos(ct_turn(K, N));
    static constexpr float s = (float)((double)SIGN * ct_sin(ct_turn(K, N)));
};

constexpr int ct_brev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
constexpr int ct_log2(int x) { return x <= 1 ? 0 : 1 + ct_log2(x >> 1); }

// one DIT butterfly with twiddle e^{SIGN*j*2*pi*K/LEN} applied to the second input
template <int K, int LEN, int SIGN>
__device__ __forceinline__ void butterfly(float2& a, float2& b) {
    if constexpr (K == 0) {
        const float2 t = b;
        b = make_float2(a.x - t.x, a.y - t.y);
        a = make_float2(a.x + t.x, a.y + t.y);
    } else if constexpr (4 * K == LEN) {  // w = SIGN*j : t = (-S*b.y, S*b.x)
        const float2 t = b;
        if constexpr (SIGN > 0) {
            b = make_float2(a.x + t.y, a.y - t.x);
            a = make_float2(a.x - t.y, a.y + t.x);
        } else {
            b = make_float2(a.x - t.y, a.y + t.x);
            a = make_float2(a.x + t.y, a.y - t.x);
        }
    } else if constexpr (8 * K == LEN || 8 * K == 3 * LEN) {
        constexpr float h = 0.70710678118654752440f;
        // w = (c + j s), |c| = |s| = h.  t = b*w
        constexpr float cs = Tw<K, LEN, SIGN>::c > 0 ? 1.0f : -1.0f;
        constexpr float ss = Tw<K, LEN, SIGN>::s > 0 ? 1.0f : -1.0f;
        // t.x = h*(cs*b.x - ss*b.y), t.y = h*(ss*b.x + cs*b.y)
        const float d0 = cs * b.x - ss * b.y;   // sign multiplies fold into FADD negations
        const float d1 = ss * b.x + cs * b.y;
        b = make_float2(fmaf(-d0, h, a.x), fmaf(-d1, h, a.y));
        a = make_float2(fmaf(d0, h, a.x), fmaf(d1, h, a.y));
    } else {
        constexpr float c = Tw<K, LEN, SIGN>::c;
        constexpr float s = Tw<K, LEN, SIGN>::s;
        const float ox = fmaf(b.x, c, fmaf(-b.y, s, a.x));
        const float oy = fmaf(b.x, s, fmaf(b.y, c, a.y));
        b = make_float2(fmaf(2.0f, a.x, -ox), fmaf(2.0f, a.y, -oy));
        a = make_float2(ox, oy);
    }
}

template <int R, int LEN, int SIGN, int IDX>
__device__ __forceinline__ void stage_one(float2 (&w)[R]) {
    constexpr int half = LEN / 2;
    constexpr int blk = IDX / half;
    constexpr int k = IDX % half;
    constexpr int i0 = blk * LEN + k;
    butterfly<k, LEN, SIGN>(w[i0], w[i0 + half]);
}
template <int R, int LEN, int SIGN, int... I>
__device__ __forceinline__ void stage_all(float2 (&w)[R], std::integer_sequence<int, I...>) {
    (stage_one<R, LEN, SIGN, I>(w), ...);
}
template <int R, int LEN, int SIGN>
__device__ __forceinline__ void stages(float2 (&w)[R]) {
    if constexpr (LEN <= R) {
        stage_all<R, LEN, SIGN>(w, std::make_integer_sequence<int, R / 2>{});
        stages<R, LEN * 2, SIGN>(w);
    }
}

template <int I, int R>
struct BrevIdx {
    static constexpr int value = ct_brev(I, ct_log2(R));
};
template <int R, int... I>
__device__ __forceinline__ void brev_copy(float2 (&dst)[R], const float2 (&src)[R],
                                          std::integer_sequence<int, I...>) {
    ((dst[I] = src[BrevIdx<I, R>::value]), ...);
}

// X[k] = sum_j v[j] e^{SIGN * j 2 pi j k / R}, natural order in and out, unnormalised.
template <int R, int SIGN>
__device__ __forceinline__ void fft_inreg(float2 (&v)[R]) {
    float2 w[R];
    brev_copy<R>(w, v, std::make_integer_sequence<int, R>{});
    stages<R, 2, SIGN>(w);
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = w[i];
}

}  // namespace rcb

namespace rcb {

// N = R*R point DFT of one "frame" spread over R lanes (lane ll holds points R*j + l, j = 0..R-1, with
// l = ll, or l = R-1-ll when REV), as two in-register radix-R passes with the W_N twiddle and a
// warp-private shared-memory transpose in between.
//   in : v[j]                      out: v[m2] = X[ll + R*m2]
//   buf: this frame's private R*(R+2) complex scratch (row stride R+2 => STS.128 / LDS.64 conflict free)
//   tws: shared twiddle table tws[ll*(R+2) + m1] = W_N^{SIGN * l * m1}
// Only __syncwarp() is used; buf may be reused by the warp as soon as the call returns.
// second half: twiddle, transpose through buf, second radix-R pass (first pass already done in v)
template <int R, int SIGN, bool REV>
__device__ __forceinline__ void warp_fft_xpose_pass2(float2 (&v)[R], float2* __restrict__ buf,
                                                     const float2* __restrict__ tws, const int ll) {
    constexpr int S = R + 2;
    {
        const float4* twp = reinterpret_cast<const float4*>(tws + ll * S);
        float4* bp = reinterpret_cast<float4*>(buf + ll * S);
#pragma unroll
        for (int m1 = 0; m1 < R; m1 += 2) {
            const float4 t = twp[m1 >> 1];
            const float2 b0 = make_float2(fmaf(v[m1].x, t.x, -v[m1].y * t.y), fmaf(v[m1].x, t.y, v[m1].y * t.x));
            const float2 b1 = make_float2(fmaf(v[m1 + 1].x, t.z, -v[m1 + 1].y * t.w),
                                          fmaf(v[m1 + 1].x, t.w, v[m1 + 1].y * t.z));
            bp[m1 >> 1] = make_float4(b0.x, b0.y, b1.x, b1.y);
        }
    }
    __syncwarp();
#pragma unroll
    for (int l2 = 0; l2 < R; ++l2) v[REV ? (R - 1 - l2) : l2] = buf[l2 * S + ll];
    fft_inreg<R, SIGN>(v);  // v[m2] = X[ll + R*m2]
    __syncwarp();
}

template <int R, int SIGN, bool REV>
__device__ __forceinline__ void warp_fft_2pass(float2 (&v)[R], float2* __restrict__ buf,
                                               const float2* __restrict__ tws, const int ll) {
    fft_inreg<R, SIGN>(v);  // v[m1] = sum_j u[R j + l] W_R^{j m1}
    warp_fft_xpose_pass2<R, SIGN, REV>(v, buf, tws, ll);
}

}  // namespace rcb
