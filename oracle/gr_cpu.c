/* oracle/gr_cpu.c - float32 "GNU-Radio-semantics" CPU restatement of the hot path.  TEST INFRASTRUCTURE.
 *
 * PARITY UNPINNED (see oracle/__init__.py): GNU Radio 3.8 is not vendored in /root/reference and not
 * installed; these functions restate its published algorithms in plain C the way the GNU Radio blocks
 * compute them (float32 data, direct-form dot products, FFT + table atan2):
 *   grc_pfb_fm          gr-filter pfb_channelizer_ccf_impl.cc / polyphase_filterbank.cc  (rc_frontend/receiver.py:249-261)
 *                       + gr-analog quadrature_demod_cf_impl.cc per bin                  (moto_control_demod.py:105 ...)
 *   grc_xlating_fir     gr-filter freq_xlating_fir_filter_impl.cc + gr-blocks rotator.h   (rc_frontend/channel.py:35)
 *   grc_quad_demod      gr-analog quadrature_demod_cf_impl.cc + fast_atan2f.cc
 *   grc_convert_iq      the host-side wire-format conversion GNU Radio's sources do before anything else: gr-osmosdr
 *                       rtl_source_c ((u8 - 127.4) / 128), UHD convert sc8 / sc16 -> fc32 (configs/config_denver_usrp.py:20)
 *   grc_fft_logpow      gr-fft fft_vcc_fftw.cc + complex_to_mag_squared + nlog10_ff + moving sum (fft_vector.py:37-60)
 * Used ONLY by tests/ (checked against the float64 numpy oracle) and by bench.py's cpu_baseline /
 * --impl reference legs (timed on the host cores, OpenMP over frames / outputs; GNU Radio itself runs
 * each block on ONE thread, so this is generous to the CPU side).  Never linked into the product.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef float complex cf;

/* ---- fast_atan2f (gnuradio-runtime/lib/math/fast_atan2f.cc): 256-interval table, linear interpolation ---- */
static float g_atan_table[257];
static int g_atan_init = 0;
static void atan_init(void) {
    if (g_atan_init) return;
    for (int i = 0; i < 256; ++i) g_atan_table[i] = (float)atan((double)i / 255.0);
    g_atan_table[256] = g_atan_table[255];
    g_atan_init = 1;
}
static inline float fast_atan2f(float y, float x) {
    float y_abs = fabsf(y), x_abs = fabsf(x), z, alpha, angle, base_angle;
    int index;
    if (!((y_abs > 0.0f) || (x_abs > 0.0f))) return 0.0f;
    z = (y_abs < x_abs) ? y_abs / x_abs : x_abs / y_abs;
    if (z < 0.003921569f) {
        base_angle = z;
    } else {
        alpha = z * 255.0f;
        index = ((int)alpha) & 0xff;
        alpha -= (float)index;
        base_angle = g_atan_table[index] + (g_atan_table[index + 1] - g_atan_table[index]) * alpha;
    }
    if (x_abs > y_abs) {
        if (x >= 0.0f) angle = (y >= 0.0f) ? base_angle : -base_angle;
        else { angle = 3.14159265358979323846f; angle = (y >= 0.0f) ? angle - base_angle : base_angle - angle; }
    } else {
        if (y >= 0.0f) { angle = 1.57079632679489661923f; angle = (x >= 0.0f) ? angle - base_angle : angle + base_angle; }
        else { angle = -1.57079632679489661923f; angle = (x >= 0.0f) ? angle + base_angle : angle - base_angle; }
    }
    return angle;
}

/* ---- in-place iterative radix-2 FFT (n power of two), sign = +1 backward / -1 forward; tw[k] = e^{sign j 2 pi k/n} ---- */
static void make_twiddles(cf* tw, int n, int sign) {
    for (int k = 0; k < n / 2; ++k) {
        double a = sign * 2.0 * M_PI * (double)k / (double)n;
        tw[k] = (float)cos(a) + (float)sin(a) * I;
    }
}
static void fft_pow2(cf* a, int n, const cf* tw) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cf t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < half; ++k) {
                cf w = tw[k * step];
                float ur = crealf(a[i + k]), ui = cimagf(a[i + k]);
                float br = crealf(a[i + k + half]), bi = cimagf(a[i + k + half]);
                float vr = br * crealf(w) - bi * cimagf(w), vi = br * cimagf(w) + bi * crealf(w);
                a[i + k] = (ur + vr) + (ui + vi) * I;
                a[i + k + half] = (ur - vr) + (ui - vi) * I;
            }
        }
    }
}
static void dft_naive(const cf* in, cf* out, int n, const cf* twfull /* [n] e^{sign j 2 pi q/n} */) {
    for (int m = 0; m < n; ++m) {
        float ar = 0.f, ai = 0.f;
        int q = 0;
        for (int i = 0; i < n; ++i) {
            cf w = twfull[q];
            ar += crealf(in[i]) * crealf(w) - cimagf(in[i]) * cimagf(w);
            ai += crealf(in[i]) * cimagf(w) + cimagf(in[i]) * crealf(w);
            q += m;
            if (q >= n) q -= n;
        }
        out[m] = ar + ai * I;
    }
}
static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* one PFB frame: u_i = sum_k h[i+kN] x[(f-k)N + N-1-i]  (rows before the block come from hist = last P rows) */
static void pfb_frame(const cf* x, const cf* hist, long f, int N, int P, const float* arms /* [P][N] */, cf* u) {
    for (int i = 0; i < N; ++i) u[i] = 0;
    for (int k = 0; k < P; ++k) {
        long r = f - k;
        const cf* row;
        if (r >= 0) row = x + (size_t)r * N;
        else if (r >= -(long)P) row = hist + (size_t)(r + P) * N;
        else continue;
        const float* h = arms + (size_t)k * N;
        for (int i = 0; i < N; ++i) u[i] += h[i] * row[N - 1 - i];
    }
}

/* Channelizer + per-bin FM.  out_iq / out_fm channel-major [N][ostride] (either may be NULL).
 * hist: P*N samples preceding x (P = ceil(ntaps/N)); updated in place to the last P rows on return. */
int grc_pfb_fm(const cf* x, long nframes, int N, const float* taps, int ntaps, float gain, cf* out_iq, float* out_fm,
               long ostride, cf* hist, int nthreads) {
    atan_init();
    const int P = (ntaps + N - 1) / N;
    float* arms = (float*)calloc((size_t)P * N, sizeof(float));
    if (!arms) return -2;
    memcpy(arms, taps, sizeof(float) * ntaps);
    const int p2 = is_pow2(N);
    cf* tw = (cf*)malloc(sizeof(cf) * (size_t)N);
    if (p2) make_twiddles(tw, N, +1);
    else for (int q = 0; q < N; ++q) { double a = 2.0 * M_PI * q / (double)N; tw[q] = (float)cos(a) + (float)sin(a) * I; }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cf* u = (cf*)malloc(sizeof(cf) * N);
        cf* y = (cf*)malloc(sizeof(cf) * N);
        cf* prev = (cf*)malloc(sizeof(cf) * N);
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        long per = (nframes + nt - 1) / nt, f0 = id * per, f1 = f0 + per;
        if (f1 > nframes) f1 = nframes;
        if (f0 < f1) {
            /* previous frame for the FM carry of this thread's run (recomputed, like the GPU kernel) */
            pfb_frame(x, hist, f0 - 1, N, P, arms, u);
            if (p2) { memcpy(prev, u, sizeof(cf) * N); fft_pow2(prev, N, tw); } else dft_naive(u, prev, N, tw);
            for (long f = f0; f < f1; ++f) {
                pfb_frame(x, hist, f, N, P, arms, u);
                if (p2) { memcpy(y, u, sizeof(cf) * N); fft_pow2(y, N, tw); } else dft_naive(u, y, N, tw);
                if (out_iq) for (int m = 0; m < N; ++m) out_iq[(size_t)m * ostride + f] = y[m];
                if (out_fm) for (int m = 0; m < N; ++m) {
                    cf p = y[m] * conjf(prev[m]);
                    out_fm[(size_t)m * ostride + f] = gain * fast_atan2f(cimagf(p), crealf(p));
                }
                cf* t = prev; prev = y; y = t;
            }
        }
        free(u); free(y); free(prev);
    }
    /* hist <- last P rows of (hist ++ x) */
    {
        size_t cap = (size_t)P * N, n = (size_t)nframes * N;
        cf* nh = (cf*)malloc(sizeof(cf) * cap);
        for (size_t i = 0; i < cap; ++i) {
            long src = (long)i + (long)n - (long)cap;
            nh[i] = (src >= 0) ? x[src] : hist[cap + src];
        }
        memcpy(hist, nh, sizeof(cf) * cap);
        free(nh);
    }
    free(arms); free(tw);
    return 0;
}

/* freq_xlating_fir_filter_ccc from a zero-history start: float composite taps, recursive complex64
 * rotator renormalised every 512 outputs (gr-blocks rotator.h).  nout = ceil(n / decim) outputs whose newest
 * input is x[i*decim].  OpenMP over output ranges (each range seeds its rotator phase exactly). */
int grc_xlating_fir(const cf* x, long n, const float* taps, int ntaps, int decim, double f0, double fs, cf* out,
                    long* nout_p, int nthreads) {
    const float fwT0 = (float)(2.0 * M_PI * f0 / fs);
    cf* ct = (cf*)malloc(sizeof(cf) * ntaps);
    for (int k = 0; k < ntaps; ++k) ct[k] = taps[k] * cexpf(I * (fwT0 * (float)k));
    const long nout = (n + decim - 1) / decim;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        id = omp_get_thread_num();
#endif
        long per = (nout + nt - 1) / nt, o0 = id * per, o1 = o0 + per;
        if (o1 > nout) o1 = nout;
        cf phase = cexpf(-I * (float)fmod((double)fwT0 * decim * (double)o0, 2.0 * M_PI));
        const cf incr = cexpf(-I * (fwT0 * (float)decim));
        unsigned counter = 0;
        for (long o = o0; o < o1; ++o) {
            float ar = 0.f, ai = 0.f;
            long s = o * decim;
            int kmax = (s + 1 < ntaps) ? (int)(s + 1) : ntaps;
            for (int k = 0; k < kmax; ++k) {
                cf xv = x[s - k], c = ct[k];
                ar += crealf(c) * crealf(xv) - cimagf(c) * cimagf(xv);
                ai += crealf(c) * cimagf(xv) + cimagf(c) * crealf(xv);
            }
            out[o] = (ar + ai * I) * phase;
            phase *= incr;
            if ((++counter % 512) == 0) phase /= cabsf(phase);
        }
    }
    free(ct);
    *nout_p = nout;
    return 0;
}

/* out[i] = (v[i] + offset) * scale over 2*n interleaved integers.  fmt: 1 u8, 2 s8, 3 s16. */
int grc_convert_iq(const void* src, int fmt, float offset, float scale, long n, float* out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const long m = 2 * n;
    if (fmt == 3) {
        const short* v = (const short*)src;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < m; ++i) out[i] = ((float)v[i] + offset) * scale;
    } else if (fmt == 2) {
        const signed char* v = (const signed char*)src;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < m; ++i) out[i] = ((float)v[i] + offset) * scale;
    } else if (fmt == 1) {
        const unsigned char* v = (const unsigned char*)src;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < m; ++i) out[i] = ((float)v[i] + offset) * scale;
    } else {
        return -1;
    }
    return 0;
}

int grc_quad_demod(const cf* x, long n, float gain, float prev_re, float prev_im, float* out) {
    atan_init();
    cf prev = prev_re + prev_im * I;
    for (long i = 0; i < n; ++i) {
        cf p = x[i] * conjf(prev);
        out[i] = gain * fast_atan2f(cimagf(p), crealf(p));
        prev = x[i];
    }
    return 0;
}

/* sum over nframes of  log10(max(|fftshift(FFT(x_f * w))|^2, 1e-18)) + 1.   L power of two. */
int grc_fft_logpow(const cf* x, int L, const float* window, long nframes, float* out_sum, int nthreads) {
    if (!is_pow2(L)) return -7;
    cf* tw = (cf*)malloc(sizeof(cf) * (size_t)L / 2);
    make_twiddles(tw, L, -1);
    memset(out_sum, 0, sizeof(float) * (size_t)L);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cf* a = (cf*)malloc(sizeof(cf) * (size_t)L);
        float* acc = (float*)calloc((size_t)L, sizeof(float));
#pragma omp for schedule(static)
        for (long f = 0; f < nframes; ++f) {
            const cf* xf = x + (size_t)f * L;
            for (int i = 0; i < L; ++i) a[i] = xf[i] * window[i];
            fft_pow2(a, L, tw);
            const int half = L / 2;
            for (int k = 0; k < L; ++k) {
                cf v = a[(k + half) % L];
                float p = crealf(v) * crealf(v) + cimagf(v) * cimagf(v);
                acc[k] += log10f(p > 1e-18f ? p : 1e-18f) + 1.0f;
            }
        }
#pragma omp critical
        for (int k = 0; k < L; ++k) out_sum[k] += acc[k];
        free(a); free(acc);
    }
    free(tw);
    return 0;
}

int grc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
