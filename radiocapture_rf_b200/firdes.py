"""Host-side filter design (one-time, per channel): numpy restatement of the GNU Radio 3.8 designers the
reference calls - firdes.low_pass_2 (rc_frontend/channel.py:33, p25_control_demod.py:107), firdes.low_pass
(channel.py:50,55; rc_frontend/receiver.py:83), fft.window.blackmanharris (fft_vector.py:38) and
optfir.low_pass (rc_frontend/receiver.py:251).  Vectorised; taps are float32 like std::vector<float>.
"""
import math

import numpy as np

WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_KAISER, WIN_BLACKMAN_hARRIS = range(6)
WIN_BLACKMAN_HARRIS = WIN_BLACKMAN_hARRIS
_COS_COEFFS = {
    WIN_HAMMING: (0.54, 0.46),
    WIN_HANN: (0.5, 0.5),
    WIN_BLACKMAN: (0.42, 0.5, 0.08),
    WIN_BLACKMAN_HARRIS: (0.35875, 0.48829, 0.14128, 0.01168),
}
_MAX_ATTEN = {WIN_HAMMING: 53.0, WIN_HANN: 44.0, WIN_BLACKMAN: 74.0, WIN_RECTANGULAR: 21.0,
              WIN_BLACKMAN_HARRIS: 92.0}


def window(win_type, ntaps, beta=6.76):
    if ntaps == 1 or win_type == WIN_RECTANGULAR:
        return np.ones(ntaps, dtype=np.float32)
    if win_type == WIN_KAISER:
        return np.kaiser(ntaps, beta).astype(np.float32)
    c = _COS_COEFFS[win_type]
    n = np.arange(ntaps, dtype=np.float64)
    m = float(ntaps - 1)
    w = np.full(ntaps, c[0])
    for k in range(1, len(c)):
        w = w + ((-1.0) ** k) * c[k] * np.cos(2.0 * math.pi * k * n / m)
    return w.astype(np.float32)


def blackmanharris(ntaps):
    return window(WIN_BLACKMAN_HARRIS, ntaps)


def _odd(n):
    return n + 1 if (n & 1) == 0 else n


def _sinc_lowpass(gain, fs, fc, ntaps, w):
    m = (ntaps - 1) // 2
    n = np.arange(-m, m + 1, dtype=np.float64)
    fwt0 = 2.0 * math.pi * fc / fs
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.where(n == 0, fwt0 / math.pi, np.sin(n * fwt0) / (n * math.pi))
    taps = (h * w.astype(np.float64)).astype(np.float32)
    fmax = float(taps[m]) + 2.0 * float(taps[m + 1:].astype(np.float64).sum())
    return (taps.astype(np.float64) * (gain / fmax)).astype(np.float32)


def max_attenuation(win_type, beta=6.76):
    """gr::fft::window::max_attenuation: fixed per window, beta / 0.1102 + 8.7 for Kaiser."""
    if win_type == WIN_KAISER:
        return beta / 0.1102 + 8.7
    return _MAX_ATTEN[win_type]


def compute_ntaps(sampling_freq, transition_width, win_type=WIN_HAMMING, beta=6.76):
    """firdes::compute_ntaps: int(A fs / (22 tw)), made odd."""
    return _odd(int(max_attenuation(win_type, beta) * sampling_freq / (22.0 * transition_width)))


def low_pass(gain, sampling_freq, cutoff_freq, transition_width, win_type=WIN_HAMMING, beta=6.76):
    ntaps = compute_ntaps(sampling_freq, transition_width, win_type, beta)
    return _sinc_lowpass(gain, sampling_freq, cutoff_freq, ntaps, window(win_type, ntaps, beta))


def low_pass_2(gain, sampling_freq, cutoff_freq, transition_width, attenuation_dB, win_type=WIN_HAMMING,
               beta=6.76):
    ntaps = _odd(int(attenuation_dB * sampling_freq / (22.0 * transition_width)))
    return _sinc_lowpass(gain, sampling_freq, cutoff_freq, ntaps, window(win_type, ntaps, beta))


def channel_decimation(samp_rate, channel_rate):
    """rc_frontend/channel.py:31 `int(samp_rate/channel_rate)/2` made an integer (Appendix C.5)."""
    d = int(samp_rate / channel_rate) // 2
    if d < 1:
        raise ValueError("channel_rate %s too high for samp_rate %s" % (channel_rate, samp_rate))
    return d


def optfir_low_pass(gain, Fs, freq1, freq2, passband_ripple_db, stopband_atten_db, nextra_taps=2):
    """optfir.low_pass via scipy.signal.remez with the remezord order estimate (Herrmann et al.)."""
    from scipy.signal import remez
    r = passband_ripple_db / 20.0
    dp = (10.0 ** r - 1) / (10.0 ** r + 1)
    ds = 10.0 ** (-stopband_atten_db / 20.0)
    f1, f2 = freq1 / float(Fs), freq2 / float(Fs)
    ldp, lds = math.log10(dp), math.log10(ds)
    dinf = ((5.309e-3 * ldp * ldp + 7.114e-2 * ldp - 4.761e-1) * lds) + (-2.66e-3 * ldp * ldp - 5.941e-1 * ldp - 4.278e-1)
    ff = 11.01217 + 0.5124401 * (ldp - lds)
    df = abs(f2 - f1)
    n = int(math.ceil(dinf / df - ff * df + 1)) - 1
    ntaps = n + nextra_taps + 1
    mx = max(dp, ds)
    wts = [mx / dp, mx / ds]
    taps = remez(ntaps, [0.0, f1, f2, 0.5], [gain, 0.0], weight=wts, fs=1.0, maxiter=200)
    return np.asarray(taps, dtype=np.float32)


def pfb_prototype(num_channels):
    """rc_frontend/receiver.py:249-254: optfir.low_pass(1, N, 0.5, 0.5+0.2, 0.1, 80)."""
    return optfir_low_pass(1.0, float(num_channels), 0.5, 0.7, 0.1, 80.0)


# ---- designs used by the post-demod chains (K6) ------------------------------------------------
def high_pass(gain, sampling_freq, cutoff_freq, transition_width, win_type=WIN_HAMMING, beta=6.76):
    """firdes::high_pass - the 300 Hz audio high-pass of logging_receiver.py:215 (2007 taps at 25 kHz)."""
    ntaps = compute_ntaps(sampling_freq, transition_width, win_type, beta)
    w = window(win_type, ntaps, beta).astype(np.float64)
    m = (ntaps - 1) // 2
    n = np.arange(-m, m + 1, dtype=np.float64)
    fwt0 = 2.0 * math.pi * cutoff_freq / sampling_freq
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.where(n == 0, 1.0 - fwt0 / math.pi, -np.sin(n * fwt0) / (n * math.pi))
    taps = (h * w).astype(np.float32)
    k = np.arange(1, m + 1, dtype=np.float64)
    fmax = float(taps[m]) + 2.0 * float((taps[m + 1:].astype(np.float64) * np.cos(k * math.pi)).sum())
    return (taps.astype(np.float64) * (gain / fmax)).astype(np.float32)


def fm_deemph(fs, tau=75e-6):
    """analog.fm_deemph(fs, tau) -> (b0, b1, a1) of y[n] = b0 x[n] + b1 x[n-1] - a1 y[n-1] (iir_filter_ffd, fp64):
    bilinear transform of H(s) = w_ca / (s + w_ca) with the corner prewarped (gr-analog fm_emph.py, 3.8)."""
    w_c = 1.0 / tau
    w_ca = 2.0 * fs * math.tan(w_c / (2.0 * fs))
    k = -w_ca / (2.0 * fs)
    p1 = (1.0 + k) / (1.0 - k)
    b0 = -k / (1.0 - k)
    return b0, b0, -p1


def rational_resampler_design(interpolation, decimation, fractional_bw=0.4):
    """rational_resampler_fff(interpolation, decimation, taps=None, fractional_bw=None) -> (I, D, taps): the gcd
    reduction and design_filter of gr-filter's rational_resampler.py (Kaiser beta 7 low-pass, gain I)."""
    g = math.gcd(int(interpolation), int(decimation))
    interpolation, decimation = int(interpolation) // g, int(decimation) // g
    if fractional_bw >= 0.5 or fractional_bw <= 0:
        raise ValueError("Invalid fractional_bandwidth, must be in (0, 0.5)")
    beta = 7.0
    halfband = 0.5
    rate = float(interpolation) / float(decimation)
    if rate >= 1.0:
        trans_width = halfband - fractional_bw
        mid_transition_band = halfband - trans_width / 2.0
    else:
        trans_width = rate * (halfband - fractional_bw)
        mid_transition_band = rate * halfband - trans_width / 2.0
    return interpolation, decimation, low_pass(interpolation, interpolation, mid_transition_band, trans_width,
                                               WIN_KAISER, beta)
