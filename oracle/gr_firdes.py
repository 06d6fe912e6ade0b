"""Oracle: GNU Radio 3.8 ``firdes`` / ``fft.window`` / ``optfir`` tap design (TEST INFRASTRUCTURE).

Restates gr-filter/lib/firdes.cc and gr-fft/lib/window.cc (GNU Radio 3.8.x, un-vendored —
"parity unpinned", see oracle/__init__.py) as used by the reference at

  rc_frontend/channel.py:33        firdes.low_pass_2(1.0, fs, rate/2, rate/2, 20.0, WIN_HAMMING)
  rc_frontend/channel.py:50,55     firdes.low_pass(1, fs, (rate-2000)/2, 4000)
  rc_frontend/receiver.py:83       firdes.low_pass(1, fs, fs/4, fs/8)           (split-2 half band)
  rc_frontend/receiver.py:251      optfir.low_pass(1, N, 0.5, 0.7, 0.1, 80)     (PFB prototype)
  p25_control_demod.py:107         firdes.low_pass_2(1.0, 25000, 6250, 500.0, 30.0, WIN_BLACKMAN)
  fft_vector.py:38                 window.blackmanharris(16384)

GNU Radio keeps taps in ``std::vector<float>``: every stored value is rounded to
float32 here as well, while the expressions are evaluated in double like the C++.
"""
import math

import numpy as np

WIN_HAMMING = 0
WIN_HANN = 1
WIN_BLACKMAN = 2
WIN_RECTANGULAR = 3
WIN_KAISER = 4
WIN_BLACKMAN_hARRIS = 5
WIN_BLACKMAN_HARRIS = 5

# gr-fft/lib/window.cc  window::max_attenuation()
_MAX_ATTEN = {
    WIN_HAMMING: 53.0,
    WIN_HANN: 44.0,
    WIN_BLACKMAN: 74.0,
    WIN_RECTANGULAR: 21.0,
    WIN_BLACKMAN_HARRIS: 92.0,
}


def _coswindow(ntaps, coeffs):
    """gr-fft/lib/window.cc coswindow(): sum_k (-1)^k c_k cos(2 pi k n / (ntaps-1))."""
    n = np.arange(ntaps, dtype=np.float64)
    m = float(ntaps - 1)
    w = np.full(ntaps, coeffs[0], dtype=np.float64)
    sign = -1.0
    for k, c in enumerate(coeffs[1:], start=1):
        w = w + sign * c * np.cos((2.0 * math.pi * k * n) / m)
        sign = -sign
    return w.astype(np.float32)


def window(win_type, ntaps, beta=6.76):
    """firdes::window() -> float32 vector (symmetric windows, denominator ntaps-1)."""
    if ntaps == 1:
        return np.ones(1, dtype=np.float32)
    if win_type == WIN_HAMMING:
        return _coswindow(ntaps, (0.54, 0.46))
    if win_type == WIN_HANN:
        return _coswindow(ntaps, (0.5, 0.5))
    if win_type == WIN_BLACKMAN:
        return _coswindow(ntaps, (0.42, 0.5, 0.08))
    if win_type == WIN_RECTANGULAR:
        return np.ones(ntaps, dtype=np.float32)
    if win_type == WIN_BLACKMAN_HARRIS:
        return _coswindow(ntaps, (0.35875, 0.48829, 0.14128, 0.01168))
    if win_type == WIN_KAISER:
        return np.kaiser(ntaps, beta).astype(np.float32)
    raise ValueError("unknown window type %r" % (win_type,))


def blackmanharris(ntaps):
    """gnuradio.fft.window.blackmanharris(ntaps) (fft_vector.py:38), default 92 dB variant."""
    return window(WIN_BLACKMAN_HARRIS, ntaps)


def compute_ntaps(sampling_freq, transition_width, win_type, beta=6.76):
    """firdes::compute_ntaps(): int(a*fs/(22*tw)), made odd; a = window::max_attenuation (Kaiser: beta/0.1102 + 8.7)."""
    a = (beta / 0.1102 + 8.7) if win_type == WIN_KAISER else _MAX_ATTEN[win_type]
    ntaps = int(a * sampling_freq / (22.0 * transition_width))
    if (ntaps & 1) == 0:
        ntaps += 1
    return ntaps


def compute_ntaps_windes(sampling_freq, transition_width, attenuation_dB):
    """firdes::compute_ntaps_windes(): int(atten*fs/(22*tw)), made odd."""
    ntaps = int(attenuation_dB * sampling_freq / (22.0 * transition_width))
    if (ntaps & 1) == 0:
        ntaps += 1
    return ntaps


def _windowed_sinc_lowpass(gain, sampling_freq, cutoff_freq, ntaps, w):
    m = (ntaps - 1) // 2
    fwt0 = 2.0 * math.pi * cutoff_freq / sampling_freq
    taps = np.zeros(ntaps, dtype=np.float32)
    for n in range(-m, m + 1):
        if n == 0:
            taps[n + m] = np.float32(fwt0 / math.pi * float(w[n + m]))
        else:
            taps[n + m] = np.float32(math.sin(n * fwt0) / (n * math.pi) * float(w[n + m]))
    # normalise for unity DC gain (accumulated in double over the float32 taps)
    fmax = float(taps[m])
    for n in range(1, m + 1):
        fmax += 2.0 * float(taps[n + m])
    g = gain / fmax
    return (taps.astype(np.float64) * g).astype(np.float32)


def low_pass(gain, sampling_freq, cutoff_freq, transition_width, win_type=WIN_HAMMING, beta=6.76):
    """firdes::low_pass()."""
    ntaps = compute_ntaps(sampling_freq, transition_width, win_type, beta)
    return _windowed_sinc_lowpass(gain, sampling_freq, cutoff_freq, ntaps, window(win_type, ntaps, beta))


def low_pass_2(gain, sampling_freq, cutoff_freq, transition_width, attenuation_dB,
               win_type=WIN_HAMMING, beta=6.76):
    """firdes::low_pass_2()."""
    ntaps = compute_ntaps_windes(sampling_freq, transition_width, attenuation_dB)
    return _windowed_sinc_lowpass(gain, sampling_freq, cutoff_freq, ntaps, window(win_type, ntaps, beta))


def high_pass(gain, sampling_freq, cutoff_freq, transition_width, win_type=WIN_HAMMING, beta=6.76):
    """firdes::high_pass() (logging_receiver.py:215, 300 Hz audio HPF)."""
    ntaps = compute_ntaps(sampling_freq, transition_width, win_type, beta)
    w = window(win_type, ntaps, beta)
    m = (ntaps - 1) // 2
    fwt0 = 2.0 * math.pi * cutoff_freq / sampling_freq
    taps = np.zeros(ntaps, dtype=np.float32)
    for n in range(-m, m + 1):
        if n == 0:
            taps[n + m] = np.float32((1.0 - fwt0 / math.pi) * float(w[n + m]))
        else:
            taps[n + m] = np.float32(-math.sin(n * fwt0) / (n * math.pi) * float(w[n + m]))
    fmax = float(taps[m])
    for n in range(1, m + 1):
        fmax += 2.0 * float(taps[n + m]) * math.cos(n * math.pi)
    g = gain / fmax
    return (taps.astype(np.float64) * g).astype(np.float32)


def channel_taps(samp_rate, channel_rate):
    """Taps + decimation exactly as rc_frontend/channel.py:31-33 builds them."""
    decim = int(samp_rate / channel_rate) // 2
    taps = low_pass_2(1.0, float(samp_rate), channel_rate / 2, channel_rate / 2, 20.0, WIN_HAMMING)
    return decim, taps


# ---------------------------------------------------------------------------------------------
# optfir.low_pass (gr-filter/python/filter/optfir.py): Parks-McClellan with the remezord order
# estimate.  GNU Radio calls its own pm_remez; scipy.signal.remez implements the same exchange
# algorithm (same optimum, different numerics) -> approximation, only used to make realistic
# PFB prototypes; kernels always take taps as explicit inputs.
# ---------------------------------------------------------------------------------------------
def _lporder(freq1, freq2, delta_p, delta_s):
    df = abs(freq2 - freq1)
    ddp = math.log10(delta_p)
    dds = math.log10(delta_s)
    a1, a2, a3 = 5.309e-3, 7.114e-2, -4.761e-1
    a4, a5, a6 = -2.66e-3, -5.941e-1, -4.278e-1
    b1, b2 = 11.01217, 0.5124401
    t1 = a1 * ddp * ddp
    t2 = a2 * ddp
    t3 = a4 * ddp * ddp
    t4 = a5 * ddp
    dinf = ((t1 + t2 + a3) * dds) + (t3 + t4 + a6)
    ff = b1 + b2 * (ddp - dds)
    return dinf / df - ff * df + 1


def _remezord(fcuts, mags, devs, fsamp=2):
    fcuts = [f / float(fsamp) for f in fcuts]
    nbands = len(mags)
    devs = list(devs)
    for i in range(nbands):
        if mags[i] != 0:
            devs[i] = devs[i] / mags[i]
    f1 = fcuts[0::2]
    f2 = fcuts[1::2]
    n = 0
    min_delta = 2
    for i in range(len(f1)):
        if f2[i] - f1[i] < min_delta:
            n = i
            min_delta = f2[i] - f1[i]
    if nbands == 2:
        l = _lporder(f1[n], f2[n], devs[0], devs[1])
    else:
        raise NotImplementedError
    n = int(math.ceil(l)) - 1
    ff = [0] + fcuts + [1]
    for i in range(1, len(ff) - 1):
        ff[i] *= 2
    aa = []
    for a in mags:
        aa = aa + [a, a]
    max_dev = max(devs)
    wts = [max_dev / d for d in devs]
    return n, ff, aa, wts


def optfir_low_pass(gain, Fs, freq1, freq2, passband_ripple_db, stopband_atten_db, nextra_taps=2):
    """optfir.low_pass() restated with scipy.signal.remez (rc_frontend/receiver.py:251)."""
    from scipy.signal import remez
    ripple = passband_ripple_db / 20.0
    passband_dev = (10.0 ** ripple - 1) / (10.0 ** ripple + 1)
    stopband_dev = 10.0 ** (-stopband_atten_db / 20.0)
    desired_ampls = (gain, 0)
    n, fo, ao, w = _remezord([freq1, freq2], desired_ampls, [passband_dev, stopband_dev], Fs)
    ntaps = n + nextra_taps + 1
    bands = [f * 0.5 for f in fo]  # remezord's ff is in units of Nyquist; scipy wants fs=1 cycles
    taps = remez(ntaps, bands, ao[0::2], weight=w, fs=1.0, maxiter=200)
    return np.asarray(taps, dtype=np.float32)


def pfb_prototype(nchans, taps_per_arm=None, atten_db=80.0):
    """A realistic PFB prototype low-pass.

    ``taps_per_arm=None`` follows the reference (rc_frontend/receiver.py:249-254:
    optfir.low_pass(1, N, 0.5, 0.7, 0.1, 80)); otherwise a Blackman-Harris windowed sinc of
    exactly ``nchans*taps_per_arm`` taps with cutoff at half a bin (used for the BASELINE
    configs that fix the tap count, e.g. 128 taps / 64 channels).
    """
    if taps_per_arm is None:
        return optfir_low_pass(1.0, float(nchans), 0.5, 0.7, 0.1, atten_db)
    ntaps = int(nchans * taps_per_arm)
    n = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    fc = 0.5 / nchans
    h = 2 * fc * np.sinc(2 * fc * n)
    from scipy.signal.windows import blackmanharris as _bh
    h = h * _bh(ntaps, sym=True)
    h = h / h.sum()
    return h.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# post-demod chain designs
# ---------------------------------------------------------------------------------------------
def fm_deemph_taps(fs, tau=75e-6):
    """gr-analog/python/analog/fm_emph.py fm_deemph (3.8): bilinear transform of H(s) = w_ca / (s + w_ca) with the
    corner prewarped; returns (btaps, ataps) of iir_filter_ffd(btaps, ataps, oldstyle=False):
    y[n] = b0 x[n] + b1 x[n-1] - a1 y[n-1]   (logging_receiver.py:214 via fm_demod_cf(tau=75e-6))."""
    w_c = 1.0 / tau
    w_ca = 2.0 * fs * math.tan(w_c / (2.0 * fs))
    k = -w_ca / (2.0 * fs)
    z1 = -1.0
    p1 = (1.0 + k) / (1.0 - k)
    b0 = -k / (1.0 - k)
    return [b0 * 1.0, b0 * -z1], [1.0, -p1]


def rational_resampler_taps(interpolation, decimation, fractional_bw=0.4):
    """gr-filter/python/filter/rational_resampler.py design_filter (3.8), after the gcd reduction the block applies
    when no taps are given (logging_receiver.py:216-221: 8000 / input_rate)."""
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
    taps = low_pass(interpolation, interpolation, mid_transition_band, trans_width, WIN_KAISER, beta)
    return interpolation, decimation, taps


def fm_demod_audio_taps(channel_rate, audio_pass, audio_stop, gain):
    """gr-analog fm_demod_cf: optfir.low_pass(gain, channel_rate, audio_pass, audio_stop, 0.1, 60)."""
    return optfir_low_pass(gain, channel_rate, audio_pass, audio_stop, 0.1, 60)
