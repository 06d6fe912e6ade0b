"""Oracle: GNU Radio 3.8 block semantics on the radiocapture-rf hot path (TEST INFRASTRUCTURE).

float64 numpy restatements ("canonical" form, exact phase / exact atan2 / exact sums) plus
float32 "GR-compat" emulations where GNU Radio's own arithmetic differs measurably from ideal
math (recursive rotator, table atan2, running-sum moving average).  PARITY UNPINNED for the GNU
Radio blocks (see oracle/__init__.py); the peak picker calls the reference's real dependency.

Block -> upstream source restated -> reference call site
  freq_xlating_fir      gr-filter/lib/freq_xlating_fir_filter_impl.cc, fir_filter.cc, gr-blocks rotator.h
                        rc_frontend/channel.py:35, rc_frontend/receiver.py:85-86, p25_control_demod.py:108
  pfb_channelizer       gr-filter/lib/pfb_channelizer_ccf_impl.cc, polyphase_filterbank.cc, python pfb.py
                        rc_frontend/receiver.py:249-261
  quadrature_demod      gr-analog/lib/quadrature_demod_cf_impl.cc, gnuradio-runtime/lib/math/fast_atan2f.cc
                        moto_control_demod.py:105, edacs_control_demod.py:82, p25_control_demod.py:121,
                        logging_receiver.py:214,234,336,346
  fft_vcc / mag2 / nlog10 / moving_average / head+skiphead
                        gr-fft/lib/fft_vcc_fftw.cc, gr-blocks/lib/{complex_to_mag_squared,nlog10_ff,
                        moving_average}_impl.cc ; fft_vector.py:37-60
  peak_detect           fft_peak_detection.py:46-72 (scipy.signal.find_peaks, the real library)
"""
import math

import numpy as np


# ---------------------------------------------------------------------------------------------
# freq_xlating_fir_filter_ccc
# ---------------------------------------------------------------------------------------------
def composite_taps(taps, center_freq, samp_rate, omega_f32=False):
    """c[k] = h[k] e^{+j w k}, w = 2 pi f0 / fs (freq_xlating_fir_filter_impl::build_composite_fir).

    ``omega_f32`` mirrors GNU Radio holding w in a C ``float``."""
    w = 2.0 * math.pi * float(center_freq) / float(samp_rate)
    if omega_f32:
        w = float(np.float32(w))
    k = np.arange(len(taps), dtype=np.float64)
    return np.asarray(taps, dtype=np.float64) * np.exp(1j * w * k), w


def fir_decimate(x, ctaps, decim, history=None):
    """v[i] = sum_k c[k] x[i*D - k], x[<0] = history (zeros by default).  float64/complex128.

    fir_filter.cc with decimation: history() = ntaps-1 items, one output per D inputs."""
    x = np.asarray(x)
    ctaps = np.asarray(ctaps)
    nt = len(ctaps)
    if history is None:
        history = np.zeros(nt - 1, dtype=np.complex128)
    history = np.asarray(history, dtype=np.complex128)
    assert len(history) == nt - 1
    xx = np.concatenate([history, x.astype(np.complex128)])
    nout = len(x) // decim
    if nout == 0:
        return np.zeros(0, dtype=np.complex128)
    from numpy.lib.stride_tricks import sliding_window_view
    win = sliding_window_view(xx, nt)[::decim][:nout]  # win[i][q] = xx[i*D+q] = x[i*D - (nt-1-q)]
    rev = ctaps[::-1].astype(np.complex128)
    out = np.empty(nout, dtype=np.complex128)
    step = max(1, (1 << 22) // nt)
    for s in range(0, nout, step):
        out[s:s + step] = win[s:s + step] @ rev
    return out


def freq_xlating_fir(x, taps, decim, center_freq, samp_rate, history=None, out_index0=0,
                     omega_f32=False):
    """Canonical freq_xlating_fir_filter_ccc: y[i] = e^{-j w D (i+i0)} sum_k h[k] e^{jwk} x[iD-k]."""
    ctaps, w = composite_taps(taps, center_freq, samp_rate, omega_f32)
    v = fir_decimate(x, ctaps, decim, history)
    i = np.arange(len(v), dtype=np.float64) + float(out_index0)
    # exact phase: reduce f0*D/fs*i modulo 1 before the exponential
    cyc = (w / (2.0 * math.pi)) * decim
    ph = np.modf(cyc * i)[0]
    return v * np.exp(-2j * math.pi * ph)


def freq_xlating_fir_grcompat(x, taps, decim, center_freq, samp_rate):
    """GR-compat float32 emulation: complex64 composite taps and dot product accumulate, recursive
    complex64 rotator renormalised every 512 outputs (gr-blocks rotator.h).  Slow; small inputs."""
    w = np.float32(2.0 * math.pi * float(center_freq) / float(samp_rate))
    k = np.arange(len(taps), dtype=np.float32)
    ctaps = (np.asarray(taps, dtype=np.float32) * np.exp(1j * (w * k).astype(np.float32))).astype(np.complex64)
    nt = len(ctaps)
    xx = np.concatenate([np.zeros(nt - 1, np.complex64), np.asarray(x, np.complex64)])
    nout = len(x) // decim
    rev = ctaps[::-1]
    out = np.empty(nout, np.complex64)
    phase = np.complex64(1.0)
    incr = np.complex64(np.exp(-1j * np.float32(w * np.float32(decim))))
    counter = 0
    for i in range(nout):
        v = np.complex64(np.dot(xx[i * decim:i * decim + nt], rev))
        out[i] = np.complex64(v * phase)
        phase = np.complex64(phase * incr)
        counter += 1
        if counter % 512 == 0:
            phase = np.complex64(phase / np.abs(phase))
    return out


# ---------------------------------------------------------------------------------------------
# pfb.channelizer_ccf (critically sampled, identity channel map)
# ---------------------------------------------------------------------------------------------
def pfb_arm_taps(taps, nchans):
    """polyphase_filterbank::set_taps: arm i gets h[i + k*N], zero padded to ceil(L/N) each.
    Returns array [P][N] (k major)."""
    taps = np.asarray(taps)
    p = int(math.ceil(len(taps) / float(nchans)))
    padded = np.zeros(p * nchans, dtype=taps.dtype)
    padded[:len(taps)] = taps
    return padded.reshape(p, nchans)


def pfb_channelizer(x, taps, nchans, history=None):
    """Y[m][n] = sum_i u_i[n] e^{+j 2 pi m i / N},  u_i[n] = sum_k h[i+kN] x[(n-k)N + N-1-i].

    stream_to_streams sends sample s to stream s mod N; stream j is filtered by arm N-1-j and
    written to FFT input N-1-j; backward unnormalised FFT; output port m = FFT bin m.
    ``history``: the (P-1)*N samples preceding x (zeros by default).  Returns complex128 [N][T]."""
    x = np.asarray(x)
    n = int(nchans)
    arms = pfb_arm_taps(np.asarray(taps, dtype=np.float64), n)  # [P][N]
    p = arms.shape[0]
    t = len(x) // n
    if history is None:
        history = np.zeros((p - 1) * n, dtype=np.complex128)
    assert len(history) == (p - 1) * n
    rows = np.concatenate([np.asarray(history, np.complex128), x[:t * n].astype(np.complex128)]).reshape(t + p - 1, n)
    rows_rev = rows[:, ::-1]  # rows_rev[r][i] = row r, position N-1-i
    u = np.zeros((t, n), dtype=np.complex128)
    for k in range(p):
        # output frame f uses row (f + p-1 - k)
        u += arms[k][None, :] * rows_rev[p - 1 - k:p - 1 - k + t, :]
    y = np.fft.ifft(u, axis=1) * n  # backward, unnormalised
    return np.ascontiguousarray(y.T)


def pfb_channelizer_direct(x, taps, nchans, m, nout):
    """Definition-level check: y_m[n] = sum_t h[t] e^{+j 2 pi m t/N} x[(n+1)N-1-t]."""
    x = np.asarray(x, dtype=np.complex128)
    taps = np.asarray(taps, dtype=np.float64)
    n = int(nchans)
    t = np.arange(len(taps))
    c = taps * np.exp(2j * math.pi * m * t / n)
    out = np.zeros(nout, dtype=np.complex128)
    for o in range(nout):
        idx = (o + 1) * n - 1 - t
        ok = idx >= 0
        out[o] = np.sum(c[ok] * x[idx[ok]])
    return out


# ---------------------------------------------------------------------------------------------
# quadrature_demod_cf
# ---------------------------------------------------------------------------------------------
def quadrature_demod(x, gain, prev=0.0):
    """out[n] = gain * atan2(Im p, Re p), p = x[n] conj(x[n-1]); x[-1] = prev (0 at stream start;
    atan2(0,0) = 0 as in fast_atan2f).  Works along the last axis.  float64."""
    x = np.asarray(x, dtype=np.complex128)
    prev = np.broadcast_to(np.asarray(prev, dtype=np.complex128), x.shape[:-1] + (1,))
    xm1 = np.concatenate([prev, x[..., :-1]], axis=-1)
    p = x * np.conj(xm1)
    # fast_atan2f returns 0 for (0,0); numpy's arctan2(0,-0.0) would give pi
    ang = np.where((p.real == 0) & (p.imag == 0), 0.0, np.arctan2(p.imag, p.real))
    return gain * ang


_FAST_ATAN_TABLE = None


def _fast_atan_table():
    global _FAST_ATAN_TABLE
    if _FAST_ATAN_TABLE is None:
        t = np.arctan(np.arange(256, dtype=np.float64) / 255.0)
        _FAST_ATAN_TABLE = np.concatenate([t, t[-1:]]).astype(np.float32)  # 257 entries
    return _FAST_ATAN_TABLE


def fast_atan2f(y, x):
    """gnuradio-runtime/lib/math/fast_atan2f.cc emulated in float32 (vectorised)."""
    tab = _fast_atan_table()
    y = np.asarray(y, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    ya = np.abs(y)
    xa = np.abs(x)
    zero = ~((ya > 0) | (xa > 0))
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(ya < xa, ya / xa, xa / ya).astype(np.float32)
    z = np.where(zero, np.float32(0), z)
    alpha = (z * np.float32(255.0)).astype(np.float32)
    index = alpha.astype(np.int32) & 0xFF
    alpha = (alpha - index.astype(np.float32)).astype(np.float32)
    base = (tab[index] + (tab[index + 1] - tab[index]) * alpha).astype(np.float32)
    base = np.where(z < np.float32(0.003921569), z, base).astype(np.float32)
    pi = np.float32(3.14159265358979323846)
    hpi = np.float32(1.57079632679489661923)
    ang_x = np.where(x >= 0, np.where(y >= 0, base, -base),
                     np.where(y >= 0, pi - base, base - pi))
    ang_y = np.where(y >= 0, np.where(x >= 0, hpi - base, hpi + base),
                     np.where(x >= 0, -hpi + base, -hpi - base))
    ang = np.where(xa > ya, ang_x, ang_y).astype(np.float32)
    return np.where(zero, np.float32(0), ang).astype(np.float32)


def quadrature_demod_grcompat(x, gain, prev=0.0):
    """float32 + table atan, as GNU Radio computes it."""
    x = np.asarray(x, dtype=np.complex64)
    xm1 = np.concatenate([np.asarray([prev], np.complex64), x[:-1]])
    p = (x * np.conj(xm1)).astype(np.complex64)
    return (np.float32(gain) * fast_atan2f(p.imag, p.real)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# moving_average_ff (AFC probe p25_control_demod.py:123-127 ; fft_vector.py:42)
# ---------------------------------------------------------------------------------------------
def moving_average(x, length, scale, history=None):
    """out[n] = scale * sum_{g=n-length+1..n} x[g]  (exact float64 windowed sum; axis 0)."""
    x = np.asarray(x, dtype=np.float64)
    if history is None:
        history = np.zeros((length - 1,) + x.shape[1:], dtype=np.float64)
    xx = np.concatenate([history, x], axis=0)
    cs = np.concatenate([np.zeros((1,) + x.shape[1:]), np.cumsum(xx, axis=0)], axis=0)
    return scale * (cs[length:] - cs[:-length])


def moving_average_grcompat(x, length, scale):
    """float32 running add / subtract exactly as moving_average_impl.cc does it."""
    x = np.asarray(x, dtype=np.float32)
    xx = np.concatenate([np.zeros((length - 1,) + x.shape[1:], np.float32), x], axis=0)
    s = np.zeros(x.shape[1:], dtype=np.float32)
    for g in range(length - 1):
        s = (s + xx[g]).astype(np.float32)
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        s = (s + xx[i + length - 1]).astype(np.float32)
        out[i] = s * np.float32(scale)
        s = (s - xx[i]).astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------
# fft_vector.py flowgraph
# ---------------------------------------------------------------------------------------------
def fft_vcc(frames, window, shift=True):
    """fft_vcc(L, forward=True, window, shift): fftshift(FFT(x*w)), unnormalised.  frames [F][L]."""
    frames = np.asarray(frames, dtype=np.complex128)
    xw = frames * np.asarray(window, dtype=np.float64)[None, :]
    out = np.fft.fft(xw, axis=1)
    if shift:
        l = frames.shape[1]
        half = int(math.ceil(l / 2.0))
        out = np.concatenate([out[:, half:], out[:, :half]], axis=1)
    return out


def nlog10(v, n=1.0, k=0.0):
    """blocks.nlog10_ff(n, vlen, k) (gr-blocks nlog10_ff_impl.cc; fft_vector.py:41): n*log10(max(v, 1e-18)) + k."""
    return n * np.log10(np.maximum(np.asarray(v, np.float64), 1e-18)) + k


def log_power(spec, n=1.0, k=1.0):
    """complex_to_mag_squared -> nlog10_ff(n, L, k): n*log10(max(|X|^2, 1e-18)) + k."""
    return nlog10(spec.real ** 2 + spec.imag ** 2, n, k)


def fft_vector_flowgraph(x, length, window, nframes=1000, avg=100):
    """fft_vector.py:37-60: the single vector the flowgraph writes = item #(nframes-1) of the
    avg-frame moving sum of log-power spectra, i.e. sum over frames nframes-avg .. nframes-1."""
    frames = np.asarray(x)[:nframes * length].reshape(nframes, length)
    first = max(0, nframes - avg)
    lp = log_power(fft_vcc(frames[first:nframes], window, True))
    return lp.sum(axis=0)


def logpower_block_sums(x, length, window, avg):
    """Streaming form used by the B200 scan kernel: one vector per block of ``avg`` frames,
    S_b[k] = sum_{f in block b} (log10(max(|X_f[k]|^2,1e-18)) + 1).  Returns [nblocks][L]."""
    nfr = len(x) // length
    nb = nfr // avg
    frames = np.asarray(x)[:nb * avg * length].reshape(nb, avg, length)
    out = np.empty((nb, length), dtype=np.float64)
    for b in range(nb):
        out[b] = log_power(fft_vcc(frames[b], window, True)).sum(axis=0)
    return out


# ---------------------------------------------------------------------------------------------
# fft_peak_detection.py:46-72
# ---------------------------------------------------------------------------------------------
def peak_detect(data, samp_rate, center_freq, fft_width=None):
    """Returns (peak_bin_indices, frequencies_hz) exactly as fft_peak_detection.py:46-72 computes
    them: float32 data, shifted by |min|, sequential float32 mean (Python ``sum`` over a float32
    array), scipy.signal.find_peaks(width=[3 kHz, 30 kHz] in bins, prominence=1), keep > 2*mean."""
    from scipy import signal
    data = np.array(data, dtype=np.float32, copy=True)
    if fft_width is None:
        fft_width = len(data)
    bandwidth = samp_rate
    hz_per_bin = bandwidth / fft_width
    min_w = 3000 / hz_per_bin
    max_w = 30000 / hz_per_bin
    data_min = data.min()
    data = (data + abs(data_min)).astype(np.float32)
    # Python's sum() over float32 scalars accumulates sequentially in float32
    data_average = np.add.accumulate(data, dtype=np.float32)[-1] / np.float32(len(data))
    peaks = signal.find_peaks(data, width=[min_w, max_w], prominence=1)
    idx = []
    freqs = []
    for line in peaks[0]:
        if data[line] > data_average * 2:
            idx.append(int(line))
            freqs.append(int((line * hz_per_bin) - (bandwidth / 2) + center_freq))
    return np.asarray(idx, dtype=np.int64), np.asarray(freqs, dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# rc_frontend/receiver.py:367-383  PFB bin / residual arithmetic (a4)
# ---------------------------------------------------------------------------------------------
def pfb_bin_for_offset(offset_hz, bin_hz, num_channels):
    chan = int(round(offset_hz / float(bin_hz)))
    pfb_offset = offset_hz - chan * bin_hz
    if chan < 0:
        chan = chan + num_channels
    return int(chan), pfb_offset


# ---------------------------------------------------------------------------------------------
# post-demod data-parallel stages: p25_control_demod.py:106-133, logging_receiver.py:210-222
# ---------------------------------------------------------------------------------------------
def fir_filter_fff(x, taps, decim=1, history=None):
    """gr-filter fir_filter_fff(decim, taps): y[i] = sum_k h[k] x[i D - k], zero history (or the ntaps-1 samples given).
    Symbol filter (1/5,)*5 (p25_control_demod.py:130-133), audio LPF / 300 Hz HPF (logging_receiver.py:214-215)."""
    h = np.asarray(taps, dtype=np.float64)
    nt = len(h)
    x = np.asarray(x, dtype=np.float64)
    hist = np.zeros(nt - 1) if history is None else np.asarray(history, dtype=np.float64)
    assert len(hist) == nt - 1
    xx = np.concatenate([hist, x])
    full = np.convolve(xx, h)[nt - 1:nt - 1 + len(x)]
    return full[::decim]


def rational_resampler_fff(x, interp, decim, taps):
    """gr-filter rational_resampler_base: zero-stuff by I, FIR, keep every D-th: y[m] = sum_n x[n] h[m D - n I]
    (phase ctr starts at 0, filter history zero).  Outputs produced while input remains: m D / I < len(x)."""
    h = np.asarray(taps, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    nout = (len(x) * interp + decim - 1) // decim
    up = np.zeros(len(x) * interp)
    up[::interp] = x
    full = np.convolve(up, h)
    return full[np.arange(nout) * decim]


def iir_filter_ffd(x, btaps, ataps, state=None):
    """gr-filter iir_filter_ffd(fftaps, fbtaps, oldstyle=False), first order: y[n] = b0 x[n] + b1 x[n-1] - a1 y[n-1]
    in double precision, float output (the fm_deemph block).  state = (x[-1], y[-1])."""
    b0, b1 = float(btaps[0]), float(btaps[1])
    a1 = float(ataps[1])
    xp, yp = (0.0, 0.0) if state is None else state
    x = np.asarray(x, dtype=np.float64)
    y = np.empty(len(x))
    for n in range(len(x)):
        yp = b0 * x[n] + b1 * xp - a1 * yp
        xp = x[n]
        y[n] = yp
    return y


def pwr_squelch_cc(x, db, alpha=0.0001, ramp=0, gate=False):
    """gr-analog pwr_squelch_cc (squelch_base_cc, ramp 0): d_pwr = single_pole_iir(alpha)(|x|^2) updated per sample,
    muted while d_pwr < 10^(db/10); unmuted samples pass, muted ones are dropped (gate) or zeroed.
    logging_receiver.py:211: pwr_squelch_cc(-100, 0.01, 0, True).  Returns (output, keep_mask)."""
    assert ramp == 0
    thr = 10.0 ** (db / 10.0)
    x = np.asarray(x, dtype=np.complex128)
    pw = x.real * x.real + x.imag * x.imag
    p = 0.0
    keep = np.zeros(len(x), dtype=bool)
    for n in range(len(x)):
        p = alpha * pw[n] + (1.0 - alpha) * p
        keep[n] = not (p < thr)
    if gate:
        return x[keep], keep
    return np.where(keep, x, 0.0), keep


def p25_c4fm_front(x, prefilter_taps, gain, sps=5):
    """p25_control_demod.py:106-133 up to the symbol filter: 1x prefilter (freq_xlating, no shift) -> quadrature demod
    -> boxcar (1/sps,)*sps.  Returns (prefiltered, fm, symbol_filtered)."""
    z = freq_xlating_fir(x, prefilter_taps, 1, 0.0, 1.0)
    fm = quadrature_demod(z, gain)
    sym = fir_filter_fff(fm, np.full(sps, 1.0 / sps))
    return z, fm, sym


def analog_fm_chain(x, rate, squelch_db=-100.0, squelch_alpha=0.01, deviation=15000.0, gain=8.0, tau=75e-6,
                    audio_taps=None, hp_taps=None, resamp=None):
    """logging_receiver.py:210-222 `analog`: pwr_squelch_cc(gate) -> fm_demod_cf (quad demod k = rate/(2 pi dev),
    fm_deemph(tau), audio LPF) -> 300 Hz high pass -> rational_resampler_fff(8000, rate).  Taps are explicit inputs
    (audio_taps, hp_taps, resamp = (I, D, taps)) so design differences cannot contaminate kernel parity."""
    from . import gr_firdes as fd
    y, keep = pwr_squelch_cc(x, squelch_db, squelch_alpha, 0, True)
    fm = quadrature_demod(y, rate / (2.0 * math.pi * deviation))
    b, a = fd.fm_deemph_taps(rate, tau)
    de = iir_filter_ffd(fm, b, a)
    lp = fir_filter_fff(de, audio_taps)
    hp = fir_filter_fff(lp, hp_taps)
    i, d, rt = resamp
    out = rational_resampler_fff(hp, i, d, rt)
    return dict(keep=keep, fm=fm, deemph=de, lpf=lp, hpf=hp, audio=out)


def rel_l2(a, b):
    """||a-b||2 / ||b||2 (the parity metric of SURVEY 8(d))."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    if den == 0:
        return float(np.linalg.norm(a.ravel()))
    return float(np.linalg.norm((a - b).ravel()) / den)
