"""CPU: pin the oracle against everything pinnable here.

The reference ships no tests / golden vectors and GNU Radio is not installed => "parity unpinned" for
the GR blocks.  What CAN be checked: (i) the tap-count table the reference's call sites imply,
(ii) independent formulations (scipy.signal.firwin, numpy FFT, definition-level sums), (iii) golden
vectors minted by scripts/make_golden.py (committed under tests/golden/) so oracle drift is caught,
(iv) the peak picker against the real scipy routine the reference calls.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_cpu, gr_firdes as fd, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_tap_count_table():
    # rc_frontend/channel.py:31-33 at the sample rates the shipped configs use
    for fs, ntaps, decim in [(2.0e6, 291, 80), (2.4e6, 349, 96), (2.85e6, 415, 114), (6e6, 873, 240),
                             (8e6, 1163, 320), (10e6, 1455, 400), (12e6, 1745, 480), (16e6, 2327, 640)]:
        d, t = fd.channel_taps(fs, 12500)
        assert (d, len(t)) == (decim, ntaps)
        assert abs(float(t.astype(np.float64).sum()) - 1.0) < 1e-5
        assert np.allclose(t, t[::-1])
    d, t = fd.channel_taps(2.4e6, 6250)
    assert len(t) == 699 or len(t) == 697 or len(t) % 2 == 1
    assert len(fd.low_pass_2(1.0, 25000, 6250, 500.0, 30.0, fd.WIN_BLACKMAN)) == 69   # p25_control_demod.py:107
    assert len(fd.low_pass(1, 8e6, 2e6, 1e6)) == 19                                    # rc_frontend/receiver.py:83
    # 10666666 sps (configs/config_denver_usrp.py:23): non-integer fs/rate -> integer decimation
    d, _ = fd.channel_taps(10666666, 12500)
    assert d == 426


def test_firdes_matches_scipy_firwin():
    from scipy.signal import firwin
    for fs, fc, tw in [(8e6, 2e6, 1e6), (2.4e6, 5250.0, 4000.0), (25000.0, 6250.0, 2000.0)]:
        t = fd.low_pass(1, fs, fc, tw)
        s = firwin(len(t), fc, window="hamming", fs=fs)
        assert np.abs(t - s).max() < 2e-7
    from scipy.signal.windows import blackmanharris
    assert np.abs(fd.blackmanharris(16384) - blackmanharris(16384, sym=True)).max() < 1e-6


def test_firdes_matches_gnuradio_qa_known_taps():
    """GNU Radio's own QA pins firdes with a known-answer vector: gr-filter/python/filter/qa_firdes.py
    test_low_pass / test_low_pass_2 compare firdes.low_pass(1, 1, 0.4, 0.2) and
    firdes.low_pass_2(1, 1, 0.4, 0.2, 60) with the 13 taps below (to 5 places).  Upstream is not vendored in the
    reference and not installed here, so the vector is quoted from the upstream test file rather than generated;
    our independent restatement reproduces it to float32 precision, odd quirks (-4.4e-18) included."""
    known = (0.0024871660862118006, -4.403502608370943e-18, -0.014456653036177158, 0.0543283149600029,
             -0.116202212870121, 0.17504146695137024, 0.7976038455963135, 0.17504146695137024,
             -0.116202212870121, 0.0543283149600029, -0.014456653036177158, -4.403502608370943e-18,
             0.0024871660862118006)
    from radiocapture_rf_b200 import firdes as product_firdes
    for mod in (fd, product_firdes):
        for taps in (mod.low_pass(1, 1, 0.4, 0.2), mod.low_pass_2(1, 1, 0.4, 0.2, 60)):
            assert len(taps) == 13
            np.testing.assert_allclose(np.asarray(taps, np.float64), known, rtol=0, atol=1e-7)


def test_fft_vcc_matches_gnuradio_qa_known_answer():
    """gr-fft/python/fft/qa_fft.py test_001: forward fft_vcc (no window, no shift, unnormalised) of
    complex(primes[2i], primes[2i+1]), i < 32; the first expected outputs as printed upstream (float32)."""
    primes = []
    k = 2
    while len(primes) < 64:
        if all(k % q for q in primes if q * q <= k):
            primes.append(k)
        k += 1
    x = np.array([complex(primes[2 * i], primes[2 * i + 1]) for i in range(32)], np.complex64)
    known = np.array([4377 + 4516j, -1706.1268310546875 + 1638.4256591796875j,
                      -915.2083740234375 + 660.69427490234375j, -660.370361328125 + 381.59600830078125j])
    X = gb.fft_vcc(x.reshape(1, 32), np.ones(32), shift=False)[0]
    np.testing.assert_allclose(X[:4], known, rtol=2e-7)
    # shift=True is what fft_vector.py:38 uses: the same bins rotated by L/2
    Xs = gb.fft_vcc(x.reshape(1, 32), np.ones(32), shift=True)[0]
    np.testing.assert_allclose(Xs[16:20], known, rtol=2e-7)


def test_quadrature_demod_matches_gnuradio_qa_case():
    """gr-analog/python/analog/qa_quadrature_demod.py: a 1 kHz tone at 8 ksps with gain fs/(2 pi f) demodulates to
    [0] + 199 * [1.0] (first output: x[-1] = 0 -> atan2(0, 0) = 0)."""
    f, fs = 1000.0, 8000.0
    x = np.exp(2j * np.pi * f / fs * np.arange(200))
    gain = fs / (2 * np.pi * f)
    for out in (gb.quadrature_demod(x, gain), gb.quadrature_demod_grcompat(x.astype(np.complex64), gain)):
        assert out[0] == 0.0
        np.testing.assert_allclose(out[1:], 1.0, atol=1e-5)


def test_nlog10_matches_gnuradio_qa_case():
    """gr-blocks/python/blocks/qa_nlog10.py: nlog10_ff(10) of (-10, 0, 10, 100, 1000, 10000, 100000) is
    (-180, -180, 10, 20, 30, 40, 50) - the 1e-18 clamp that fft_vector.py:41 relies on for empty bins."""
    out = gb.nlog10([-10, 0, 10, 100, 1000, 10000, 100000], 10.0, 0.0)
    np.testing.assert_allclose(out, [-180, -180, 10, 20, 30, 40, 50], atol=1e-9)


def test_pfb_prototype_reference_shape():
    # rc_frontend/receiver.py:249-254: ~17-19 taps per arm for the 80 dB optfir prototype
    for n in (5, 20, 40):
        t = fd.pfb_prototype(n)
        assert 15 <= len(t) / n <= 20


def test_pfb_equals_definition_and_xlating():
    rng = np.random.default_rng(0)
    n = 16
    h = fd.pfb_prototype(n, 4).astype(np.float64)
    x = rng.standard_normal(n * 40) + 1j * rng.standard_normal(n * 40)
    y = gb.pfb_channelizer(x, h, n)
    for m in (0, 3, 9, 15):
        assert gb.rel_l2(y[m], gb.pfb_channelizer_direct(x, h, n, m, 40)) < 1e-12
    # PFB bin m == xlating FIR at m*fs/N, decim N, input advanced by N-1
    m = 3
    xa = np.concatenate([x[n - 1:], np.zeros(n - 1)])
    hist = np.concatenate([np.zeros(len(h) - 1 - (n - 1)), x[:n - 1]])
    z = gb.freq_xlating_fir(xa, h, n, m / n, 1.0, history=hist)
    assert gb.rel_l2(z[:39], y[m][:39]) < 1e-12


def test_pfb_split_history():
    rng = np.random.default_rng(1)
    n, p = 8, 3
    h = rng.standard_normal(n * p)
    x = rng.standard_normal(n * 50) + 1j * rng.standard_normal(n * 50)
    full = gb.pfb_channelizer(x, h, n)
    a = gb.pfb_channelizer(x[:n * 20], h, n)
    b = gb.pfb_channelizer(x[n * 20:], h, n, history=x[n * 20 - (p - 1) * n:n * 20])
    assert np.allclose(np.concatenate([a, b], axis=1), full, atol=1e-12)


def test_single_tone_fm_constant():
    # tone through DDC -> constant FM output gain*2*pi*df/fs_out
    fs, f0, df = 2.4e6, -62500.0, 1000.0
    decim, taps = fd.channel_taps(fs, 12500)
    t = np.arange(96 * 3000) / fs
    x = np.exp(2j * np.pi * (f0 + df) * t)
    y = gb.freq_xlating_fir(x, taps, decim, f0, fs)
    fm = gb.quadrature_demod(y, 5.0)
    assert np.allclose(fm[10:], 5.0 * 2 * np.pi * df / 25000.0, atol=1e-9)


def test_quadrature_demod_zero_and_table_atan():
    assert gb.quadrature_demod(np.zeros(4, complex), 5.0).tolist() == [0, 0, 0, 0]
    rng = np.random.default_rng(2)
    a = rng.uniform(-np.pi, np.pi, 200000)
    r = rng.uniform(0.1, 2.0, 200000)
    e = gb.fast_atan2f(r * np.sin(a), r * np.cos(a)).astype(np.float64) - a
    assert np.abs(e).max() < 2.0e-6          # table interpolation error ~1.25e-6 + float32 rounding
    assert gb.fast_atan2f(0.0, 0.0) == 0.0
    assert abs(float(gb.fast_atan2f(0.0, -1.0)) - math.pi) < 1e-6


def test_grcompat_rotator_drift_is_what_survey_says():
    x, fs, _ = synth.cfg1(96 * 1200, seed=1)
    decim, taps = fd.channel_taps(fs, 12500)
    e32 = gb.freq_xlating_fir(x, taps, decim, -62500.0, fs, omega_f32=True)
    e64 = gb.freq_xlating_fir(x, taps, decim, -62500.0, fs)
    gr = gb.freq_xlating_fir_grcompat(x, taps, decim, -62500.0, fs)
    # GNU Radio's float32 recursive rotator + float32 dot products stay within 1e-5 of exact arithmetic
    # done with the SAME float32-rounded w ...
    assert gb.rel_l2(gr, e32) < 1e-5
    # ... but holding w in a C float is itself a ~2e-4 departure from ideal math after 1200 outputs
    # (48 ms): GNU Radio is not 1e-5-exact against float64 ideal, hence the gr_float_omega option.
    assert 1e-5 < gb.rel_l2(e64, e32) < 5e-3


def test_moving_average_forms_agree():
    rng = np.random.default_rng(3)
    v = rng.standard_normal((300, 5))
    a = gb.moving_average(v, 100, 1.0)
    b = gb.moving_average_grcompat(v, 100, 1.0)
    assert np.abs(a - b).max() < 5e-4
    assert np.allclose(a[150], v[51:151].sum(axis=0))


def test_fft_flowgraph_is_sum_of_last_100_frames():
    length = 256
    x, _ = synth.scan_stream(length * 120, 2.4e6, length, seed=5, ncarriers=2)
    w = fd.blackmanharris(length)
    vec = gb.fft_vector_flowgraph(x, length, w, nframes=120, avg=100)
    frames = x.reshape(120, length)
    lp = gb.log_power(gb.fft_vcc(frames, w))
    ma = gb.moving_average(lp, 100, 1.0)
    assert np.allclose(vec, ma[119], atol=1e-9)           # head(120) -> skiphead(119) keeps item #119
    blocks = gb.logpower_block_sums(x[:length * 100], length, w, 100)
    assert np.allclose(blocks[0], lp[:100].sum(axis=0))


def test_fft_tone_bin_after_shift():
    length, k = 1024, 100
    t = np.arange(length)
    x = np.exp(2j * np.pi * k * t / length)[None, :]
    spec = gb.fft_vcc(x, np.ones(length))
    assert int(np.argmax(np.abs(spec[0]))) == (k + length // 2) % length


def test_pfb_bin_arithmetic():
    # rc_frontend/receiver.py:367-377
    assert gb.pfb_bin_for_offset(810e3, 400e3, 6) == (2, 10e3)
    assert gb.pfb_bin_for_offset(-390e3, 400e3, 6) == (5, 10e3)
    assert gb.pfb_bin_for_offset(-1.0e6, 400e3, 6)[0] == 3 or gb.pfb_bin_for_offset(-1.0e6, 400e3, 6)[0] == 4


def test_c_restatement_matches_numpy_oracle():
    n = 64
    taps = fd.pfb_prototype(n, 4)
    x, _ = synth.pfb_stream(n * 300, 16e6, n, 1)
    iq, fm, hist = gr_cpu.pfb_fm(x, n, taps, 5.0)
    ref = gb.pfb_channelizer(x, taps.astype(np.float64), n)
    assert gb.rel_l2(iq, ref) < 1e-6
    fr = gb.quadrature_demod(ref, 5.0)
    d = (fm[1::2] - fr[1::2]) / 5.0
    d = (d + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() < 1e-4           # table atan + float32
    # streaming: second block with carried history == one shot
    iq2a, _, h = gr_cpu.pfb_fm(x[:n * 100], n, taps, 5.0)
    iq2b, _, h = gr_cpu.pfb_fm(x[n * 100:], n, taps, 5.0, hist=h)
    assert np.array_equal(np.concatenate([iq2a, iq2b], axis=1), iq)
    for ng in (6, 20):
        tg = fd.pfb_prototype(ng)
        xg, _ = synth.pfb_stream(ng * 200, 2.4e6, ng, 2)
        iqg, _, _ = gr_cpu.pfb_fm(xg, ng, tg, 5.0)
        assert gb.rel_l2(iqg, gb.pfb_channelizer(xg, tg.astype(np.float64), ng)) < 1e-6
    x, fs, _ = synth.cfg1(1 << 16, 1)
    d_, t_ = fd.channel_taps(fs, 12500)
    y = gr_cpu.xlating_fir(x, t_, d_, -62500.0, fs)
    r = gb.freq_xlating_fir(x, t_, d_, -62500.0, fs, omega_f32=True)
    assert gb.rel_l2(y[:len(r)], r) < 5e-6
    q = gr_cpu.quad_demod(y, 5.0)
    assert np.abs(q - gb.quadrature_demod_grcompat(y, 5.0)).max() < 1e-5


def test_golden_vectors():
    """Golden vectors minted by scripts/make_golden.py from the oracle + the real scipy peak picker."""
    with open(os.path.join(GOLD, "manifest.json")) as f:
        man = json.load(f)
    g = np.load(os.path.join(GOLD, "hotpath_golden.npz"))
    # taps
    assert np.array_equal(fd.channel_taps(2.4e6, 12500)[1], g["taps_349"])
    assert np.array_equal(fd.low_pass_2(1.0, 25000, 6250, 500.0, 30.0, fd.WIN_BLACKMAN), g["taps_69"])
    assert np.array_equal(fd.low_pass(1, 8e6, 2e6, 1e6), g["taps_19"])
    # cfg1 DDC + FM on the committed input
    x = g["cfg1_x"]
    y = gb.freq_xlating_fir(x, g["taps_349"], 96, -62500.0, 2.4e6)
    assert gb.rel_l2(y, g["cfg1_y"]) < 1e-12
    assert np.allclose(gb.quadrature_demod(y, 5.0), g["cfg1_fm"], atol=1e-9)
    # PFB
    yp = gb.pfb_channelizer(g["pfb_x"], g["pfb_taps"].astype(np.float64), 16)
    assert gb.rel_l2(yp, g["pfb_y"]) < 1e-12
    # scan vector + peaks (scipy.signal.find_peaks is the reference's own call, fft_peak_detection.py:65)
    idx, freqs = gb.peak_detect(g["scan_vec"], 2.4e6, 855.05e6)
    assert idx.tolist() == man["scan_peaks_idx"]
    assert freqs.tolist() == man["scan_peaks_hz"]


# ---------------------------------------------------------------------------------------------------------------------
# More of GNU Radio's own QA restated against the oracle (the upstream tests compute their expectations with small Python
# reference loops rather than golden numbers; the same loops are written out here and must agree with the oracle).
# ---------------------------------------------------------------------------------------------------------------------
def _qa_fir_filter(x, taps, decim):
    """gr-filter/python/filter/qa_freq_xlating_fir_filter.py `fir_filter`: zero history, newest tap first."""
    y = []
    x2 = (len(taps) - 1) * [0, ] + list(x)
    for i in range(0, len(x), decim):
        yi = 0
        for j in range(len(taps)):
            yi += taps[len(taps) - 1 - j] * x2[i + j]
        y.append(yi)
    return np.array(y)


def _qa_sig_source_c(samp_rate, freq, amp, n):
    t = np.arange(n) / float(samp_rate)
    return amp * np.exp(2j * np.pi * freq * t)


def _qa_mix(lo, data):
    return np.asarray(lo) * np.asarray(data)


@pytest.mark.parametrize("decim", [1, 4])
def test_freq_xlating_matches_gnuradio_qa_construction(decim):
    """qa_freq_xlating_fir_filter.py test_fir_filter_ccf_00x / ccc_00x: the block equals "mix down with a rotator at
    -f0, then FIR-decimate" computed by the QA file's Python loops (fs 1.0, f0 0.2 / -0.2 there, real low-pass taps
    (ccf) and the same taps made complex-bandpass (ccc))."""
    fs, f0 = 1.0, 0.2
    bw = 0.1
    taps = fd.low_pass(1, fs, bw, bw / 4.0).astype(np.float64)
    src = _qa_sig_source_c(fs, -0.2, 1, 600) + _qa_sig_source_c(fs, 0.21, 0.5, 600)
    despun = _qa_mix(_qa_sig_source_c(fs, -f0, 1, len(src)), src)       # rotator(-2 pi f0 / fs), phase 0 at n = 0
    expected = _qa_fir_filter(despun, list(taps), decim)
    got = gb.freq_xlating_fir(src.astype(np.complex64), taps, decim, f0, fs)
    assert gb.rel_l2(got, expected[:len(got)]) < 2e-7                     # input rounded to complex64 only
    # ccc variant: complex band-pass taps t[k] e^{j w k} centred on f0, then the same construction
    ctaps = taps * np.exp(2j * np.pi * f0 * np.arange(len(taps)))
    expected_c = _qa_fir_filter(despun, list(ctaps * np.exp(-2j * np.pi * f0 * np.arange(len(taps)))), decim)
    assert gb.rel_l2(got, expected_c[:len(got)]) < 2e-7


def test_rotator_matches_gnuradio_qa_definition():
    """gr-blocks/python/blocks/qa_rotator.py: ones through rotator_cc(phase_inc) give e^{j n phase_inc}, the first
    output unrotated - the phase convention the xlating filter's derotation inherits (oracle: exact phase)."""
    n, f, fs = 1000, 0.0125, 1.0
    inc = 2 * np.pi * f / fs
    expected = np.exp(1j * inc * np.arange(n))
    # unit taps, decimation 1, centre frequency -f: y[i] = x[i] e^{+j inc i}
    got = gb.freq_xlating_fir(np.ones(n, np.complex64), np.array([1.0]), 1, -f, fs)
    assert np.abs(got - expected).max() < 1e-12


def test_moving_average_matches_gnuradio_qa_construction():
    """gr-blocks/python/blocks/qa_moving_average.py test_01 / test_02: N samples through moving_average(length, scale)
    = scale * running sum of the last `length` inputs (zero history)."""
    rng = np.random.default_rng(0)
    for length, scale in ((100, 1.0 / 100), (10000, 1e-4), (7, 3.0)):
        x = rng.standard_normal(2500)
        pad = np.concatenate([np.zeros(length - 1), x])
        expected = scale * np.array([pad[i:i + length].sum() for i in range(len(x))])
        assert np.abs(gb.moving_average(x, length, scale) - expected).max() < 1e-9


def test_pfb_channelizer_matches_gnuradio_qa_construction():
    """gr-filter/python/filter/qa_pfb_channelizer.py test_0000: M = 5 channels of fs = 5000, one tone per channel at
    freqs = [-230, 121, 110, -513, 203] Hz around the channel centres, low_pass_2(1, M fs, fs/2, fs/10, 80,
    BLACKMAN_hARRIS) prototype; every output port carries its tone, delayed by the filter: expected = e^{j 2 pi f t}
    with t starting (tpf - 1) / 2 samples early - compared after the filter has filled."""
    M, fs, N = 5, 5000.0, 1000
    ifs = M * fs
    taps = fd.low_pass_2(1, ifs, fs / 2, fs / 10, 80.0, fd.WIN_BLACKMAN_HARRIS).astype(np.float64)
    freqs = [-230.0, 121.0, 110.0, -513.0, 203.0]
    t = np.arange(N * M) / ifs
    x = np.zeros(N * M, complex)
    for i, f in enumerate(freqs):
        # channel i sits at i * fs (bins above M/2 are the negative frequencies, rc_frontend/receiver.py:373-375)
        centre = i * fs if i <= M // 2 else (i - M) * fs
        x += np.exp(2j * np.pi * (centre + f) * t)
    y = gb.pfb_channelizer(x.astype(np.complex64), taps, M)
    tpf = int(np.ceil(len(taps) / float(M)))
    L = y.shape[1]
    for i, f in enumerate(freqs):
        # group delay of the prototype = (len(taps) - 1) / 2 input samples; the commutator advances the input by M - 1
        tt = (np.arange(L) * M + (M - 1) - (len(taps) - 1) / 2.0) / ifs
        centre = i * fs if i <= M // 2 else (i - M) * fs
        # output n of bin i = the tone at the (delayed) time of frame n with the bin centre mixed out at frame rate:
        # e^{j 2 pi (centre + f) tt} e^{-j 2 pi centre n M / ifs}
        expected = np.exp(2j * np.pi * f * tt) * np.exp(2j * np.pi * centre * ((M - 1) - (len(taps) - 1) / 2.0) / ifs)
        got = y[i]
        err = np.abs(got[2 * tpf:] - expected[2 * tpf:]).max()
        assert err < 2e-3, (i, err)       # QA: assertComplexTuplesAlmostEqual(..., 3); stop-band leakage of the others


# ---------------------------------------------------------------------------------------------------------------------
# post-demod stages: the restatements against independent library implementations
# ---------------------------------------------------------------------------------------------------------------------
def test_post_demod_restatements_match_scipy():
    from scipy.signal import lfilter, upfirdn
    rng = np.random.default_rng(4)
    x = rng.standard_normal(3000)
    hp = fd.high_pass(1, 25000, 300, 30, fd.WIN_HAMMING, 6.76)
    assert len(hp) == 2007 and abs(hp.astype(np.float64).sum()) < 1e-3          # DC removed
    assert np.abs(gb.fir_filter_fff(x, hp) - lfilter(hp.astype(np.float64), 1, x)).max() < 1e-12
    assert np.abs(gb.fir_filter_fff(x, hp, decim=3) - lfilter(hp.astype(np.float64), 1, x)[::3]).max() < 1e-12
    i, d, rt = fd.rational_resampler_taps(8000, 25000)
    assert (i, d, len(rt)) == (8, 25, 821)
    y = gb.rational_resampler_fff(x, i, d, rt)
    assert len(y) == (len(x) * 8 + 24) // 25
    assert np.abs(y - upfirdn(rt.astype(np.float64), x, i, d)[:len(y)]).max() < 1e-12
    b, a = fd.fm_deemph_taps(25000.0, 75e-6)
    assert np.abs(gb.iir_filter_ffd(x, b, a) - lfilter(b, a, x)).max() < 1e-12
    # de-emphasis: unity gain at DC, -3 dB near 1 / (2 pi tau) = 2122 Hz
    w = 2 * np.pi * 2122.0 / 25000.0
    hz = (b[0] + b[1] * np.exp(-1j * w)) / (1 + a[1] * np.exp(-1j * w))
    assert abs((b[0] + b[1]) / (1 + a[1]) - 1.0) < 1e-12 and abs(20 * np.log10(abs(hz)) + 3.0) < 0.1
    # symbol filter (p25_control_demod.py:130-133) and the squelch state machine
    assert np.allclose(gb.fir_filter_fff(np.ones(10), np.full(5, 0.2)), [0.2, 0.4, 0.6, 0.8, 1, 1, 1, 1, 1, 1])
    z = np.concatenate([np.zeros(50), 0.1 * np.ones(500), np.zeros(3000)]).astype(complex)
    out, keep = gb.pwr_squelch_cc(z, -40.0, 0.01, 0, True)
    assert not keep[:50].any() and keep[60:550].all() and not keep[-1000:].any() and len(out) == keep.sum()
    # Kaiser tap count of the resampler prototype: firdes.compute_ntaps with max_attenuation = beta / 0.1102 + 8.7
    assert fd.compute_ntaps(8.0, 0.032, fd.WIN_KAISER, 7.0) == 821
