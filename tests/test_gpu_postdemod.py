"""K6 parity: the data-parallel post-demod stages through the C ABI vs the float64 oracle.

Tolerance: relative L2 <= 1e-5 per row for filtered float outputs whose energy is not removed by the filter; the FM
stages are compared on the wrapped phase difference like K1 / K2.  Streaming: any split of the rows into calls must
give the samples of one call (bit exact for the FIR stages; the IIR scans differ by fp64 rounding only)."""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd
from radiocapture_rf_b200 import firdes
from radiocapture_rf_b200.engine import PostDemod

pytestmark = pytest.mark.gpu


def _c4fm_rows(rows, n, rate=25000.0, seed=1):
    """4-level FSK at 4800 baud (+-600 / +-1800 Hz) + noise per row, small per-row carrier offsets."""
    rng = np.random.default_rng(seed)
    out = np.empty((rows, n), np.complex64)
    sps = rate / 4800.0
    for r in range(rows):
        nsym = int(n / sps) + 2
        sym = rng.integers(0, 4, nsym)
        dev = np.array([-1800.0, -600.0, 600.0, 1800.0])[sym]
        f = np.repeat(dev, int(np.ceil(sps)))[:n] + rng.uniform(-200, 200)
        ph = 2 * np.pi * np.cumsum(f) / rate
        x = np.exp(1j * ph) * 0.5 + (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01
        out[r] = x.astype(np.complex64)
    return out


def _fm_voice_rows(rows, n, rate=25000.0, seed=2):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / rate
    out = np.empty((rows, n), np.complex64)
    for r in range(rows):
        audio = 0.5 * np.sin(2 * np.pi * (700 + 150 * r) * t) + 0.3 * np.sin(2 * np.pi * (1900 - 100 * r) * t) + \
            0.2 * np.sin(2 * np.pi * 120.0 * t)       # 120 Hz hum: removed by the 300 Hz high-pass
        ph = 2 * np.pi * 2500.0 * np.cumsum(audio) / rate
        x = 0.4 * np.exp(1j * ph) + (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.003
        out[r] = x.astype(np.complex64)
    return out


def _fm_err(a, ref, gain):
    d = (np.asarray(a, np.float64) - ref) / gain
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(d) / max(np.linalg.norm(ref / gain), 1e-30)


@pytest.mark.parametrize("rows,n,blocks", [(3, 4000, None), (5, 3001, [1, 7, 64, 500, 1429, 1000]), (1, 300, None)])
def test_p25_c4fm_front_matches_oracle(engine, rows, n, blocks):
    rate = 25000.0
    x = _c4fm_rows(rows, n)
    taps = fd.low_pass_2(1.0, rate, 6250.0, 500.0, 30.0, fd.WIN_BLACKMAN)
    assert len(taps) == 69
    gain = rate / (2 * np.pi * 600.0)
    pd = PostDemod.p25_c4fm(engine, rows, rate, 4800, 600.0, prefilter_taps=taps, probe_len=1000, probe_scale=1e-3)
    if blocks is None:
        sym, probe = pd.process(x, want_probe=True)
    else:
        parts, pos = [], 0
        for b in blocks:
            s, probe = pd.process(x[:, pos:pos + b], want_probe=True)
            parts.append(s)
            pos += b
        assert pos == n
        sym = [np.concatenate([p[r] for p in parts]) for r in range(rows)]
    for r in range(rows):
        _, fm, ref = gb.p25_c4fm_front(x[r], taps, gain, 5)
        assert len(sym[r]) == n
        # boxcar of FM: linear in the phase differences -> compare like FM (wrapped per sample is not meaningful after
        # the average; the signals here never wrap)
        err = np.linalg.norm(sym[r] - ref) / np.linalg.norm(ref)
        assert err <= 1e-5, "row %d symbol-filter err %g" % (r, err)
        want = 1e-3 * fm[max(0, n - 1000):].sum()
        assert abs(probe[r] - want) <= 1e-5 * max(1.0, abs(want)) + 2e-5 * gain
    pd.close()


def test_p25_split_is_bit_exact(engine):
    x = _c4fm_rows(2, 2000, seed=9)
    a = PostDemod.p25_c4fm(engine, 2)
    b = PostDemod.p25_c4fm(engine, 2)
    one = a.process(x)
    parts = [b.process(x[:, :777]), b.process(x[:, 777:778]), b.process(x[:, 778:])]
    for r in range(2):
        assert np.array_equal(one[r], np.concatenate([p[r] for p in parts]))
    a.close()
    b.close()


@pytest.mark.parametrize("rows,n,blocks", [(2, 6000, None), (3, 5000, [1, 24, 975, 2500, 1500])])
def test_analog_fm_chain_matches_oracle(engine, rows, n, blocks):
    rate = 25000.0
    x = _fm_voice_rows(rows, n)
    audio_taps = fd.fm_demod_audio_taps(rate, rate * 0.25, rate * 0.25 + 2000, 8.0)
    hp = fd.high_pass(1, rate, 300, 30, fd.WIN_HAMMING, 6.76)
    resamp = fd.rational_resampler_taps(8000, 25000)
    assert len(hp) == 2007 and resamp[0] == 8 and resamp[1] == 25 and len(resamp[2]) == 821
    pd = PostDemod.analog_fm(engine, rows, rate, 8000, audio_taps=audio_taps, hp_taps=hp, resampler=resamp)
    if blocks is None:
        out = pd.process(x)
    else:
        parts, pos = [], 0
        for b in blocks:
            parts.append(pd.process(x[:, pos:pos + b]))
            pos += b
        out = [np.concatenate([p[r] for p in parts]) for r in range(rows)]
    for r in range(rows):
        ref = gb.analog_fm_chain(x[r], rate, audio_taps=audio_taps, hp_taps=hp, resamp=resamp)
        assert ref["keep"].all()                       # -100 dB squelch: open from the first sample on
        assert len(out[r]) == len(ref["audio"]) == (n * 8 + 24) // 25
        # the high-pass removes the hum and the group delay of 2007 + 821 taps is still filling: compare where the
        # chain has settled, relative to the settled audio
        lo = 700
        err = np.linalg.norm(out[r][lo:] - ref["audio"][lo:]) / np.linalg.norm(ref["audio"][lo:])
        assert err <= 1e-5, "row %d audio err %g" % (r, err)
        assert np.abs(out[r][:lo] - ref["audio"][:lo]).max() <= 1e-5 * np.abs(ref["audio"]).max()
    pd.close()


def test_squelch_gate_drops_muted_samples(engine):
    """pwr_squelch_cc gate=True: samples are dropped while the smoothed power is below the threshold - the stream
    gets shorter (logging_receiver.py:211)."""
    rate, n = 25000.0, 4000
    x = _fm_voice_rows(1, n, seed=5)
    x[0, 1000:2500] *= 1e-4                                   # a fade 80 dB down
    audio_taps = fd.fm_demod_audio_taps(rate, 6250, 8250, 8.0)
    hp = fd.high_pass(1, rate, 300, 300, fd.WIN_HAMMING, 6.76)    # short high-pass: this test is about the gate
    resamp = fd.rational_resampler_taps(8000, 25000)
    pd = PostDemod.analog_fm(engine, 1, rate, 8000, squelch_db=-40.0, squelch_alpha=0.01, audio_taps=audio_taps,
                             hp_taps=hp, resampler=resamp)
    out = pd.process(x)[0]
    ref = gb.analog_fm_chain(x[0], rate, squelch_db=-40.0, squelch_alpha=0.01, audio_taps=audio_taps, hp_taps=hp,
                             resamp=resamp)
    kept = int(ref["keep"].sum())
    assert 2000 < kept < n - 500
    assert len(out) == len(ref["audio"]) == (kept * 8 + 24) // 25
    err = np.linalg.norm(out - ref["audio"]) / np.linalg.norm(ref["audio"])
    assert err <= 1e-5
    pd.close()


def test_post_demod_from_device_resident_pfb_rows(engine):
    """Rows of a device-resident PFB IQ output feed the chain without leaving the GPU (row_map picks the bins)."""
    from oracle import synth
    from radiocapture_rf_b200.engine import PfbChannelizer, OUT_IQ
    nch, frames = 64, 3000
    taps = fd.pfb_prototype(nch, 8)
    x, _ = synth.pfb_stream(nch * frames, 25000.0 * nch, nch, 31, active_every=2)
    ch = PfbChannelizer(engine, nch, taps, OUT_IQ, 1.0)
    d_in = engine.to_device(x)
    d_iq = engine.dev_alloc(nch * frames * 8)
    assert ch.process_device(d_in, len(x), d_iq, None, frames) == frames
    bins = [1, 5, 33]
    pd = PostDemod.p25_c4fm(engine, len(bins))
    d_out = engine.dev_alloc(len(bins) * frames * 4)
    nout = pd.process_device(d_iq, frames, frames, d_out, frames, row_map=bins)
    assert nout == [frames] * len(bins)
    got = engine.to_host(d_out, (len(bins), frames), np.float32)
    iq = engine.to_host(d_iq, (nch, frames), np.complex64)
    ptaps = firdes.low_pass_2(1.0, 25000.0, 6250.0, 500.0, 30.0, firdes.WIN_BLACKMAN)
    for i, m in enumerate(bins):
        _, _, ref = gb.p25_c4fm_front(iq[m], ptaps, 25000.0 / (2 * np.pi * 600.0), 5)
        err = np.linalg.norm(got[i] - ref) / np.linalg.norm(ref)
        assert err <= 2e-5, "bin %d err %g" % (m, err)
    pd.close()
