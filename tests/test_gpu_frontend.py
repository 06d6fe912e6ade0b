"""GPU: the drop-in surfaces end to end - receiver RPC -> channel -> GPU DDC bank -> sink, and the
fft_vector / fft_peak_detection mirrors - checked against the oracle on the same synthetic IQ."""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.receiver import receiver

pytestmark = pytest.mark.gpu


class Cfg(object):
    frontend_mode = "xlat"
    receiver_split2 = False
    scan_mode = False
    sources = {0: {"type": "push", "center_freq": 855050000, "samp_rate": 2400000}}


def test_create_channel_rpc_delivers_oracle_samples(built_lib):
    """`create,<cid>,12500,854987500` (configs/config_denver_dev_den817.py:32,127) -> the narrowband stream
    the backend demod would receive == freq_xlating_fir_filter_ccc(96, 349 taps, -62.5 kHz) of the oracle."""
    cfg = Cfg()
    cfg.sources = {0: dict(Cfg.sources[0])}
    tb = receiver(config=cfg, sink="capture", use_zmq=False)
    try:
        assert tb.handler("connect") == "connect,0"
        r = tb.handler("create,0,12500,854987500").split(",")
        assert r[0] == "create"
        ch = tb.channels[r[1]]
        r2 = tb.handler("create,0,12500,855487500").split(",")       # a second channel on the same source
        ch2 = tb.channels[r2[1]]
        x, fs, offs = synth.cfg1(96 * 3000, seed=1)
        for blk in np.array_split(x, 7):                              # ragged blocks, like ZMQ messages
            tb.push(0, blk)
        y = ch.sink.data()
        decim, taps = fd.channel_taps(fs, 12500)
        ref = gb.freq_xlating_fir(x, taps, decim, -62500.0, fs)
        n = min(len(y), len(ref))
        assert n >= 2999
        assert gb.rel_l2(y[:n], ref[:n]) <= 1e-5
        y2 = ch2.sink.data()
        ref2 = gb.freq_xlating_fir(x, taps, decim, 437500.0, fs)
        assert gb.rel_l2(y2[:n], ref2[:n]) <= 1e-5
        # release + re-create on another frequency reuses the parked channel via set_offset (receiver.py:311-319)
        assert tb.handler("release,0,%s" % r[1]).startswith("release")
        r3 = tb.handler("create,0,12500,855062500").split(",")
        assert r3[1] == r[1]
        before = len(ch.sink.data())
        tb.push(0, x[:96 * 500])
        y3 = ch.sink.data()[before:]
        assert len(y3) == 500
        # the +12.5 kHz carrier of the synthetic scene now sits inside the channel
        assert np.abs(y3[50:]).mean() > 0.01
    finally:
        tb.stop()


def test_pfb_mode_two_stage_channel_matches_oracle(built_lib):
    """frontend_mode = 'pfb' (rc_frontend/receiver.py:242-261, 343-423): pfb.channelizer_ccf with 400 kHz bins, then a
    second-stage channel (59-tap xlating FIR, decimation 16) on the bin - both stages on the GPU, the bin rows never
    leave the device.  Oracle = the same two GNU Radio blocks in float64."""
    cfg = Cfg()
    cfg.frontend_mode = "pfb"
    cfg.sources = {0: dict(Cfg.sources[0])}
    tb = receiver(config=cfg, sink="capture", use_zmq=False)
    try:
        assert tb.handler("connect") == "connect,0"
        want = {855487500: (1, 37500.0),      # +437.5 kHz -> bin 1, +37.5 kHz
                854987500: (0, -62500.0),     # -62.5 kHz  -> bin 0
                854012500: (3, 162500.0)}     # -1.0375 MHz -> bin -3 -> wraps to bin 3 of 6, +162.5 kHz
        chans = {}
        for f, (b, off) in want.items():
            r = tb.handler("create,0,12500,%d" % f).split(",")
            assert r[0] == "create"
            ch = tb.channels[r[1]]
            assert ch.pfb_id == b and abs(ch.offset - off) < 1e-6 and ch.decim == 16 and len(ch.taps) == 59
            chans[f] = ch
        x, fs, offs = synth.cfg1(6 * 16 * 1500, seed=1)
        for blk in np.array_split(x, 7):      # ragged blocks: frames are cut at arbitrary samples
            tb.push(0, blk)
        stream = tb.sources[0]["block"]
        bins = gb.pfb_channelizer(x, np.asarray(stream.pfb_taps, np.float64), 6)
        for f, (b, off) in want.items():
            y = chans[f].sink.data()
            ref = gb.freq_xlating_fir(bins[b], chans[f].taps, 16, off, 400000.0)
            n = min(len(y), len(ref))
            assert n >= 1490
            assert gb.rel_l2(y[:n], ref[:n]) <= 1e-5, (f, gb.rel_l2(y[:n], ref[:n]))
        # release + re-create inside the same bin reuses the parked second stage via set_offset (:386-394)
        first = [k for k, v in tb.channels.items() if v is chans[855487500]][0]
        assert tb.handler("release,0,%s" % first).startswith("release")
        r = tb.handler("create,0,12500,855500000").split(",")
        assert r[1] == first and abs(tb.channels[first].offset - 50000.0) < 1e-6
    finally:
        tb.stop()


def test_fft_vector_and_peak_detection_mirror(engine, tmp_path):
    """fft_vector.py -> /tmp/fft_source_<i> -> fft_peak_detection.py, GPU vs oracle: identical peak indices."""
    from radiocapture_rf_b200.fft_peak_detection import detect_peaks, load_vector
    from radiocapture_rf_b200.fft_vector import fft_vector
    length, nframes, fs, centre = 16384, 200, 2.4e6, 855.05e6
    x, truth = synth.scan_stream(length * nframes, fs, length, seed=44, ncarriers=8)
    tb = fft_vector(index=0, samp_rate=fs, length=length, nframes=nframes, avg=100, engine=engine)
    tb.path = str(tmp_path / "fft_source_0")
    vec = tb.run(x, write=True)
    assert np.array_equal(load_vector(tb.path), vec)                  # raw float32 file like blocks.file_sink
    ref = gb.fft_vector_flowgraph(x, length, fd.blackmanharris(length), nframes, 100)
    idx, hz = detect_peaks(vec, fs, centre)
    idx_ref, hz_ref = gb.peak_detect(ref.astype(np.float32), fs, centre)
    assert len(idx_ref) >= 4 and np.array_equal(idx, idx_ref) and np.array_equal(hz, hz_ref)
    # transforming every frame like GNU Radio does gives the same vector
    tb2 = fft_vector(index=0, samp_rate=fs, length=length, nframes=nframes, avg=100, engine=engine,
                     skip_discarded=False)
    vec2 = tb2.run(x)
    np.testing.assert_allclose(vec2, vec, atol=1e-4)


def test_file_source_in_u8_wire_format_matches_oracle(tmp_path):
    """f4 through the server's own ingest: a source entry with "format": "u8" (an RTL-SDR capture) - the reader thread
    hands the bytes to the GPU bank untouched, rcb_ddc_set_input_format converts on the device, and the channel's sink
    receives the DDC of the converted stream (float64 oracle, 1e-5)."""
    import time
    from oracle import gr_blocks as gb, gr_firdes as fd
    from radiocapture_rf_b200 import channel as channel_mod
    from radiocapture_rf_b200.receiver import SourceStream
    fs = 2400000
    n = 96 * 600
    t = np.arange(n)
    sig = 0.5 * np.exp(2j * np.pi * (-62500.0 / fs * t + 0.2 * np.sin(2 * np.pi * 2e-4 * t)))
    raw = np.clip(np.round(np.stack([sig.real, sig.imag], 1) * 127.0 + 127.4), 0, 255).astype(np.uint8).reshape(-1)
    path = tmp_path / "cap.u8"
    raw.tofile(path)
    cfg = {"type": "file", "path": str(path), "format": "u8", "center_freq": 855050000, "samp_rate": fs}
    src = SourceStream(90, cfg, block_samples=96 * 200)
    try:
        ch = channel_mod.channel(src, 0, 12500, fs, -62500, sink="capture")
        ch.start()
        src.start()
        for _ in range(500):
            if src.samples_in >= n:
                break
            time.sleep(0.01)
        assert src.samples_in == n
    finally:
        src.stop()
    y = ch.sink.data()
    xf = ((raw.astype(np.float32) - np.float32(127.4)) * np.float32(1 / 128.0)).view(np.complex64)
    decim, taps = fd.channel_taps(fs, 12500)
    ref = gb.freq_xlating_fir(xf, taps, decim, -62500.0, fs)
    m = min(len(y), len(ref))
    assert m >= 590
    assert gb.rel_l2(y[:m], ref[:m]) <= 1e-5
