"""CPU: host logic of the drop-in frontend (rc_frontend/receiver.py + channel.py mirrors) with a recording
test double instead of the GPU engine - the control plane, not the DSP, is what is checked here."""
import importlib.util
import os
import threading
import time

import numpy as np
import pytest

from radiocapture_rf_b200 import firdes
from radiocapture_rf_b200.receiver import receiver


class FakeBank(object):
    def __init__(self):
        self.calls = []
        self.chans = {}
        self.next = 1

    def open(self, decim, taps, center_freq, samp_rate, out_mask=1, fm_gain=1.0):
        cid = self.next
        self.next += 1
        self.chans[cid] = dict(decim=decim, ntaps=len(taps), f0=center_freq, fs=samp_rate)
        self.calls.append(("open", cid, decim, len(taps), center_freq))
        return cid

    def retune(self, cid, f0):
        self.chans[cid]["f0"] = f0
        self.calls.append(("retune", cid, f0))

    def set_taps(self, cid, taps):
        self.chans[cid]["ntaps"] = len(taps)
        self.calls.append(("set_taps", cid, len(taps)))

    def close(self, cid):
        del self.chans[cid]
        self.calls.append(("close", cid))

    def set_input_format(self, fmt, offset=0.0, scale=1.0):
        self.fmt = (fmt, offset, scale)
        self.calls.append(("set_input_format", fmt, offset, scale))

    def process(self, iq):
        raw = getattr(self, "fmt", None) is not None
        self.last_n = len(iq) // 2 if raw else len(iq)
        self.blocks = getattr(self, "blocks", [])
        self.blocks.append((np.asarray(iq).dtype, len(iq)))

    def pull(self, cid, which=1):
        return np.zeros(self.last_n // self.chans[cid]["decim"], np.complex64)


class FakeEngine(object):
    def __init__(self, device=0):
        self.device = device
        self.bank = FakeBank()

    def make_bank(self):
        return self.bank

    def close(self):
        pass


class Cfg(object):
    frontend_mode = "xlat"
    receiver_split2 = False
    scan_mode = False
    sources = {
        0: {"type": "push", "center_freq": 855050000, "samp_rate": 2400000},
        1: {"type": "push", "center_freq": 857450000, "samp_rate": 2400000},
    }


def make(**kw):
    cfg = Cfg()
    cfg.sources = {k: dict(v) for k, v in Cfg.sources.items()}
    return receiver(config=cfg, sink="capture", engine_factory=FakeEngine, **kw)


def test_channel_request_lifecycle_and_reuse():
    tb = make(use_zmq=False)
    try:
        assert tb.handler("create,0,12500,854987500") .startswith("na,")       # connect must precede create
        assert len([c for c in tb.channels.values() if c.in_use]) == 0
        r = tb.handler("connect")
        assert r == "connect,0"
        r = tb.handler("create,0,12500,854987500").split(",")
        assert r[0] == "create" and 10000 <= int(r[2]) <= 60000
        bid = r[1]
        ch = tb.channels[bid]
        assert ch.source_id == 0 and ch.offset == 854987500 - 855050000 and ch.in_use
        assert ch.decim == 96 and len(ch.taps) == 349                         # channel.py:31-33
        bank = tb.sources[0]["block"].bank
        # the channel built for the refused create was parked (Appendix C.10) and is now reused
        assert bank.calls[0][:4] == ("open", 1, 96, 349) and bank.calls[-1] == ("retune", 1, -62500.0)
        # nearest-centre source pick: 856.3 MHz is inside both sources, closer to source 1
        r2 = tb.handler("create,0,12500,856300000").split(",")
        assert tb.channels[r2[1]].source_id == 1
        # out of range
        assert tb.handler("create,0,12500,900000000") == "na,900000000"
        # release parks the channel; the next create of the same (source, rate) reuses it via set_offset
        assert tb.handler("release,0,%s" % bid) == "release,%s" % bid
        assert not tb.channels[bid].in_use and tb.channels[bid].channel_close_time > 0
        r3 = tb.handler("create,0,12500,855000000").split(",")
        assert r3[1] == bid and int(r3[2]) == int(r[2])
        assert bank.calls[-1] == ("retune", 1, -50000.0)
        assert tb.channels[bid].channel_close_time == 0
        # unknown release is idempotent and lock-balanced (Appendix C.3)
        assert tb.handler("release,0,nonexistent") == "release,nonexistent"
        assert tb.access_lock.acquire(blocking=False)
        tb.access_lock.release()
        assert tb.handler("hb,0") == "hb,0"
        assert tb.handler("hb,7") == "fail,7"
        assert tb.handler("hb,x") == "fail,0"
        assert tb.handler("offset,0,%s,0.1" % bid) == "offset,0"
        assert tb.handler("scan_mode_set_freq,855000000") == "success"
        assert tb.handler("quit,0") == "quit,0"
        assert 0 not in tb.clients and not tb.channels[bid].in_use
        assert tb.handler("garbage") == "na"
    finally:
        tb.stop()


def test_heartbeat_expiry_and_idle_reaping():
    tb = make(use_zmq=False)
    try:
        tb.handler("connect")
        bid = tb.handler("create,0,12500,854987500").split(",")[1]
        now = time.time()
        tb.housekeeping(now + 4.0)
        assert 0 in tb.clients and tb.channels[bid].in_use
        tb.housekeeping(now + 6.0)                      # > 5 s silent: client dropped, channels released
        assert 0 not in tb.clients and not tb.channels[bid].in_use
        closed = tb.channels[bid].channel_close_time
        tb.last_channel_cleanup = closed - 100
        tb.housekeeping(closed + 5.0)                   # idle < 10 s: kept
        assert bid in tb.channels
        tb.last_channel_cleanup = closed - 100
        tb.housekeeping(closed + 11.0)                  # idle > 10 s and sweep due: destroyed
        assert bid not in tb.channels
        assert tb.sources[0]["block"].bank.calls[-1][0] == "close"
    finally:
        tb.stop()


def test_scan_mode_relative_offsets_and_afc():
    cfg = Cfg()
    cfg.sources = {0: dict(Cfg.sources[0])}
    cfg.scan_mode = True
    tb = receiver(config=cfg, sink="capture", engine_factory=FakeEngine, use_zmq=False)
    try:
        tb.handler("connect")
        bid = tb.handler("create,0,12500,-300000").split(",")[1]   # < 10 MHz: offset relative to source 0
        assert tb.channels[bid].offset == -300000
        assert tb.source_offset(bid, 2.0) is False                   # AFC disabled in scan mode
    finally:
        tb.stop()
    tb = make(use_zmq=False)
    try:
        tb.handler("connect")
        bid = tb.handler("create,0,12500,854987500").split(",")[1]
        assert tb.source_offset(bid, 0.1) is True                    # 0.4 Hz: below the 5 Hz deadband
        assert "accumulated_offset" not in tb.sources[0]
        assert tb.source_offset(bid, 2.0) is True                    # 100 Hz step
        assert tb.sources[0]["accumulated_offset"] == 100.0
        assert tb.source_offset("nope", 2.0) is False
    finally:
        tb.stop()


def test_data_plane_delivers_per_channel_and_channel_surface():
    tb = make(use_zmq=False)
    try:
        tb.handler("connect")
        bid = tb.handler("create,0,12500,854987500").split(",")[1]
        ch = tb.channels[bid]
        tb.push(0, np.zeros(96 * 100, np.complex64))
        assert ch.sink.samples == 100
        # method surface of rc_frontend/channel.py:39-67
        assert ch.get_samp_rate() == 2400000 and ch.get_channel_rate() == 12500 and ch.get_offset() == -62500
        ch.set_channel_rate(12500)      # channel.py:53-55 redesign with low_pass(1, fs, (rate-2000)/2, 4000)
        assert len(ch.taps) == len(firdes.low_pass(1, 2400000, (12500 - 2000) / 2, 4000))
        ch.set_offset(1000)
        assert tb.sources[0]["block"].bank.calls[-1] == ("retune", 1, 1000.0)
        assert "port:%s" % ch.port in str(ch) and repr(ch).startswith("<Channel")
        blob = tb.describe()                                          # redis_channel_publisher.py:63-90 schema
        for k in ("instance_uuid", "start_time", "current_time", "hostname", "pid", "address", "port",
                  "channel_count", "source_count", "sources"):
            assert k in blob
        assert blob["sources"] == [[855050000, 2400000], [857450000, 2400000]] and blob["channel_count"] == 1
        assert tb.pfb_bin_for(0, 855050000 + 810000) == (2, 10000, False)
        assert tb.pfb_bin_for(0, 855050000 - 390000)[0] == 5
    finally:
        tb.stop()


def test_sinks_zmq_and_udp_wire_format():
    import socket
    import zmq
    from radiocapture_rf_b200.sinks import UdpSink, ZmqPubSink
    x = (np.arange(400) + 1j * np.arange(400)).astype(np.complex64)
    port = 23000 + os.getpid() % 20000
    s = ZmqPubSink(port, bind_host="127.0.0.1")
    sub = zmq.Context.instance().socket(zmq.SUB)
    sub.setsockopt(zmq.SUBSCRIBE, b"")
    sub.setsockopt(zmq.RCVTIMEO, 2000)
    sub.connect("tcp://127.0.0.1:%d" % port)
    time.sleep(0.3)
    s.send(x)
    got = np.frombuffer(sub.recv(), np.complex64)        # raw complex64, no framing (channel.py:36)
    assert np.array_equal(got, x)
    with pytest.raises(RuntimeError):
        ZmqPubSink(port, bind_host="127.0.0.1")          # port busy -> RuntimeError like the GNU Radio block
    s.close()
    sub.close()
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(2.0)
    u = UdpSink(rx.getsockname()[1])
    u.send(x)
    sizes, data = [], b""
    for _ in range(3):
        d = rx.recv(4096)
        sizes.append(len(d))
        data += d
    assert sizes == [1472, 1472, 256]                    # 184 samples per datagram (moto_control_demod.py:119)
    assert np.array_equal(np.frombuffer(data, np.complex64), x)
    u.close()
    assert rx.recv(4096) == b""                          # zero-length datagram = EOF
    rx.close()


REF_CONNECTOR = "/root/reference/frontend_connector.py"


@pytest.mark.skipif(not os.path.exists(REF_CONNECTOR), reason="reference tree not mounted (GPU box)")
def test_unmodified_reference_frontend_connector_works_against_us():
    """The reference's own client stub (frontend_connector.py:13-229), loaded UNMODIFIED from the mounted
    reference tree, drives our REP server: connect/create/hb/release over ZMQ REQ/REP."""
    spec = importlib.util.spec_from_file_location("ref_frontend_connector", REF_CONNECTOR)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tb = make(bind="tcp://127.0.0.1:0")
    tb.serve_in_thread()
    host, port = tb.endpoint().replace("tcp://", "").rsplit(":", 1)

    class Rcm(object):  # stands in for redis_channelizer_manager.get_channelizer_for_frequency (:52-76)
        def get_channelizer_for_frequency(self, freq):
            return host, int(port)

    fc = mod.frontend_connector("test-uuid", Rcm())
    try:
        channel_id, chan_port = fc.create_channel(12500, 854987500)
        assert channel_id in tb.channels and int(chan_port) == tb.channels[channel_id].port
        time.sleep(0.8)                                    # a few 4 Hz heartbeats
        assert fc.my_client_id in tb.client_hb
        assert time.time() - tb.client_hb[fc.my_client_id] < 0.7
        assert fc.report_offset(0.0) is True
        assert fc.release_channel() == channel_id
        assert not tb.channels[channel_id].in_use
        assert fc.create_channel(12500, 990000000) == (False, False)   # out of range -> 'na'
    finally:
        fc.exit()
        time.sleep(0.4)
        tb.stop()


def test_pfb_mode_bin_arithmetic_and_host_fallback():
    """frontend_mode 'pfb' (rc_frontend/receiver.py:365-383): bin / residual arithmetic incl. the negative-bin wrap
    and non-integer bin widths; with an injected (CPU) engine the request is served by the DDC bank directly."""
    cfg = Cfg()
    cfg.frontend_mode = "pfb"
    cfg.sources = {0: {"type": "push", "center_freq": 855050000, "samp_rate": 2400000},
                   1: {"type": "push", "center_freq": 860000000, "samp_rate": 2850000}}
    tb = receiver(config=cfg, sink="capture", engine_factory=FakeEngine, use_zmq=False)
    try:
        assert tb.pfb_bin_for(0, 855050000 + 437500) == (1, 37500, False)
        assert tb.pfb_bin_for(0, 855050000 - 1037500) == (3, 162500, False)      # bin -3 wraps to 3 of 6
        assert tb.pfb_bin_for(0, 855050000 + 1190000)[0] == 3                     # +3 is the same (Nyquist) bin
        # 2.85 Msps: 7 bins of 407142.857 Hz (the reference divides by the nominal 400 kHz, Appendix C.4)
        w = 2850000 / 7.0
        chan, off, edge = tb.pfb_bin_for(1, 860000000 - 500000, w)
        assert chan == 6 and abs(off - (-500000 + w)) < 1e-6 and not edge
        assert tb.handler("connect") == "connect,0"
        r = tb.handler("create,0,12500,855487500").split(",")
        assert r[0] == "create"
        ch = tb.channels[r[1]]
        assert ch.source_id == 0 and ch.decim == 96 and ch.offset == 437500      # fallback: single-stage channel
    finally:
        tb.stop()


def test_file_source_in_the_sdr_wire_format(tmp_path):
    """A source whose samples are the SDR's own bytes ("format": "u8", an RTL-SDR capture): the reader hands the bank
    interleaved uint8 (I, Q) pairs untouched (rcb_ddc_set_input_format converts on the device), counts complex samples,
    drops a trailing odd byte; an unknown format is refused loudly (so is pfb mode, which needs the GPU engine)."""
    import time
    from radiocapture_rf_b200 import _lib
    from radiocapture_rf_b200.receiver import SourceStream
    raw = (np.arange(2 * 5000 + 1) % 251).astype(np.uint8)
    path = tmp_path / "capture.u8"
    raw.tofile(path)
    cfg = {"type": "file", "path": str(path), "format": "u8", "center_freq": 855e6, "samp_rate": 2400000}
    src = SourceStream(7, cfg, block_samples=2048, engine_factory=FakeEngine)
    try:
        bank = src.bank
        assert bank.calls[0] == ("set_input_format", _lib.FMT_U8, -127.4, 1.0 / 128.0)
        src.start()
        for _ in range(200):
            if src.samples_in >= 5000:
                break
            time.sleep(0.01)
        assert src.samples_in == 5000
        assert [b[0] for b in bank.blocks] == [np.dtype(np.uint8)] * 3
        assert [b[1] for b in bank.blocks] == [4096, 4096, 2 * 5000 - 8192]
    finally:
        src.stop()
    with pytest.raises(ValueError):
        SourceStream(8, dict(cfg, format="u12"), engine_factory=FakeEngine)
