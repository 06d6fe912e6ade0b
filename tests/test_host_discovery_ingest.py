"""CPU: the boundary pieces around the DSP that SURVEY 8(b) / (f-2) / a5 / a12 name - Redis discovery writer, ZMQ SUB
ingest of the wideband stream, receiver_split2 wiring - with in-process fakes (no redis server, no GPU)."""
import json
import threading
import time

import numpy as np
import pytest

from radiocapture_rf_b200 import firdes
from radiocapture_rf_b200.receiver import SourceStream, receiver
from radiocapture_rf_b200.redis_channel_publisher import redis_channel_publisher
from test_frontend_host import Cfg, FakeEngine, make


class FakeRedis(object):
    """The slice of redis.StrictRedis the reference uses: pipeline().sadd().set().execute()."""

    def __init__(self, fail=False):
        self.sets = {}
        self.kv = {}
        self.fail = fail
        self.executed = 0

    def pipeline(self):
        return _Pipe(self)


class _Pipe(object):
    def __init__(self, r):
        self.r = r
        self.ops = []

    def sadd(self, key, member):
        self.ops.append(("sadd", key, member))
        return self

    def set(self, key, value):
        self.ops.append(("set", key, value))
        return self

    def execute(self):
        if self.r.fail:
            raise ConnectionError("redis is down")
        for op, k, v in self.ops:
            if op == "sadd":
                self.r.sets.setdefault(k, set()).add(v)
            else:
                self.r.kv[k] = v
        self.r.executed += 1
        return [1] * len(self.ops)


def test_redis_publisher_writes_the_reference_blob():
    """rc_frontend/redis_channel_publisher.py:63-91: SADD channelizers <uuid>; SET <uuid> json{...} each second."""
    r = FakeRedis()
    tb = make(redis_client=r, index=1)
    try:
        pub = tb.redis_channel_publisher
        assert pub is not None and pub.index == 1
        tb.handler("connect")
        resp = tb.handler("create,0,12500,857400000").split(",")
        assert resp[0] == "create"
        assert pub.publish_once()
        assert r.sets["channelizers"] == {tb.instance_uuid}
        blob = json.loads(r.kv[tb.instance_uuid])
        for key in ("instance_uuid", "start_time", "current_time", "hostname", "pid", "address", "port",
                    "channel_count", "source_count", "sources", "index"):
            assert key in blob, key
        assert blob["instance_uuid"] == tb.instance_uuid
        assert blob["port"] == int(tb.endpoint().rsplit(":", 1)[1])          # what frontend_connector dials
        assert blob["channel_count"] == 1 and blob["source_count"] == 1 and blob["index"] == 1
        assert blob["sources"] == [[857450000, 2400000]]
        # the background thread publishes on its own (every `interval` seconds)
        n0 = r.executed
        t0 = time.time()
        while r.executed == n0 and time.time() - t0 < 5:
            time.sleep(0.05)
        assert r.executed > n0
        # a dead redis is logged, never raised (reference :90-93)
        r.fail = True
        assert pub.publish_once() is False
    finally:
        tb.stop()
    assert not pub.continue_running


def test_redis_publisher_requires_the_reference_arguments():
    with pytest.raises(Exception):
        redis_channel_publisher(channels={}, zmq_socket=object(), client=FakeRedis(), start=False)
    with pytest.raises(Exception):
        redis_channel_publisher(sources={}, zmq_socket=object(), client=FakeRedis(), start=False)
    with pytest.raises(Exception):
        redis_channel_publisher(sources={}, channels={}, client=FakeRedis(), start=False)


def test_wideband_zmq_ingest_reader(tmp_path):
    """a12: the SDR side publishes raw complex64 on ipc:///tmp/rx_source_<id> (rc_frontend/receiver.py:201, one
    message per work() chunk, no framing); the source reader re-blocks it and pushes every sample exactly once."""
    import zmq
    addr = "ipc://%s" % (tmp_path / "rx_source_t")
    cfg = {"type": "zmq", "address": addr, "center_freq": 855e6, "samp_rate": 2400000}
    got = []

    class Recorder(SourceStream):
        def push(self, iq):
            got.append(np.array(iq, copy=True))

    s = Recorder("t", cfg, engine_factory=FakeEngine, block_samples=4096)
    ctx = zmq.Context.instance()
    pub = ctx.socket(zmq.PUB)
    pub.bind(addr)
    s.start()
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)).astype(np.complex64)
    try:
        time.sleep(0.3)                      # PUB/SUB slow joiner
        # chunks of uneven size, like GNU Radio's scheduler produces
        pos, sizes = 0, [1000, 37, 4096, 8191, 5000, 3, 12000, 9673]
        assert sum(sizes) == 40000
        for n in sizes:
            pub.send(x[pos:pos + n].tobytes())
            pos += n
            time.sleep(0.01)
        t0 = time.time()
        while sum(len(g) for g in got) < 36000 and time.time() - t0 < 5:
            time.sleep(0.05)
    finally:
        s.stop()
        pub.close()
    y = np.concatenate(got)
    assert len(y) >= 36000 and all(len(g) >= 4096 for g in got)
    assert np.array_equal(y, x[:len(y)])     # in order, nothing dropped or duplicated, bytes untouched


def test_receiver_split2_builds_two_half_rate_sources():
    """a5, rc_frontend/receiver.py:205-237: every source becomes centre -+ fs/4 at fs/2, fed by a 19-tap half-band
    decimate-by-2 DDC pair on the parent; channel requests land on the half that contains them."""
    cfg = Cfg()
    cfg.sources = {0: {"type": "push", "center_freq": 855000000, "samp_rate": 8000000}}
    cfg.receiver_split2 = True
    tb = receiver(config=cfg, sink="capture", engine_factory=FakeEngine, use_zmq=False)
    try:
        assert len(tb.sources) == 2
        lo, hi = tb.sources[0], tb.sources[1]
        assert (lo["center_freq"], lo["samp_rate"]) == (853000000, 4000000)
        assert (hi["center_freq"], hi["samp_rate"]) == (857000000, 4000000)
        parent = tb.realsources[0]["block"]
        opens = [c for c in parent.bank.calls if c[0] == "open"]
        assert [(c[2], c[3], c[4]) for c in opens] == [(2, 19, -2000000.0), (2, 19, 2000000.0)]
        taps = firdes.low_pass(1, 8e6, 2e6, 1e6)
        assert len(taps) == 19
        tb.handler("connect")
        r = tb.handler("create,0,12500,856500000").split(",")
        assert r[0] == "create"
        ch = tb.channels[r[1]]
        assert ch.source_id == 1 and ch.offset == 856500000 - 857000000
        assert ch.decim == firdes.channel_decimation(4000000, 12500)
        # one wideband block in: both halves see half as many samples, the channel gets its share
        parent.push(np.zeros(64000, np.complex64))      # the SDR feeds the real source; the halves follow
        assert lo["block"].samples_in == 32000 and hi["block"].samples_in == 32000
        assert ch.sink.samples == 32000 // ch.decim
    finally:
        tb.stop()
