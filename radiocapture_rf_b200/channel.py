"""`channel` - the per-channel digital down-converter block of the frontend.

Mirrors rc_frontend/channel.py:17-67 (class channel(gr.top_block)): same constructor arguments, same
attributes (in_use, source_id, block_id, port, pfb_id, init_time, channel_close_time) and the same
method surface (get/set_samp_rate, get/set_channel_rate, set_taps, get/set_offset, start, stop,
destroy).  Instead of a private GNU Radio flowgraph (ZMQ SUB of the whole wideband stream ->
freq_xlating_fir_filter_ccc -> ZMQ PUB, channel.py:29-38) a channel here is one entry of the shared
GPU DDC bank of its source (K2, rcb_ddc_*): every channel of a source reads the one staged copy of
the wideband block.  The sample sink is unchanged: raw complex64 on a ZMQ PUB socket bound to
tcp://0.0.0.0:<port> (channel.py:36), or the legacy UDP plane (1472-byte datagrams of raw complex64,
moto_control_demod.py:119, top-level receiver.py:146).

Both historical constructor shapes are accepted:
    channel(parent, port, channel_rate, samp_rate, offset)     today's 5-argument form (channel.py:18);
                                                               `parent` is the source stream object or the
                                                               ipc:///tmp/rx_source_<k> address it serves
    channel(port, channel_rate, samp_rate, offset)             hier_block2 form of the pfb call site
                                                               (rc_frontend/receiver.py:403), bound later
                                                               with  receiver.connect(source, block)
"""
import time

from . import firdes
from .sinks import make_sink

_SOURCES_BY_ADDRESS = {}


def register_source(address, source):
    """Frontend registers each wideband source under the address the reference would publish it on
    (ipc:///tmp/rx_source_<id>, rc_frontend/receiver.py:201) so channel(address, ...) can find it."""
    _SOURCES_BY_ADDRESS[address] = source


def unregister_source(address):
    _SOURCES_BY_ADDRESS.pop(address, None)


class channel(object):
    def __init__(self, *args, **kwargs):
        sink = kwargs.pop("sink", "zmq")
        demod_gain = kwargs.pop("fm_gain", None)
        if kwargs:
            raise TypeError("unexpected arguments %r" % (kwargs,))
        if len(args) == 5:
            parent, port, channel_rate, samp_rate, offset = args
        elif len(args) == 4:
            parent = None
            port, channel_rate, samp_rate, offset = args
        else:
            raise TypeError("channel(parent_zmq_address, port, channel_rate, samp_rate, offset)")
        self.parent_zmq_address = parent
        self.samp_rate = samp_rate
        self.channel_rate = channel_rate
        self.port = port
        self.offset = offset
        self.in_use = False
        self.source_id = None
        self.block_id = None
        self.pfb_id = None
        self.fm_gain = demod_gain

        # channel.py:31-33 (integer decimation)
        self.decim = firdes.channel_decimation(samp_rate, channel_rate)
        self.taps = firdes.low_pass_2(1.0, float(samp_rate), channel_rate / 2, channel_rate / 2, 20.0,
                                      firdes.WIN_HAMMING)
        self.sink = make_sink(sink, port)
        self._source = None
        self._chan_id = None
        self._running = False
        if parent is not None:
            try:
                self._bind(parent)
            except Exception:   # unknown source / refused rcb_ddc_open: do not leak the bound PUB socket and its port
                self.sink.close()
                raise
        self.init_time = time.time()
        self.channel_close_time = 0

    def __str__(self):
        return "Channel: port:%s channel_rate:%s samp_rate:%s offset:%s init_time:%s" % (
            self.port, self.channel_rate, self.samp_rate, self.offset, self.init_time)

    def __repr__(self):
        return "<Channel port:%s channel_rate:%s samp_rate:%s offset:%s init_time:%s>" % (
            self.port, self.channel_rate, self.samp_rate, self.offset, self.init_time)

    # ---- binding to a wideband source (what self.connect(sub_source, prefilter, sink) did) -------
    def _bind(self, parent):
        src = _SOURCES_BY_ADDRESS.get(parent) if isinstance(parent, str) else parent
        if src is None:
            raise RuntimeError("no wideband source is published on %r" % (parent,))
        self._source = src
        self._chan_id = src.open_channel(self)

    def connect_source(self, source):
        """hier_block2 form: receiver.connect(source, block)."""
        if self._source is not None:
            raise RuntimeError("channel already connected")
        self._bind(source)

    # ---- reference method surface ----------------------------------------------------------------
    def get_samp_rate(self):
        return self.samp_rate

    def set_samp_rate(self, samp_rate):
        self.samp_rate = samp_rate
        # channel.py:50 redesigns with a different spec than the constructor (Appendix C.6): keep verbatim
        self.set_taps(firdes.low_pass(1, self.samp_rate, (self.channel_rate - 2000) / 2, 4000))

    def get_channel_rate(self):
        return self.channel_rate

    def set_channel_rate(self, channel_rate):
        self.channel_rate = channel_rate
        self.set_taps(firdes.low_pass(1, self.samp_rate, (self.channel_rate - 2000) / 2, 4000))

    def set_taps(self, taps):
        self.taps = taps
        if self._source is not None:
            self._source.set_channel_taps(self._chan_id, taps)

    def get_offset(self):
        return self.offset

    def set_offset(self, offset):
        self.offset = offset
        if self._source is not None:
            self._source.retune_channel(self._chan_id, offset)   # prefilter.set_center_freq(offset)

    def start(self):
        self._running = True

    def stop(self):
        self._running = False

    def wait(self):
        return None

    def destroy(self):
        self.stop()
        if self._source is not None and self._chan_id is not None:
            self._source.close_channel(self._chan_id)
        self._source = None
        self._chan_id = None
        if self.sink is not None:
            self.sink.close()
        self.sink = None

    # ---- data plane (called by the source pump with this block's narrowband samples) -------------
    def deliver(self, samples):
        if self._running and self.sink is not None and len(samples):
            self.sink.send(samples)
