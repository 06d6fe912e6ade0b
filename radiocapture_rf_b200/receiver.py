"""Frontend server - drop-in for rc_frontend/receiver.py behind the unmodified frontend_connector.py.

Mirrors `class receiver` (rc_frontend/receiver.py:29-475) and its `__main__` RPC loop (:477-700):
  * sources come from the same `rc_config.sources` schema (:54-61); instead of opening SDR hardware
    (:88-204, out of scope) every source is a SourceStream = one GPU handle fed with complex64 blocks
    from ZMQ (`ipc:///tmp/rx_source_<id>`, what the reference's SDR process publishes, :201), a raw
    .dat file (logging_receiver.py:107-109 format), a synthetic generator, or push();
  * connect_channel / connect_channel_xlat / release_channel / source_offset keep the reference's
    behaviour: nearest-centre source pick (:286-294), scan-mode relative offsets (:295-305), idle-channel
    reuse by (source, rate) with set_offset (:311-319), random port 10000-60000 with 3 tries (:321-329),
    uuid block ids (:332), parked-not-destroyed release (:424-435), 10 s idle reaping (:635-648);
  * handler() speaks the same comma-separated text RPC over ZMQ REP bound to tcp://0.0.0.0:0 (:44-46,
    :503-614): connect / create / release / hb / quit / offset / scan_mode_set_freq, errors in band,
    exactly one reply per request; clients silent for 5 s lose their channels (:651-680);
  * the Redis discovery blob (rc_frontend/redis_channel_publisher.py:63-90) is produced by
    describe() and handed to a pluggable publisher (redis is not a dependency of this package).
The DSP (one freq_xlating_fir_filter_ccc flowgraph + full-rate ZMQ copy per channel in the reference)
is the GPU DDC bank: one staged copy of each wideband block, all channels in one launch.
"""
import json
import logging
import os
import random
import socket
import threading
import time
import uuid

import numpy as np

from . import channel as channel_mod
from . import firdes
from .channel import channel
from ._lib import COPY_H2D, check
from .engine import DdcBank, Engine, PfbChannelizer, OUT_IQ


class SourceStream(object):
    """One wideband source: GPU handle + DDC bank + its channels.  Thread safe via a lock (the reference
    serialises on receiver.access_lock)."""

    def __init__(self, source_id, cfg, device=0, block_samples=None, engine_factory=None):
        self.source_id = source_id
        self.cfg = cfg
        self.center_freq = cfg["center_freq"]
        self.samp_rate = cfg["samp_rate"]
        self.address = "ipc:///tmp/rx_source_%s" % (source_id,)
        self.lock = threading.RLock()
        self.engine = (engine_factory or Engine)(device)
        self.bank = DdcBank(self.engine) if engine_factory is None else self.engine.make_bank()
        self.channels = {}          # chan_id -> channel object
        self.block_samples = block_samples or max(1 << 16, int(self.samp_rate // 20))
        self.samples_in = 0
        self._thread = None
        self._stop = threading.Event()
        self.post_push = []         # callables run at the end of push() under the lock (split2 feeds its halves here)
        # "format": "u8" | "s8" | "s16" - the source delivers the SDR's wire format (an RTL-SDR capture, a UHD sc8 / sc16
        # recording): the bytes go to the GPU as they are (rcb_ddc_set_input_format), a quarter / half of the PCIe traffic
        self.raw_format = cfg.get("format")
        self._raw_dtype = None
        if self.raw_format not in (None, "fc32", "complex64"):
            if self.raw_format not in Engine._FMT:
                raise ValueError("source %s: unknown sample format %r" % (source_id, self.raw_format))
            if not hasattr(self.bank, "set_input_format") or type(self) is not SourceStream:
                raise ValueError("source %s: format %r needs the xlat-mode GPU bank" % (source_id, self.raw_format))
            code, self._raw_dtype, off, scale = Engine._FMT[self.raw_format]
            self.bank.set_input_format(code, off, scale)
        channel_mod.register_source(self.address, self)

    # ---- channel management (called by channel objects) ------------------------------------------
    def open_channel(self, ch):
        with self.lock:
            cid = self.bank.open(ch.decim, ch.taps, float(ch.offset), float(ch.samp_rate), OUT_IQ, 1.0)
            self.channels[cid] = ch
            return cid

    def retune_channel(self, cid, offset):
        with self.lock:
            self.bank.retune(cid, float(offset))

    def set_channel_taps(self, cid, taps):
        with self.lock:
            self.bank.set_taps(cid, taps)

    def close_channel(self, cid):
        with self.lock:
            self.channels.pop(cid, None)
            self.bank.close(cid)

    def set_center_freq(self, freq, chan=0):
        """SDR retune (AFC / scan mode).  No hardware here: remember it so offsets are computed right."""
        self.center_freq = freq

    # ---- data plane ------------------------------------------------------------------------------
    def push(self, iq):
        """Channelise one wideband block and deliver each channel's narrowband samples to its sink."""
        with self.lock:
            self.bank.process(iq)
            self.samples_in += len(iq) // 2 if self._raw_dtype is not None else len(iq)   # raw blocks: interleaved I, Q
            if hasattr(self.bank, "pull_all") and len(self.channels) > 1:
                outs = self.bank.pull_all(OUT_IQ)     # one device-to-host transfer for every channel of the source
                for cid, ch in list(self.channels.items()):
                    ch.deliver(outs.get(cid, np.zeros(0, np.complex64)))
            else:
                for cid, ch in list(self.channels.items()):
                    ch.deliver(self.bank.pull(cid, OUT_IQ))
            for hook in self.post_push:
                hook()

    def _reader(self):
        kind = self.cfg.get("type")
        try:
            if kind == "file":
                path = self.cfg["path"]
                with open(path, "rb") as fh:
                    while not self._stop.is_set():
                        if self._raw_dtype is not None:
                            blk = np.fromfile(fh, dtype=self._raw_dtype, count=2 * self.block_samples)
                            blk = blk[:len(blk) & ~1]            # whole (I, Q) pairs
                            nblk = len(blk) // 2
                        else:
                            blk = np.fromfile(fh, dtype=np.complex64, count=self.block_samples)
                            nblk = len(blk)
                        if nblk == 0:
                            if not self.cfg.get("loop", False):
                                break
                            fh.seek(0)
                            continue
                        self.push(blk)
                        self._pace(nblk)
            elif kind == "synthetic":
                gen = self.cfg.get("generator")
                n0 = 0
                while not self._stop.is_set():
                    blk = gen(n0, self.block_samples) if gen else _default_generator(self, n0, self.block_samples)
                    n0 += len(blk)
                    self.push(blk)
                    self._pace(len(blk))
            elif kind in ("zmq", "rtlsdr", "usrp", "usrp2x", "bladerf", "osmosdr"):
                # hardware drivers are out of scope: take the wideband stream from the ZMQ PUB the
                # reference's SDR-side flowgraph publishes (rc_frontend/receiver.py:201)
                import zmq
                ctx = zmq.Context.instance()
                sub = ctx.socket(zmq.SUB)
                sub.setsockopt(zmq.SUBSCRIBE, b"")
                sub.setsockopt(zmq.RCVTIMEO, 100)
                sub.connect(self.cfg.get("address", self.address))
                pending = []
                have = 0
                while not self._stop.is_set():
                    try:
                        msg = sub.recv()
                    except zmq.Again:
                        continue
                    a = np.frombuffer(msg, dtype=np.complex64)
                    pending.append(a)
                    have += len(a)
                    if have >= self.block_samples:
                        self.push(np.concatenate(pending))
                        pending, have = [], 0
                sub.close()
        except Exception:
            logging.getLogger("frontend").exception("source %s reader died", self.source_id)

    def _pace(self, nsamples):
        if self.cfg.get("realtime", False):
            time.sleep(nsamples / float(self.samp_rate))

    def start(self):
        if self.cfg.get("type") in (None, "push") or self._thread is not None:
            return
        self._thread = threading.Thread(target=self._reader, name="source-%s" % self.source_id, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        channel_mod.unregister_source(self.address)
        for _, child in getattr(self, "split_children", []):
            child.stop()
        with self.lock:
            self.engine.close()


class BinStream(object):
    """One output bin of a source's polyphase channelizer, acting as the (400 kHz) source of the second-stage
    `channel` blocks the reference connects to it: `channel.channel(port, channel_rate, pfb_samp_rate, pfb_offset)`
    + `self.connect((pfb, pfb_id), block)` (rc_frontend/receiver.py:403-417).  Own GPU handle = own DDC bank."""

    def __init__(self, parent, bin_index, device):
        self.parent = parent
        self.bin_index = bin_index
        self.lock = parent.lock
        self.samp_rate = parent.pfb_samp_rate
        self.engine = Engine(device)
        self.bank = DdcBank(self.engine)
        self.channels = {}

    def open_channel(self, ch):
        with self.lock:
            cid = self.bank.open(ch.decim, ch.taps, float(ch.offset), float(ch.samp_rate), OUT_IQ, 1.0)
            self.channels[cid] = ch
            return cid

    def retune_channel(self, cid, offset):
        with self.lock:
            self.bank.retune(cid, float(offset))

    def set_channel_taps(self, cid, taps):
        with self.lock:
            self.bank.set_taps(cid, taps)

    def close_channel(self, cid):
        with self.lock:
            self.channels.pop(cid, None)
            self.bank.close(cid)

    def launch_device(self, d_ptr, nsamples):
        """Queue the second stage over `nsamples` complex64 samples of this bin, resident on the device (a row of the
        PFB output).  Asynchronous: the bins of a source are queued one after the other on their own CUDA streams and
        run side by side; `collect` then fetches the results."""
        self.bank.process_device(d_ptr, nsamples)

    def collect(self):
        """Every channel's block in one transfer (rcb_ddc_pull_all), delivered to the sinks."""
        got = self.bank.pull_all(OUT_IQ, copy=False)
        for cid, ch in list(self.channels.items()):
            y = got.get(cid)
            if y is not None and len(y):
                ch.deliver(np.array(y))
        # the parent reuses (or frees) the PFB output row after this call: nothing of this bin may still be reading it
        # (a block that produced no output returns without a sync)
        self.engine.sync()

    def push_device(self, d_ptr, nsamples):
        self.launch_device(d_ptr, nsamples)
        self.collect()

    def close(self):
        self.engine.close()


class PfbSourceStream(SourceStream):
    """frontend_mode 'pfb' (rc_frontend/receiver.py:242-261): the wideband stream goes once through
    `pfb.channelizer_ccf(num_channels, optfir.low_pass(1, N, 0.5, 0.7, 0.1, 80), 1.0, 100)` with
    num_channels = samp_rate / 400 kHz (K1, device resident), and each requested channel is a second-stage
    `channel` (freq_xlating_fir_filter_ccc at the 400 kHz bin rate: 59 taps, decimation 16 for 12.5 kHz channels)
    on its bin (K2 on the bin's row of the PFB output - the samples never leave the GPU in between)."""

    target_size = 400000  # rc_frontend/receiver.py:244

    def __init__(self, source_id, cfg, device=0, block_samples=None, engine_factory=None):
        if engine_factory is not None:
            raise NotImplementedError("pfb mode needs the GPU engine")
        SourceStream.__init__(self, source_id, cfg, device=device, block_samples=block_samples)
        self.device = device
        if self.samp_rate % self.target_size and not cfg.get("pfb_allow_odd_rate", False):
            raise Exception("samp_rate not round enough")   # rc_frontend/receiver.py:245-246
        self.num_channels = int(self.samp_rate // self.target_size)
        if self.num_channels < 1:
            raise ValueError("samp_rate %s is below one %s Hz pfb bin" % (self.samp_rate, self.target_size))
        self.pfb_samp_rate = self.samp_rate / float(self.num_channels)
        self.pfb_taps = firdes.pfb_prototype(self.num_channels)
        self.pfb = PfbChannelizer(self.engine, self.num_channels, self.pfb_taps, OUT_IQ, 1.0)
        self.bins = {}
        self._rem = np.zeros(0, np.complex64)
        self._d_in = None
        self._d_iq = None
        self._cap = 0

    def bin_stream(self, chan):
        with self.lock:
            if chan not in self.bins:
                self.bins[chan] = BinStream(self, chan, self.device)
            return self.bins[chan]

    def push(self, iq):
        with self.lock:
            if self.channels:  # channels opened directly on the wideband stream (xlat style) keep working
                SourceStream.push(self, iq)
            else:
                self.samples_in += len(iq)
            n = self.num_channels
            x = np.concatenate([self._rem, np.asarray(iq, np.complex64)]) if len(self._rem) else np.asarray(iq, np.complex64)
            nfr = len(x) // n        # stream_to_streams granularity: whole frames only, the rest waits
            self._rem = x[nfr * n:].copy()
            if nfr == 0:
                return
            x = np.ascontiguousarray(x[:nfr * n])
            if self._cap < nfr:
                for b in (self._d_in, self._d_iq):
                    if b is not None:
                        b.free()
                self._cap = nfr + nfr // 4
                self._d_in = self.engine.dev_alloc(self._cap * n * 8)
                self._d_iq = self.engine.dev_alloc(self._cap * n * 8)
            check(self.engine.lib.rcb_memcpy(self.engine.h, self._d_in.ptr, x.ctypes.data, x.nbytes, COPY_H2D),
                  "rcb_memcpy h2d", self.engine.h)
            self.pfb.process_device(self._d_in, nfr * n, self._d_iq, None, nfr)   # bin m = row m, nfr samples
            self.engine.sync()
            # second stage: queue every bin that has channels (each on its own handle / CUDA stream, so the banks
            # overlap on the GPU), then fetch each bin's outputs with one transfer
            active = [(m, bs) for m, bs in list(self.bins.items()) if bs.channels]
            for m, bs in active:
                bs.launch_device(self._d_iq.ptr + m * nfr * 8, nfr)
            for m, bs in active:
                bs.collect()

    def stop(self):
        # reader first (it pushes into the bins), then the bin engines under the lock, then the parent engine
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        with self.lock:
            for bs in list(self.bins.values()):
                bs.close()
            self.bins = {}
        SourceStream.stop(self)


def _default_generator(src, n0, n):
    rng = np.random.default_rng(n0 & 0xffffffff)
    return (rng.standard_normal(2 * n, dtype=np.float32) * np.float32(0.05)).view(np.complex64)


class receiver(object):
    def __init__(self, index=None, config=None, sink="zmq", devices=None, bind="tcp://0.0.0.0:0",
                 engine_factory=None, publisher=None, start=True, use_zmq=True, redis_client=None):
        self.log = logging.getLogger("frontend" if index is None else "frontend-%s" % (index,))
        self.index = index
        self.sink_kind = sink
        self.zmq_context = None
        self.zmq_socket = None
        if use_zmq:
            import zmq
            self.zmq_context = zmq.Context()
            self.zmq_socket = self.zmq_context.socket(zmq.REP)
            self.zmq_socket.setsockopt(zmq.LINGER, 0)
            self.zmq_socket.bind(bind)
        self.access_lock = threading.RLock()
        self.last_channel_cleanup = time.time()
        self.channel_idle_timeout = 10
        self.client_timeout = 5

        if config is None:
            from config import rc_config  # the reference's config.py symlink (README.md:83-85)
            config = rc_config()
        self.config = config
        self.scan_mode = getattr(config, "scan_mode", False)
        self.frontend_mode = getattr(config, "frontend_mode", "xlat")
        self.realsources = dict(config.sources)
        if index is not None:
            for i in list(self.realsources):
                if i != int(index):
                    del self.realsources[i]
        self.receiver_split2 = bool(getattr(config, "receiver_split2", False))
        ndev = len(devices) if devices else 1
        self.sources = {}
        numsources = 0
        for source in sorted(self.realsources):
            cfg = dict(self.realsources[source])
            dev = (devices[numsources % ndev] if devices else 0)
            if self.frontend_mode == "pfb" and engine_factory is None:
                stream = PfbSourceStream(source, cfg, device=dev)
            else:
                stream = SourceStream(source, cfg, device=dev, engine_factory=engine_factory)
            cfg["block"] = stream
            cfg["source_id"] = source
            self.realsources[source]["block"] = stream
            if self.receiver_split2:
                # rc_frontend/receiver.py:205-237: two half-band DDCs (+-fs/4, decimation 2) become two fs/2 sources
                from .split2 import split_source
                for half in split_source(stream, cfg, device=dev, engine_factory=engine_factory):
                    self.sources[numsources] = half
                    numsources += 1
                continue
            self.sources[numsources] = cfg
            numsources += 1
        self.channels = {}
        self.clients = {}
        self.client_hb = {}
        self.client_num = 0
        self.start_time = time.time()
        self.instance_uuid = "%s" % uuid.uuid4()
        self.publisher = publisher
        self.redis_channel_publisher = None
        if redis_client is not None or publisher == "redis":
            # rc_frontend/receiver.py:268: SADD channelizers / SET <uuid> every second from a background thread
            from .redis_channel_publisher import redis_channel_publisher
            self.publisher = None
            self.redis_channel_publisher = redis_channel_publisher(
                sources=self.sources, channels=self.channels, zmq_socket=self.zmq_socket, index=index,
                client=redis_client, instance_uuid=self.instance_uuid)
        self._serving = False
        if start:
            self.start()

    # ---- lifecycle -------------------------------------------------------------------------------
    def start(self):
        for s in self.sources.values():
            s["block"].start()

    def stop(self):
        self._serving = False
        if self.redis_channel_publisher is not None:
            self.redis_channel_publisher.stop()
        with self.access_lock:
            for c in list(self.channels):
                self.channels[c].destroy()
                del self.channels[c]
        for s in self.sources.values():
            s["block"].stop()
        if self.zmq_socket is not None:
            self.zmq_socket.close()
            self.zmq_context.term()
            self.zmq_socket = None

    def push(self, source_index, iq):
        """Feed one wideband block to a 'push' source (tests, benchmarks, embedding)."""
        self.sources[source_index]["block"].push(iq)

    def endpoint(self):
        import zmq
        return self.zmq_socket.getsockopt_string(zmq.LAST_ENDPOINT)

    # ---- channel requests ------------------------------------------------------------------------
    def connect_channel(self, channel_rate, freq):
        if self.frontend_mode == "xlat":
            return self.connect_channel_xlat(channel_rate, freq)
        if self.frontend_mode == "pfb":
            return self.connect_channel_pfb(channel_rate, freq)
        raise Exception("No frontend_mode selected")

    def connect_channel_xlat(self, channel_rate, freq):
        source_id = None
        source_distance = None
        if not self.scan_mode:
            for i in list(self.sources):
                d = abs(freq - self.sources[i]["center_freq"])
                if d < self.sources[i]["samp_rate"] / 2:
                    if source_distance is None or d < source_distance:
                        source_id = i
                        source_distance = d
            if source_id is None:
                raise Exception("Unable to find source for frequency %s" % freq)
        else:
            source_id = 0
        source_center_freq = self.sources[source_id]["center_freq"]
        source_samp_rate = self.sources[source_id]["samp_rate"]
        stream = self.sources[source_id]["block"]
        offset = freq - source_center_freq
        if freq < 10000000:
            offset = freq  # scan mode, relative freq (rc_frontend/receiver.py:304-305)

        with self.access_lock:
            block = None
            for c in list(self.channels):
                ch = self.channels[c]
                if ch.source_id == source_id and ch.channel_rate == channel_rate and not ch.in_use:
                    block = ch
                    port = block.port
                    block.set_offset(offset)
                    block.channel_close_time = 0
                    break
            if block is None:
                for x in range(0, 3):
                    port = random.randint(10000, 60000)
                    try:
                        block = channel(stream, port, channel_rate, source_samp_rate, offset, sink=self.sink_kind)
                        break
                    except RuntimeError:   # port in use, unknown source, or a refused rcb_ddc_open (B200ChanError)
                        self.log.error("Failed to build channel on port: %s attempt: %s" % (port, x))
                        block = None
                if block is None:
                    return False, False
                block.source_id = source_id
                block_id = "%s" % uuid.uuid4()
                self.channels[block_id] = block
                block.block_id = block_id
                block.start()
            block.in_use = True
            return block.block_id, port

    def connect_channel_pfb(self, channel_rate, freq):
        """rc_frontend/receiver.py:343-423 as intended (the shipped code is bit-rotted): pick the
        source, the 400 kHz bin and the residual offset (:365-377), reuse an idle second-stage channel of that bin
        (:386-394) or build `channel(port, channel_rate, pfb_samp_rate, pfb_offset)` and connect it to the bin
        (:396-417)."""
        source_id = None
        if not self.scan_mode:
            for i in list(self.sources):
                if abs(freq - self.sources[i]["center_freq"]) < self.sources[i]["samp_rate"] / 2:
                    source_id = i
                    break
            if source_id is None:
                raise Exception("Unable to find source for frequency %s" % freq)
        else:
            source_id = 0
        stream = self.sources[source_id]["block"]
        if not isinstance(stream, PfbSourceStream):  # fake engines (host tests): the DDC bank does the whole job
            return self.connect_channel_xlat(channel_rate, freq)
        if freq < 10000000:  # scan mode, relative freq (:295-305)
            freq = freq + self.sources[source_id]["center_freq"]
        chan, pfb_offset, _ = self.pfb_bin_for(source_id, freq, stream.pfb_samp_rate)
        half = stream.pfb_samp_rate / 2.0
        if pfb_offset < -half + channel_rate / 2.0 or pfb_offset > half - channel_rate / 2.0:
            self.log.warning("warning: %s edge boundary" % freq)
        with self.access_lock:
            block = None
            for c in list(self.channels):
                ch = self.channels[c]
                if ch.source_id == source_id and ch.pfb_id == chan and ch.channel_rate == channel_rate and not ch.in_use:
                    block = ch
                    port = block.port
                    block.set_offset(pfb_offset)
                    block.channel_close_time = 0
                    break
            if block is None:
                for x in range(0, 3):
                    port = random.randint(10000, 60000)
                    try:
                        block = channel(port, channel_rate, stream.pfb_samp_rate, pfb_offset, sink=self.sink_kind)
                        break
                    except RuntimeError:
                        self.log.error("Failed to build channel on port: %s attempt: %s" % (port, x))
                        block = None
                if block is None:
                    return False, False
                block.source_id = source_id
                block.pfb_id = chan
                block_id = "%s" % uuid.uuid4()
                self.channels[block_id] = block   # keyed by block id (the reference keys by port here, Appendix C.4)
                block.block_id = block_id
                block.connect_source(stream.bin_stream(chan))   # self.connect((pfb, pfb_id), block)
                block.start()
            block.in_use = True
            return block.block_id, port

    def pfb_bin_for(self, source_id, freq, target_size=400000):
        """rc_frontend/receiver.py:365-383 bin / residual arithmetic (row a4), with the negative-bin
        residual computed before the wrap (the reference subtracts the wrapped bin, Appendix C.4)."""
        src = self.sources[source_id]
        num_channels = int(round(src["samp_rate"] / float(target_size)))
        offset = freq - src["center_freq"]
        chan = int(round(offset / float(target_size)))
        pfb_offset = offset - chan * target_size
        chan %= num_channels   # negative bins wrap (:373-375); so does +N/2 of an even channel count
        edge = (pfb_offset < (-target_size / 2) or pfb_offset > (target_size / 2))
        return chan, pfb_offset, edge

    def release_channel(self, block_id):
        with self.access_lock:  # lock-balanced and idempotent (Appendix C.3)
            if block_id not in self.channels:
                return True
            self.channels[block_id].in_use = False
            self.channels[block_id].channel_close_time = time.time()
            return True

    def source_offset(self, block_id, offset):
        """AFC, rc_frontend/receiver.py:436-475."""
        if self.scan_mode:
            return False
        try:
            src = self.sources[self.channels[block_id].source_id]
        except Exception:
            return False
        center_freq = src["center_freq"]
        accumulated_offset = src.get("accumulated_offset", 0)
        base_offset = src.get("offset", 0)
        if offset > 1 or offset < -1:
            hz_offset = offset * 50
        elif offset > 0.5 or offset < -0.5:
            hz_offset = offset * 10
        else:
            hz_offset = offset * 4
        if -5 < hz_offset < 5:
            return True
        total_offset = accumulated_offset + hz_offset
        if abs(total_offset) > self.channels[block_id].channel_rate / 2:
            total_offset = (total_offset / 2) * -1
        new_center_freq = center_freq + total_offset
        with self.access_lock:
            src["block"].set_center_freq(new_center_freq + base_offset, 0)
            src["accumulated_offset"] = total_offset
        return True

    # ---- RPC (rc_frontend/receiver.py:503-614) ---------------------------------------------------
    def handler(self, msg):
        data = msg.strip().split(",")
        cmd = data[0]
        try:
            if cmd == "create":
                c = int(data[1])
                channel_rate = int(data[2])
                freq = int(data[3])
                try:
                    block_id, port = self.connect_channel(channel_rate, freq)
                except Exception as e:
                    block_id = -1
                    self.log.error("Exception: %s" % e)
                if block_id == -1 or block_id is False:
                    return "na,%s" % freq
                if c not in self.clients:  # connect must precede create (Appendix C.10)
                    self.release_channel(block_id)
                    return "na,%s" % freq
                self.clients[c].append(block_id)
                return "create,%s,%s" % (block_id, port)
            elif cmd == "release":
                try:
                    c = int(data[1])
                    block_id = data[2]
                    self.release_channel(block_id)
                    try:
                        self.clients[c].remove(block_id)
                    except (ValueError, KeyError):
                        pass
                    return "release,%s" % block_id
                except Exception:
                    return "na\n"
            elif cmd == "scan_mode_set_freq":
                freq = int(data[1])
                src = self.realsources[sorted(self.realsources)[0]]
                src["block"].set_center_freq(freq + src.get("offset", 0), 0)
                self.sources[0]["center_freq"] = freq
                return "success"
            elif cmd == "quit":
                c = int(data[1])
                for x in self.clients.get(c, []):
                    self.release_channel(x)
                self.client_hb.pop(c, None)
                self.clients.pop(c, None)
                return "quit,%s" % c
            elif cmd == "connect":
                c = self.client_num
                self.client_num += 1
                self.clients[c] = []
                self.client_hb[c] = time.time()
                return "connect,%s" % c
            elif cmd == "hb":
                try:
                    c = int(data[1])
                except Exception:
                    return "fail,0"
                if c not in self.client_hb:
                    return "fail,%s" % c
                self.client_hb[c] = time.time()
                return "hb,%s" % c
            elif cmd == "offset":
                client_id = int(data[1])
                self.source_offset(data[2], float(data[3]))
                return "offset,%s" % client_id
        except Exception as e:  # always answer exactly once (Appendix C.9)
            self.log.error("Exception in handler: (%s) %s" % (type(e), e))
            return "na"
        return "na"

    def housekeeping(self, now=None):
        """Heartbeat expiry (5 s, :651-680) and idle-channel reaping (10 s idle, swept every 20 s, :635-648)."""
        now = time.time() if now is None else now
        for client in list(self.client_hb):
            if now - self.client_hb[client] > self.client_timeout:
                self.log.warning("Client heartbeat timeout %s" % client)
                for x in self.clients.get(client, []):
                    self.release_channel(x)
                self.client_hb.pop(client, None)
                self.clients.pop(client, None)
        if now - self.last_channel_cleanup > self.channel_idle_timeout * 2:
            self.last_channel_cleanup = now
            with self.access_lock:
                for c in list(self.channels):
                    ch = self.channels[c]
                    if ch.channel_close_time != 0 and now - ch.channel_close_time > self.channel_idle_timeout:
                        self.log.info("disconnecting channel %s" % ch.block_id)
                        ch.destroy()
                        del self.channels[c]

    def describe(self):
        """The JSON blob rc_frontend/redis_channel_publisher.py:63-90 stores under <instance_uuid>."""
        address, port = "0.0.0.0", 0
        if self.zmq_socket is not None:
            ep = self.endpoint()
            address, port = ep.replace("tcp://", "").rsplit(":", 1)
        blob = {
            "instance_uuid": self.instance_uuid,
            "start_time": self.start_time,
            "current_time": time.time(),
            "hostname": socket.gethostname(),
            "pid": os.getpid(),
            "address": address,
            "port": int(port),
            "channel_count": len(self.channels),
            "source_count": len(self.sources),
            "sources": [[self.sources[s]["center_freq"], self.sources[s]["samp_rate"]] for s in self.sources],
        }
        if self.index is not None:
            blob["index"] = self.index
        return blob

    def serve_forever(self, poll_sleep=0.001):
        """The __main__ loop of rc_frontend/receiver.py:617-699."""
        import zmq
        self._serving = True
        last_status = time.time()
        last_publish = 0.0
        while self._serving:
            now = time.time()
            if now - last_status > 10:
                self.log.info("Frontend Status: client: %s client_hb: %s channels: %s uptime: %s msps_in: %.3f" % (
                    len(self.clients), len(self.client_hb), len(self.channels), int(now - self.start_time),
                    sum(s["block"].samples_in for s in self.sources.values()) / 1e6 / max(now - self.start_time, 1e-9)))
                last_status = now
            if self.publisher is not None and now - last_publish > 1.0:
                try:
                    self.publisher(self.instance_uuid, json.dumps(self.describe()))
                except Exception as e:
                    self.log.error("publisher failed: %s" % e)
                last_publish = now
            self.housekeeping(now)
            try:
                msg = self.zmq_socket.recv_string(flags=zmq.NOBLOCK)
            except zmq.Again:
                time.sleep(poll_sleep)
                continue
            except zmq.ZMQError:
                break
            resp = self.handler(msg)
            for _ in range(3):
                try:
                    self.zmq_socket.send_string(resp)
                    break
                except Exception as e:
                    self.log.error("Exception in send_string: (%s) %s" % (type(e), e))

    def serve_in_thread(self):
        t = threading.Thread(target=self.serve_forever, name="frontend-rpc", daemon=True)
        t.start()
        return t


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--index",
                        help="Device config index, if specified, all other configured sources will be deleted")
    parser.add_argument("--sink", default="zmq", choices=["zmq", "udp", "null"])
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    tb = receiver(args.index, sink=args.sink)
    print(json.dumps(tb.describe()))
    tb.serve_forever()


if __name__ == "__main__":
    main()
