"""Narrowband sample sinks of the frontend data plane (SURVEY.md 8(b) "Data plane").

ZmqPubSink  = zeromq.pub_sink(gr.sizeof_gr_complex, 1, 'tcp://0.0.0.0:<port>')   rc_frontend/channel.py:36
              each send() is one ZMQ message of raw little-endian complex64, no framing / tags
UdpSink     = blocks.udp_sink(gr.sizeof_gr_complex, host, port, 1472, True)      moto_control_demod.py:119,
              top-level receiver.py:146: datagrams of <= 1472 payload bytes (184 samples), zero-length
              datagram = EOF
NullSink    = blocks.null_sink (benchmarks / tests)
"""
import socket

import numpy as np


class NullSink(object):
    def __init__(self, port=None):
        self.port = port
        self.samples = 0

    def send(self, samples):
        self.samples += len(samples)

    def close(self):
        pass


class CaptureSink(NullSink):
    """Keeps everything it is sent (tests)."""

    def __init__(self, port=None):
        NullSink.__init__(self, port)
        self.chunks = []

    def send(self, samples):
        NullSink.send(self, samples)
        self.chunks.append(np.array(samples, copy=True))

    def data(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.complex64)


class ZmqPubSink(object):
    def __init__(self, port, bind_host="0.0.0.0", context=None):
        import zmq
        self.port = port
        self._ctx = context or zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.PUB)
        self._sock.setsockopt(zmq.LINGER, 0)
        try:
            self._sock.bind("tcp://%s:%s" % (bind_host, port))
        except zmq.ZMQError as e:  # the reference surfaces a RuntimeError from the GNU Radio block
            self._sock.close()
            raise RuntimeError("cannot bind tcp://%s:%s: %s" % (bind_host, port, e))

    def send(self, samples):
        self._sock.send(np.ascontiguousarray(samples, dtype=np.complex64).tobytes(), copy=False)

    def close(self):
        try:
            self._sock.close()
        except Exception:
            pass


class UdpSink(object):
    PAYLOAD = 1472

    def __init__(self, port, host="127.0.0.1"):
        self.port = port
        self._addr = (host, int(port))
        self._sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)

    def send(self, samples):
        raw = np.ascontiguousarray(samples, dtype=np.complex64).tobytes()
        step = self.PAYLOAD - (self.PAYLOAD % 8)
        for i in range(0, len(raw), step):
            self._sock.sendto(raw[i:i + step], self._addr)

    def close(self):
        try:
            self._sock.sendto(b"", self._addr)  # EOF marker (udp_sink eof=True)
            self._sock.close()
        except Exception:
            pass


def make_sink(kind, port):
    if kind is None or kind == "null":
        return NullSink(port)
    if kind == "capture":
        return CaptureSink(port)
    if kind == "zmq":
        return ZmqPubSink(port)
    if kind == "udp":
        return UdpSink(port)
    if hasattr(kind, "send"):
        return kind
    if callable(kind):
        return kind(port)
    raise ValueError("unknown sink %r" % (kind,))
