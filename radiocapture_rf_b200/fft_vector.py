"""`fft_vector` - the scan flowgraph, mirroring fft_vector.py:22-95 of the reference.

Reference: ZMQ SUB ipc:///tmp/rx_source_<index> (:37) -> stream_to_vector(16384) (:39) ->
fft_vcc(16384, True, blackmanharris, True) (:38) -> |.|^2 (:46) -> nlog10_ff(1, L, 1) (:41) ->
moving_average_ff(100, 1, 1200, L) (:42) -> head(1000) (:43) -> skiphead(999) (:40) ->
file_sink /tmp/fft_source_<index> (:44): ONE float32 vector = sum over frames 900..999 of (log10|X|^2 + 1).

Here the whole chain is the GPU scan kernel pair (K3, rcb_fft_*).  `run(samples)` consumes the first
nframes*L samples and writes / returns that vector; `skip_discarded=True` (default) does not spend FFTs
on frames 0..nframes-avg-1 whose only effect in the reference is to be added to and then subtracted from
the running sum (set False to transform every frame like GNU Radio does).
"""
import numpy as np

from . import firdes
from .engine import Engine, FftScanner
from .fft_peak_detection import save_vector


class fft_vector(object):
    def __init__(self, index=0, samp_rate=2400000, length=1024 * 16, nframes=1000, avg=100, engine=None,
                 window=None, skip_discarded=True):
        self.index = index
        self.samp_rate = samp_rate
        self.length = length
        self.nframes = nframes
        self.avg = avg
        self.skip_discarded = skip_discarded
        self._own = engine is None
        self.engine = engine or Engine(0)
        self.window = firdes.blackmanharris(length) if window is None else np.asarray(window, np.float32)
        self.scanner = FftScanner(self.engine, length, self.window, avg)
        self.path = "/tmp/fft_source_%s" % index

    def get_samp_rate(self):
        return self.samp_rate

    def set_samp_rate(self, samp_rate):
        self.samp_rate = samp_rate

    def get_length(self):
        return self.length

    def set_length(self, length):
        self.length = length
        self.window = firdes.blackmanharris(length)
        self.scanner = FftScanner(self.engine, length, self.window, self.avg)

    def run(self, samples, write=False):
        """samples: >= nframes*length complex64.  Returns the float32 vector fft_vector.py writes."""
        need = self.nframes * self.length
        if len(samples) < need:
            raise ValueError("need %d samples (head(%d) frames of %d)" % (need, self.nframes, self.length))
        self.scanner.reset()
        first = self.nframes - self.avg
        if self.skip_discarded:
            out = self.scanner.process(samples[first * self.length:need])
            vec = out[-1]
        else:
            # every frame transformed; block b = frames b*avg..; the written item is the last block when
            # nframes is a multiple of avg, as in the reference (1000 / 100)
            if self.nframes % self.avg:
                raise ValueError("nframes must be a multiple of avg when skip_discarded=False")
            out = self.scanner.process(samples[:need])
            vec = out[-1]
        if write:
            save_vector(self.path, vec)
        return vec

    def run_from_zmq(self, address=None, write=True, timeout_s=60.0):
        """Subscribe to the wideband stream like fft_vector.py:37 and produce the vector."""
        import time
        import zmq
        address = address or "ipc:///tmp/rx_source_%s" % self.index
        ctx = zmq.Context.instance()
        sub = ctx.socket(zmq.SUB)
        sub.setsockopt(zmq.SUBSCRIBE, b"")
        sub.setsockopt(zmq.RCVTIMEO, 200)
        sub.connect(address)
        need = self.nframes * self.length
        buf = np.empty(need, np.complex64)
        have = 0
        t0 = time.time()
        while have < need and time.time() - t0 < timeout_s:
            try:
                msg = sub.recv()
            except zmq.Again:
                continue
            a = np.frombuffer(msg, np.complex64)
            k = min(len(a), need - have)
            buf[have:have + k] = a[:k]
            have += k
        sub.close()
        if have < need:
            raise RuntimeError("timed out with %d of %d samples" % (have, need))
        return self.run(buf, write=write)

    def close(self):
        if self._own:
            self.engine.close()


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--index", help="Device config index", default=0)
    args = parser.parse_args(argv)
    tb = fft_vector(args.index)
    tb.run_from_zmq()
    tb.close()


if __name__ == "__main__":
    main()
