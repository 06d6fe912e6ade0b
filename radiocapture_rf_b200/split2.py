"""`receiver_split2` - rc_frontend/receiver.py:74-86,205-237.

The reference splits every configured source into two half-rate sources with a pair of half-band DDCs,

    taps  = firdes.low_pass(1, samp_rate, samp_rate/4, samp_rate/8)          (19 taps, Hamming)
    filt1 = freq_xlating_fir_filter_ccc(2, taps, -samp_rate/4, samp_rate)    -> centre - fs/4, fs/2
    filt2 = freq_xlating_fir_filter_ccc(2, taps, +samp_rate/4, samp_rate)    -> centre + fs/4, fs/2

and registers `newsource1/2` (block = the filter) in `self.sources` so channel requests land on the half that
contains them.  Here the two filters are two decimate-by-2 channels of the parent's GPU DDC bank (K2: one staged copy
of the wideband block serves both), and each half is a SourceStream of its own whose wideband input is that channel's
output - handed over on the device (rcb_ddc_pull into device memory -> rcb_ddc_process from device memory), never
through the host.  (In the reference as shipped this mode raises KeyError in xlat mode; the
intent is implemented.)
"""
import numpy as np

from . import firdes
from .engine import OUT_IQ


def half_band_taps(samp_rate):
    decim = 2
    channel_rate = (samp_rate / decim) / 2.0
    transition = channel_rate * 0.5
    return decim, firdes.low_pass(1, float(samp_rate), channel_rate, transition)


def split_source(stream, cfg, device=0, engine_factory=None):
    """Open the two half-band channels on `stream` (a receiver.SourceStream) and return the two source dicts that
    replace it in receiver.sources (low half first, like newsource1 / newsource2)."""
    from .receiver import SourceStream
    fs = float(cfg["samp_rate"])
    decim, taps = half_band_taps(fs)
    halves = []
    children = []
    for name, sign in (("lo", -1.0), ("hi", +1.0)):
        ccfg = dict(cfg)
        ccfg["type"] = "push"
        ccfg["center_freq"] = cfg["center_freq"] + sign * fs / 4.0
        ccfg["samp_rate"] = fs / decim
        child = SourceStream("%s.%s" % (stream.source_id, name), ccfg, device=device, engine_factory=engine_factory)
        cid = stream.bank.open(decim, taps, sign * fs / 4.0, fs, OUT_IQ, 1.0)
        ccfg["block"] = child
        ccfg["source_id"] = stream.source_id
        ccfg["split_half"] = name
        halves.append(ccfg)
        children.append((cid, child))
    state = {"buf": None, "cap": 0}

    def feed_children():
        # runs at the end of the parent's push(), under its lock
        for cid, child in children:
            if hasattr(stream.bank, "pull_device"):
                n = stream.bank.nout(cid)
                if n == 0:
                    continue
                if state["cap"] < n:
                    if state["buf"] is not None:
                        state["buf"].free()
                    state["cap"] = n + n // 4
                    state["buf"] = stream.engine.dev_alloc(state["cap"] * 8)
                stream.bank.pull_device(cid, state["buf"], state["cap"])   # synchronises the parent's stream
                with child.lock:
                    child.bank.process_device(state["buf"], n)
                    child.samples_in += n
                    for ccid, ch in list(child.channels.items()):
                        ch.deliver(child.bank.pull(ccid, OUT_IQ))
                    child.engine.sync()
            else:   # recording test doubles: host hand-over
                y = np.asarray(stream.bank.pull(cid, OUT_IQ))
                if len(y):
                    child.push(y)

    stream.post_push.append(feed_children)
    stream.split_children = children
    return halves
