"""GPU check of a "format": "u8" file source through the real engine: the channel's delivered samples equal the DDC of the
converted stream (float64 oracle, 1e-5)."""
import sys, time, tempfile, os
import numpy as np
sys.path.insert(0, ".")
from oracle import gr_blocks as gb, gr_firdes as fd
from radiocapture_rf_b200.receiver import SourceStream
from radiocapture_rf_b200 import channel as channel_mod

fs = 2400000
n = 96 * 600
t = np.arange(n)
sig = 0.5 * np.exp(2j * np.pi * (-62500.0 / fs * t + 0.2 * np.sin(2 * np.pi * 2e-4 * t)))
raw = np.clip(np.round(np.stack([sig.real, sig.imag], 1) * 127.0 + 127.4), 0, 255).astype(np.uint8).reshape(-1)
d = tempfile.mkdtemp()
path = os.path.join(d, "cap.u8")
raw.tofile(path)
cfg = {"type": "file", "path": path, "format": "u8", "center_freq": 855050000, "samp_rate": fs}
src = SourceStream(0, cfg, block_samples=96 * 200)
ch = channel_mod.channel(src, 0, 12500, fs, -62500, sink="capture")
ch.start()
src.start()
for _ in range(300):
    if src.samples_in >= n:
        break
    time.sleep(0.01)
src.stop()
y = ch.sink.data()
xf = ((raw.astype(np.float32) - np.float32(127.4)) * np.float32(1 / 128.0)).view(np.complex64)
decim, taps = fd.channel_taps(fs, 12500)
ref = gb.freq_xlating_fir(xf, taps, decim, -62500.0, fs)
m = min(len(y), len(ref))
err = gb.rel_l2(y[:m], ref[:m])
print("raw source: %d samples in, %d out, rel err %.2e" % (src.samples_in, len(y), err))
assert m >= 590 and err <= 1e-5
print("OK")
