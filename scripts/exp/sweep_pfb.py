"""Device-resident throughput of the FM-only channelizer per (N, taps per arm): python sweep_pfb.py [log2n]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from radiocapture_rf_b200.engine import Engine, PfbChannelizer, OUT_FM

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
shapes = [(1024, p) for p in (1, 2, 4, 8, 16)] + [(256, p) for p in (1, 2, 4, 8, 16)]
if len(sys.argv) > 2:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
e = Engine(0)
n = 1 << log2n
base = bench.synth_block(1 << 22, 1024, 3)
d_in = e.dev_alloc(n * 8)
from radiocapture_rf_b200._lib import COPY_H2D, check
check(e.lib.rcb_memcpy(e.h, d_in.ptr, base.ctypes.data, base.nbytes, COPY_H2D), "h2d", e.h)
filled = len(base)
while filled < n:
    c = min(filled, n - filled)
    e.copy_d2d(d_in.ptr + filled * 8, d_in.ptr, c * 8)
    filled += c
d_fm = e.dev_alloc(n * 4)
for nch, tpa in shapes:
    ch = PfbChannelizer(e, nch, bench.make_taps(nch, nch * tpa), OUT_FM, 5.0)
    ch.set_out_block(1024)
    frames = n // nch
    for _ in range(3):
        ch.process_device(d_in, n, None, d_fm, frames)
    e.sync()
    e.timer_start()
    reps = 8
    for _ in range(reps):
        ch.process_device(d_in, n, None, d_fm, frames)
    ms = e.timer_stop()
    print("N %4d taps/arm %2d : %7.1f Gsps  frac %.3f" % (nch, tpa, n * reps / ms / 1e6, n * reps * 12 / ms / 1e6 / 6551.0), flush=True)
e.close()
