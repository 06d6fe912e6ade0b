import sys, time
import numpy as np
sys.path.insert(0, ".")
from radiocapture_rf_b200.engine import Engine, FftScanner
from radiocapture_rf_b200 import firdes
import bench
L, avg = 16384, 100
n = 1 << 26
for mode in (0, 2, 0):
    e = Engine(0)
    sc = FftScanner(e, L, firdes.blackmanharris(L), avg)
    if mode:
        sc.set_pipeline(mode)
    hin = e.pinned((n,), np.complex64)
    base = bench.synth_block(1 << 22, 1024, 5)
    for i in range(0, n, len(base)):
        hin[i:i + len(base)] = base
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        out = sc.process(hin)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("mode", mode, "ms per call:", " ".join("%.1f" % t for t in ts), "vectors", len(out), flush=True)
    # same with a block-aligned number of frames (no carry between calls)
    m = (n // L // avg) * avg * L
    sc.reset()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        out = sc.process(hin[:m])
        ts.append((time.perf_counter() - t0) * 1e3)
    print("mode", mode, "aligned ms per call:", " ".join("%.1f" % t for t in ts), flush=True)
    e.close()
