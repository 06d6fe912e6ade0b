import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radiocapture_rf_b200.engine import Engine, FftScanner
from radiocapture_rf_b200 import firdes
e = Engine(0)
for L, avg in [(16384, 100), (16384, 128), (1 << 20, 64)]:
    n = 1 << 26
    sc = FftScanner(e, L, firdes.blackmanharris(L), avg)
    hin = e.pinned((n,), np.complex64)
    hin[:] = 0.1
    sc.process(hin)
    t0 = time.perf_counter(); out = sc.process(hin); t1 = time.perf_counter()
    d_in = e.to_device(hin); d_out = e.dev_alloc((n // L // avg + 2) * L * 4)
    sc.process_device(d_in, n, d_out, n // L // avg + 2); e.sync()
    t2 = time.perf_counter(); sc.process_device(d_in, n, d_out, n // L // avg + 2); e.sync(); t3 = time.perf_counter()
    print(L, avg, 'host call %.1f ms  device call %.1f ms  vectors %d' % ((t1 - t0) * 1e3, (t3 - t2) * 1e3, len(out)))
    d_in.free(); d_out.free()
