import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from radiocapture_rf_b200.engine import Engine, FftScanner
from radiocapture_rf_b200 import firdes
e = Engine(0)
L, avg, n = 16384, 100, 1 << 26
sc = FftScanner(e, L, firdes.blackmanharris(L), avg)
hin = e.pinned((n,), np.complex64)
for name, fill in (("const", None), ("synth", bench.synth_block(1 << 22, 1024, 5))):
    if fill is None:
        hin[:] = 0.1
    else:
        for i in range(0, n, len(fill)):
            hin[i:i + len(fill)] = fill
    sc.reset(); sc.process(hin)
    for rep in range(3):
        t0 = time.perf_counter(); out = sc.process(hin); t1 = time.perf_counter()
        print(name, rep, 'host call %.1f ms vectors %d' % ((t1 - t0) * 1e3, len(out)))
t0 = time.perf_counter(); r = bench.run_e2e_fft(0, "cfg4_16k", 2, None, 0); t1 = time.perf_counter()
print('bench.run_e2e_fft ->', r[0], 'Msps; total wall', t1 - t0)
