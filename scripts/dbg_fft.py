import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import Engine, FftScanner
e = Engine(0)
for length, avg in [(4096, 1), (16384, 1), (65536, 1), (262144, 1), (1048576, 1), (16384, 100)]:
    x, truth = synth.scan_stream(length * avg, 2.4e6, length, seed=4, ncarriers=6)
    w = fd.blackmanharris(length)
    sc = FftScanner(e, length, w, avg)
    out = sc.process(x)[0].astype(np.float64)
    ref = gb.logpower_block_sums(x, length, w, avg)[0]
    err = np.abs(out - ref)
    if avg == 1:
        pl, pr = 10 ** (out - 1), 10 ** (ref - 1)
        print(length, avg, 'log err max %.3g mean %.3g' % (err.max(), err.mean()), 'linear power relL2 %.3g' % (np.linalg.norm(pl - pr) / np.linalg.norm(pr)),
              'dyn range dB %.1f' % (10 * (ref.max() - ref.min())), 'argmax', int(np.argmax(out)), int(np.argmax(ref)))
    else:
        print(length, avg, 'log err max %.3g mean %.3g' % (err.max(), err.mean()))
