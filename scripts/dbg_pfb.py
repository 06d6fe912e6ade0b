import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import Engine, PfbChannelizer, OUT_FM, OUT_IQ
e = Engine(0)
n, tpa, frames = 64, 2, 1024
taps = fd.pfb_prototype(n, tpa)
x, offs = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, n + tpa)
ch = PfbChannelizer(e, n, taps, OUT_IQ | OUT_FM, 5.0)
iq, fm = ch.process(x)
ref = gb.pfb_channelizer(x, np.asarray(taps, np.float64), n)
fref = gb.quadrature_demod(ref, 5.0)
for m in (13, 15, 17):
    d = (fm[m].astype(np.float64) - fref[m]) / 5.0
    d = (d + np.pi) % (2 * np.pi) - np.pi
    bad = np.nonzero(np.abs(d) > 1e-4)[0]
    print(m, 'nbad', len(bad), bad[:20], 'maxd', np.abs(d).max())
    for b in bad[:5]:
        print('   t', b, 'fm', fm[m][b], 'ref', fref[m][b], '|Y|', abs(ref[m][b]), abs(ref[m][b-1]), 'iqerr', abs(iq[m][b]-ref[m][b]))
    fm_from_gpu_iq = gb.quadrature_demod(iq[m].astype(np.complex128), 5.0)
    print('   fm vs fm(gpu iq):', np.abs(((fm[m]-fm_from_gpu_iq)/5+np.pi)%(2*np.pi)-np.pi).max())
