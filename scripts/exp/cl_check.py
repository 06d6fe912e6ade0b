"""First-light diagnostic of pfb_cl_kernel: error metrics per shape against the oracle (prints, does not assert)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import Engine, PfbChannelizer, OUT_FM


def fm_err(fm, ref, gain):
    d = (fm.astype(np.float64) - ref) / gain
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(d) / max(np.linalg.norm(ref / gain), 1e-30)


def main():
    cases = [(256, 1, 64), (256, 16, 64), (256, 2, 33), (1024, 1, 64), (1024, 16, 96), (1024, 4, 131), (1024, 16, 7),
             (256, 8, 1), (1024, 2, 16), (1024, 16, 3000)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    e = Engine(0)
    for n, tpa, frames in cases:
        taps = fd.pfb_prototype(n, tpa)
        x, _ = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, 5, active_every=2)
        ch = PfbChannelizer(e, n, taps, OUT_FM, 5.0)
        _, fm = ch.process(x)
        ref = gb.quadrature_demod(gb.pfb_channelizer(x, np.asarray(taps, np.float64), n), 5.0)
        act = list(range(1, n, 2))
        errs = np.array([fm_err(fm[m], ref[m], 5.0) for m in act])
        bad = [act[i] for i in np.argsort(errs)[::-1][:6]]
        print("N %d tpa %s frames %d: max act err %.3e (worst chans %s) aggregate %.3e nan %d" %
              (n, tpa, frames, errs.max(), bad, fm_err(fm, ref, 5.0), int(np.isnan(fm).sum())), flush=True)
        if errs.max() > 1e-5 and frames <= 131:
            m = bad[0]
            terr = np.abs(((fm[m] - ref[m]) / 5.0 + np.pi) % (2 * np.pi) - np.pi)
            print("   chan %d per-frame err:" % m, np.array2string(terr[:40], precision=2, max_line_width=200))
    e.close()


main()
