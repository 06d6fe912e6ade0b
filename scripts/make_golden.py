#!/usr/bin/env python
"""Mint tests/golden/ (run in the build container; scipy is the reference's own peak-picking library).

The reference ships no golden vectors and its GNU Radio dependency is absent, so these pin (a) the
oracle against drift and (b) the peak indices against the real scipy.signal.find_peaks call of
fft_peak_detection.py:65.  Inputs are small so the fixtures stay < 1 MB.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gr_blocks as gb, gr_firdes as fd, synth  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
g = {}
g["taps_349"] = fd.channel_taps(2.4e6, 12500)[1]
g["taps_69"] = fd.low_pass_2(1.0, 25000, 6250, 500.0, 30.0, fd.WIN_BLACKMAN)
g["taps_19"] = fd.low_pass(1, 8e6, 2e6, 1e6)
x, fs, offs = synth.cfg1(96 * 256, seed=1)
g["cfg1_x"] = x
y = gb.freq_xlating_fir(x, g["taps_349"], 96, -62500.0, 2.4e6)
g["cfg1_y"] = y
g["cfg1_fm"] = gb.quadrature_demod(y, 5.0)
px, _ = synth.pfb_stream(16 * 128, 6.4e6, 16, 2)
g["pfb_x"] = px
g["pfb_taps"] = fd.pfb_prototype(16, 4)
g["pfb_y"] = gb.pfb_channelizer(px, g["pfb_taps"].astype(np.float64), 16)
length = 16384
sx, truth = synth.scan_stream(length * 100, 2.4e6, length, seed=44, ncarriers=8)
vec = gb.fft_vector_flowgraph(sx, length, fd.blackmanharris(length), nframes=100, avg=100).astype(np.float32)
g["scan_vec"] = vec
idx, hz = gb.peak_detect(vec, 2.4e6, 855.05e6)
np.savez_compressed(os.path.join(out, "hotpath_golden.npz"), **g)
man = {"scan_peaks_idx": idx.tolist(), "scan_peaks_hz": hz.tolist(), "scan_truth_bins": [int(t[0]) for t in truth],
       "made_with": {"numpy": np.__version__, "scipy": __import__("scipy").__version__}}
json.dump(man, open(os.path.join(out, "manifest.json"), "w"), indent=1)
print("peaks", idx.tolist(), "truth", man["scan_truth_bins"])
