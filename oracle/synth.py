"""Seeded synthetic wideband IQ (TEST INFRASTRUCTURE; SURVEY.md 8(d) "Concrete synthetic inputs").

The reference ships no captures (only the raw-IQ .dat format, logging_receiver.py:107-109 /
file_to_wav.py:42: headerless interleaved little-endian float32 I/Q), so every input is synthetic:
a sum of narrowband FM carriers (1 kHz tone + 4-level 4800-baud FSK, like the P25 C4FM control
channels the reference decodes) plus AWGN, amplitude-normalised to RMS 0.25, complex64.
"""
import math

import numpy as np


def fm_carrier(n, fs, offset_hz, rng, dev_hz=2500.0, tone_hz=1000.0, baud=4800.0, start=0):
    """One NBFM carrier at ``offset_hz``: phase = 2pi*(offset*t + integral of modulation)."""
    t = (np.arange(n, dtype=np.float64) + start) / fs
    nsym = int(n * baud / fs) + 2
    levels = rng.choice(np.array([-1800.0, -600.0, 600.0, 1800.0]), size=nsym)
    sym_idx = np.minimum((np.arange(n) * (baud / fs)).astype(np.int64), nsym - 1)
    inst = dev_hz * np.sin(2 * math.pi * tone_hz * t) * 0.5 + levels[sym_idx]
    ph = 2 * math.pi * (np.cumsum(inst) / fs + offset_hz * t)
    return np.exp(1j * ph)


def wideband(n, fs, offsets_hz, seed, noise_dbc=-30.0, rms=0.25):
    """Sum of FM carriers at ``offsets_hz`` + AWGN ``noise_dbc`` below one carrier."""
    rng = np.random.default_rng(seed)
    x = np.zeros(n, dtype=np.complex128)
    for off in offsets_hz:
        x += fm_carrier(n, fs, off, rng) * rng.uniform(0.5, 1.0)
    sigma = 10.0 ** (noise_dbc / 20.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / math.sqrt(2.0)
    x *= rms / math.sqrt(np.mean(np.abs(x) ** 2))
    return x.astype(np.complex64)


def cfg1(n=1 << 20, seed=1):
    """BASELINE config 1: fs 2.4 Msps, channel under test at -62.5 kHz (854.9875 MHz vs 855.05 MHz,
    configs/config_denver_dev_den817.py:32,127), rate 12500 -> D 96, 349 taps."""
    fs = 2.4e6
    offs = [-1.0375e6, -62.5e3, 0.0, 12.5e3, 437.5e3, 600e3, -300e3, 900e3]
    return wideband(n, fs, offs, seed), fs, offs


def pfb_stream(n, fs, nchans, seed, active_every=2, jitter_hz=None, noise_dbc=-30.0):
    """BASELINE configs 2/3/5: one FM carrier per ``active_every``-th bin at bin centre + U(-j,j)."""
    rng = np.random.default_rng(seed)
    bin_hz = fs / nchans
    if jitter_hz is None:
        jitter_hz = min(50e3, bin_hz * 0.125)
    bins = np.arange(1, nchans, active_every)
    offs = []
    for b in bins:
        f = b * bin_hz + rng.uniform(-jitter_hz, jitter_hz)
        if f >= fs / 2:
            f -= fs
        offs.append(f)
    # cheap generator for many carriers: constant-envelope tones with slow sinusoidal FM
    t = np.arange(n, dtype=np.float64) / fs
    x = np.zeros(n, dtype=np.complex128)
    for f in offs:
        dev = rng.uniform(0.02, 0.08) * bin_hz
        fmod = rng.uniform(0.001, 0.01) * bin_hz
        ph0 = rng.uniform(0, 2 * math.pi)
        x += np.exp(1j * (2 * math.pi * f * t + (dev / fmod) * np.sin(2 * math.pi * fmod * t + ph0)))
    sigma = 10.0 ** (noise_dbc / 20.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / math.sqrt(2.0)
    x *= 0.25 / math.sqrt(np.mean(np.abs(x) ** 2))
    return x.astype(np.complex64), offs


def noise_block(n, seed, rms=0.25):
    """Cheap complex Gaussian block (bench fill: arithmetic cost is data independent)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(2 * n, dtype=np.float32) * np.float32(rms / math.sqrt(2.0))
    return x.view(np.complex64)


def scan_stream(n, fs, length, seed, ncarriers=8, noise_dbc=-40.0):
    """BASELINE config 4 style: ``ncarriers`` carriers 3-30 kHz wide at known bins over a noise
    floor.  Returns (x, list of (bin_index_after_fftshift, width_hz))."""
    rng = np.random.default_rng(seed)
    hz_per_bin = fs / length
    t = np.arange(n, dtype=np.float64) / fs
    x = np.zeros(n, dtype=np.complex128)
    truth = []
    # spread carriers over the middle 80 % of the band
    centres = np.linspace(-0.4 * fs, 0.4 * fs, ncarriers) + rng.uniform(-0.01 * fs, 0.01 * fs, ncarriers)
    for fc in centres:
        width = rng.uniform(6e3, 20e3)
        # band-limited noise-like FM: wideband-ish FM with random tone set gives a flat-topped hump
        nt = 12
        fm = rng.uniform(200.0, 1500.0, nt)
        ph = rng.uniform(0, 2 * math.pi, nt)
        beta = (width / 2.0) / fm / nt * 2.2
        arg = 2 * math.pi * fc * t
        for q in range(nt):
            arg = arg + beta[q] * np.sin(2 * math.pi * fm[q] * t + ph[q])
        x += np.exp(1j * arg)
        truth.append((int(round(fc / hz_per_bin)) + length // 2, width))
    sigma = 10.0 ** (noise_dbc / 20.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / math.sqrt(2.0)
    x *= 0.25 / math.sqrt(np.mean(np.abs(x) ** 2))
    return x.astype(np.complex64), truth
