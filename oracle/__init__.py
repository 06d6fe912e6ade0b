"""CPU oracle for the radiocapture-rf hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy float64 + a float32 GNU-Radio-semantics
C library, ``oracle/gr_cpu.c``) of the arithmetic the reference delegates to
GNU Radio 3.8 blocks on the channelizer / FM-demod / FFT-scan path.

PARITY UNPINNED: the reference ships no tests, golden vectors or captures, and
GNU Radio 3.8 (where the arithmetic lives; pinned only by
``fft_vector.py:9`` "GNU Radio version: 3.8.1.0" and ``README.md:56``) is neither
vendored under /root/reference nor installed.  The block semantics below are
restated from GNU Radio 3.8's published algorithms (file names cited per
function) and anchored on the reference's call sites.  Only the peak picking
(``fft_peak_detection.py:46-72``) is pinned against the real library the
reference calls (scipy.signal.find_peaks), via tests/golden/.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``radiocapture_rf_b200``) never does.
"""
from . import gr_firdes, gr_blocks, synth  # noqa: F401
