"""radiocapture_rf_b200 - B200-native channelizer / FM demod / FFT scan behind radiocapture-rf's frontend API.

Host side is Python + ctypes over libb200chan.so (hand-written sm_100a CUDA, include/b200chan.h);
no PyTorch, no CPU fallback.  See DESIGN.md.
"""
__version__ = "0.1.0"
