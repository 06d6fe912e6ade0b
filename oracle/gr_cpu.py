"""ctypes wrapper of oracle/gr_cpu.c (TEST INFRASTRUCTURE / CPU baseline - see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libgrcpu.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "gr_cpu.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        # -march=native is resolved on the box that builds; rebuild on the GPU box if the ISA differs
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgrcpu.so"], stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(LIB)
            _lib.grc_num_threads()
        except OSError:
            build(force=True)
            _lib = C.CDLL(LIB)
        vp, ll = C.c_void_p, C.c_long
        _lib.grc_pfb_fm.argtypes = [vp, ll, C.c_int, vp, C.c_int, C.c_float, vp, vp, ll, vp, C.c_int]
        _lib.grc_xlating_fir.argtypes = [vp, ll, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, C.POINTER(ll), C.c_int]
        _lib.grc_quad_demod.argtypes = [vp, ll, C.c_float, C.c_float, C.c_float, vp]
        _lib.grc_fft_logpow.argtypes = [vp, C.c_int, vp, ll, vp, C.c_int]
        _lib.grc_convert_iq.argtypes = [vp, C.c_int, C.c_float, C.c_float, ll, vp, C.c_int]
    return _lib


def num_threads():
    return int(load().grc_num_threads())


def pfb_fm(x, nchans, taps, gain, want_iq=True, want_fm=True, hist=None, nthreads=0, out_iq=None, out_fm=None):
    """out_iq / out_fm: optional preallocated [nchans][t] outputs (a timed loop should not pay first-touch page
    faults of fresh buffers on every step)."""
    lib = load()
    x = np.ascontiguousarray(x, np.complex64)
    taps = np.ascontiguousarray(taps, np.float32)
    t = len(x) // nchans
    p = -(-len(taps) // nchans)
    if hist is None:
        hist = np.zeros(p * nchans, np.complex64)
    iq = (out_iq if out_iq is not None else np.empty((nchans, t), np.complex64)) if want_iq else None
    fm = (out_fm if out_fm is not None else np.empty((nchans, t), np.float32)) if want_fm else None
    assert iq is None or (iq.shape == (nchans, t) and iq.dtype == np.complex64 and iq.flags.c_contiguous)
    assert fm is None or (fm.shape == (nchans, t) and fm.dtype == np.float32 and fm.flags.c_contiguous)
    rc = lib.grc_pfb_fm(x.ctypes.data, t, nchans, taps.ctypes.data, len(taps), gain,
                        iq.ctypes.data if want_iq else None, fm.ctypes.data if want_fm else None, max(t, 1),
                        hist.ctypes.data, nthreads)
    assert rc == 0
    return iq, fm, hist


def convert_iq(raw, fmt, offset, scale, out=None, nthreads=0):
    """Interleaved integer I/Q (fmt 1 u8, 2 s8, 3 s16) -> complex64 = (v + offset) * scale."""
    lib = load()
    raw = np.ascontiguousarray(raw).reshape(-1)
    n = len(raw) // 2
    if out is None:
        out = np.empty(n, np.complex64)
    assert out.dtype == np.complex64 and len(out) >= n
    rc = lib.grc_convert_iq(raw.ctypes.data, int(fmt), float(offset), float(scale), n, out.ctypes.data, nthreads)
    assert rc == 0
    return out[:n]


def xlating_fir(x, taps, decim, f0, fs, nthreads=0):
    lib = load()
    x = np.ascontiguousarray(x, np.complex64)
    taps = np.ascontiguousarray(taps, np.float32)
    out = np.empty((len(x) + decim - 1) // decim, np.complex64)
    n = C.c_long(0)
    rc = lib.grc_xlating_fir(x.ctypes.data, len(x), taps.ctypes.data, len(taps), decim, f0, fs, out.ctypes.data,
                             C.byref(n), nthreads)
    assert rc == 0
    return out[:n.value]


def quad_demod(x, gain, prev=0j):
    lib = load()
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty(len(x), np.float32)
    lib.grc_quad_demod(x.ctypes.data, len(x), gain, float(np.real(prev)), float(np.imag(prev)), out.ctypes.data)
    return out


def fft_logpow(x, length, window, nthreads=0):
    lib = load()
    x = np.ascontiguousarray(x, np.complex64)
    window = np.ascontiguousarray(window, np.float32)
    nfr = len(x) // length
    out = np.empty(length, np.float32)
    rc = lib.grc_fft_logpow(x.ctypes.data, length, window.ctypes.data, nfr, out.ctypes.data, nthreads)
    assert rc == 0
    return out
