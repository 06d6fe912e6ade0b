"""ctypes binding of libb200chan.so (include/b200chan.h).  No PyTorch anywhere on this path.

The library is built in-tree by ``radiocapture_rf_b200.build.build_library()`` (also called from
``__graft_entry__.build()``).  Loading fails LOUDLY when the shared object is missing - there is no
CPU fallback in the product path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCB_LIBRARY points at an alternative build of the same C ABI (e.g. an -DRCB_EXPERIMENTS tuning build)
LIB_PATH = os.environ.get("RCB_LIBRARY") or os.path.join(_HERE, "libb200chan.so")

RCB_OK = 0
RCB_EINVAL = -1
RCB_ENOMEM = -2
RCB_ECUDA = -3
RCB_ESTATE = -4
RCB_ENODEV = -5
RCB_ERANGE = -6
RCB_EUNSUPPORTED = -7

MEM_HOST = 0
MEM_DEVICE = 1
OUT_IQ = 1
OUT_FM = 2
FMT_U8 = 1
FMT_S8 = 2
FMT_S16 = 3
COPY_H2D = 1
COPY_D2H = 2
COPY_D2D = 3


class rcb_stats_t(C.Structure):
    _fields_ = [("samples_in", C.c_uint64), ("channel_samples", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


POST_P25_C4FM = 1
POST_ANALOG_FM = 2


class rcb_post_cfg(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_int), ("gain", C.c_float),
                ("taps0", C.c_void_p), ("ntaps0", C.c_int),
                ("taps1", C.c_void_p), ("ntaps1", C.c_int),
                ("taps2", C.c_void_p), ("ntaps2", C.c_int),
                ("interp", C.c_int), ("decim", C.c_int),
                ("squelch_db", C.c_double), ("squelch_alpha", C.c_double), ("squelch_gate", C.c_int),
                ("deemph_b0", C.c_double), ("deemph_b1", C.c_double), ("deemph_a1", C.c_double),
                ("probe_len", C.c_int), ("probe_scale", C.c_float)]


class B200ChanError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = "%s failed: %s (%d)" % (where, _strerror(status), status)
        if detail:
            msg += " - " + detail
        RuntimeError.__init__(self, msg)


_vp = C.c_void_p
_sz = C.c_size_t
_PROTOTYPES = {
    # name: (restype, argtypes)           -- must list every symbol declared in include/b200chan.h
    "rcb_version": (C.c_int, []),
    "rcb_strerror": (C.c_char_p, [C.c_int]),
    "rcb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rcb_open": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "rcb_close": (C.c_int, [_vp]),
    "rcb_sync": (C.c_int, [_vp]),
    "rcb_last_error": (C.c_char_p, [_vp]),
    "rcb_stats": (C.c_int, [_vp, C.POINTER(rcb_stats_t)]),
    "rcb_device_name": (C.c_int, [_vp, C.c_char_p, _sz, C.POINTER(C.c_int)]),
    "rcb_dev_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "rcb_dev_free": (C.c_int, [_vp, _vp]),
    "rcb_host_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "rcb_host_free": (C.c_int, [_vp, _vp]),
    "rcb_memcpy": (C.c_int, [_vp, _vp, _vp, _sz, C.c_int]),
    "rcb_memset": (C.c_int, [_vp, _vp, C.c_int, _sz]),
    "rcb_l2_flush": (C.c_int, [_vp]),
    "rcb_timer_start": (C.c_int, [_vp]),
    "rcb_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rcb_copy_ceiling": (C.c_int, [_vp, _sz, _sz, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]),
    "rcb_pfb_config": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_float]),
    "rcb_pfb_reset": (C.c_int, [_vp]),
    "rcb_pfb_set_out_block": (C.c_int, [_vp, C.c_int]),
    "rcb_pfb_set_input_format": (C.c_int, [_vp, C.c_int, C.c_float, C.c_float]),
    "rcb_ddc_set_input_format": (C.c_int, [_vp, C.c_int, C.c_float, C.c_float]),
    "rcb_pfb_process": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _vp, _sz, C.c_int, C.POINTER(_sz)]),
    "rcb_pfb_process_multi": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz]),
    "rcb_ddc_open": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_float,
                               C.POINTER(C.c_int)]),
    "rcb_ddc_retune": (C.c_int, [_vp, C.c_int, C.c_double]),
    "rcb_ddc_set_taps": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "rcb_ddc_close": (C.c_int, [_vp, C.c_int]),
    "rcb_ddc_process": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "rcb_ddc_pull": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _sz, C.c_int, C.POINTER(_sz)]),
    "rcb_ddc_set_tensor_cores": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rcb_ddc_tensor_core_launches": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "rcb_ddc_pull_all": (C.c_int, [_vp, C.c_int, _vp, _sz, C.c_int, _vp, _vp, _sz, C.POINTER(_sz)]),
    "rcb_quad_demod": (C.c_int, [_vp, _vp, _sz, _sz, _sz, C.c_float, _vp, _vp, _sz, C.c_int]),
    "rcb_probe_mean": (C.c_int, [_vp, _vp, _sz, _sz, _sz, _sz, C.c_float, _vp, C.c_int]),
    "rcb_convert_iq": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_float, _sz, C.c_int, _vp, C.c_int]),
    "rcb_post_open": (C.c_int, [_vp, C.POINTER(rcb_post_cfg), C.POINTER(C.c_int)]),
    "rcb_post_process": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, C.c_int, _vp, _sz, C.c_int, _vp, _vp]),
    "rcb_post_close": (C.c_int, [_vp, C.c_int]),
    "rcb_fft_config": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "rcb_fft_reset": (C.c_int, [_vp]),
    "rcb_fft_set_pipeline": (C.c_int, [_vp, C.c_int]),
    "rcb_fft_process": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _sz, C.c_int, C.POINTER(_sz)]),
}

_lib = None


def load():
    """Load libb200chan.so and bind every prototype.  Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libb200chan.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or radiocapture_rf_b200.build.build_library(); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError = missing export: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _strerror(status):
    try:
        return load().rcb_strerror(status).decode()
    except Exception:  # pragma: no cover
        return "status %d" % status


def check(status, where, handle=None):
    if status != RCB_OK:
        detail = ""
        if handle is not None and status == RCB_ECUDA:
            detail = load().rcb_last_error(handle).decode()
        raise B200ChanError(status, where, detail)
