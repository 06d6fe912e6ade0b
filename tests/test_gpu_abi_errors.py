"""GPU: error behaviour of the C ABI - every misuse returns a negative rcb_status, never crashes, and the
handle stays usable (the reference turns failures into in-band replies, rc_frontend/receiver.py:511-534)."""
import ctypes as C

import numpy as np
import pytest

from radiocapture_rf_b200 import _lib
from radiocapture_rf_b200.engine import DdcBank, Engine, FftScanner, PfbChannelizer, OUT_FM, OUT_IQ

pytestmark = pytest.mark.gpu


def test_call_order_and_argument_errors(engine, built_lib):
    L = built_lib
    h = engine.h
    n = C.c_size_t()
    x = np.zeros(1024, np.complex64)
    o = np.zeros(1024, np.complex64)
    # process before config
    assert L.rcb_pfb_process(h, x.ctypes.data, 1024, 0, o.ctypes.data, None, 16, 0, C.byref(n)) == _lib.RCB_ESTATE
    assert L.rcb_pfb_reset(h) == _lib.RCB_ESTATE
    assert L.rcb_pfb_set_out_block(h, 64) == _lib.RCB_ESTATE
    assert L.rcb_fft_process(h, x.ctypes.data, 1024, 0, o.ctypes.data, 1, 0, C.byref(n)) == _lib.RCB_ESTATE
    # bad configs
    t = np.ones(8, np.float32)
    assert L.rcb_pfb_config(h, 0, t.ctypes.data, 8, OUT_IQ, 1.0) == _lib.RCB_EINVAL
    assert L.rcb_pfb_config(h, 64, None, 8, OUT_IQ, 1.0) == _lib.RCB_EINVAL
    assert L.rcb_pfb_config(h, 64, t.ctypes.data, 8, 0, 1.0) == _lib.RCB_EINVAL
    assert L.rcb_pfb_config(h, 64, t.ctypes.data, 8, 8, 1.0) == _lib.RCB_EINVAL
    assert L.rcb_pfb_config(h, 1 << 20, t.ctypes.data, 8, OUT_IQ, 1.0) == _lib.RCB_EUNSUPPORTED
    cid = C.c_int()
    assert L.rcb_ddc_open(h, 0, t.ctypes.data, 8, 0.0, 1.0, OUT_IQ, 1.0, C.byref(cid)) == _lib.RCB_EINVAL
    assert L.rcb_ddc_open(h, 4, t.ctypes.data, 8, 0.0, 0.0, OUT_IQ, 1.0, C.byref(cid)) == _lib.RCB_EINVAL
    big = np.ones(20000, np.float32)
    assert L.rcb_ddc_open(h, 4, big.ctypes.data, 20000, 0.0, 1.0, OUT_IQ, 1.0, C.byref(cid)) == _lib.RCB_EUNSUPPORTED
    assert L.rcb_ddc_retune(h, 12345, 1.0) == _lib.RCB_ERANGE
    assert L.rcb_ddc_close(h, 12345) == _lib.RCB_ERANGE
    assert L.rcb_ddc_pull(h, 12345, OUT_IQ, o.ctypes.data, 10, 0, C.byref(n)) == _lib.RCB_ERANGE
    assert L.rcb_convert_iq(h, x.ctypes.data, 9, 0.0, 1.0, 4, 0, o.ctypes.data, 0) == _lib.RCB_EINVAL
    assert L.rcb_quad_demod(h, None, 1, 4, 4, 1.0, None, o.ctypes.data, 4, 0) == _lib.RCB_EINVAL
    w = np.ones(4096, np.float32)
    assert L.rcb_fft_config(h, 4096, w.ctypes.data, 0) == _lib.RCB_EINVAL
    assert L.rcb_fft_config(h, 5000, w.ctypes.data, 4) == _lib.RCB_EUNSUPPORTED
    for code in (_lib.RCB_EINVAL, _lib.RCB_ESTATE, _lib.RCB_ERANGE, _lib.RCB_EUNSUPPORTED):
        assert L.rcb_strerror(code)
    # the handle still works after all of that
    ch = PfbChannelizer(engine, 64, np.ones(64, np.float32) / 64, OUT_IQ)
    iq, _ = ch.process(np.ones(64 * 8, np.complex64))
    assert iq.shape == (64, 8) and np.isfinite(iq).all()


def test_pull_semantics_and_missing_fm(engine, built_lib):
    bank = DdcBank(engine)
    taps = np.ones(5, np.float32) / 5
    c = bank.open(4, taps, 0.0, 1.0, OUT_IQ)
    assert len(bank.pull(c)) == 0                                # nothing processed yet
    bank.process(np.ones(40, np.complex64))
    y = bank.pull(c)
    assert len(y) == 10 and np.allclose(y[2:], 1.0, atol=1e-6)
    n = C.c_size_t()
    small = np.zeros(4, np.complex64)
    st = built_lib.rcb_ddc_pull(engine.h, c, OUT_IQ, small.ctypes.data, 4, 0, C.byref(n))
    assert st == _lib.RCB_ERANGE and n.value == 10               # too small: size reported, nothing copied
    st = built_lib.rcb_ddc_pull(engine.h, c, OUT_FM, small.ctypes.data, 4, 0, C.byref(n))
    assert st == _lib.RCB_ESTATE                                 # FM was not requested for this channel
    bank.close(c)
    with pytest.raises(_lib.B200ChanError):
        bank.retune(c, 1.0)


def test_fft_capacity_and_partial_blocks(engine, built_lib):
    w = np.ones(4096, np.float32)
    sc = FftScanner(engine, 4096, w, 4)
    x = (np.ones(4096 * 3) * 0.5).astype(np.complex64)
    assert sc.process(x).shape == (0, 4096)                      # 3 of 4 frames: nothing emitted yet
    assert sc.process(x[:4096]).shape == (1, 4096)               # the 4th frame completes the block
    n = C.c_size_t()
    xx = np.zeros(4096 * 8, np.complex64)
    out = np.zeros((1, 4096), np.float32)
    st = built_lib.rcb_fft_process(engine.h, xx.ctypes.data, len(xx), 0, out.ctypes.data, 1, 0, C.byref(n))
    assert st == _lib.RCB_ERANGE                                 # 2 vectors would be produced, room for 1
    with pytest.raises(ValueError):
        sc.process(np.zeros(100, np.complex64))


def test_two_handles_are_independent(built_lib):
    a, b = Engine(0), Engine(0)
    try:
        t1, t2 = np.ones(64, np.float32) / 64, np.ones(128, np.float32) / 128
        ca = PfbChannelizer(a, 64, t1, OUT_IQ)
        cb = PfbChannelizer(b, 64, t2, OUT_IQ | OUT_FM, 2.0)
        rng = np.random.default_rng(0)
        x = (rng.standard_normal(64 * 40) + 1j * rng.standard_normal(64 * 40)).astype(np.complex64)
        ya1, _ = ca.process(x)
        yb1, fb1 = cb.process(x)
        ca.reset()
        ya2, _ = ca.process(x)
        assert np.array_equal(ya1, ya2) and not np.array_equal(ya1, yb1) and fb1.shape == (64, 40)
        sa, sb = a.stats(), b.stats()
        assert sa["samples_in"] == 2 * len(x) and sb["samples_in"] == len(x)
        assert sa["kernel_launches"] > 0 and sa["h2d_bytes"] >= 2 * x.nbytes
    finally:
        a.close()
        b.close()
