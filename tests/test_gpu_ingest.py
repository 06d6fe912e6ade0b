"""f4: the SDR wire formats - RTL-SDR u8, UHD sc8 / sc16 - read directly by the channelizer.

`rcb_pfb_set_input_format` makes rcb_pfb_process take interleaved integer I/Q.  The 1024-channel one-tap-per-arm FM kernel
converts inside its first FFT pass (2-4x fewer HBM / PCIe bytes in); other shapes convert the block once on the device.
Parity: against the float64 oracle fed with the same samples converted by K5's rule (v + offset) * scale, at the usual
1e-5; against the complex64 path of the same kernel bit-for-bit when the scale is a power of two (the real formats)."""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd
from radiocapture_rf_b200 import _lib
from radiocapture_rf_b200.engine import PfbChannelizer, OUT_FM, OUT_IQ

pytestmark = pytest.mark.gpu

FORMATS = {
    "u8": (_lib.FMT_U8, np.uint8, -127.4, 1.0 / 128.0),            # gr-osmosdr rtl_source_c
    "s8": (_lib.FMT_S8, np.int8, 0.0, 1.0 / 128.0),                # UHD sc8 (configs/config_denver_usrp.py:20)
    "s16": (_lib.FMT_S16, np.int16, 0.0, 1.0 / 32768.0),           # UHD sc16
}


def _raw_stream(n, nch, dtype, seed):
    """Quantised multi-carrier stream: FM carriers in every 8th bin + noise, scaled to use the integer range."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01
    for m in range(1, nch, 8):
        f = (m + rng.uniform(-0.2, 0.2)) / nch
        dev = 0.02 * np.sin(2 * np.pi * rng.uniform(1e-4, 4e-4) * t)
        x += 0.05 * np.exp(2j * np.pi * (f * t + np.cumsum(dev)))
    x /= np.abs(x).max() * 1.05
    info = np.iinfo(dtype)
    if dtype == np.uint8:
        q = np.clip(np.round(np.stack([x.real, x.imag], 1) * 127.0 + 127.4), 0, 255)
    else:
        q = np.clip(np.round(np.stack([x.real, x.imag], 1) * info.max), info.min, info.max)
    return q.astype(dtype).reshape(-1)


def _fm_err(fm, ref, gain):
    d = (np.asarray(fm, np.float64) - ref) / gain
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(d) / max(np.linalg.norm(ref / gain), 1e-30)


@pytest.mark.parametrize("name", ["u8", "s8", "s16"])
@pytest.mark.parametrize("nch,tpa,mode", [(1024, 0.25, OUT_FM), (1024, 1, OUT_FM), (1024, 4, OUT_FM), (64, 2, OUT_IQ | OUT_FM)])
def test_raw_input_matches_oracle_and_float_path(engine, name, nch, tpa, mode):
    fmt, dtype, off, scale = FORMATS[name]
    frames = 203
    raw = _raw_stream(nch * frames, nch, dtype, seed=11)
    xf = ((raw.astype(np.float32) + np.float32(off)) * np.float32(scale)).view(np.complex64)   # K5's rule
    taps = fd.pfb_prototype(nch, tpa)
    ch = PfbChannelizer(engine, nch, taps, mode, 5.0)
    iq_f, fm_f = ch.process(xf)
    ch.set_input_format(fmt, off, scale)
    # streamed in ragged blocks: the carried history (raw and complex64 copies) must line up
    parts, pos = [], 0
    for b in (1, 50, 7, 145):
        parts.append(ch.process(raw[2 * pos * nch:2 * (pos + b) * nch]))
        pos += b
    assert pos == frames
    fm = np.concatenate([p[1] for p in parts], axis=1)
    ref = gb.pfb_channelizer(xf, np.asarray(taps, np.float64), nch)
    fref = gb.quadrature_demod(ref, 5.0)
    # (one tap per arm is no channel filter: every bin also carries the leakage of all other carriers and, for the 8-bit
    # formats, the quantisation noise - float32 rounding is relative to that total, so the bar is 5e-5 there; the fused
    # conversion itself is pinned bit-for-bit against the complex64 path below)
    tol = 1e-5 if tpa >= 2 else 5e-5
    for m in range(1, nch, 8):
        assert _fm_err(fm[m], fref[m], 5.0) <= tol, (name, m)
    # same arithmetic as the complex64 path (scale is a power of two: folding it into the taps is exact)
    assert np.array_equal(fm, fm_f)
    if mode & OUT_IQ:
        iq = np.concatenate([p[0] for p in parts], axis=1)
        assert np.array_equal(iq, iq_f)
    ch.set_input_format(0)
    _, again = ch.process(xf)
    assert np.array_equal(again, fm_f)


def test_raw_u8_zero_history_is_true_zero(engine):
    """Stream start: the samples before the first one are 0.0, not the u8 code for 'about zero' (offset -127.4):
    the first FM output of every channel is atan2(0, 0) -> 0 exactly like the complex64 path."""
    fmt, dtype, off, scale = FORMATS["u8"]
    nch, frames = 1024, 24
    raw = _raw_stream(nch * frames, nch, dtype, seed=3)
    ch = PfbChannelizer(engine, nch, fd.pfb_prototype(4, 64), OUT_FM, 5.0)
    ch.set_input_format(fmt, off, scale)
    _, fm = ch.process(raw)
    assert not fm[:, 0].any()
    assert np.abs(fm[1::8, 1:]).max() > 0


def test_raw_input_device_resident_full_iteration_blocks(engine):
    """Device-resident raw input with the blocked output layout (what bench.py's ingest lines run)."""
    fmt, dtype, off, scale = FORMATS["s16"]
    nch, frames, block = 1024, 4096 + 24, 1024
    raw = _raw_stream(nch * frames, nch, dtype, seed=5)
    xf = ((raw.astype(np.float32) + np.float32(off)) * np.float32(scale)).view(np.complex64)
    ch = PfbChannelizer(engine, nch, fd.pfb_prototype(4, 64), OUT_FM, 5.0)
    _, ref = ch.process(xf)
    ch.set_input_format(fmt, off, scale)
    ch.set_out_block(block)
    d_in = engine.dev_alloc(raw.nbytes)
    from radiocapture_rf_b200._lib import COPY_H2D, check
    check(engine.lib.rcb_memcpy(engine.h, d_in.ptr, raw.ctypes.data, raw.nbytes, COPY_H2D), "h2d", engine.h)
    nb = -(-frames // block)
    d_fm = engine.dev_alloc(nb * nch * block * 4)
    assert ch.process_device(d_in, nch * frames, None, d_fm, 0) == frames
    engine.sync()
    got = PfbChannelizer.unblock(engine.to_host(d_fm, (nb * nch * block,), np.float32), nch, frames, block)
    assert np.array_equal(got, ref)
