"""K1 parity: CUDA polyphase channelizer (+fused FM) vs the float64 oracle, through the C ABI.

Tolerance (BASELINE.json north_star): ||gpu - oracle||2 / ||oracle||2 <= 1e-5 per channel for float
outputs (after discarding the first ceil(L/N) outputs is NOT needed here because history is zero in
both); FM compared on the wrapped phase difference.
"""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import PfbChannelizer, OUT_FM, OUT_IQ

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _fm_err(fm, ref, gain):
    d = (fm.astype(np.float64) - ref) / gain
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(d) / max(np.linalg.norm(ref / gain), 1e-30)


def _check(engine, n, taps, frames, seed, mode=OUT_IQ | OUT_FM, gain=5.0, blocks=None, active_every=2):
    x, offs = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, seed, active_every=active_every)
    ch = PfbChannelizer(engine, n, taps, mode, gain)
    if blocks is None:
        iq, fm = ch.process(x)
    else:
        outs = []
        pos = 0
        for b in blocks:
            outs.append(ch.process(x[pos * n:(pos + b) * n]))
            pos += b
        assert pos == frames
        iq = np.concatenate([o[0] for o in outs], axis=1) if mode & OUT_IQ else None
        fm = np.concatenate([o[1] for o in outs], axis=1) if mode & OUT_FM else None
    ref = gb.pfb_channelizer(x, np.asarray(taps, np.float64), n)
    if mode & OUT_IQ:
        active = list(range(1, n, active_every))
        errs = [gb.rel_l2(iq[m], ref[m]) for m in active]
        assert max(errs) <= TOL, "IQ rel err %g at channel %d" % (max(errs), active[int(np.argmax(errs))])
        assert gb.rel_l2(iq, ref) <= TOL          # whole output
        # empty (noise-only, 30 dB down) bins: float32 rounding is relative to the strong bins that are
        # filtered out, so the per-channel ratio is looser there (same for GNU Radio's float32 VOLK path)
        errs = [gb.rel_l2(iq[m], ref[m]) for m in range(n)]
        assert max(errs) <= 1e-4, "IQ rel err %g at empty channel %d" % (max(errs), int(np.argmax(errs)))
    if mode & OUT_FM:
        fref = gb.quadrature_demod(ref, gain)
        active = list(range(1, n, active_every))
        errs = [_fm_err(fm[m], fref[m], gain) for m in active]
        assert max(errs) <= TOL, "FM rel err %g at channel %d" % (max(errs), active[int(np.argmax(errs))])
        # noise-only bins: atan2 of small products is ill conditioned; bound the aggregate loosely
        allerr = _fm_err(fm, fref, gain)
        assert allerr <= 2e-4, "FM aggregate err %g" % allerr
    return iq, fm


@pytest.mark.parametrize("n,tpa,frames", [(64, 2, 1024), (64, 16, 520), (256, 16, 256), (256, 1, 333),
                                          (1024, 16, 96), (1024, 4, 131)])
def test_pfb_iq_fm_parity(engine, n, tpa, frames):
    taps = fd.pfb_prototype(n, tpa)
    _check(engine, n, taps, frames, seed=n + tpa)


@pytest.mark.parametrize("n,tpa,frames", [(1024, 16, 96), (1024, 3, 131), (1024, 2, 200), (1024, 8, 77),
                                          (256, 16, 300), (256, 12, 257), (256, 2, 130),
                                          (64, 8, 700), (64, 2, 1024), (64, 16, 333)])
def test_pfb_fm_only_multi_tap_fast_kernel(engine, n, tpa, frames):
    """FM-only output with 2..16 taps per arm runs the time-blocked-FIR variant of the headline kernel
    (P rounded up to a power of two with zero taps): same parity bar as every other path."""
    taps = fd.pfb_prototype(n, tpa)
    _check(engine, n, taps, frames, seed=1000 + n + tpa, mode=OUT_FM)


@pytest.mark.parametrize("n,tpa", [(1024, 16), (256, 5), (64, 4)])
def test_pfb_fm_only_multi_tap_split_invariance(engine, n, tpa):
    taps = fd.pfb_prototype(n, tpa)
    frames = 257
    _, a_fm = _check(engine, n, taps, frames, seed=8, mode=OUT_FM)
    _, b_fm = _check(engine, n, taps, frames, seed=8, mode=OUT_FM, blocks=[1, 2, 3, 100, 9, 1, 141])
    assert np.array_equal(a_fm, b_fm)


@pytest.mark.parametrize("n,tpa,frames", [(1024, 16, 1), (1024, 16, 15), (1024, 16, 16), (1024, 16, 17), (1024, 1, 33),
                                          (1024, 0.25, 48), (256, 16, 1), (256, 16, 31), (256, 1, 32), (256, 8, 49),
                                          (1024, 16, 2500), (256, 16, 5000)])
def test_pfb_cluster_kernel_iteration_edges(engine, n, tpa, frames):
    """pfb_cl_kernel works in 16-frame iterations with one warm-up iteration per CTA run, TMA-fed rows (zero fill
    outside the stream) and TMA tile stores: block lengths around the iteration size, single frames, and blocks long
    enough that every cluster owns a run (2500 frames = 157 iterations over 74 clusters)."""
    taps = fd.pfb_prototype(n, tpa)
    _check(engine, n, taps, frames, seed=2000 + n + frames, mode=OUT_FM)


@pytest.mark.parametrize("n,tpa,block", [(1024, 16, 16), (1024, 1, 32), (256, 16, 16), (256, 4, 1024)])
def test_pfb_cluster_kernel_blocked_layouts(engine, n, tpa, block):
    """Time-blocked output through the 4-D TMA store (block >= 16 frames), ragged last block, vs the plain layout."""
    taps = fd.pfb_prototype(n, tpa)
    frames = 1000 + 7
    x, _ = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, 23)
    ch = PfbChannelizer(engine, n, taps, OUT_FM, 5.0)
    _, fm_ref = ch.process(x)
    ch.reset()
    ch.set_out_block(block)
    nb = -(-frames // block)
    d_in = engine.to_device(x)
    d_fm = engine.dev_alloc(nb * n * block * 4)
    assert ch.process_device(d_in, len(x), None, d_fm, 0) == frames
    engine.sync()
    fm = PfbChannelizer.unblock(engine.to_host(d_fm, (nb * n * block,), np.float32), n, frames, block)
    assert np.array_equal(fm, fm_ref)


def test_pfb_cluster_kernel_unaligned_output_falls_back(engine):
    """An output row stride that is not a multiple of 16 bytes cannot be described by a TMA tensor map: the call is
    served by the round-1 kernel and must give the same samples."""
    n, frames = 1024, 100
    taps = fd.pfb_prototype(n, 4)
    x, _ = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, 29)
    ch = PfbChannelizer(engine, n, taps, OUT_FM, 5.0)
    _, fm_ref = ch.process(x)
    ch.reset()
    d_in = engine.to_device(x)
    stride = frames + 1
    d_fm = engine.dev_alloc(n * stride * 4)
    assert ch.process_device(d_in, len(x), None, d_fm, stride) == frames
    engine.sync()
    fm = engine.to_host(d_fm, (n, stride), np.float32)[:, :frames]
    # different kernel, different (equally valid) float32 rounding: same parity bar as against the oracle
    for m in range(1, n, 2):
        assert _fm_err(fm[m], fm_ref[m].astype(np.float64), 5.0) <= 2e-5
    assert _fm_err(fm, fm_ref.astype(np.float64), 5.0) <= 2e-4


def test_pfb_host_pipeline_blocked_layout_spans_chunks(engine):
    """Host in / host out through the chunked H2D | kernel | D2H pipeline with the blocked output layout: several
    4 Mi-sample chunks, a ragged last block, against the plain layout of the same stream."""
    n, frames, block = 1024, 3 * 4096 + 1000 + 5, 1024
    taps = fd.pfb_prototype(4, 64)
    x, _ = synth.pfb_stream(n * 2048, 1.0e6 * n / 4.0, n, 41, active_every=8)
    x = np.tile(x, -(-frames // 2048))[:n * frames]
    ch = PfbChannelizer(engine, n, taps, OUT_FM, 5.0)
    _, ref = ch.process(x)
    ch.reset()
    ch.set_out_block(block)
    _, got = ch.process(x)
    assert got.shape == (-(-frames // block), n, block)
    assert np.array_equal(PfbChannelizer.unblock(got, n, frames, block), ref)


def test_copy_ceiling_reports_both_directions(engine):
    from radiocapture_rf_b200.engine import copy_ceiling
    h2d, d2h, wall = copy_ceiling(engine, 64 << 20, 32 << 20, iters=3)
    assert h2d > 1.0 and d2h > 1.0 and wall > 0


@pytest.mark.parametrize("n,tpa,frames,mode", [(64, 20, 300, OUT_IQ | OUT_FM), (256, 40, 100, OUT_FM), (1024, 24, 64, OUT_FM),
                                               (1024, 256, 300, OUT_FM)])
def test_pfb_more_than_16_taps_per_arm(engine, n, tpa, frames, mode):
    """P > 16 runs the register-prefetch kernel (pfb_fm_kernel, taps streamed from global memory): the third reading
    of BASELINE's "256-tap / 1024-channel" (256 taps PER ARM, bench.py also.tap_readings) included."""
    taps = fd.pfb_prototype(n, tpa)
    assert -(-len(taps) // n) == tpa
    _check(engine, n, taps, frames, seed=3000 + tpa, mode=mode, active_every=8 if n == 1024 else 2)


@pytest.mark.parametrize("n,tpa,streams,block", [(256, 16, 8, 0), (256, 4, 3, 64), (1024, 2, 2, 0)])
def test_pfb_multi_stream_launch_equals_per_stream_calls(built_lib, n, tpa, streams, block):
    """BASELINE config 5 shape: several independent 256-channel streams on one GPU through rcb_pfb_process_multi (one
    persistent launch that walks the streams) give bit for bit what per-stream calls give, block after block (streaming state per
    handle); other shapes take the stream-after-stream path of the same call."""
    from radiocapture_rf_b200.engine import Engine, pfb_process_multi
    taps = fd.pfb_prototype(n, tpa)
    frames = [200, 56]      # (multiples of 8: a plain-layout row stride the sector-store / TMA kernels can address)
    engines = [Engine(0) for _ in range(streams)]
    try:
        xs = [synth.pfb_stream(n * sum(frames), 1.0e6 * n / 4.0, n, 50 + i)[0] for i in range(streams)]
        refs = []
        for i, e in enumerate(engines):
            ch = PfbChannelizer(e, n, taps, OUT_FM, 5.0)
            refs.append(ch.process(xs[i])[1])
        chs = [PfbChannelizer(e, n, taps, OUT_FM, 5.0) for e in engines]     # reconfigure: fresh streaming state
        if block:
            for c in chs:
                c.set_out_block(block)
        pos = 0
        got = [[] for _ in range(streams)]
        for f in frames:
            d_ins = [e.to_device(xs[i][pos * n:(pos + f) * n]) for i, e in enumerate(engines)]
            nb = -(-f // block) if block else 0
            d_fms = [e.dev_alloc((nb * n * block if block else n * f) * 4) for e in engines]
            pfb_process_multi(chs, d_ins, n * f, d_fms, f)
            for i, e in enumerate(engines):
                e.sync()
                if block:
                    got[i].append(PfbChannelizer.unblock(e.to_host(d_fms[i], (nb * n * block,), np.float32), n, f, block))
                else:
                    got[i].append(e.to_host(d_fms[i], (n, f), np.float32))
            pos += f
        for i in range(streams):
            assert np.array_equal(np.concatenate(got[i], axis=1), refs[i]), i
    finally:
        for e in engines:
            e.close()


def test_pfb_cfg3_literal_256_taps_1024_channels(engine):
    """BASELINE config 3, literal reading: 256-tap prototype, 1024 channels -> 1 tap/arm, 768 zero arms."""
    taps = fd.pfb_prototype(4, 64)  # any 256-tap low-pass
    assert len(taps) == 256
    _check(engine, 1024, taps, 200, seed=3, mode=OUT_FM)
    _check(engine, 1024, taps, 64, seed=33, mode=OUT_IQ)


def test_pfb_cfg2_64ch_128taps(engine):
    taps = fd.pfb_prototype(64, 2)
    assert len(taps) == 128
    _check(engine, 64, taps, 4096, seed=2, mode=OUT_IQ)


def test_pfb_ragged_tap_count(engine):
    """L not a multiple of N: arms are zero padded (polyphase_filterbank::set_taps)."""
    taps = fd.pfb_prototype(64, 5)[:-17]
    _check(engine, 64, taps, 300, seed=5)


@pytest.mark.parametrize("n", [5, 6, 20, 25, 40, 96])
def test_pfb_generic_channel_counts(engine, n):
    """The reference's own PFB shapes: N = fs/400 kHz (rc_frontend/receiver.py:244-249), optfir prototype."""
    taps = fd.pfb_prototype(n)
    _check(engine, n, taps, 400, seed=100 + n)


@pytest.mark.parametrize("n,tpa", [(64, 4), (1024, 2), (20, None)])
def test_pfb_split_invariance(engine, n, tpa):
    """Block-wise processing == one shot (streaming state = last P input rows)."""
    taps = fd.pfb_prototype(n, tpa) if tpa else fd.pfb_prototype(n)
    frames = 257
    a_iq, a_fm = _check(engine, n, taps, frames, seed=7)
    b_iq, b_fm = _check(engine, n, taps, frames, seed=7, blocks=[1, 2, 3, 100, 9, 1, 141])
    assert np.array_equal(a_iq, b_iq)
    # FM of the first frame of a block is recomputed from the carried input rows: identical arithmetic
    assert np.array_equal(a_fm, b_fm)


def test_pfb_edge_cases(engine, built_lib):
    import ctypes as C
    from radiocapture_rf_b200 import _lib
    taps = fd.pfb_prototype(64, 2)
    ch = PfbChannelizer(engine, 64, taps, OUT_IQ)
    iq, fm = ch.process(np.zeros(0, np.complex64))  # empty input
    assert iq.shape == (64, 0) and fm is None
    with pytest.raises(ValueError):
        ch.process(np.zeros(63, np.complex64))
    n = C.c_size_t()
    x = np.zeros(63, np.complex64)
    st = built_lib.rcb_pfb_process(engine.h, x.ctypes.data, 63, 0, x.ctypes.data, None, 1, 0, C.byref(n))
    assert st == _lib.RCB_EINVAL
    # single frame, all-zero input -> zeros, FM atan2(0,0) = 0
    ch2 = PfbChannelizer(engine, 1024, fd.pfb_prototype(1024, 2), OUT_IQ | OUT_FM, 5.0)
    iq, fm = ch2.process(np.zeros(1024, np.complex64))
    assert not iq.any() and not fm.any()


def test_pfb_tone_lands_in_its_bin(engine):
    """Tone at bin centre m -> energy only in output m; FM of a tone offset by df is constant
    gain*2*pi*df/fs_out (SURVEY 8(c) golden (iii))."""
    n, frames = 64, 2048
    taps = fd.pfb_prototype(n, 8)
    fs = 64.0
    m = 11
    df = 0.05  # Hz with fs_out = 1 Hz
    t = np.arange(n * frames)
    x = np.exp(2j * np.pi * (m * fs / n + df) * t / fs).astype(np.complex64)
    ch = PfbChannelizer(engine, n, taps, OUT_IQ | OUT_FM, 5.0)
    iq, fm = ch.process(x)
    p = (np.abs(iq[:, 64:]) ** 2).mean(axis=1)
    assert int(np.argmax(p)) == m
    assert p[m] / (p.sum() - p[m]) > 1e4
    np.testing.assert_allclose(fm[m, 64:], 5.0 * 2 * np.pi * df, rtol=0, atol=2e-4)


@pytest.mark.parametrize("n,tpa,mode,block", [(1024, 0.25, OUT_FM, 64), (256, 4, OUT_IQ | OUT_FM, 64), (20, None, OUT_FM, 64),
                                              (1024, 0.25, OUT_FM, 8), (1024, 2, OUT_IQ | OUT_FM, 8), (64, 2, OUT_FM, 8)])
def test_pfb_blocked_device_output_layout(engine, n, tpa, mode, block):
    """rcb_pfb_set_out_block: channel-major inside time blocks (device-resident outputs) carries exactly
    the same samples as the plain [N][T] layout, including a ragged last block."""
    taps = fd.pfb_prototype(n, tpa) if tpa else fd.pfb_prototype(n)
    frames = 300
    x, _ = synth.pfb_stream(n * frames, 1.0e6 * n / 4.0, n, 21)
    ch = PfbChannelizer(engine, n, taps, mode, 5.0)
    iq_ref, fm_ref = ch.process(x)
    ch.reset()
    ch.set_out_block(block)
    nb = -(-frames // block)
    d_in = engine.to_device(x)
    d_iq = engine.dev_alloc(nb * n * block * 8) if mode & OUT_IQ else None
    d_fm = engine.dev_alloc(nb * n * block * 4) if mode & OUT_FM else None
    nout = ch.process_device(d_in, len(x), d_iq, d_fm, 0)
    engine.sync()
    assert nout == frames
    if mode & OUT_FM:
        fm = PfbChannelizer.unblock(engine.to_host(d_fm, (nb * n * block,), np.float32), n, frames, block)
        assert np.array_equal(fm, fm_ref)
    if mode & OUT_IQ:
        iq = PfbChannelizer.unblock(engine.to_host(d_iq, (nb * n * block,), np.complex64), n, frames, block)
        assert np.array_equal(iq, iq_ref)
    # host outputs use the same blocked layout (one contiguous D2H per pipeline chunk)
    h_iq, h_fm = ch.process(x) if (ch.reset() or True) else (None, None)
    if mode & OUT_FM:
        assert h_fm.shape == (nb, n, block)
        assert np.array_equal(PfbChannelizer.unblock(h_fm, n, frames, block), fm_ref)
    if mode & OUT_IQ:
        assert np.array_equal(PfbChannelizer.unblock(h_iq, n, frames, block), iq_ref)
    ch.set_out_block(0)
    with pytest.raises(Exception):
        ch.set_out_block(12)     # must be a power of two >= 8
