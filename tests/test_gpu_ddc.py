"""K2 / K4 parity: DDC bank (freq_xlating_fir_filter_ccc per channel) + quad demod vs the float64 oracle."""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import DdcBank, OUT_FM, OUT_IQ

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _fm_err(fm, ref, gain):
    d = (fm.astype(np.float64) - ref) / gain
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return np.linalg.norm(d) / max(np.linalg.norm(ref / gain), 1e-30)


def test_cfg1_single_channel_xlat_fm(engine):
    """BASELINE config 1: fs 2.4 Msps, offset -62.5 kHz, rate 12500 -> D 96, 349 taps, gains 5 and 6.6315."""
    x, fs, offs = synth.cfg1(1 << 20, seed=1)
    decim, taps = fd.channel_taps(fs, 12500)
    assert decim == 96 and len(taps) == 349
    bank = DdcBank(engine)
    g2 = 25000.0 / (2 * np.pi * 600.0)
    c1 = bank.open(decim, taps, -62500.0, fs, OUT_IQ | OUT_FM, 5.0)
    c2 = bank.open(decim, taps, -62500.0, fs, OUT_IQ | OUT_FM, g2)
    bank.process(x)
    y1 = bank.pull(c1, OUT_IQ)
    f1 = bank.pull(c1, OUT_FM)
    f2 = bank.pull(c2, OUT_FM)
    ref = gb.freq_xlating_fir(x, taps, decim, -62500.0, fs)
    assert len(y1) == len(ref) == (1 << 20) // 96 + 1 - (1 if ((1 << 20) % 96 == 0) else 0) or len(y1) == len(ref) + 1
    n = min(len(y1), len(ref))
    assert gb.rel_l2(y1[:n], ref[:n]) <= TOL
    fref = gb.quadrature_demod(ref[:n], 1.0)
    assert _fm_err(f1[:n], 5.0 * fref, 5.0) <= TOL
    assert _fm_err(f2[:n], g2 * fref, g2) <= TOL


def test_ddc_output_count_and_decimation_phase(engine):
    """Output i's newest sample is x[i*D]: a block of n samples yields ceil(n/D) outputs from a fresh channel."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(np.complex64)
    taps = fd.low_pass(1, 8.0, 1.0, 1.0)
    bank = DdcBank(engine)
    c = bank.open(7, taps, 0.3, 8.0)
    bank.process(x)
    y = bank.pull(c)
    assert len(y) == (1000 + 6) // 7
    # oracle convention: one output per D inputs, first output uses x[0] as newest sample
    xr = np.concatenate([x, np.zeros(6, np.complex64)])
    ref = gb.freq_xlating_fir(xr, taps, 7, 0.3, 8.0)
    assert gb.rel_l2(y, ref[:len(y)]) <= TOL


def test_ddc_many_channels_share_one_block(engine):
    x, fs, offs = synth.cfg1(1 << 18, seed=11)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    ids = [bank.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in offs]
    bank.process(x)
    for cid, f in zip(ids, offs):
        y = bank.pull(cid)
        ref = gb.freq_xlating_fir(x, taps, decim, f, fs)
        n = min(len(y), len(ref))
        assert gb.rel_l2(y[:n], ref[:n]) <= TOL, f


def test_ddc_split_invariance_and_ragged_blocks(engine):
    from radiocapture_rf_b200.engine import Engine
    x, fs, offs = synth.cfg1(200000, seed=5)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    a = bank.open(decim, taps, 437500.0, fs, OUT_IQ | OUT_FM, 5.0)
    bank.process(x)
    ya, fa = bank.pull(a, OUT_IQ), bank.pull(a, OUT_FM)
    e2 = Engine(0)                      # a fresh stream: same samples fed in ragged blocks
    try:
        bank2 = DdcBank(e2)
        b = bank2.open(decim, taps, 437500.0, fs, OUT_IQ | OUT_FM, 5.0)
        ys, fs_ = [], []
        pos = 0
        for blk in [1, 95, 96, 97, 1000, 12345, 7, 0, 50000]:
            bank2.process(x[pos:pos + blk])
            pos += blk
            ys.append(bank2.pull(b, OUT_IQ))
            fs_.append(bank2.pull(b, OUT_FM))
        bank2.process(x[pos:])
        ys.append(bank2.pull(b, OUT_IQ))
        fs_.append(bank2.pull(b, OUT_FM))
    finally:
        e2.close()
    yb, fb = np.concatenate(ys), np.concatenate(fs_)
    assert len(yb) == len(ya)
    # the first outputs of a channel run through the zero-history path, later blocks through the tiled
    # path: same samples to float32 rounding of a different summation order
    assert gb.rel_l2(yb, ya) <= 2e-6
    assert _fm_err(fb[1:], fa[1:].astype(np.float64), 5.0) <= 2e-5


def test_ddc_channel_opened_mid_stream_joins_the_decimation_grid(engine):
    """Channels of one source share a decimation grid: a channel opened after n samples starts at the next
    multiple of D with zero filter history (a GNU Radio channel starts at whatever sample its ZMQ SUB sees
    first), and from then on is computed together with the older channels by the tiled kernel."""
    x, fs, offs = synth.cfg1(96 * 1500 + 37, seed=6)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    a = bank.open(decim, taps, -62500.0, fs)
    n0 = 96 * 500 + 37
    bank.process(x[:n0])
    ya0 = bank.pull(a)
    b = bank.open(decim, taps, 12500.0, fs, OUT_IQ | OUT_FM, 5.0)
    bank.process(x[n0:n0 + 96 * 400])
    ya1, yb1 = bank.pull(a), bank.pull(b)
    bank.process(x[n0 + 96 * 400:])
    ya2, yb2 = bank.pull(a), bank.pull(b)
    ya = np.concatenate([ya0, ya1, ya2])
    ref_a = gb.freq_xlating_fir(np.concatenate([x, np.zeros(96, np.complex64)]), taps, decim, -62500.0, fs)
    assert gb.rel_l2(ya, ref_a[:len(ya)]) <= TOL
    start = -(-n0 // 96) * 96                                  # next grid point
    yb = np.concatenate([yb1, yb2])
    ref_b = gb.freq_xlating_fir(np.concatenate([x[start:], np.zeros(96, np.complex64)]), taps, decim, 12500.0, fs)
    assert len(yb) == len(ya) - start // 96
    assert gb.rel_l2(yb, ref_b[:len(yb)]) <= TOL


def test_ddc_retune_keeps_phase_continuous(engine):
    """set_center_freq (channel.py:61-63): composite taps and rotator increment change, phase continues."""
    fs = 2.4e6
    n = 96 * 2000
    t = np.arange(2 * n) / fs
    x = np.exp(2j * np.pi * 100e3 * t).astype(np.complex64)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    c = bank.open(decim, taps, 100e3, fs)
    bank.process(x[:n])
    y0 = bank.pull(c)
    bank.retune(c, 101e3)
    bank.process(x[n:])
    y1 = bank.pull(c)
    # first segment: tone at DC of the channel -> constant phase
    ph0 = np.angle(y0[50:])
    assert np.ptp(np.unwrap(ph0)) < 1e-3
    # after retune the tone sits at -1 kHz in the channel: phase slope = -2 pi 1000 / 25000 per sample
    ph1 = np.unwrap(np.angle(y1[50:]))
    slope = np.polyfit(np.arange(len(ph1)), ph1, 1)[0]
    assert abs(slope + 2 * np.pi * 1000.0 / 25000.0) < 1e-5
    # continuity: the phase right after the retune continues from the end of the first segment
    # (filter transient aside) - compare extrapolated phases
    end0 = ph0[-1]
    ph1_full = np.unwrap(np.angle(y1))
    start1 = ph1_full[50] - slope * 50
    # ... plus the FIR group delay seen by the now -1 kHz tone: +2 pi 1000 (ntaps-1)/2 / fs
    expect = 2 * np.pi * 1000.0 * (len(taps) - 1) / 2.0 / fs
    d = (start1 - end0 - expect + np.pi) % (2 * np.pi) - np.pi
    assert abs(d) < 0.02


def test_prefilter_1x_69_taps(engine):
    """p25_control_demod.py:106-108: freq_xlating_fir_filter_ccc(1, low_pass_2(1,25000,6250,500,30,BLACKMAN), 0, 25000)."""
    taps = fd.low_pass_2(1.0, 25000, 6250, 500.0, 30.0, fd.WIN_BLACKMAN)
    assert len(taps) == 69
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(50000) + 1j * rng.standard_normal(50000)).astype(np.complex64) * 0.1
    bank = DdcBank(engine)
    c = bank.open(1, taps, 0.0, 25000.0, OUT_IQ | OUT_FM, 25000 / (2 * np.pi * 600))
    bank.process(x)
    y = bank.pull(c)
    ref = gb.freq_xlating_fir(x, taps, 1, 0.0, 25000.0)
    assert len(y) == len(ref)
    assert gb.rel_l2(y, ref) <= TOL


def test_split2_half_band_pair(engine):
    """rc_frontend/receiver.py:78-86: two decim-2 DDCs at -fs/4 and +fs/4 with firdes.low_pass(1,fs,fs/4,fs/8)."""
    fs = 8.0e6
    taps = fd.low_pass(1, fs, fs / 4, fs / 8)
    assert len(taps) == 19
    x = synth.wideband(1 << 16, fs, [-2.5e6, -1e6, 0.4e6, 3e6], seed=9)
    bank = DdcBank(engine)
    lo = bank.open(2, taps, -fs / 4, fs)
    hi = bank.open(2, taps, fs / 4, fs)
    bank.process(x)
    for cid, f0 in ((lo, -fs / 4), (hi, fs / 4)):
        y = bank.pull(cid)
        ref = gb.freq_xlating_fir(x, taps, 2, f0, fs)
        assert gb.rel_l2(y, ref) <= TOL


def test_quad_demod_and_probe(engine):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((7, 5001)) + 1j * rng.standard_normal((7, 5001))).astype(np.complex64)
    fm, last = engine.quad_demod(x, 5.0)
    ref = gb.quadrature_demod(x, 5.0)
    assert _fm_err(fm, ref, 5.0) <= TOL
    assert np.array_equal(last, x[:, -1])
    # streaming carry
    fm_a, last_a = engine.quad_demod(x[:, :1234], 5.0)
    fm_b, _ = engine.quad_demod(x[:, 1234:], 5.0, prev=last_a)
    assert np.array_equal(np.concatenate([fm_a, fm_b], axis=1), fm)
    # (0,0) -> 0 like fast_atan2f
    z, _ = engine.quad_demod(np.zeros(16, np.complex64), 5.0)
    assert not z.any()
    # table-atan GR emulation differs from exact by <= 1.6e-6 rad * gain
    gr = gb.quadrature_demod_grcompat(x[0], 5.0)
    assert np.abs(gr - fm[0]).max() <= 5.0 * 2.0e-6 + 1e-6
    # AFC probe: moving_average_ff(10000,1,40000) -> *1e-4
    pm = engine.probe_mean(fm, 1000, 1e-3)
    ref_pm = gb.moving_average(ref.T, 1000, 1e-3)[-1]
    np.testing.assert_allclose(pm, ref_pm, atol=1e-5)


def test_gr_float_omega_mode_matches_gnuradio_emulation(engine):
    """With w rounded to float32 like GNU Radio, the GPU (exact phase) matches a float32 emulation of the
    GNU Radio block (complex64 taps, recursive rotator renormalised every 512) to 1e-5."""
    x, fs, _ = synth.cfg1(96 * 1200, seed=1)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    c = bank.open(decim, taps, -62500.0, fs, gr_float_omega=True)
    bank.process(x)
    y = bank.pull(c)
    gr = gb.freq_xlating_fir_grcompat(x, taps, decim, -62500.0, fs)
    assert gb.rel_l2(y[:len(gr)], gr) <= TOL


@pytest.mark.parametrize("fmt,dt,off,sc", [("u8", np.uint8, -127.4, 1 / 128.0), ("s8", np.int8, 0.0, 1 / 128.0),
                                           ("s16", np.int16, 0.0, 1 / 32768.0)])
def test_ingest_conversion_bit_exact(engine, fmt, dt, off, sc):
    """K5: SDR wire formats -> complex64, bit-exact against (v + offset) * scale in float32
    (u8: gr-osmosdr rtl_source_c's (v - 127.4)/128 mapping), including a ragged tail and empty input."""
    rng = np.random.default_rng(7)
    info = np.iinfo(dt)
    for n in (0, 1, 5, 4096, 100003):
        raw = rng.integers(info.min, info.max + 1, size=2 * n).astype(dt)
        y = engine.convert_iq(raw, fmt)
        ref = ((raw.astype(np.float32) + np.float32(off)) * np.float32(sc)).astype(np.float32).view(np.complex64)
        assert y.shape == (n,) and np.array_equal(y, ref)
    # converted samples feed the channelizer directly from device memory
    raw = rng.integers(0, 256, size=2 * 96 * 200).astype(np.uint8)
    d = engine.dev_alloc(96 * 200 * 8)
    assert engine.convert_iq(raw, "u8", out_device=d) == 96 * 200
    x = engine.to_host(d, (96 * 200,), np.complex64)
    assert np.array_equal(x, ((raw.astype(np.float32) - np.float32(127.4)) * np.float32(1 / 128.0)).view(np.complex64))


def test_pull_all_matches_per_channel_pulls_and_chunked_host_input(engine):
    """rcb_ddc_pull_all = every channel's block in one transfer; host input longer than one staging chunk (2^22 samples)
    is pipelined in chunks whose outputs are appended - same samples as per-channel pulls / one device-resident call."""
    fs = 2.4e6
    n = (1 << 22) + (1 << 20) + 12345
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * np.float32(0.1)
    decim, taps = fd.channel_taps(fs, 12500)
    bank = DdcBank(engine)
    ids = [bank.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in (-62500.0, 12500.0, 437500.0)]
    ids.append(bank.open(decim * 2, taps, 100000.0, fs, OUT_IQ, 1.0))     # another decimation grid, no FM
    bank.process(x)
    iq_all, fm_all = bank.pull_all(OUT_IQ), bank.pull_all(OUT_FM)
    assert sorted(iq_all) == sorted(ids)
    for c in ids:
        assert np.array_equal(iq_all[c], bank.pull(c, OUT_IQ))
    for c in ids[:3]:
        assert np.array_equal(fm_all[c], bank.pull(c, OUT_FM))
    assert len(fm_all[ids[3]]) == 0
    # reference: the same stream from device memory in one piece on a fresh handle (a handle = one wideband stream).
    # The first block of a channel runs the gated kernel (zero filter history), later chunks the tiled one: same
    # samples up to float32 rounding of the two kernels
    from radiocapture_rf_b200.engine import Engine
    e2 = Engine(0)
    try:
        bank2 = DdcBank(e2)
        ids2 = [bank2.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in (-62500.0, 12500.0, 437500.0)]
        d_x = e2.to_device(x)
        bank2.process_device(d_x, n)
        for a, b in zip(ids[:3], ids2):
            y2 = bank2.pull(b, OUT_IQ)
            assert len(y2) == len(iq_all[a])
            assert gb.rel_l2(iq_all[a], y2) <= 2e-6
            f2 = bank2.pull(b, OUT_FM)
            d = (fm_all[a] - f2) / 5.0
            d = (d + np.pi) % (2 * np.pi) - np.pi
            assert np.abs(d).max() <= 5e-3 and np.sqrt(np.mean(d * d)) <= 2e-5    # noise input: atan2 of small products
    finally:
        e2.close()
    y = iq_all[ids[0]]
    ref = gb.freq_xlating_fir(x[:96 * 4000], taps, decim, -62500.0, fs)
    assert gb.rel_l2(y[:len(ref)], ref) <= 1e-5


def test_two_engines_on_one_device_keep_their_shared_memory_opt_in(built_lib):
    """cudaFuncSetAttribute is per function and per device - a second handle with a smaller tile must not
    lower the limit the first handle's launches rely on (lone 12.5 kHz channel: 100 KB tile; group of 3: 26 KB)."""
    from radiocapture_rf_b200.engine import Engine
    fs = 2.4e6
    decim, taps = fd.channel_taps(fs, 12500)
    a, b = Engine(0), Engine(0)
    try:
        rng = np.random.default_rng(3)
        x = (rng.standard_normal(96 * 500) + 1j * rng.standard_normal(96 * 500)).astype(np.complex64)
        ba, bb = DdcBank(a), DdcBank(b)
        ca = ba.open(decim, taps, -62500.0, fs)
        cbs = [bb.open(decim, taps, f, fs) for f in (1e4, 2e4, 3e4)]
        outs = []
        for _ in range(3):     # alternate: A (large tile), B (small tile), A again ...
            ba.process(x)
            outs.append(ba.pull(ca))
            bb.process(x)
            bb.pull(cbs[0])
        ref = gb.freq_xlating_fir(x, taps, decim, -62500.0, fs)
        assert gb.rel_l2(outs[0][:len(ref)], ref[:len(outs[0])]) <= 1e-5
        assert len(outs[2]) > 0
    finally:
        a.close()
        b.close()


def _open_bank(engine, decim, taps, offs, fs, tensor_cores, seg=0):
    bank = DdcBank(engine)
    bank.set_tensor_cores(tensor_cores, seg)
    ids = [bank.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in offs]
    return bank, ids


def _second_block(bank, ids, x, cut):
    """Feed x[:cut] (a fresh channel's first block runs on the zero-history kernel), then x[cut:]; return the second
    call's IQ / FM per channel and the index of its first output."""
    bank.process(x[:cut])
    first = {c: len(bank.pull(c, OUT_IQ)) for c in ids}
    l0 = bank.tensor_core_launches()
    bank.process(x[cut:])
    return ({c: bank.pull(c, OUT_IQ) for c in ids}, {c: bank.pull(c, OUT_FM) for c in ids}, first,
            bank.tensor_core_launches() - l0)


@pytest.mark.parametrize("mode,seg", [(1, 0), (1, 4), (1, 1000), (2, 3), (2, 1)])
def test_tensor_core_bank_matches_oracle_and_cuda_core_kernel(engine, mode, seg):
    """24 channels sharing (D 96, 349 taps): the bucket runs on ddc_mma_kernel (tcgen05 kind::tf32, hi/lo split), the
    block's first outputs on the CUDA-core kernel.  Channels on a carrier meet the 1e-5 bar against the float64 oracle;
    EMPTY channels (output 40-60 dB below the input: any fp32 summation order differs from float64 at the 1e-5 level
    there) are held to the error of the CUDA-core kernel on the same channel, both measured against the strongest
    channel's norm."""
    from radiocapture_rf_b200.engine import Engine
    x, fs, carriers = synth.cfg1((1 << 18) + (1 << 16), seed=21)
    cut = 1 << 16
    decim, taps = fd.channel_taps(fs, 12500)
    rng = np.random.default_rng(3)
    offs = list(carriers) + list(rng.uniform(-0.45 * fs, 0.45 * fs, 16))
    bank, ids = _open_bank(engine, decim, taps, offs, fs, mode, seg)
    ys, fms, first, nl = _second_block(bank, ids, x, cut)
    assert nl == 1
    e2 = Engine(0)
    try:
        bank2, ids2 = _open_bank(e2, decim, taps, offs, fs, False)
        ys2, _, first2, nl2 = _second_block(bank2, ids2, x, cut)
        assert nl2 == 0
        refs = [gb.freq_xlating_fir(x, taps, decim, f, fs) for f in offs]
        scale = max(np.linalg.norm(r) for r in refs)
        rows = []
        for k, (cid, cid2, f) in enumerate(zip(ids, ids2, offs)):
            y, y2 = ys[cid], ys2[cid2]
            ref = refs[k][first[cid]:]
            n = min(len(y), len(ref))
            assert n >= (len(x) - cut) // decim - 1 and len(y) == len(y2)
            rel = gb.rel_l2(y[:n], ref[:n])
            e_mma = np.linalg.norm(y[:n] - ref[:n]) / scale
            e_cc = np.linalg.norm(y2[:n] - ref[:n]) / scale
            rows.append((f, np.linalg.norm(ref[:n]) / scale, rel, e_mma, e_cc))
            if k < len(carriers):
                assert rel <= TOL, (f, rel)
                fref = gb.quadrature_demod(refs[k], 5.0)[first[cid]:]
                assert _fm_err(fms[cid][1:n], fref[1:n], 5.0) <= 2e-5
            assert e_mma <= max(3e-6, 3.0 * e_cc), (f, e_mma, e_cc)
    finally:
        e2.close()
    print("\ntensor-core bank, mode %d seg %d:  offset Hz | level | rel_l2 vs oracle | err/scale mma | err/scale cuda-core" % (mode, seg))
    for r in rows[:10]:
        print("  %10.0f  %.2e  %.2e  %.2e  %.2e" % r)


def test_tensor_core_bank_ragged_blocks_and_two_column_groups(engine):
    """70 channels (two column groups: 64 + 6) fed in ragged blocks - odd lengths flip the 16-byte alignment of the
    window base (the `lead` row of the B operand), short blocks fall back to the CUDA-core kernel - equal the one-shot
    result and the oracle."""
    from radiocapture_rf_b200.engine import Engine
    x, fs, _ = synth.cfg1(300000, seed=8)
    decim, taps = fd.channel_taps(fs, 12500)
    rng = np.random.default_rng(5)
    offs = list(rng.uniform(-0.45 * fs, 0.45 * fs, 70))
    offs[:8] = [-1.0375e6, -62.5e3, 0.0, 12.5e3, 437.5e3, 600e3, -300e3, 900e3]   # the carriers of synth.cfg1
    bank, ids = _open_bank(engine, decim, taps, offs, fs, False)
    bank.process(x)
    one = {c: bank.pull(c, OUT_IQ) for c in ids}
    scale = max(np.linalg.norm(v) for v in one.values())
    e2 = Engine(0)
    try:
        bank2, ids2 = _open_bank(e2, decim, taps, offs, fs, True)
        parts = {c: [] for c in ids2}
        pos = 0
        for blk in [50001, 96 * 700 + 1, 333, 100000, 7, 0, 40003]:
            bank2.process(x[pos:pos + blk])
            pos += blk
            for c in ids2:
                parts[c].append(bank2.pull(c, OUT_IQ))
        bank2.process(x[pos:])
        for c in ids2:
            parts[c].append(bank2.pull(c, OUT_IQ))
        assert bank2.tensor_core_launches() >= 4
        for k, (c, c2, f) in enumerate(zip(ids, ids2, offs)):
            yb = np.concatenate(parts[c2])
            assert len(yb) == len(one[c])
            assert np.linalg.norm(yb - one[c]) / scale <= 2e-6, f      # vs the CUDA-core kernels, one block
            if k < 8 or k >= 64:
                ref = gb.freq_xlating_fir(x, taps, decim, f, fs)
                n = min(len(yb), len(ref))
                if k < 8:
                    assert gb.rel_l2(yb[:n], ref[:n]) <= TOL, f
                assert np.linalg.norm(yb[:n] - ref[:n]) / scale <= 2e-6, f
    finally:
        e2.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_tensor_core_bank_wideband_2327_taps(engine, mode):
    """The many-channel bench shape: fs 16 Msps, D 640, 2327 taps (147 k-chunks of the MMA pipeline), 64 channels."""
    fs, rate = 16.0e6, 12500
    decim, taps = fd.channel_taps(fs, rate)
    assert decim == 640 and len(taps) == 2327
    n = 1 << 19
    rng = np.random.default_rng(9)
    offs = list(rng.uniform(-0.45 * fs, 0.45 * fs, 64))
    cut = 1 << 16
    x = synth.wideband(n + cut, fs, [offs[k] for k in (0, 1, 31, 62, 63)], seed=13)
    bank, ids = _open_bank(engine, decim, taps, offs, fs, mode)
    ys, _, first, nl = _second_block(bank, ids, x, cut)
    assert nl == 1
    for k in [0, 1, 31, 62, 63]:
        y = ys[ids[k]]
        ref = gb.freq_xlating_fir(x, taps, decim, offs[k], fs)[first[ids[k]]:]
        m = min(len(y), len(ref))
        assert m >= n // decim - 1
        err = gb.rel_l2(y[:m], ref[:m])
        print("2327 taps, mode %d, channel %d: rel_l2 %.2e" % (mode, k, err))
        assert err <= TOL, offs[k]


@pytest.mark.parametrize("decim,ntaps", [(96, 349), (96, 96), (96, 97), (8, 64), (9, 41), (33, 100), (50, 349),
                                         (200, 701), (640, 2327), (97, 500)])
def test_lone_channel_kernel_shapes(engine, decim, ntaps):
    """ddc_lone_kernel (a lone channel of a decimation grid: frame-per-lane, P = ceil(ntaps / decim) in 1..8, two frames
    per lane) and its fall-backs: odd and even P, decimations that are not multiples of the warp, a window shorter than
    one frame, tiles that do not fit shared memory (D 640); blocks that start inside the history, interior tiles and a
    ragged tail, streamed in uneven blocks."""
    rng = np.random.default_rng(decim * 1000 + ntaps)
    n = 122 * decim * 5 + 3 * decim + 7
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1).astype(np.complex64)
    x += (0.5 * np.exp(2j * np.pi * 0.1003 * np.arange(n))).astype(np.complex64)
    taps = (rng.standard_normal(ntaps) / np.sqrt(ntaps)).astype(np.float32)
    fs, f0 = 1.0e6, 0.1e6
    bank = DdcBank(engine)
    c = bank.open(decim, taps, f0, fs, OUT_IQ | OUT_FM, 3.0)
    ys, fms, pos = [], [], 0
    for b in (n // 3 + 1, 5, n // 2, n):
        b = min(b, n - pos)
        if b <= 0:
            break
        bank.process(x[pos:pos + b])
        ys.append(bank.pull(c, OUT_IQ))
        fms.append(bank.pull(c, OUT_FM))
        pos += b
    y, fm = np.concatenate(ys), np.concatenate(fms)
    ref = gb.freq_xlating_fir(x, taps, decim, f0, fs)
    m = min(len(y), len(ref))
    assert m >= n // decim
    assert gb.rel_l2(y[:m], ref[:m]) <= TOL
    assert _fm_err(fm[:m], gb.quadrature_demod(ref[:m], 3.0), 3.0) <= TOL


@pytest.mark.parametrize("name,dt,off,sc", [("u8", np.uint8, -127.4, 1 / 128.0), ("s8", np.int8, 0.0, 1 / 128.0),
                                            ("s16", np.int16, 0.0, 1 / 32768.0)])
@pytest.mark.parametrize("nchan,device_input", [(1, False), (3, True), (16, False)])
def test_ddc_wire_format_input(engine, name, dt, off, sc, nchan, device_input):
    """rcb_ddc_set_input_format (f4 for K2): the block crosses PCIe as u8 / sc8 / sc16 and is converted on the device.
    Same samples bit-for-bit as feeding the complex64 conversion (K5's rule) to a fresh bank - lone-channel, tiled and
    tensor-core kernels - streamed in ragged blocks incl. one longer than a staging chunk; <= 1e-5 against the oracle."""
    from radiocapture_rf_b200 import _lib
    from radiocapture_rf_b200.engine import Engine
    fmt = {"u8": _lib.FMT_U8, "s8": _lib.FMT_S8, "s16": _lib.FMT_S16}[name]
    fs = 2.4e6
    decim, taps = fd.channel_taps(fs, 12500)
    n = (1 << 22) + 96 * 300 + 11 if nchan == 1 else 96 * 3000 + 11
    rng = np.random.default_rng(5)
    info = np.iinfo(dt)
    t = np.arange(n)
    offs = [-62500.0, 437500.0, 12500.0] + list(rng.uniform(-1.0e6, 1.0e6, 13))
    offs = offs[:nchan]
    sig = sum(0.3 / len(offs) * np.exp(2j * np.pi * (f / fs * t + 0.3 * np.sin(2 * np.pi * 1e-4 * t))) for f in offs[:3])
    sig = sig + (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.02
    amp = 127.0 if info.max < 256 else float(info.max)
    raw = np.clip(np.round(np.stack([sig.real, sig.imag], 1) * amp - (off if name == "u8" else 0.0)), info.min,
                  info.max).astype(dt).reshape(-1)
    xf = ((raw.astype(np.float32) + np.float32(off)) * np.float32(sc)).view(np.complex64)
    blocks = (n // 5, 7, n) if nchan > 1 else (96 * 100 + 3, n)
    bank = DdcBank(engine)
    ids = [bank.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in offs]
    bank.set_input_format(fmt, off, sc)
    e2 = Engine(0)
    try:
        bank2 = DdcBank(e2)
        ids2 = [bank2.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in offs]
        got = {c: [] for c in ids}
        want = {c: [] for c in ids2}
        pos = 0
        for b in blocks:
            b = min(b, n - pos)
            if device_input:
                chunk = np.ascontiguousarray(raw[2 * pos:2 * (pos + b)])
                d_raw = engine.dev_alloc(max(chunk.nbytes, 16))
                _lib.check(engine.lib.rcb_memcpy(engine.h, d_raw.ptr, chunk.ctypes.data, chunk.nbytes, _lib.COPY_H2D),
                           "h2d", engine.h)
                bank.process_device(d_raw, b)
                d_x = e2.to_device(xf[pos:pos + b])
                bank2.process_device(d_x, b)
            else:
                bank.process(raw[2 * pos:2 * (pos + b)])
                bank2.process(xf[pos:pos + b])
            for c, c2 in zip(ids, ids2):
                got[c].append((bank.pull(c, OUT_IQ), bank.pull(c, OUT_FM)))
                want[c2].append((bank2.pull(c2, OUT_IQ), bank2.pull(c2, OUT_FM)))
            pos += b
        for c, c2 in zip(ids, ids2):
            y = np.concatenate([p[0] for p in got[c]])
            y2 = np.concatenate([p[0] for p in want[c2]])
            assert np.array_equal(y, y2)
            assert np.array_equal(np.concatenate([p[1] for p in got[c]]), np.concatenate([p[1] for p in want[c2]]))
        for c, f in list(zip(ids, offs))[:3]:
            y = np.concatenate([p[0] for p in got[c]])
            m = min(len(y), 3000)
            ref = gb.freq_xlating_fir(xf[:96 * m], taps, decim, f, fs)
            assert gb.rel_l2(y[:len(ref)], ref) <= TOL
        # back to complex64 without losing the stream position
        bank.set_input_format(0)
        bank.process(xf[:96 * 10])
        assert len(bank.pull(ids[0], OUT_IQ)) in (10, 11)
    finally:
        e2.close()
