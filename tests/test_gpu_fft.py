"""K3 parity: streaming FFT + log-power sums vs the float64 oracle; peak indices bit-exact vs the
reference's own scipy.signal.find_peaks call (fft_peak_detection.py:46-72)."""
import numpy as np
import pytest

from oracle import gr_blocks as gb, gr_firdes as fd, synth
from radiocapture_rf_b200.engine import FftScanner

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("length,avg,nblocks", [(4096, 10, 3), (16384, 100, 1), (65536, 4, 2), (262144, 3, 1),
                                                 (1048576, 2, 1)])
def test_fft_logpow_parity(engine, length, avg, nblocks):
    fs = 2.4e6
    n = length * avg * nblocks
    x, truth = synth.scan_stream(n, fs, length, seed=4, ncarriers=6)
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, avg)
    out = sc.process(x)
    assert out.shape == (nblocks, length)
    ref = gb.logpower_block_sums(x, length, w, avg)
    _check_logsum(out, ref, avg)


@pytest.mark.parametrize("length,avg,nblocks,frames_extra", [(4096, 10, 3, 4), (16384, 7, 2, 0), (262144, 3, 2, 1),
                                                              (1048576, 2, 2, 1)])
def test_fft_persistent_pipeline_matches_oracle_and_default(engine, length, avg, nblocks, frames_extra):
    """rcb_fft_set_pipeline(1): ONE persistent launch per call (task queue over column / row tiles, L2-resident scratch
    ring, block sums accumulated in frame order).  Same parity bar as the three-kernel pipeline, the same samples as it
    bit for bit, and split invariance (the running block sum is carried across calls)."""
    n = length * (avg * nblocks + frames_extra)
    x, _ = synth.scan_stream(n, 2.4e6, length, seed=14, ncarriers=5)
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, avg)
    sc.set_pipeline(2)          # the tiled three-kernel pipeline (16384-point frames default to fft_frame_kernel)
    base = sc.process(x)
    sc.set_pipeline(1)
    out = sc.process(x)
    assert out.shape == (nblocks, length)
    _check_logsum(out, gb.logpower_block_sums(x[:length * avg * nblocks], length, w, avg), avg)
    np.testing.assert_allclose(out, base, atol=2e-5 * avg)
    sc.reset()
    parts, pos = [], 0
    for nfr in [1, avg - 1, 2, avg * nblocks + frames_extra - avg - 2]:
        parts.append(sc.process(x[pos * length:(pos + nfr) * length]))
        pos += nfr
    split = np.concatenate([p for p in parts if len(p)], axis=0)
    assert np.array_equal(split, out)       # frame-ordered accumulation: any split gives the same bits
    # device-resident input and output
    sc.reset()
    d_in = engine.to_device(x)
    d_out = engine.dev_alloc(nblocks * length * 4)
    assert sc.process_device(d_in, n, d_out, nblocks) == nblocks
    engine.sync()
    assert np.array_equal(engine.to_host(d_out, (nblocks, length), np.float32), out)


def _every_frame_strong(x, length, w, avg, nblocks):
    """[nblocks][L] mask of the bins that stay within 40 dB of the frame maximum in EVERY frame of their block.  (A bin
    whose block sum is strong can still hold a deep null in one frame - 100 dB down happens on the synthetic streams -
    and that one log10 then carries the float32 rounding floor of the whole transform, in any float32 FFT.)"""
    fr = np.asarray(x[:length * avg * nblocks], np.complex128).reshape(nblocks, avg, length) * np.asarray(w, np.float64)
    p = np.abs(np.fft.fftshift(np.fft.fft(fr, axis=2), axes=2)) ** 2
    lp = np.log10(np.maximum(p, 1e-18))
    return (lp - lp.max(axis=2, keepdims=True)).min(axis=1) >= -4.0


def _check_logsum(out, ref, avg, mask=None):
    """float32 FFT rounding is relative to the STRONGEST bins of a frame (~5e-7 of their amplitude), so
    bins 100+ dB down (Blackman-Harris leaves >110 dB of dynamic range on synthetic data) carry large
    relative errors in any float32 implementation, GNU Radio's FFTW float path included.  Parity is
    therefore asserted where it is meaningful: bins within 40 dB of the block maximum agree to 1e-4 per
    frame in log10 (2.3e-4 relative in power), and the mean error over ALL bins stays tiny."""
    out = np.asarray(out, np.float64)
    err = np.abs(out - ref)
    for b in range(ref.shape[0]):
        strong = (ref[b] >= ref[b].max() - 4.0 * avg) if mask is None else mask[b]
        assert strong.sum() > 0
        assert err[b][strong].max() <= 1e-4 * avg, (b, err[b][strong].max())
    assert err.mean() <= 2e-4 * np.sqrt(avg), err.mean()


@pytest.mark.parametrize("length", [4096, 16384, 65536, 262144, 1048576])
def test_fft_linear_power_parity(engine, length):
    """One frame per vector: ||P_gpu - P_oracle||2 / ||P_oracle||2 <= 1e-5 in linear power
    (north_star tolerance), and the strongest bin index is identical."""
    x, _ = synth.scan_stream(length * 2, 2.4e6, length, seed=14, ncarriers=5)
    w = fd.blackmanharris(length)
    out = FftScanner(engine, length, w, 1).process(x).astype(np.float64)
    ref = gb.logpower_block_sums(x, length, w, 1)
    pg, pr = 10.0 ** (out - 1.0), 10.0 ** (ref - 1.0)
    for b in range(2):
        assert np.linalg.norm(pg[b] - pr[b]) / np.linalg.norm(pr[b]) <= 1e-5
        assert int(np.argmax(out[b])) == int(np.argmax(ref[b]))


def test_fft_tone_bin_after_shift(engine):
    """Tone at FFT bin k -> peak at (k + L/2) mod L after fftshift (SURVEY 8(c) golden (v))."""
    length = 16384
    k = 1234
    t = np.arange(length * 2)
    x = (0.5 * np.exp(2j * np.pi * k * t / length)).astype(np.complex64)
    sc = FftScanner(engine, length, fd.blackmanharris(length), 2)
    out = sc.process(x)
    assert int(np.argmax(out[0])) == (k + length // 2) % length
    x2 = (0.5 * np.exp(-2j * np.pi * 77 * t / length)).astype(np.complex64)
    out = sc.process(x2)
    assert int(np.argmax(out[0])) == (-77 + length // 2) % length


def test_fft_vector_flowgraph_and_peaks_bit_exact(engine):
    """fft_vector.py end to end at the reference's own size: 16384-pt, 1000 frames, last-100 sum ->
    the vector fft_vector.py writes; then fft_peak_detection.py:46-72 on GPU vs oracle spectrum."""
    length, fs, centre = 16384, 2.4e6, 855.05e6
    nframes = 1000
    x, truth = synth.scan_stream(length * nframes, fs, length, seed=44, ncarriers=8)
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, 100)
    out = sc.process(x)
    assert out.shape == (10, length)
    vec = out[9]  # frames 900..999 == item #999 of the moving sum
    ref = gb.fft_vector_flowgraph(x, length, w, nframes, 100)
    _check_logsum(vec[None, :], ref[None, :], 100)
    from radiocapture_rf_b200.fft_peak_detection import detect_peaks
    idx_gpu, f_gpu = detect_peaks(vec, fs, centre)
    idx_ref, f_ref = gb.peak_detect(ref.astype(np.float32), fs, centre)
    assert len(idx_ref) >= 4
    assert np.array_equal(idx_gpu, idx_ref)
    assert np.array_equal(f_gpu, f_ref)


def test_fft_streaming_split_invariance(engine):
    length, avg = 4096, 8
    x, _ = synth.scan_stream(length * avg * 2, 2.4e6, length, seed=8, ncarriers=3)
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, avg)
    a = sc.process(x)
    sc.reset()
    parts = []
    pos = 0
    for nfr in [1, 3, 5, 7]:
        parts.append(sc.process(x[pos * length:(pos + nfr) * length]))
        pos += nfr
    b = np.concatenate([p for p in parts if len(p)], axis=0)
    assert a.shape == b.shape == (2, length)
    np.testing.assert_allclose(a, b, atol=2e-5)


def test_fft_rejects_unsupported(engine, built_lib):
    from radiocapture_rf_b200 import _lib
    w = np.ones(1000, np.float32)
    st = built_lib.rcb_fft_config(engine.h, 1000, w.ctypes.data, 10)
    assert st == _lib.RCB_EUNSUPPORTED


@pytest.mark.parametrize("avg,nblocks,frames_extra", [(100, 2, 37), (7, 5, 3), (1, 9, 0), (12, 3, 11), (25, 2, 24), (13, 4, 5)])
def test_fft_frame_resident_kernel_16384(engine, avg, nblocks, frames_extra):
    """fft_frame_kernel (the reference's own scan length, fft_vector.py:32 `length = 1024*16`): whole frame in one SM's
    shared memory, sums per group of frames, one fold.  Parity bar of the tiled pipeline against the float64 oracle;
    agreement with the three-kernel pipeline (different association of the sum); the same BITS for any split of the
    stream into calls - blocks and groups cut anywhere, calls that end inside a group twice in a row; device-resident
    input at an address the bulk copies cannot use (8-byte aligned); device-resident output."""
    length = 16384
    nfr = avg * nblocks + frames_extra
    n = length * nfr
    x, _ = synth.scan_stream(n, 2.4e6, length, seed=21 + avg, ncarriers=5)
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, avg)
    out = sc.process(x)
    assert out.shape == (nblocks, length)
    ref = gb.logpower_block_sums(x[:length * avg * nblocks], length, w, avg)
    mask = _every_frame_strong(x, length, w, avg, nblocks)
    _check_logsum(out, ref, avg, mask)
    # the tiled pipeline on the same stream
    sc.set_pipeline(2)
    tiled = sc.process(x)
    # (two different FFT factorisations: bins far below the strongest carry each one's own float32 rounding noise, so they
    # are compared where _check_logsum compares - within 40 dB of the frame maximum in every frame - and on average)
    for b in range(nblocks):
        strong = mask[b]
        assert np.abs(out[b][strong] - tiled[b][strong]).max() <= 2e-4 * avg
    assert np.abs(out - tiled).mean() <= 4e-4 * np.sqrt(avg)
    sc.set_pipeline(0)
    # ragged calls
    cuts = [1, 2, 1, max(avg - 3, 1), 5, avg + 1, 3 * avg + 2, 1]
    parts, pos = [], 0
    for c in cuts:
        c = min(c, nfr - pos)
        if c <= 0:
            break
        parts.append(sc.process(x[pos * length:(pos + c) * length]))
        pos += c
    if pos < nfr:
        parts.append(sc.process(x[pos * length:]))
    split = np.concatenate([p for p in parts if len(p)], axis=0)
    assert np.array_equal(split, out)
    # device-resident, input one sample off 16-byte alignment, device output
    sc.reset()
    d_in = engine.dev_alloc((n + 1) * 8)
    from radiocapture_rf_b200._lib import COPY_H2D, check
    check(engine.lib.rcb_memcpy(engine.h, d_in.ptr + 8, x.ctypes.data, x.nbytes, COPY_H2D), "h2d", engine.h)
    d_out = engine.dev_alloc(nblocks * length * 4)
    assert sc.process_device(d_in.ptr + 8, n, d_out, nblocks) == nblocks
    engine.sync()
    assert np.array_equal(engine.to_host(d_out, (nblocks, length), np.float32), out)
    # aligned device input
    sc.reset()
    d_al = engine.to_device(x)
    assert sc.process_device(d_al, n, d_out, nblocks) == nblocks
    engine.sync()
    assert np.array_equal(engine.to_host(d_out, (nblocks, length), np.float32), out)


def test_fft_frame_resident_kernel_many_frames(engine):
    """A call of several thousand frames (more groups than SMs, several waves) and the linear-power / argmax check
    through avg = 1."""
    length, avg, nblocks = 16384, 100, 21
    n = length * avg * nblocks
    base, _ = synth.scan_stream(length * 300, 2.4e6, length, seed=5, ncarriers=7)
    x = np.tile(base, avg * nblocks // 300)
    assert len(x) == n
    w = fd.blackmanharris(length)
    sc = FftScanner(engine, length, w, avg)
    out = sc.process(x)
    assert out.shape == (nblocks, length)
    ref = gb.logpower_block_sums(x[:length * 300], length, w, avg)      # blocks repeat with period 3
    for b in range(nblocks):
        _check_logsum(out[b:b + 1], ref[b % 3:b % 3 + 1], avg)
    assert np.array_equal(out[0], out[3]) and np.array_equal(out[1], out[19])
