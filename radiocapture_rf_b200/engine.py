"""Python host over the C ABI (ctypes, numpy only).  One ``Engine`` = one rcb handle = one wideband
stream on one GPU (rc_frontend/receiver.py runs one frontend process per SDR source, :67-70).

Classes
  Engine          handle, device/pinned memory, timers
  PfbChannelizer  K1  (pfb.channelizer_ccf + per-bin quadrature_demod_cf; rc_frontend/receiver.py:249-261)
  DdcBank         K2  (freq_xlating_fir_filter_ccc per channel; rc_frontend/channel.py:35)
  FftScanner      K3  (fft_vector.py:37-60)
  Engine.quad_demod / Engine.probe_mean  K4
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MEM_DEVICE, MEM_HOST, OUT_FM, OUT_IQ, COPY_D2D, COPY_D2H, COPY_H2D, B200ChanError, check)


def device_count():
    n = C.c_int(0)
    st = _lib.load().rcb_device_count(C.byref(n))
    return n.value if st == 0 else 0


class DeviceBuffer(object):
    """Raw device allocation owned by an Engine."""

    def __init__(self, engine, nbytes):
        self.engine = engine
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(engine.lib.rcb_dev_alloc(engine.h, self.nbytes, C.byref(p)), "rcb_dev_alloc", engine.h)
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.engine.lib.rcb_dev_free(self.engine.h, self.ptr)
            self.ptr = None

    def offset(self, nbytes):
        return self.ptr + int(nbytes)


class Engine(object):
    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.rcb_open(int(device), C.byref(h)), "rcb_open(device=%d)" % device)
        self.h = h
        self.device = device
        self._pinned = []

    # ---- lifecycle -------------------------------------------------------------------------------
    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.rcb_host_free(self.h, p)
            self._pinned = []
            self.lib.rcb_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(self.lib.rcb_sync(self.h), "rcb_sync", self.h)

    def device_name(self):
        buf = C.create_string_buffer(256)
        sm = C.c_int(0)
        check(self.lib.rcb_device_name(self.h, buf, 256, C.byref(sm)), "rcb_device_name", self.h)
        return buf.value.decode(), sm.value

    def stats(self):
        s = _lib.rcb_stats_t()
        check(self.lib.rcb_stats(self.h, C.byref(s)), "rcb_stats", self.h)
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    # ---- memory ----------------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def pinned(self, shape, dtype):
        """numpy array backed by pinned host memory (freed with the engine)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(self.lib.rcb_host_alloc(self.h, max(n, 1), C.byref(p)), "rcb_host_alloc", self.h)
        self._pinned.append(p.value)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.dev_alloc(max(arr.nbytes, 1))
        check(self.lib.rcb_memcpy(self.h, d.ptr, arr.ctypes.data, arr.nbytes, COPY_H2D), "rcb_memcpy h2d", self.h)
        return d

    def to_host(self, dev_ptr, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        ptr = dev_ptr.ptr if isinstance(dev_ptr, DeviceBuffer) else dev_ptr
        check(self.lib.rcb_memcpy(self.h, out.ctypes.data, ptr, out.nbytes, COPY_D2H), "rcb_memcpy d2h", self.h)
        return out

    def copy_d2d(self, dst_ptr, src_ptr, nbytes):
        check(self.lib.rcb_memcpy(self.h, dst_ptr, src_ptr, nbytes, COPY_D2D), "rcb_memcpy d2d", self.h)

    def l2_flush(self):
        check(self.lib.rcb_l2_flush(self.h), "rcb_l2_flush", self.h)

    def timer_start(self):
        check(self.lib.rcb_timer_start(self.h), "rcb_timer_start", self.h)

    def timer_stop(self):
        ms = C.c_float(0)
        check(self.lib.rcb_timer_stop(self.h, C.byref(ms)), "rcb_timer_stop", self.h)
        return ms.value

    # ---- K5 --------------------------------------------------------------------------------------
    _FMT = {"u8": (_lib.FMT_U8, np.uint8, -127.4, 1.0 / 128.0),     # gr-osmosdr rtl_source_c
            "s8": (_lib.FMT_S8, np.int8, 0.0, 1.0 / 128.0),         # UHD sc8
            "s16": (_lib.FMT_S16, np.int16, 0.0, 1.0 / 32768.0)}    # UHD sc16

    def convert_iq(self, raw, fmt, offset=None, scale=None, out_device=None):
        """Interleaved integer I/Q (host array) -> complex64.  Returns a host array, or fills `out_device`."""
        code, dt, off, sc = self._FMT[fmt]
        raw = np.ascontiguousarray(raw, dtype=dt)
        n = raw.size // 2
        off = off if offset is None else float(offset)
        sc = sc if scale is None else float(scale)
        if out_device is not None:
            ptr = out_device.ptr if isinstance(out_device, DeviceBuffer) else out_device
            check(self.lib.rcb_convert_iq(self.h, raw.ctypes.data, code, off, sc, n, MEM_HOST, ptr, MEM_DEVICE),
                  "rcb_convert_iq", self.h)
            return n
        out = np.empty(n, dtype=np.complex64)
        check(self.lib.rcb_convert_iq(self.h, raw.ctypes.data, code, off, sc, n, MEM_HOST, out.ctypes.data, MEM_HOST),
              "rcb_convert_iq", self.h)
        return out

    # ---- K4 --------------------------------------------------------------------------------------
    def quad_demod(self, x, gain, prev=None):
        """analog.quadrature_demod_cf(gain) over rows of x (complex64 [rows][n] or [n]).
        Returns (fm float32 same shape, last samples per row)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        one = (x.ndim == 1)
        x2 = x.reshape(1, -1) if one else x
        rows, n = x2.shape
        out = np.empty((rows, n), dtype=np.float32)
        pv = np.zeros(rows, dtype=np.complex64) if prev is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(prev, np.complex64), (rows,)))
        pv = pv.copy()
        if n:
            check(self.lib.rcb_quad_demod(self.h, x2.ctypes.data, rows, n, n, float(gain), pv.ctypes.data,
                                          out.ctypes.data, n, MEM_HOST), "rcb_quad_demod", self.h)
        return (out[0] if one else out), (pv[0] if one else pv)

    def probe_mean(self, x, length, scale):
        """moving_average_ff(length,1,..) -> multiply_const(scale) -> probe value after the block."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        x2 = x.reshape(1, -1) if x.ndim == 1 else x
        rows, n = x2.shape
        out = np.empty(rows, dtype=np.float32)
        check(self.lib.rcb_probe_mean(self.h, x2.ctypes.data, rows, n, n, int(length), float(scale),
                                      out.ctypes.data, MEM_HOST), "rcb_probe_mean", self.h)
        return out[0] if x.ndim == 1 else out


def pfb_process_multi(channelizers, d_ins, nsamples, d_fms, out_stride):
    """One call (one kernel launch where the shape allows) over several independent streams of one shape on one GPU:
    channelizers[i] (its own Engine handle = its own streaming state), device input d_ins[i], device FM output d_fms[i]."""
    n = len(channelizers)
    assert n == len(d_ins) == len(d_fms) and n >= 1

    def _p(x):
        return x.ptr if isinstance(x, DeviceBuffer) else int(x)
    hs = (C.c_void_p * n)(*[c.e.h for c in channelizers])
    ins = (C.c_void_p * n)(*[_p(x) for x in d_ins])
    outs = (C.c_void_p * n)(*[_p(x) for x in d_fms])
    lead = channelizers[0].e
    check(lead.lib.rcb_pfb_process_multi(hs, n, ins, int(nsamples), outs, int(out_stride)), "rcb_pfb_process_multi", lead.h)


def copy_ceiling(engine, h2d_bytes, d2h_bytes, iters=4):
    """Bare pinned host <-> device copy rate of the engine's GPU (both directions concurrently, no kernels):
    returns (h2d GB/s, d2h GB/s, wall seconds)."""
    a, b, w = C.c_double(0), C.c_double(0), C.c_double(0)
    check(engine.lib.rcb_copy_ceiling(engine.h, int(h2d_bytes), int(d2h_bytes), int(iters), C.byref(a), C.byref(b),
                                      C.byref(w)), "rcb_copy_ceiling", engine.h)
    return a.value, b.value, w.value


class PfbChannelizer(object):
    """K1: N-channel critically sampled polyphase channelizer with fused FM demod."""

    def __init__(self, engine, nchans, taps, out_mask=OUT_IQ, fm_gain=1.0):
        self.e = engine
        self.nchans = int(nchans)
        self.out_mask = int(out_mask)
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.ntaps = len(taps)
        check(engine.lib.rcb_pfb_config(engine.h, self.nchans, taps.ctypes.data, len(taps), self.out_mask,
                                        float(fm_gain)), "rcb_pfb_config", engine.h)

    def reset(self):
        check(self.e.lib.rcb_pfb_reset(self.e.h), "rcb_pfb_reset", self.e.h)

    def set_out_block(self, frames):
        """Device outputs as channel-major blocks of `frames` (power of two >= 8; 0 = plain [N][stride])."""
        check(self.e.lib.rcb_pfb_set_out_block(self.e.h, int(frames)), "rcb_pfb_set_out_block", self.e.h)
        self.out_block = int(frames)

    _RAW_DTYPES = {_lib.FMT_U8: np.uint8, _lib.FMT_S8: np.int8, _lib.FMT_S16: np.int16}

    def set_input_format(self, fmt, offset=0.0, scale=1.0):
        """Fused ingest: process() / process_device() then take interleaved integer I/Q (FMT_U8 / FMT_S8 / FMT_S16),
        sample = (v + offset) * scale; fmt 0 = complex64 again.  Resets the streaming state."""
        check(self.e.lib.rcb_pfb_set_input_format(self.e.h, int(fmt), float(offset), float(scale)),
              "rcb_pfb_set_input_format", self.e.h)
        self.in_fmt = int(fmt)

    @staticmethod
    def unblock(arr, nchans, frames, block):
        """[nblocks][nchans][block] device layout (flat) -> [nchans][frames]."""
        nb = -(-frames // block)
        a = np.asarray(arr).reshape(nb, nchans, block)
        return np.ascontiguousarray(a.transpose(1, 0, 2).reshape(nchans, nb * block)[:, :frames])

    def process(self, iq, out_iq=None, out_fm=None):
        """iq: complex64 host array, len multiple of nchans.  Returns (iq_out, fm_out) (None where not produced):
        [N][T] arrays, or - after set_out_block(b) - [ceil(T / b)][N][b] arrays (each channel as contiguous b-sample
        messages; `unblock` gives the [N][T] view back)."""
        fmt = getattr(self, "in_fmt", 0)
        if fmt:
            iq = np.ascontiguousarray(iq, dtype=self._RAW_DTYPES[fmt]).reshape(-1)
            nsamp = len(iq) // 2
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            nsamp = len(iq)
        if nsamp % self.nchans:
            raise ValueError("nsamples must be a multiple of nchans")
        t = nsamp // self.nchans
        b = getattr(self, "out_block", 0)
        shape = (-(-t // b), self.nchans, b) if b else (self.nchans, t)
        if (self.out_mask & OUT_IQ) and out_iq is None:
            out_iq = np.empty(shape, dtype=np.complex64)
        if (self.out_mask & OUT_FM) and out_fm is None:
            out_fm = np.empty(shape, dtype=np.float32)
        nout = C.c_size_t(0)
        check(self.e.lib.rcb_pfb_process(
            self.e.h, iq.ctypes.data, nsamp, MEM_HOST,
            out_iq.ctypes.data if out_iq is not None else None,
            out_fm.ctypes.data if out_fm is not None else None,
            max(t, 1), MEM_HOST, C.byref(nout)), "rcb_pfb_process", self.e.h)
        return out_iq, out_fm

    def process_device(self, d_in, nsamples, d_iq, d_fm, out_stride):
        """All pointers device resident (ints / DeviceBuffer).  Asynchronous: returns after the launch."""
        def _p(x):
            return x.ptr if isinstance(x, DeviceBuffer) else x
        nout = C.c_size_t(0)
        check(self.e.lib.rcb_pfb_process(self.e.h, _p(d_in), int(nsamples), MEM_DEVICE, _p(d_iq), _p(d_fm),
                                         int(out_stride), MEM_DEVICE, C.byref(nout)), "rcb_pfb_process", self.e.h)
        return nout.value


class DdcBank(object):
    """K2: bank of freq_xlating_fir_filter_ccc channels over one wideband stream."""

    def __init__(self, engine):
        self.e = engine

    @staticmethod
    def gr_float_omega(center_freq, samp_rate):
        """GNU Radio 3.8 holds w = 2*pi*f0/fs in a C float (freq_xlating_fir_filter_impl.cc): return the
        frequency whose exact w equals that float32 value, for bit-level studies against GNU Radio."""
        w = float(np.float32(2.0 * np.pi * float(center_freq) / float(samp_rate)))
        return w * float(samp_rate) / (2.0 * np.pi)

    def open(self, decim, taps, center_freq, samp_rate, out_mask=OUT_IQ, fm_gain=1.0, gr_float_omega=False):
        if gr_float_omega:
            center_freq = self.gr_float_omega(center_freq, samp_rate)
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        cid = C.c_int(0)
        check(self.e.lib.rcb_ddc_open(self.e.h, int(decim), taps.ctypes.data, len(taps), float(center_freq),
                                      float(samp_rate), int(out_mask), float(fm_gain), C.byref(cid)),
              "rcb_ddc_open", self.e.h)
        return cid.value

    def retune(self, chan, center_freq):
        check(self.e.lib.rcb_ddc_retune(self.e.h, int(chan), float(center_freq)), "rcb_ddc_retune", self.e.h)

    def set_taps(self, chan, taps):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        check(self.e.lib.rcb_ddc_set_taps(self.e.h, int(chan), taps.ctypes.data, len(taps)), "rcb_ddc_set_taps",
              self.e.h)

    def close(self, chan):
        check(self.e.lib.rcb_ddc_close(self.e.h, int(chan)), "rcb_ddc_close", self.e.h)

    def set_tensor_cores(self, mode=1, seg=0):
        """Buckets of >= 12 channels sharing (decim, ntaps) run on tcgen05 (3 x tf32 split).  mode 0 / False: CUDA-core
        kernel for every bucket; 1 / True (default): A operand in tensor memory, ``seg`` = k-chunks per accumulation
        segment; 2: A operand in shared memory, ``seg`` = 1..3 accumulators.  seg 0 keeps the current setting."""
        check(self.e.lib.rcb_ddc_set_tensor_cores(self.e.h, int(mode), int(seg)), "rcb_ddc_set_tensor_cores", self.e.h)

    def tensor_core_launches(self):
        n = C.c_uint64(0)
        check(self.e.lib.rcb_ddc_tensor_core_launches(self.e.h, C.byref(n)), "rcb_ddc_tensor_core_launches", self.e.h)
        return int(n.value)

    _RAW_DTYPES = {_lib.FMT_U8: np.uint8, _lib.FMT_S8: np.int8, _lib.FMT_S16: np.int16}

    def set_input_format(self, fmt, offset=0.0, scale=1.0):
        """process() / process_device() then take interleaved integer I/Q (FMT_U8 / FMT_S8 / FMT_S16), sample =
        (v + offset) * scale; the block crosses PCIe in the wire format.  fmt 0 = complex64 again.  Channel state is
        kept."""
        check(self.e.lib.rcb_ddc_set_input_format(self.e.h, int(fmt), float(offset), float(scale)),
              "rcb_ddc_set_input_format", self.e.h)
        self.in_fmt = int(fmt)

    def process(self, iq):
        fmt = getattr(self, "in_fmt", 0)
        if fmt:
            iq = np.ascontiguousarray(iq, dtype=self._RAW_DTYPES[fmt]).reshape(-1)
            n = len(iq) // 2
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            n = len(iq)
        check(self.e.lib.rcb_ddc_process(self.e.h, iq.ctypes.data, n, MEM_HOST), "rcb_ddc_process", self.e.h)

    def process_device(self, d_in, nsamples):
        p = d_in.ptr if isinstance(d_in, DeviceBuffer) else d_in
        check(self.e.lib.rcb_ddc_process(self.e.h, p, int(nsamples), MEM_DEVICE), "rcb_ddc_process", self.e.h)

    def nout(self, chan, which=OUT_IQ):
        """Items the last process() call produced for this channel."""
        n = C.c_size_t(0)
        st = self.e.lib.rcb_ddc_pull(self.e.h, int(chan), int(which), None, 0, MEM_HOST, C.byref(n))
        if st not in (0, _lib.RCB_ERANGE):
            check(st, "rcb_ddc_pull", self.e.h)
        return n.value

    def pull_device(self, chan, d_dst, cap_items, which=OUT_IQ):
        """Copy the channel's last outputs into device memory (chaining stages on the GPU).  Returns the item count."""
        n = C.c_size_t(0)
        p = d_dst.ptr if isinstance(d_dst, DeviceBuffer) else d_dst
        check(self.e.lib.rcb_ddc_pull(self.e.h, int(chan), int(which), p, int(cap_items), MEM_DEVICE, C.byref(n)),
              "rcb_ddc_pull", self.e.h)
        return n.value

    def pull_all(self, which=OUT_IQ, max_items=None, copy=True):
        """Every open channel's outputs of the last process() call in ONE device-to-host transfer into a pinned
        staging block kept by this object.  Returns {chan_id: array}; with copy=False the arrays are views of the
        staging block, valid until the next pull_all of the same kind (what a sink that serialises at once needs)."""
        n = C.c_size_t(0)
        st = self.e.lib.rcb_ddc_pull_all(self.e.h, int(which), None, 0, MEM_HOST, None, None, 0, C.byref(n))
        if st not in (0, _lib.RCB_ERANGE):
            check(st, "rcb_ddc_pull_all", self.e.h)
        m = n.value
        if m == 0:
            return {}
        if max_items is None:
            max_items = getattr(self, "_last_max", 4096)
        dt = np.complex64 if which == OUT_IQ else np.float32
        ids = (C.c_int * m)()
        counts = (C.c_size_t * m)()
        pins = self.__dict__.setdefault("_pins", {})
        while True:
            out = pins.get(which)
            if out is None or out.shape[0] < m or out.shape[1] < max_items:
                out = self.e.pinned((m, max_items), dt)
                pins[which] = out
            st = self.e.lib.rcb_ddc_pull_all(self.e.h, int(which), out.ctypes.data, out.shape[1], MEM_HOST, ids, counts, m,
                                             C.byref(n))
            if st == _lib.RCB_ERANGE and max(counts) > out.shape[1]:   # rows longer than guessed: counts are valid, retry
                max_items = int(max(counts))
                continue
            check(st, "rcb_ddc_pull_all", self.e.h)
            break
        self._last_max = max(int(max(counts)), 1)
        if copy:
            return {int(ids[r]): out[r, :counts[r]].copy() for r in range(m)}
        return {int(ids[r]): out[r, :counts[r]] for r in range(m)}

    def pull(self, chan, which=OUT_IQ):
        n = C.c_size_t(0)
        # first ask for the size (dst NULL is only legal when there is nothing to fetch)
        st = self.e.lib.rcb_ddc_pull(self.e.h, int(chan), int(which), None, 0, MEM_HOST, C.byref(n))
        if st not in (0, _lib.RCB_ERANGE):
            check(st, "rcb_ddc_pull", self.e.h)
        dt = np.complex64 if which == OUT_IQ else np.float32
        out = np.empty(n.value, dtype=dt)
        if n.value:
            check(self.e.lib.rcb_ddc_pull(self.e.h, int(chan), int(which), out.ctypes.data, n.value, MEM_HOST,
                                          C.byref(n)), "rcb_ddc_pull", self.e.h)
        return out


class PostDemod(object):
    """K6: the data-parallel post-demod stages of the backend demods over `rows` channels (rcb_post_*).

    PostDemod.p25_c4fm(...)  = p25_control_demod.py:106-133 / logging_receiver.py:229-245 up to the symbol filter
    PostDemod.analog_fm(...) = logging_receiver.py:210-222 (squelch, fm_demod_cf, 300 Hz high-pass, resampler to 8 kHz)
    Taps default to the reference's own designs (firdes.py) and can be passed explicitly."""

    def __init__(self, engine, cfg, keep):
        self.e = engine
        self.rows = cfg.rows
        self.kind = cfg.kind
        self._keep = keep     # tap arrays referenced by cfg during rcb_post_open
        cid = C.c_int(0)
        check(self.e.lib.rcb_post_open(self.e.h, C.byref(cfg), C.byref(cid)), "rcb_post_open", self.e.h)
        self.id = cid.value
        self._keep = None
        self.interp, self.decim = max(cfg.interp, 1), max(cfg.decim, 1)

    @classmethod
    def p25_c4fm(cls, engine, rows, channel_rate=25000.0, symbol_rate=4800, symbol_deviation=600.0, prefilter_taps=None,
                 symbol_taps=None, probe_len=10000, probe_scale=1e-4):
        from . import firdes
        if prefilter_taps is None:   # p25_control_demod.py:107
            prefilter_taps = firdes.low_pass_2(1.0, channel_rate, channel_rate / 4.0, 500.0, 30.0, firdes.WIN_BLACKMAN)
        if symbol_taps is None:      # p25_control_demod.py:130-133
            sps = int(channel_rate // symbol_rate)
            symbol_taps = np.full(sps, 1.0 / sps, np.float32)
        t0 = np.ascontiguousarray(prefilter_taps, np.float32)
        t1 = np.ascontiguousarray(symbol_taps, np.float32)
        cfg = _lib.rcb_post_cfg()
        cfg.kind, cfg.rows = _lib.POST_P25_C4FM, int(rows)
        cfg.gain = float(channel_rate / (2.0 * np.pi * symbol_deviation))   # p25_control_demod.py:120
        cfg.taps0, cfg.ntaps0 = t0.ctypes.data, len(t0)
        cfg.taps1, cfg.ntaps1 = t1.ctypes.data, len(t1)
        cfg.probe_len, cfg.probe_scale = int(probe_len), float(probe_scale)
        return cls(engine, cfg, (t0, t1))

    @classmethod
    def analog_fm(cls, engine, rows, input_rate=25000.0, audio_rate=8000, deviation=15000.0, gain=8.0, tau=75e-6,
                  squelch_db=-100.0, squelch_alpha=0.01, squelch_gate=True, audio_taps=None, hp_taps=None, resampler=None):
        from . import firdes
        if audio_taps is None:   # fm_demod_cf: optfir.low_pass(gain, rate, audio_pass, audio_stop, 0.1, 60)
            audio_taps = firdes.optfir_low_pass(gain, input_rate, input_rate * 0.25, input_rate * 0.25 + 2000, 0.1, 60)
        if hp_taps is None:      # logging_receiver.py:215
            hp_taps = firdes.high_pass(1, input_rate, 300, 30, firdes.WIN_HAMMING, 6.76)
        if resampler is None:    # logging_receiver.py:216-221
            resampler = firdes.rational_resampler_design(int(audio_rate), int(input_rate))
        interp, decim, rt = resampler
        t0 = np.ascontiguousarray(audio_taps, np.float32)
        t1 = np.ascontiguousarray(hp_taps, np.float32)
        t2 = np.ascontiguousarray(rt, np.float32)
        b0, b1, a1 = firdes.fm_deemph(input_rate, tau)
        cfg = _lib.rcb_post_cfg()
        cfg.kind, cfg.rows = _lib.POST_ANALOG_FM, int(rows)
        cfg.gain = float(input_rate / (2.0 * np.pi * deviation))
        cfg.taps0, cfg.ntaps0 = t0.ctypes.data, len(t0)
        cfg.taps1, cfg.ntaps1 = t1.ctypes.data, len(t1)
        cfg.taps2, cfg.ntaps2 = t2.ctypes.data, len(t2)
        cfg.interp, cfg.decim = int(interp), int(decim)
        cfg.squelch_db, cfg.squelch_alpha, cfg.squelch_gate = float(squelch_db), float(squelch_alpha), int(bool(squelch_gate))
        cfg.deemph_b0, cfg.deemph_b1, cfg.deemph_a1 = float(b0), float(b1), float(a1)
        return cls(engine, cfg, (t0, t1, t2))

    def _cap(self, n):
        return int(n) * self.interp // self.decim + 2

    def process(self, iq_rows, want_probe=False):
        """iq_rows: complex64 [rows][n] on the host.  Returns a list of float32 arrays (one per row) [, probe]."""
        x = np.ascontiguousarray(iq_rows, dtype=np.complex64)
        assert x.ndim == 2 and x.shape[0] == self.rows
        n = x.shape[1]
        cap = self._cap(n)
        out = np.empty((self.rows, cap), np.float32)
        nout = (C.c_int * self.rows)()
        probe = (C.c_float * self.rows)()
        check(self.e.lib.rcb_post_process(self.e.h, self.id, x.ctypes.data, n, n, None, MEM_HOST, out.ctypes.data, cap,
                                          MEM_HOST, nout, probe), "rcb_post_process", self.e.h)
        res = [out[r, :nout[r]].copy() for r in range(self.rows)]
        return (res, np.array(probe[:], np.float32)) if want_probe else res

    def process_device(self, d_iq, n, in_stride, d_out, out_stride, row_map=None):
        """Rows resident on the device (e.g. rows of a PFB IQ output).  Returns the per-row output counts."""
        def _p(x):
            return x.ptr if isinstance(x, DeviceBuffer) else x
        rm = None
        if row_map is not None:
            rm = (C.c_int * self.rows)(*[int(v) for v in row_map])
        nout = (C.c_int * self.rows)()
        check(self.e.lib.rcb_post_process(self.e.h, self.id, _p(d_iq), int(n), int(in_stride), rm, MEM_DEVICE, _p(d_out),
                                          int(out_stride), MEM_DEVICE, nout, None), "rcb_post_process", self.e.h)
        return list(nout)

    def close(self):
        if self.id:
            check(self.e.lib.rcb_post_close(self.e.h, self.id), "rcb_post_close", self.e.h)
            self.id = 0


class FftScanner(object):
    """K3: windowed streaming FFT + log-power block sums (fft_vector.py flowgraph)."""

    def __init__(self, engine, length, window, avg_frames=100):
        self.e = engine
        self.length = int(length)
        self.avg = int(avg_frames)
        window = np.ascontiguousarray(window, dtype=np.float32)
        if len(window) != self.length:
            raise ValueError("window length != fft length")
        check(engine.lib.rcb_fft_config(engine.h, self.length, window.ctypes.data, self.avg), "rcb_fft_config",
              engine.h)

    def reset(self):
        check(self.e.lib.rcb_fft_reset(self.e.h), "rcb_fft_reset", self.e.h)

    def set_pipeline(self, mode):
        """0 / False (default): frame-resident kernel for 16384-point frames, three kernels per L2-resident sub-batch for
        the other lengths; 1 / True: one persistent fused launch per call; 2: the three-kernel pipeline for every length.
        Restarts the current averaging block."""
        check(self.e.lib.rcb_fft_set_pipeline(self.e.h, int(mode)), "rcb_fft_set_pipeline", self.e.h)

    def process(self, iq):
        """Returns float32 [nvec][L]: one vector per completed block of avg frames."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if len(iq) % self.length:
            raise ValueError("nsamples must be a multiple of the FFT length")
        cap = len(iq) // self.length // self.avg + 1
        out = np.empty((cap, self.length), dtype=np.float32)
        n = C.c_size_t(0)
        check(self.e.lib.rcb_fft_process(self.e.h, iq.ctypes.data, len(iq), MEM_HOST, out.ctypes.data, cap, MEM_HOST,
                                         C.byref(n)), "rcb_fft_process", self.e.h)
        return out[:n.value]

    def process_device(self, d_in, nsamples, d_out, cap_vectors):
        def _p(x):
            return x.ptr if isinstance(x, DeviceBuffer) else x
        n = C.c_size_t(0)
        check(self.e.lib.rcb_fft_process(self.e.h, _p(d_in), int(nsamples), MEM_DEVICE, _p(d_out), int(cap_vectors),
                                         MEM_DEVICE, C.byref(n)), "rcb_fft_process", self.e.h)
        return n.value
