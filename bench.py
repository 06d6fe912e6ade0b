#!/usr/bin/env python
"""bench.py - wideband IQ Msps channelised + FM-demodulated on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg2|cfg5|cfg4]

A "step" is one pass of the hot path over one batch of synthetic wideband IQ.  Default workload =
BASELINE config 3 (the one the 70 %-of-HBM-roofline target is quoted on): ONE 1024-channel polyphase
channelizer with a 256-tap prototype (GNU Radio API reading: pfb.channelizer_ccf(1024, taps[256])) and
the fused quadrature FM demod, 2^28 complex64 samples per step resident in HBM (2 GiB in, 1 GiB FM out
=> inputs far larger than the 126 MB L2, no flush needed).  Reported on one JSON line:
  value      whole-job Msps-in with inputs resident in HBM (device timed, CUDA events, max over ranks)
  e2e        same metric through the C ABI with pinned HOST buffers (H2D + kernel + D2H pipelined inside)
  roofline   algorithmic bytes (12 B/sample: 8 read + 4 FM written) / kernel time vs measured HBM peak
  cpu_baseline  the GR-semantics C restatement (oracle/gr_cpu.c, kind "port") on a bounded sample
  also       all three readings of "256-tap / 1024-channel" (256-tap prototype = 1 tap/arm [the headline,
             an FFT-only upper bound: 768 of the 1024 arms are zero], 16 taps/arm [the reference's own filter shape],
             256 taps/arm [compute bound, also quoted against the fp32 roofline]), the plain [N][T] output layout,
             the same launch sustained for >= 2.5 s with clocks / power, fused integer ingest (sc16 / u8), and
             BASELINE config 5 (8 x 256-channel streams per GPU)
N > 1: one process per GPU (torchrun), independent streams per rank, no data-path collective ("weak").
`--impl reference` times the CPU restatement only (GNU Radio itself is not installable here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "wideband_iq_msps_channelized_demodulated"

WORKLOADS = {
    # name: nchans, ntaps, out (fm/iq), log2 samples per step per stream, streams per GPU, fs label
    "cfg3": dict(nchans=1024, ntaps=256, out="fm", log2n=28, streams=1,
                 desc="cfg3: 1024-channel PFB, 256-tap prototype, fused FM demod, one 200 Msps stream"),
    "cfg3_p16": dict(nchans=1024, ntaps=16384, out="fm", log2n=28, streams=1,
                     desc="cfg3 variant: 1024-channel PFB, 16 taps/arm (16384-tap prototype), fused FM demod"),
    "cfg3_p256": dict(nchans=1024, ntaps=262144, out="fm", log2n=26, streams=1,
                      desc="cfg3 variant: 1024-channel PFB, 256 taps/arm (262144-tap prototype), fused FM demod"),
    "cfg3_p8": dict(nchans=1024, ntaps=8192, out="fm", log2n=28, streams=1,
                    desc="cfg3 variant: 1024-channel PFB, 8 taps/arm (8192-tap prototype), fused FM demod"),
    "cfg3_iqfm_p16": dict(nchans=1024, ntaps=16384, out="iq+fm", log2n=27, streams=1,
                          desc="1024-channel PFB, 16 taps/arm, IQ + fused FM out (what pfb-mode channel requests consume)"),
    "cfg2_p16_iqfm": dict(nchans=64, ntaps=1024, out="iq+fm", log2n=27, streams=1,
                          desc="64-channel PFB, 16 taps/arm, IQ + fused FM out"),
    "cfg2": dict(nchans=64, ntaps=128, out="iq", log2n=27, streams=1,
                 desc="cfg2: 64-channel PFB, 128-tap prototype, IQ out, one 25 Msps stream"),
    "cfg1": dict(kind="ddc", fs=2.4e6, rate=12500, nchan=1, log2n=24, streams=1, nchans=1, ntaps=349, out="iq+fm",
                 desc="cfg1: 1-channel freq_xlating_fir (D 96, 349 taps, -62.5 kHz) + FM quad demod on 2.4 Msps IQ"),
    "ddc64": dict(kind="ddc", fs=16.0e6, rate=12500, nchan=64, log2n=24, streams=1, nchans=64, ntaps=2327, out="iq+fm",
                  desc="64 simultaneous channel.py channels (D 640, 2327 taps) + FM on one 16 Msps source"),
    "cfg4": dict(kind="fft", length=1 << 20, avg=64, log2n=27, streams=1, nchans=0, ntaps=0, out="logpow",
                 desc="cfg4: 2^20-point streaming FFT (Blackman-Harris) + |.|^2 + log10 + 64-frame sums, 1 Gsps scan"),
    "cfg4_16k": dict(kind="fft", length=1 << 14, avg=100, log2n=27, streams=1, nchans=0, ntaps=0, out="logpow",
                     desc="fft_vector.py shape: 16384-point FFT + log-power, 100-frame sums"),
    "cfg5": dict(nchans=256, ntaps=4096, out="fm", log2n=25, streams=8,
                 desc="cfg5: 8 independent 100 Msps streams per GPU x 256 channels, 16 taps/arm, fused FM demod"),
}


def load_tensor_peak():
    """Dense tf32 ceiling for the DDC contraction: half the measured bf16 rate (MEASURED_PEAKS.json), else the
    profiling recipe's nominal 1.1 PFLOP/s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]) / 2.0, "measured bf16_tflops / 2 (MEASURED_PEAKS.json; tf32 runs at half the bf16 rate)"
        except Exception:
            pass
    return 1100.0, "fallback (B200_PROFILING.md 1.1 PFLOP/s dense tf32)"


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_name(wl):
    """The kernel that dominates a step of this workload (the dispatch of rcb_pfb_process / rcb_ddc_process /
    rcb_fft_process, csrc/b200chan.cu)."""
    cfg = WORKLOADS[wl]
    if cfg.get("kind") == "fft":
        return "fft_frame_kernel" if cfg["length"] == (1 << 14) else "fft_cols_tma_kernel+fft_rows_kernel"
    if cfg.get("kind") == "ddc":
        if cfg["nchan"] >= 12 and DDC_TENSOR_CORES:
            return "ddc_mma2_kernel" if DDC_TENSOR_CORES == 1 else "ddc_mma_kernel"
        return "ddc_lone_kernel" if cfg["nchan"] == 1 else "ddc_tile_kernel"
    n, tpa = cfg["nchans"], -(-cfg["ntaps"] // cfg["nchans"])
    pt = 1
    while pt < tpa:
        pt *= 2
    if cfg["out"] == "fm" and n == 1024:
        if pt == 1:
            return "pfb_fm1_kernel"
        if pt <= 8:
            return "pfb_cl_kernel<32,%d>" % pt
        if pt == 16:
            return "pfb_fm_ws_kernel<32,16>"
        return "pfb_arm_fir_kernel+pfb_fm1_kernel"
    if n == 1024 and pt >= 8:
        return "pfb_fm_ws_kernel<32,%d,IQ>" % pt
    r = {64: 8, 256: 16, 1024: 32}[n]
    return "pfb_fm_tma_kernel<%d,PT=%d,%s>" % (r, pt, cfg["out"]) if pt <= 16 else "pfb_fm_kernel<%d>" % r


def config_dict(wl, world, out_block, log2n=None):
    """The `config` object of the JSON line - identical for the b200 and the reference arm."""
    cfg = WORKLOADS[wl]
    n = 1 << (log2n or cfg["log2n"])
    blocked = bool(out_block) and cfg.get("kind") is None
    return {"workload": cfg["desc"], "nchans": cfg["nchans"], "ntaps": cfg["ntaps"], "out": cfg["out"],
            "samples_per_step_per_gpu": n * cfg["streams"], "streams_per_gpu": cfg["streams"],
            "channels_out": cfg["nchans"] * cfg["streams"] * world,
            "out_layout": ("channel-major in blocks of %d frames (rcb_pfb_set_out_block), device and host outputs" % out_block)
            if blocked else "channel-major [N][T]",
            "l2": "inputs (%d MiB/step) larger than L2, no flush" % (n * cfg["streams"] * 8 >> 20),
            "parallelism": "independent streams, %d GPU(s), no collective" % world}


def load_traffic(workload):
    """dram bytes per input sample from the committed ncu --set full capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(p))[workload]["dram_bytes_per_sample"])
    except Exception:
        return None


class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            self.proc.wait(timeout=2)
        except Exception:
            pass

    def power(self, t0, t1):
        pw = []
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 4 and t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    pw.append(float(f[3]))
                except ValueError:
                    pass
        return max(pw) if pw else None

    def summary(self, t0, t1):
        sm, smax, reasons = [], 0.0, set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = max(smax, mx)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take everything we saw
            for ts, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_taps(nchans, ntaps):
    """Blackman-Harris windowed-sinc prototype of exactly `ntaps` taps, cutoff half a bin."""
    n = np.arange(ntaps, dtype=np.float64) - (ntaps - 1) / 2.0
    fc = 0.5 / nchans
    h = 2 * fc * np.sinc(2 * fc * n)
    k = np.arange(ntaps, dtype=np.float64)
    m = max(ntaps - 1, 1)
    w = 0.35875 - 0.48829 * np.cos(2 * np.pi * k / m) + 0.14128 * np.cos(4 * np.pi * k / m) - 0.01168 * np.cos(6 * np.pi * k / m)
    h = h * w
    return (h / h.sum()).astype(np.float32)


RAW_FORMATS = {   # name: (RCB_FMT_*, numpy dtype, offset, scale)   sample = (v + offset) * scale
    "u8": (1, np.uint8, -127.4, 1.0 / 128.0),      # RTL-SDR (gr-osmosdr rtl_source_c)
    "s8": (2, np.int8, 0.0, 1.0 / 128.0),          # UHD otw_format sc8 (configs/config_denver_usrp.py:20)
    "sc16": (3, np.int16, 0.0, 1.0 / 32768.0),     # UHD otw_format sc16
}


def quantise(x, name):
    """complex64 block -> interleaved integers of the named wire format (full scale = 1.0)."""
    _, dt, off, scale = RAW_FORMATS[name]
    v = np.round(x.view(np.float32) / np.float32(scale) - np.float32(off))
    info = np.iinfo(dt)
    return np.clip(v, info.min, info.max).astype(dt)


def synth_block(n, nchans, seed):
    """Seeded synthetic IQ: tones in 1/4 of the bins + AWGN, RMS 0.25 (arithmetic cost is data independent)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(2 * n, dtype=np.float32).view(np.complex64) * np.float32(0.02)
    t = np.arange(n, dtype=np.float64)
    for b in rng.choice(nchans, size=min(16, nchans // 4), replace=False):
        f = (b + rng.uniform(-0.1, 0.1)) / nchans
        x += (0.05 * np.exp(2j * np.pi * f * t)).astype(np.complex64)
    x *= np.float32(0.25 / np.sqrt(np.mean(np.abs(x) ** 2)))
    return x


def dist_setup(ngpus):
    """One process per GPU under torchrun; torch.distributed (NCCL) only for barrier + reductions."""
    from radiocapture_rf_b200 import sharding
    world, rank, local = sharding.world_from_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return world, rank, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def allreduce_max(dist, local, v):
    from radiocapture_rf_b200 import sharding
    return sharding.Reducer(dist, "cuda:%d" % local if dist is not None else None).max(v)


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: GR-semantics C restatement on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_run(wl, steps, warmup, log2n_cpu=None, budget_s=None, streams=1):
    from oracle import gr_cpu
    cfg = WORKLOADS[wl]
    if cfg.get("kind") == "fft":
        from oracle import gr_firdes
        L = cfg["length"]
        n = max(L * 8, 1 << (log2n_cpu or 23))
        x = synth_block(n, 1024, 3)
        w = gr_firdes.blackmanharris(L)
        threads = os.cpu_count() or gr_cpu.num_threads()
        gr_cpu.fft_logpow(x[:L * 2], L, w, nthreads=threads)
        t0 = time.perf_counter()
        done = 0
        for s in range(steps):
            gr_cpu.fft_logpow(x, L, w, nthreads=threads)
            done += 1
            if budget_s and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return done * n / dt / 1e6, threads, done, n, dt
    if cfg.get("kind") == "ddc":
        from oracle import gr_firdes
        fs, rate = cfg["fs"], cfg["rate"]
        decim, taps = gr_firdes.channel_taps(fs, rate)
        n = 1 << (log2n_cpu or 22)
        x = synth_block(n, 64, 3)
        threads = os.cpu_count() or gr_cpu.num_threads()
        rng = np.random.default_rng(3)
        offs = [-62500.0] if cfg["nchan"] == 1 else list(rng.uniform(-0.45 * fs, 0.45 * fs, cfg["nchan"]))
        gr_cpu.xlating_fir(x[:1 << 16], taps, decim, offs[0], fs, nthreads=threads)
        t0 = time.perf_counter()
        done = 0
        for s in range(steps):
            for f in offs:  # GNU Radio: one flowgraph (and one full-rate copy) per channel
                y = gr_cpu.xlating_fir(x, taps, decim, f, fs, nthreads=threads)
                gr_cpu.quad_demod(y, 5.0)
            done += 1
            if budget_s and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return done * n / dt / 1e6, threads, done, n, dt
    nch, ntaps = cfg["nchans"], cfg["ntaps"]
    taps = make_taps(nch, ntaps)
    log2n = log2n_cpu or 22
    n = 1 << log2n
    base = synth_block(min(n, 1 << 22), nch, 3)
    x = np.tile(base, n // len(base)) if n > len(base) else base
    want_fm = "fm" in cfg["out"]
    threads = os.cpu_count() or gr_cpu.num_threads()   # torchrun exports OMP_NUM_THREADS=1: ask explicitly
    hist = [None] * streams
    want_iq = "iq" in cfg["out"]
    o_iq = np.zeros((nch, n // nch), np.complex64) if want_iq else None   # touched once: no page faults in the timed loop
    o_fm = np.zeros((nch, n // nch), np.float32) if want_fm else None
    for _ in range(max(warmup, 1)):
        _, _, hist[0] = gr_cpu.pfb_fm(x, nch, taps, 5.0, want_iq=want_iq, want_fm=want_fm, hist=hist[0], nthreads=threads,
                                      out_iq=o_iq, out_fm=o_fm)
        if budget_s and n >= (1 << 26):
            break   # one full-size warm-up pass is plenty (page faults, thread start-up)
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        for k in range(streams):   # independent streams: each with its own history, same host buffers
            _, _, hist[k] = gr_cpu.pfb_fm(x, nch, taps, 5.0, want_iq=want_iq, want_fm=want_fm, hist=hist[k],
                                          nthreads=threads, out_iq=o_iq, out_fm=o_fm)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    msps = done * n * streams / dt / 1e6
    return msps, threads, done, n, dt


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = WORKLOADS[args.workload]
    # the same step as the b200 arm: every stream's full block (2^28 samples for cfg3 = 2 GiB of IQ, ~0.6 s on 16
    # cores), outputs preallocated and touched before the timed loop; bounded to ~2.5 minutes of wall clock
    log2n = args.log2n or cfg["log2n"]
    msps, threads, done, n, dt = cpu_run(args.workload, args.steps, max(args.warmup, 1), log2n_cpu=log2n, budget_s=150.0,
                                         streams=cfg["streams"])
    sample = "%d steps x %d stream(s) x 2^%d samples of the %s workload (oracle/gr_cpu.c, OpenMP over frames)" % (
        done, cfg["streams"], int(np.log2(n)), args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": msps, "unit": "Msps", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": dt / max(done, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.workload, max(args.gpus, 1), args.out_block, args.log2n),
        "reference_note": "GNU Radio 3.8 is not installable here (DESIGN.md section 3): GR-semantics C restatement of its "
                          "blocks on all host threads; one process regardless of --gpus; writes the plain [N][T] layout",
        "cpu_baseline": {"value": msps, "unit": "Msps", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": msps, "unit": "Msps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
class StreamCtx(object):
    """One wideband stream resident on the GPU: engine + channelizer + device buffers."""

    def __init__(self, device, wl, seed, log2n=None, ntaps=None, out_block=0, in_fmt=None):
        from radiocapture_rf_b200.engine import Engine, PfbChannelizer, OUT_FM, OUT_IQ
        cfg = WORKLOADS[wl]
        self.cfg = cfg
        self.nch = cfg["nchans"]
        self.ntaps = ntaps or cfg["ntaps"]
        self.n = 1 << (log2n or cfg["log2n"])
        self.frames = self.n // self.nch
        self.fm = "fm" in cfg["out"]
        self.iq = "iq" in cfg["out"]
        self.e = Engine(device)
        self.ch = PfbChannelizer(self.e, self.nch, make_taps(self.nch, self.ntaps),
                                 (OUT_FM if self.fm else 0) | (OUT_IQ if self.iq else 0), 5.0)
        base_n = min(self.n, 1 << 24)
        base = synth_block(base_n, self.nch, seed)
        isz = 8
        if in_fmt:   # the same stream in the SDR's wire format, read directly by the kernel
            fmt, _, off, scale = RAW_FORMATS[in_fmt]
            self.ch.set_input_format(fmt, off, scale)
            base = quantise(base, in_fmt)
            isz = base.itemsize * 2
        self.d_in = self.e.dev_alloc(self.n * isz)
        check_copy = self.e.lib.rcb_memcpy
        from radiocapture_rf_b200._lib import COPY_H2D, check
        check(check_copy(self.e.h, self.d_in.ptr, base.ctypes.data, base.nbytes, COPY_H2D), "h2d", self.e.h)
        filled = base_n
        while filled < self.n:  # replicate on device
            c = min(filled, self.n - filled)
            self.e.copy_d2d(self.d_in.ptr + filled * isz, self.d_in.ptr, c * isz)
            filled += c
        self.base = base
        self.out_block = out_block
        if out_block:
            self.ch.set_out_block(out_block)
        nb = -(-self.frames // out_block) if out_block else 0
        out_elems = nb * self.nch * out_block if out_block else self.n
        self.d_fm = self.e.dev_alloc(out_elems * 4) if self.fm else None
        self.d_iq = self.e.dev_alloc(out_elems * 8) if self.iq else None
        self.bytes_per_sample = isz + (4 if self.fm else 0) + (8 if self.iq else 0)

    def step(self):
        self.ch.process_device(self.d_in, self.n, self.d_iq, self.d_fm, self.frames)

    def close(self):
        self.e.close()


class FftCtx(object):
    """Scan path: K3 over a device-resident stream."""

    def __init__(self, device, wl, seed, log2n=None, **kw):
        from radiocapture_rf_b200.engine import Engine, FftScanner
        from radiocapture_rf_b200 import firdes
        cfg = WORKLOADS[wl]
        self.cfg = cfg
        self.n = 1 << (log2n or cfg["log2n"])
        self.L = cfg["length"]
        self.e = Engine(device)
        self.sc = FftScanner(self.e, self.L, firdes.blackmanharris(self.L), cfg["avg"])
        base = synth_block(min(self.n, 1 << 24), 1024, seed)
        self.d_in = self.e.dev_alloc(self.n * 8)
        from radiocapture_rf_b200._lib import COPY_H2D, check
        check(self.e.lib.rcb_memcpy(self.e.h, self.d_in.ptr, base.ctypes.data, base.nbytes, COPY_H2D), "h2d", self.e.h)
        filled = len(base)
        while filled < self.n:
            c = min(filled, self.n - filled)
            self.e.copy_d2d(self.d_in.ptr + filled * 8, self.d_in.ptr, c * 8)
            filled += c
        self.cap = self.n // self.L // cfg["avg"] + 2
        self.d_out = self.e.dev_alloc(self.cap * self.L * 4)
        self.bytes_per_sample = 8

    def step(self):
        self.sc.process_device(self.d_in, self.n, self.d_out, self.cap)

    def close(self):
        self.e.close()


class DdcCtx(object):
    """xlat mode: M rc_frontend/channel.py channels over one device-resident wideband block (K2)."""

    def __init__(self, device, wl, seed, log2n=None, **kw):
        from radiocapture_rf_b200.engine import Engine, DdcBank, OUT_FM, OUT_IQ
        from radiocapture_rf_b200 import firdes
        cfg = WORKLOADS[wl]
        self.cfg = cfg
        self.n = 1 << (log2n or cfg["log2n"])
        self.e = Engine(device)
        self.bank = DdcBank(self.e)
        if DDC_TENSOR_CORES != 1:
            self.bank.set_tensor_cores(DDC_TENSOR_CORES)
        fs, rate = cfg["fs"], cfg["rate"]
        decim = firdes.channel_decimation(fs, rate)
        taps = firdes.low_pass_2(1.0, fs, rate / 2, rate / 2, 20.0, firdes.WIN_HAMMING)
        rng = np.random.default_rng(seed)
        offs = [-62500.0] if cfg["nchan"] == 1 else list(rng.uniform(-0.45 * fs, 0.45 * fs, cfg["nchan"]))
        self.decim = decim
        self.ids = [self.bank.open(decim, taps, f, fs, OUT_IQ | OUT_FM, 5.0) for f in offs]
        x = synth_block(self.n, 64, seed)
        self.d_in = self.e.to_device(x)
        self.bytes_per_sample = 8.0 + cfg["nchan"] * 12.0 / decim

    def step(self):
        self.bank.process_device(self.d_in, self.n)

    def close(self):
        self.e.close()


DDC_TENSOR_CORES = 1  # rcb_ddc_set_tensor_cores mode: 0 (--no-tensor-cores) CUDA-core ddc_tile_kernel for every bucket,
                      # 1 ddc_mma2_kernel (default), 2 (--ddc-mode 2) ddc_mma_kernel
USE_MULTI = False  # --multi: one rcb_pfb_process_multi call per step instead of one rcb_pfb_process call per stream
                   # (measured on cfg5: 202 vs 218 Gsps - the per-stream launches on their own CUDA streams overlap better)


def step_all(ctxs):
    """One step of every stream of this rank: multi-stream PFB workloads (BASELINE config 5) go through ONE
    rcb_pfb_process_multi call (one persistent kernel launch that walks the streams)."""
    if len(ctxs) > 1 and USE_MULTI and all(isinstance(c, StreamCtx) for c in ctxs):
        from radiocapture_rf_b200.engine import pfb_process_multi
        pfb_process_multi([c.ch for c in ctxs], [c.d_in for c in ctxs], ctxs[0].n, [c.d_fm for c in ctxs], ctxs[0].frames)
    else:
        for c in ctxs:
            c.step()


def timed_loop(ctxs, steps, warmup, dist, local):
    for _ in range(warmup):
        step_all(ctxs)
    for c in ctxs:
        c.e.sync()
    barrier(dist, local)
    l0 = sum(c.e.stats()["kernel_launches"] for c in ctxs)
    for c in ctxs:
        c.e.timer_start()
    t0 = time.time()
    for _ in range(steps):
        step_all(ctxs)
    ms = max(c.e.timer_stop() for c in ctxs)   # events on each stream's own CUDA stream; streams run concurrently
    t1 = time.time()
    for c in ctxs:
        c.e.sync()
    barrier(dist, local)
    launches = sum(c.e.stats()["kernel_launches"] for c in ctxs) - l0
    return ms, launches, t0, t1


def side_run(device, wl, world, dist, local, peak, steps, out_block, log2n=None, in_fmt=None, bytes_per_sample=None):
    """One extra device-resident measurement of a PFB workload for the `also` object."""
    from radiocapture_rf_b200 import sharding
    cfg = WORKLOADS[wl]
    w_, r_, _ = sharding.world_from_env()
    ctxs = [StreamCtx(device, wl, seed=3 + sid, log2n=log2n, out_block=out_block, in_fmt=in_fmt)
            for sid in sharding.assign_streams(cfg["streams"] * world, world, r_ if world > 1 else 0)]
    ms, _, _, _ = timed_loop(ctxs, steps, 3, dist, local)
    tot = sum(c.n for c in ctxs) * steps
    rate, ms = sharding.whole_job_rate(tot, ms, sharding.Reducer(dist, "cuda:%d" % local if dist is not None else None))
    bps = bytes_per_sample or ctxs[0].bytes_per_sample
    for c in ctxs:
        c.close()
    return {"workload": cfg["desc"], "kernel": kernel_name(wl), "unit": "Msps", "value": rate / 1e6,
            "algorithmic_bytes_per_sample": bps, "roofline_frac": tot * bps / (ms * 1e-3) / 1e9 / peak}


def fft_side_run(device, wl, steps, peak):
    """Device-resident rate of a scan workload (K3) on this rank."""
    cfg = WORKLOADS[wl]
    ctx = FftCtx(device, wl, seed=3)
    for _ in range(3):
        ctx.step()
    ctx.e.sync()
    ctx.e.timer_start()
    for _ in range(steps):
        ctx.step()
    ms = ctx.e.timer_stop()
    v = ctx.n * steps / (ms * 1e-3) / 1e6
    out = {"workload": cfg["desc"], "kernel": kernel_name(wl), "unit": "Msps", "value": v, "per_rank": True,
           "algorithmic_bytes_per_sample": 8, "roofline_frac": v * 1e6 * 8 / 1e9 / peak}
    ctx.close()
    return out


def ddc_lone_side_run(device, wl, steps, e2e_steps, world, dist, local, peak):
    """BASELINE config 1 (one rc_frontend/channel.py channel on a 2.4 Msps RTL-SDR source): device rate and the
    end-to-end rate with complex64 and with the dongle's own u8 samples crossing PCIe."""
    cfg = WORKLOADS[wl]
    ctx = DdcCtx(device, wl, seed=3)
    for _ in range(3):
        ctx.step()
    ctx.e.sync()
    ctx.e.timer_start()
    for _ in range(steps):
        ctx.step()
    ms = ctx.e.timer_stop()
    v = ctx.n * steps / (ms * 1e-3) / 1e6
    out = {"workload": cfg["desc"], "kernel": "ddc_lone_kernel", "unit": "Msps", "value": v, "per_rank": True,
           "algorithmic_bytes_per_sample": ctx.bytes_per_sample,
           "roofline_frac": v * 1e6 * ctx.bytes_per_sample / 1e9 / peak}
    ctx.close()
    for key, fmt in (("e2e", None), ("e2e_u8", "u8")):
        ev, h2d, d2h, _ = run_e2e_ddc(device, wl, max(e2e_steps, 2), dist, local, in_fmt=fmt)
        ev = ev * world if dist is None else allreduce_sum_min(dist, local, ev, world)
        out[key] = {"value": ev, "unit": "Msps", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "rcb_ddc_process(host pinned in%s) + rcb_ddc_pull_all" % (", u8 wire format" if fmt else "")}
    return out


def ddc_side_run(device, wl, steps):
    """Device-resident rate of a DDC-bank workload on this rank, tensor-core path and CUDA-core path."""
    global DDC_TENSOR_CORES
    cfg = WORKLOADS[wl]
    out = {"workload": cfg["desc"], "unit": "Msps"}
    keep = DDC_TENSOR_CORES
    for name, mode in (("value", 1), ("cuda_core_value", 0)):
        DDC_TENSOR_CORES = mode
        ctx = DdcCtx(device, wl, seed=3)
        for _ in range(3):
            ctx.step()
        ctx.e.sync()
        ctx.e.timer_start()
        for _ in range(steps):
            ctx.step()
        ms = ctx.e.timer_stop()
        out[name] = ctx.n * steps / (ms * 1e-3) / 1e6
        if mode == 1:
            out["kernel"] = kernel_name(wl)
            tpeak, _ = load_tensor_peak()
            alg = out[name] * 1e6 * 8.0 * cfg["nchan"] * cfg["ntaps"] / ctx.decim / 1e12
            out["algorithmic_tflops"] = alg
            out["tensor_roofline_frac"] = 3.0 * alg / tpeak
        ctx.close()
    DDC_TENSOR_CORES = keep
    out["per_rank"] = True
    return out


def sustained_run(device, wl, dist, local, peak, out_block, log2n, seconds=2.5):
    """The headline launch looped for >= `seconds` with the clock / power sampler running (burst vs sustained)."""
    ctx = StreamCtx(device, wl, seed=3, log2n=log2n, out_block=out_block)
    for _ in range(3):
        ctx.step()
    ctx.e.sync()
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.25)
    barrier(dist, local)
    ctx.e.timer_start()
    t0 = time.time()
    steps = 0
    while True:
        for _ in range(50):
            ctx.step()
        steps += 50
        ctx.e.sync()
        if time.time() - t0 >= seconds:
            break
    ms = ctx.e.timer_stop()
    t1 = time.time()
    sampler.stop()
    clk = sampler.summary(t0, t1)
    pw = sampler.power(t0, t1)
    n = ctx.n
    bps = ctx.bytes_per_sample
    ctx.close()
    return {"seconds": ms * 1e-3, "steps": steps, "unit": "Msps", "value": n * steps / (ms * 1e-3) / 1e6,
            "roofline_frac": n * steps * bps / (ms * 1e-3) / 1e9 / peak, "clocks": clk, "power_w_max": pw}


def run_e2e(device, wl, steps, warmup, dist, local, log2n=26, out_block=0, in_fmt=None):
    """Same metric through the public host-buffer call: pinned host in -> H2D -> kernel -> D2H -> pinned host out."""
    from radiocapture_rf_b200.engine import Engine, PfbChannelizer, OUT_FM, OUT_IQ
    cfg = WORKLOADS[wl]
    nch = cfg["nchans"]
    fm, iq = "fm" in cfg["out"], "iq" in cfg["out"]
    n = 1 << log2n
    frames = n // nch
    e = Engine(device)
    ch = PfbChannelizer(e, nch, make_taps(nch, cfg["ntaps"]), (OUT_FM if fm else 0) | (OUT_IQ if iq else 0), 5.0)
    base = synth_block(min(n, 1 << 22), nch, 5)
    if in_fmt:
        fmt, dt, off, scale = RAW_FORMATS[in_fmt]
        ch.set_input_format(fmt, off, scale)
        hin = e.pinned((2 * n,), dt)
        rb = quantise(base, in_fmt)
        for i in range(0, 2 * n, len(rb)):
            hin[i:i + len(rb)] = rb[:min(len(rb), 2 * n - i)]
    else:
        hin = e.pinned((n,), np.complex64)
        for i in range(0, n, len(base)):
            hin[i:i + len(base)] = base[:min(len(base), n - i)]
    if out_block:
        ch.set_out_block(out_block)
    shape = (-(-frames // out_block), nch, out_block) if out_block else (nch, frames)
    hfm = e.pinned(shape, np.float32) if fm else None
    hiq = e.pinned(shape, np.complex64) if iq else None
    for _ in range(warmup):
        ch.process(hin, out_iq=hiq, out_fm=hfm)
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(steps):
        ch.process(hin, out_iq=hiq, out_fm=hfm)   # returns after the D2H completed
    dt_s = time.perf_counter() - t0
    barrier(dist, local)
    hout = hfm if fm else hiq
    chk = float(np.abs(hout.reshape(-1)[-4096:]).sum())   # the result is really on the host
    d2h = (hfm.nbytes if fm else 0) + (hiq.nbytes if iq else 0)
    h2d = hin.nbytes
    e.close()
    return n * max(steps, 1) / max(dt_s, 1e-9) / 1e6, h2d, d2h, chk


def run_e2e_ddc(device, wl, steps, dist, local, log2n=24, in_fmt=None):
    """Host block in -> all channels' IQ+FM pulled back to the host (what SourceStream.push does).  in_fmt: the block
    crosses PCIe in the SDR's wire format (rcb_ddc_set_input_format)."""
    from radiocapture_rf_b200.engine import OUT_FM, OUT_IQ
    ctx = DdcCtx(device, wl, seed=5, log2n=log2n)
    n = ctx.n
    if in_fmt:
        fmt, dt, off, scale = RAW_FORMATS[in_fmt]
        ctx.bank.set_input_format(fmt, off, scale)
        hin = ctx.e.pinned((2 * n,), dt)
        hin[:] = quantise(synth_block(n, 64, 5), in_fmt)
    else:
        hin = ctx.e.pinned((n,), np.complex64)
        hin[:] = synth_block(n, 64, 5)
    for _ in range(3):   # warm-up incl. the pulls (staging buffers, row-length guess, PCIe link out of its idle state)
        ctx.bank.process(hin)
        ctx.bank.pull_all(OUT_IQ)
        ctx.bank.pull_all(OUT_FM)
    barrier(dist, local)
    t0 = time.perf_counter()
    d2h = 0
    chk = 0.0
    for _ in range(steps):
        ctx.bank.process(hin)
        ys = ctx.bank.pull_all(OUT_IQ, copy=False)     # one transfer per output kind for all channels, pinned staging
        fs = ctx.bank.pull_all(OUT_FM, copy=False)
        d2h = sum(v.nbytes for v in ys.values()) + sum(v.nbytes for v in fs.values())
        chk += float(sum(v[-1] for v in fs.values() if len(v)))
    dt = time.perf_counter() - t0
    barrier(dist, local)
    ctx.close()
    return n * steps / dt / 1e6, hin.nbytes, d2h, chk


def run_e2e_fft(device, wl, steps, dist, local, log2n=26):
    from radiocapture_rf_b200.engine import Engine, FftScanner
    from radiocapture_rf_b200 import firdes
    cfg = WORKLOADS[wl]
    n = 1 << log2n
    e = Engine(device)
    sc = FftScanner(e, cfg["length"], firdes.blackmanharris(cfg["length"]), min(cfg["avg"], n // cfg["length"]))
    hin = e.pinned((n,), np.complex64)
    base = synth_block(min(n, 1 << 22), 1024, 5)
    for i in range(0, n, len(base)):
        hin[i:i + len(base)] = base[:min(len(base), n - i)]
    for _ in range(3):   # staging buffers, lazy kernel loading, PCIe link out of its idle state (measured: the first two calls
        out = sc.process(hin)   # of a process take 50 - 85 ms instead of 10)
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = sc.process(hin)
    dt = time.perf_counter() - t0
    barrier(dist, local)
    chk = float(out[-1][:8].sum()) if len(out) else 0.0
    e.close()
    return n * steps / dt / 1e6, n * 8, int(out.nbytes), chk


def copy_ceiling_line(device, h2d, d2h, dist, local):
    """Bare pinned H2D || D2H of one e2e step's bytes on every rank at once (no kernels): the platform ceiling."""
    from radiocapture_rf_b200.engine import Engine, copy_ceiling
    e = Engine(device)
    barrier(dist, local)
    a, b, wall = copy_ceiling(e, h2d, d2h, iters=4)
    barrier(dist, local)
    e.close()
    return a, b, wall


def run_b200(args):
    world, rank, local, dist = dist_setup(args.gpus)
    device = local if world > 1 else 0
    wl = args.workload
    cfg = WORKLOADS[wl]
    peak, peak_src = load_peak()

    is_fft = cfg.get("kind") == "fft"
    is_ddc = cfg.get("kind") == "ddc"
    is_pfb = not (is_fft or is_ddc)
    out_block = args.out_block if is_pfb else 0
    Ctx = FftCtx if is_fft else (DdcCtx if is_ddc else StreamCtx)
    # stream s of the job runs on rank s mod world (sharding.assign_streams): streams_per_gpu x world streams in total
    from radiocapture_rf_b200 import sharding
    my_streams = sharding.assign_streams(cfg["streams"] * world, world, rank)
    ctxs = [Ctx(device, wl, seed=3 + sid, log2n=args.log2n, out_block=out_block) for sid in my_streams]
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.3)
    ms, launches, t0, t1 = timed_loop(ctxs, args.steps, args.warmup, dist, local)
    sampler.stop()
    clocks = sampler.summary(t0, t1)
    clocks["power_w_max"] = sampler.power(t0, t1)
    ms = allreduce_max(dist, local, ms)
    samples_rank = sum(c.n for c in ctxs) * args.steps
    total = samples_rank * world
    msps = total / (ms * 1e-3) / 1e6
    bps = ctxs[0].bytes_per_sample
    n_launch_samples = ctxs[0].n
    kern_ms = ms / args.steps if cfg["streams"] == 1 else None
    achieved = (samples_rank * bps) / (ms * 1e-3) / 1e9
    tr = load_traffic(wl)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (tr * n_launch_samples) if tr else None, "peak_source": peak_src,
                "kernel": kernel_name(wl),
                "algorithmic_bytes_per_sample": bps,
                "algorithmic_bytes_per_launch": n_launch_samples * bps,
                "kernel_ms_per_launch": kern_ms}
    if is_ddc and kernel_name(wl).startswith("ddc_mma"):
        # many-channel bank: a contraction, compute bound.  8 real flops per complex tap MAC; every MAC is issued as
        # three tf32 MMAs (hi*hi, hi*lo, lo*hi), so the tensor pipe executes 3x the algorithmic flops
        tpeak, tsrc = load_tensor_peak()
        decim = ctxs[0].decim
        flop_per_sample = 8.0 * cfg["nchan"] * cfg["ntaps"] / decim
        alg = samples_rank * flop_per_sample / (ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": 3.0 * alg, "peak": tpeak, "unit": "TFLOP/s", "frac": 3.0 * alg / tpeak,
                    "traffic": None, "peak_source": tsrc, "kernel": kernel_name(wl),
                    "algorithmic_tflops": alg, "algorithmic_flop_per_sample": flop_per_sample,
                    "note": "achieved = tf32 flops the tensor pipe executes (3 MMAs per algorithmic MAC for fp32 parity); "
                            "algorithmic_tflops is the fp32 work of the filter bank",
                    "hbm_frac": achieved / peak, "algorithmic_bytes_per_sample": bps, "kernel_ms_per_launch": kern_ms}
    spg = sum(c.n for c in ctxs)
    for c in ctxs:
        c.close()

    also = None
    if wl == "cfg3" and not args.no_also:
        also = {}
        half = max(3, args.steps // 2)
        # ---- the three readings of "256-tap / 1024-channel"  ----
        readings = {"prototype_256_taps_1_per_arm": {
            "workload": cfg["desc"], "kernel": kernel_name(wl), "unit": "Msps", "value": msps, "roofline_frac": achieved / peak,
            "note": "literal GNU Radio API reading (the headline): 768 of the 1024 arms have no tap - an FFT-only upper "
                    "bound, not a working channel filter"}}
        readings["16_taps_per_arm"] = side_run(device, "cfg3_p16", world, dist, local, peak, half, out_block, args.log2n)
        readings["16_taps_per_arm"]["note"] = ("the reference's own filter shape (optfir.low_pass(1, N, 0.5, 0.7, 0.1, 80) is "
                                               "17-19 taps per arm): 32 extra fp32 lane-operations per sample put the FP32 pipe, "
                                               "not HBM, on the critical path (DESIGN.md section 5)")
        r256 = side_run(device, "cfg3_p256", world, dist, local, peak, 3, out_block, min(args.log2n or 26, 26))
        flops = 4.0 * 256 + 5.0 * 10 + 30.0     # FIR 4 P + FFT 5 log2 N + demod
        r256["fp32_tflops"] = r256["value"] * 1e6 * flops / 1e12 / world
        r256["fp32_roofline_frac"] = r256["fp32_tflops"] / 74.4   # 148 SMs x 128 FMA lanes x 2 x 1.965 GHz
        r256["note"] = "256 taps PER ARM: compute bound (about 1100 flop per sample), quoted against the fp32 roofline too"
        readings["256_taps_per_arm"] = r256
        also["tap_readings"] = readings
        also["taps_per_arm_16"] = readings["16_taps_per_arm"]
        also["taps_per_arm_8"] = side_run(device, "cfg3_p8", world, dist, local, peak, half, out_block, args.log2n)
        # ---- the plain [N][T] layout (rows 1 MB apart) ----
        if out_block:
            also["plain_layout"] = side_run(device, wl, world, dist, local, peak, half, 0, args.log2n)
            also["plain_layout"]["workload"] += ", plain [N][T] output"
        # ---- burst vs sustained ----
        also["sustained"] = sustained_run(device, wl, dist, local, peak, out_block, args.log2n)
        # ---- fused integer ingest: fewer HBM read bytes per sample (and fewer PCIe bytes on the host path) ----
        for name, rd in (("sc16", 4), ("u8", 2)):
            r = side_run(device, wl, world, dist, local, peak, half, out_block, args.log2n, in_fmt=name,
                         bytes_per_sample=rd + 4)
            r["workload"] += ", %s wire format read directly (rcb_pfb_set_input_format)" % name
            ev, h2d_i, d2h_i, _ = run_e2e(device, wl, max(args.e2e_steps, 1), 3, dist, local, out_block=out_block,
                                          in_fmt=name)
            ev = ev * world if dist is None else allreduce_sum_min(dist, local, ev, world)
            r["e2e"] = {"value": ev, "unit": "Msps", "h2d_bytes_per_step": h2d_i, "d2h_bytes_per_step": d2h_i}
            also["ingest_" + name] = r
        # ---- BASELINE config 5: 8 independent 256-channel streams per GPU (64 over 8 GPUs) ----
        global USE_MULTI
        keep = USE_MULTI
        USE_MULTI = False
        also["cfg5"] = side_run(device, "cfg5", world, dist, local, peak, half, out_block)
        also["cfg5"]["streams_total"] = WORKLOADS["cfg5"]["streams"] * world
        also["cfg5"]["launch"] = "one rcb_pfb_process call per stream, each on its handle's own CUDA stream"
        USE_MULTI = True
        m = side_run(device, "cfg5", world, dist, local, peak, half, out_block)
        also["cfg5"]["single_launch"] = {"value": m["value"], "unit": "Msps", "roofline_frac": m["roofline_frac"],
                                         "launch": "one rcb_pfb_process_multi call per step (one persistent launch walks the 8 streams)"}
        USE_MULTI = keep
        # ---- K2: 64 xlat channels (rc_frontend/channel.py) on one 16 Msps source: tensor cores vs CUDA cores ----
        also["ddc64"] = ddc_side_run(device, "ddc64", half)
        # ---- K3: the reference's own scan (fft_vector.py:32, 16384 points x 100 frames) and BASELINE config 4 (2^20) ----
        also["fft_scan_16k"] = fft_side_run(device, "cfg4_16k", half, peak)
        also["cfg4"] = fft_side_run(device, "cfg4", half, peak)
        # ---- K2: BASELINE config 1, one channel on a 2.4 Msps source ----
        also["cfg1"] = ddc_lone_side_run(device, "cfg1", half, args.e2e_steps, world, dist, local, peak)

    api = "rcb_pfb_process(host pinned in, host pinned out)"
    if is_fft:
        e2e_v, h2d, d2h, chk = run_e2e_fft(device, wl, args.e2e_steps, dist, local)
        api = "rcb_fft_process(host pinned in, host out)"
    elif is_ddc:
        e2e_v, h2d, d2h, chk = run_e2e_ddc(device, wl, args.e2e_steps, dist, local)
        api = "rcb_ddc_process(host pinned in) + rcb_ddc_pull_all(host out) for IQ and FM"
    else:
        e2e_v, h2d, d2h, chk = run_e2e(device, wl, args.e2e_steps, 3, dist, local, out_block=out_block)
    e2e_v = e2e_v * world if dist is None else allreduce_sum_min(dist, local, e2e_v, world)
    e2e = {"value": e2e_v, "unit": "Msps", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": api,
           "checksum": chk}
    if not args.no_ceiling:
        # all ranks copy at once: the ceiling includes what they take from each other (host memory, PCIe switches)
        a, b, wall = copy_ceiling_line(device, h2d, d2h, dist, local)
        a_min = a if dist is None else allreduce_sum_min(dist, local, a, 1)
        b_min = b if dist is None else allreduce_sum_min(dist, local, b, 1)
        step_s = max(h2d / max(a_min, 1e-9), d2h / max(b_min, 1e-9)) / 1e9   # both directions overlap
        n_e2e = h2d // 8 if is_pfb or is_fft or is_ddc else 0
        ceil_msps = (n_e2e / step_s / 1e6) * world if step_s > 0 else None
        e2e["pcie_ceiling"] = {"h2d_gbs_per_gpu_min": a_min, "d2h_gbs_per_gpu_min": b_min, "msps": ceil_msps,
                               "frac": (e2e_v / ceil_msps) if ceil_msps else None,
                               "how": "rcb_copy_ceiling: pinned H2D || D2H of one step's bytes on all ranks at once, no kernels"}

    cpu = None
    if rank == 0 and not args.no_cpu:
        cm, threads, done, n, dt = cpu_run(wl, 100000, 1, log2n_cpu=(None if (is_fft or is_ddc) else 24), budget_s=12.0)
        cpu = {"value": cm, "unit": "Msps", "cores": threads, "kind": "port",
               "sample": "%d x 2^%d samples of the same workload (oracle/gr_cpu.c, %.1f s)" % (done, int(np.log2(n)), dt)}
    barrier(dist, local)

    if rank == 0:
        line = {
            "metric": METRIC, "value": msps, "unit": "Msps", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, world, out_block, args.log2n),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "also": also,
        }
        assert line["config"]["samples_per_step_per_gpu"] == spg
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def allreduce_sum_min(dist, local, v, world):
    """whole-job e2e = world * min over ranks (ranks run concurrently, each bounded by its own PCIe link)."""
    from radiocapture_rf_b200 import sharding
    return sharding.Reducer(dist, "cuda:%d" % local).min(v) * world


def _claim_stdout():
    """Library chatter (NCCL version banners, torchrun notices) must not pollute the one JSON line:
    fd 1 is pointed at stderr for the whole run and the line is written to the saved original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


_OUT = None


def emit(line):
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--log2n", type=int, default=None, help="override samples per step per stream (log2)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--out-block", type=int, default=1024,
                    help="device output layout: channel-major in blocks of this many frames (0 = plain [N][T])")
    ap.add_argument("--multi", action="store_true", help="multi-stream workloads: ONE rcb_pfb_process_multi launch per step instead of one call per stream")
    ap.add_argument("--no-tensor-cores", action="store_true", help="DDC workloads: CUDA-core kernel for every bucket")
    ap.add_argument("--ddc-mode", type=int, default=1, choices=[1, 2], help="DDC tensor-core kernel generation")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ceiling", action="store_true", help="skip the bare-copy ceiling measurement of the e2e path")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    global USE_MULTI, DDC_TENSOR_CORES
    USE_MULTI = bool(args.multi)
    DDC_TENSOR_CORES = 0 if args.no_tensor_cores else args.ddc_mode
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
