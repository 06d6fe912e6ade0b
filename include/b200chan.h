/* b200chan.h - C ABI of libb200chan.so : B200-native wideband channelizer / FM demod / FFT scan.
 *
 * This is the drop-in boundary for the ONE hot path of MattMills/radiocapture-rf.
 * The reference is pure Python wiring GNU Radio 3.8 C++ blocks; it has no FFI of its own, so each
 * entry point below names the GNU Radio block call (reference file:line) whose arithmetic it replaces.
 * The Python host (radiocapture_rf_b200/, ctypes - no PyTorch) mirrors rc_frontend/channel.py,
 * rc_frontend/receiver.py and fft_vector.py / fft_peak_detection.py on top of these calls.
 *
 * Conventions: every function returns RCB_OK (0) or a negative rcb_status; nothing throws; buffers
 * are caller owned; sample format is the reference's raw little-endian complex64 (interleaved
 * float32 I,Q - rc_frontend/channel.py:36, logging_receiver.py:107-109).  A handle owns one CUDA
 * device + streams and carries the streaming state of ONE wideband stream (one SDR source,
 * rc_frontend/receiver.py:67-70); it is not thread safe (the reference serialises on access_lock,
 * rc_frontend/receiver.py:48).  Handles on different GPUs are independent (no NCCL on this path).
 * `mem` arguments say where a caller buffer lives: RCB_MEM_HOST (copied with cudaMemcpyAsync, pinned
 * memory from rcb_host_alloc overlaps best) or RCB_MEM_DEVICE (used in place).
 */
#ifndef B200CHAN_H
#define B200CHAN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rcb_ctx rcb_t;

typedef enum {
    RCB_OK = 0,
    RCB_EINVAL = -1,       /* bad argument */
    RCB_ENOMEM = -2,       /* host or device allocation failed */
    RCB_ECUDA = -3,        /* CUDA runtime error (see rcb_last_error) */
    RCB_ESTATE = -4,       /* call sequence error (e.g. process before config) */
    RCB_ENODEV = -5,       /* no such CUDA device / no GPU */
    RCB_ERANGE = -6,       /* id / size out of range */
    RCB_EUNSUPPORTED = -7  /* shape not supported by any kernel */
} rcb_status;

enum { RCB_MEM_HOST = 0, RCB_MEM_DEVICE = 1 };
enum { RCB_OUT_IQ = 1, RCB_OUT_FM = 2 };
enum { RCB_FMT_C64 = 0, RCB_FMT_U8 = 1, RCB_FMT_S8 = 2, RCB_FMT_S16 = 3 };  /* input sample formats */
enum { RCB_COPY_H2D = 1, RCB_COPY_D2H = 2, RCB_COPY_D2D = 3 };

typedef struct {
    uint64_t samples_in;       /* wideband samples consumed by pfb/ddc/fft process calls */
    uint64_t channel_samples;  /* narrowband samples produced (all channels) */
    uint64_t kernel_launches;  /* CUDA kernels launched by this handle */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} rcb_stats_t;

/* ---- library / handle ------------------------------------------------------------------------ */
int rcb_version(void);
const char* rcb_strerror(int status);
int rcb_device_count(int* count);
int rcb_open(int device, rcb_t** out);          /* one handle per wideband stream */
int rcb_close(rcb_t* h);
int rcb_sync(rcb_t* h);                         /* wait for all work queued on the handle */
const char* rcb_last_error(rcb_t* h);           /* text of the last CUDA error on this handle */
int rcb_stats(rcb_t* h, rcb_stats_t* out);
int rcb_device_name(rcb_t* h, char* buf, size_t cap, int* sm_count);

/* ---- memory / timing helpers (the Python host has no torch) ------------------------------------ */
int rcb_dev_alloc(rcb_t* h, size_t bytes, void** p);
int rcb_dev_free(rcb_t* h, void* p);
int rcb_host_alloc(rcb_t* h, size_t bytes, void** p);   /* pinned */
int rcb_host_free(rcb_t* h, void* p);
int rcb_memcpy(rcb_t* h, void* dst, const void* src, size_t bytes, int kind);  /* synchronous */
int rcb_memset(rcb_t* h, void* dev, int value, size_t bytes);
int rcb_l2_flush(rcb_t* h);                     /* overwrite a > L2-sized scratch buffer */
int rcb_timer_start(rcb_t* h);                  /* cudaEventRecord on the compute stream */
int rcb_timer_stop(rcb_t* h, float* ms);        /* record + synchronize + elapsed */
/* Bare pinned-host <-> device copy rate (H2D of h2d_bytes and D2H of d2h_bytes running concurrently, `iters`
 * rounds, no kernels): the ceiling of every host-facing figure on this box.  GB/s per direction and wall seconds. */
int rcb_copy_ceiling(rcb_t* h, size_t h2d_bytes, size_t d2h_bytes, int iters, double* h2d_gbs, double* d2h_gbs,
                     double* wall_s);

/* ---- K1: polyphase channelizer + fused FM demod -------------------------------------------------
 * Replaces  pfb.channelizer_ccf(nchans, taps, 1.0, ...)           rc_frontend/receiver.py:249-261
 * and, per output bin,  analog.quadrature_demod_cf(fm_gain)       moto_control_demod.py:105,
 *   edacs_control_demod.py:82-84, p25_control_demod.py:120-121, logging_receiver.py:234,336,346.
 * taps = prototype filter (any length; arm i gets taps[i + k*nchans], zero padded).  Output port m
 * carries FFT bin m (identity channel map; bins > nchans/2 are negative frequencies,
 * rc_frontend/receiver.py:373-375).  Outputs are channel-major: element (m, n) at m*out_stride + n.
 * nsamples must be a multiple of nchans (stream_to_streams granularity); *nout = nsamples/nchans.
 * Streaming state (last ceil(ntaps/nchans) input rows) is carried between calls, so any split of a
 * stream into calls gives the same samples as one call.  out_iq / out_fm may be NULL per out_mask. */
int rcb_pfb_config(rcb_t* h, int nchans, const float* taps, int ntaps, int out_mask, float fm_gain);
int rcb_pfb_reset(rcb_t* h);
/* Optional device-output layout: channel-major inside time blocks of `frames` (power of two >= 8, 0 = plain
 * [N][out_stride]): element (m, n) at ((n / frames) * nchans + m) * frames + n % frames.  Each channel is then
 * delivered as contiguous `frames`-sample messages - the unit a zeromq.pub_sink sends (channel.py:36) - and the
 * rows one kernel iteration writes stay within a few MB (TLB / DRAM page locality).  The output buffer must hold
 * ceil(nout / frames) * nchans * frames elements; out_stride is ignored.  Host outputs use the same layout (the
 * internal H2D | kernel | D2H pipeline then moves whole blocks: one contiguous copy per 4 Mi-sample chunk instead of
 * nchans row pieces); a time block must then fit one pipeline chunk (frames * nchans <= 2^22) or be the chunk.
 * frames = 8 is the kernel's native granularity (one CTA iteration = one contiguous nchans*8-element piece,
 * full-line stores: +7 % on the 1024-channel FM path) for consumers that run on the GPU themselves. */
int rcb_pfb_set_out_block(rcb_t* h, int frames);
/* Fused ingest: after this call `iq` of rcb_pfb_process is the SDR's wire format - interleaved
 * integer I/Q, fmt = RCB_FMT_U8 (RTL-SDR: offset -127.4, scale 1/128), RCB_FMT_S8 / RCB_FMT_S16 (UHD otw_format sc8 /
 * sc16, configs/config_denver_usrp.py:20), sample = (v + offset) * scale exactly like rcb_convert_iq - and nsamples
 * still counts complex samples.  The 1024-channel one-tap-per-arm FM kernel converts inside its first FFT pass (2-4x
 * fewer HBM read and PCIe bytes); every other shape converts the block once on the device.  fmt = 0 restores
 * complex64.  Resets the streaming state. */
int rcb_pfb_set_input_format(rcb_t* h, int fmt, float offset, float scale);
int rcb_pfb_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem,
                    void* out_iq, void* out_fm, size_t out_stride, int out_mem, size_t* nout);
/* Several independent wideband streams of ONE shape on one GPU in one call (one stream per SDR source,
 * rc_frontend/receiver.py:67-70; systemd/radiocapture-channelizer@.service:11 runs one frontend per source): handle
 * hs[i] carries stream i's configuration and streaming state exactly as for rcb_pfb_process, iq[i] / out_fm[i] are its
 * device-resident input block (nsamples complex64) and FM output.  256-channel FM-only configurations with <= 16 taps
 * per arm run as ONE persistent kernel launch (the grid walks the streams) plus one history update; any other shape is processed stream
 * after stream.  All handles must live on the same device and share nchans, ntaps, out_mask and output layout.
 * Asynchronous; ordered after everything queued on every handle's stream, and every handle's later calls after it. */
int rcb_pfb_process_multi(rcb_t* const* hs, int nstreams, const void* const* iq, size_t nsamples,
                          void* const* out_fm, size_t out_stride);

/* ---- K2: bank of arbitrary-offset DDC channels (+ optional FM demod) ----------------------------
 * One channel = filter.freq_xlating_fir_filter_ccc(decim, taps, center_freq, samp_rate)
 *   rc_frontend/channel.py:35 (the `channel` top_block), split-2 pair rc_frontend/receiver.py:85-86,
 *   1x prefilter p25_control_demod.py:108 / logging_receiver.py:231;
 * rcb_ddc_retune = set_center_freq (channel.py:61-63), rcb_ddc_set_taps = set_taps (channel.py:56-57).
 * rcb_ddc_process runs every open channel over the same staged wideband block (one H2D copy instead
 * of one ZMQ copy of the full-rate stream per channel, channel.py:29); outputs of the block are then
 * fetched per channel with rcb_ddc_pull (which = RCB_OUT_IQ: complex64, RCB_OUT_FM: float32). */
int rcb_ddc_open(rcb_t* h, int decim, const float* taps, int ntaps, double center_freq,
                 double samp_rate, int out_mask, float fm_gain, int* chan_id);
int rcb_ddc_retune(rcb_t* h, int chan_id, double center_freq);
int rcb_ddc_set_taps(rcb_t* h, int chan_id, const float* taps, int ntaps);
int rcb_ddc_close(rcb_t* h, int chan_id);
int rcb_ddc_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem);
/* Wire-format input for the bank (the RTL-SDR sources of configs/ deliver u8, UHD sc8 / sc16 -
 * configs/config_denver_usrp.py:20): after this call `iq` of rcb_ddc_process is interleaved integer I/Q, fmt / offset /
 * scale as in rcb_pfb_set_input_format, nsamples still counts complex samples.  The block crosses PCIe in the wire
 * format (2-4x fewer bytes: the end-to-end rate of a DDC bank is PCIe bound) and becomes complex64 on the device, chunk
 * by chunk, in the staging buffer the DDC kernels read (same arithmetic as rcb_convert_iq: bit-identical to converting
 * on the host first).  Streaming state is kept; fmt = 0 restores complex64. */
int rcb_ddc_set_input_format(rcb_t* h, int fmt, float offset, float scale);
int rcb_ddc_pull(rcb_t* h, int chan_id, int which, void* dst, size_t cap_items, int dst_mem,
                 size_t* nitems);
/* Buckets of >= 12 channels that share (decim, ntaps) with an even decim run on the tensor cores (tcgen05 kind::tf32,
 * every operand split hi + lo -> 3 MMAs per k-step, fp32 accumulators in TMEM; same 1e-5 parity bar as the CUDA-core
 * kernel).  mode 0: CUDA-core ddc_tile_kernel for every bucket (what gr::filter::freq_xlating_fir_filter_ccc instances
 * do one by one, rc_frontend/channel.py:35).  mode 1 (default): ddc_mma2_kernel, A operand in tensor memory; seg =
 * k-chunks (32 floats of the window each) per accumulation segment before the accumulator is drained to registers
 * (0 = keep; default 16).  mode 2: ddc_mma_kernel, A operand in shared memory; seg = 1..3 TMEM accumulators the main
 * accumulation is cut into (0 = keep; default 3). */
int rcb_ddc_set_tensor_cores(rcb_t* h, int mode, int seg);
/* ddc_mma_kernel launches on this handle so far (which path a block took is otherwise invisible to the caller) */
int rcb_ddc_tensor_core_launches(rcb_t* h, uint64_t* launches);

/* Every open channel's outputs of the last rcb_ddc_process call in ONE transfer (the per-channel pulls of a source
 * with many channels cost more than the kernels): row r of dst (row_stride_items apart) receives counts[r] items of
 * channel ids[r] (ascending chan_id; ids may be NULL).  *nrows = channels; cap_rows < *nrows -> RCB_ERANGE. */
int rcb_ddc_pull_all(rcb_t* h, int which, void* dst, size_t row_stride_items, int dst_mem, int* ids, size_t* counts,
                     size_t cap_rows, size_t* nrows);

/* ---- K4: stand-alone quadrature demod / AFC probe on narrowband rows ---------------------------
 * out[r][n] = gain * atan2(Im p, Re p), p = x[r][n] conj(x[r][n-1]); prev (rows complex64, may be
 * NULL = zeros) supplies x[r][-1] and receives the last sample of each row (streaming carry).
 * Replaces analog.quadrature_demod_cf (call sites above).  All pointers in `mem` space. */
int rcb_quad_demod(rcb_t* h, const void* iq, size_t rows, size_t n, size_t in_stride, float gain,
                   void* prev, void* out_fm, size_t out_stride, int mem);
/* scale * sum of the last `length` floats of each row: what moving_average_ff(length,1,..) ->
 * multiply_const(scale) -> probe_signal_f holds (p25_control_demod.py:123-127). */
int rcb_probe_mean(rcb_t* h, const void* x, size_t rows, size_t n, size_t stride, size_t length,
                   float scale, void* out, int mem);

/* ---- K5: ingest conversion -----------------------------------------------------
 * Interleaved integer I/Q as the SDR delivers it -> complex64: out = (v + offset) * scale.
 * RCB_FMT_U8 with (offset -127.4, scale 1/128) is gr-osmosdr's rtl_source_c mapping (RTL-SDR sources of every
 * shipped rtlsdr config), RCB_FMT_S8 / RCB_FMT_S16 are UHD's sc8 / sc16 wire formats
 * (configs/config_denver_usrp.py:20 otw_format).  src holds 2*nsamples integers; dst nsamples complex64.
 * src_mem / dst_mem: RCB_MEM_HOST or RCB_MEM_DEVICE (host src = 2-4x fewer PCIe bytes than complex64). */
int rcb_convert_iq(rcb_t* h, const void* src, int fmt, float offset, float scale, size_t nsamples, int src_mem,
                   void* dst, int dst_mem);

/* ---- K6: post-demod data-parallel stages, batched over channel rows -------------
 * Everything the backend demods run between the frontend's narrowband complex64 stream and their first
 * sample-serial loop (fsk4_demod_ff, clock recovery, vocoders: out of scope):
 *   RCB_POST_P25_C4FM  p25_control_demod.py:106-133, logging_receiver.py:229-245
 *       freq_xlating_fir_filter_ccc(1, taps0, 0, rate)  ->  analog.quadrature_demod_cf(gain)  ->
 *       fir_filter_fff(1, taps1)   [taps1 = (1/sps,)*sps symbol filter]         -> out (float, n per row)
 *       and the AFC probe  moving_average_ff(probe_len, 1, ..) -> multiply_const(probe_scale) -> probe_signal_f
 *       on the demod output (p25_control_demod.py:123-127): `probe[r]` = its value after this block.
 *   RCB_POST_ANALOG_FM  logging_receiver.py:210-222
 *       analog.pwr_squelch_cc(squelch_db, squelch_alpha, 0, squelch_gate) -> analog.fm_demod_cf = quadrature_demod_cf
 *       (gain) -> fm_deemph = iir_filter_ffd([deemph_b0, deemph_b1], [1, deemph_a1]) -> fir_filter_fff(1, taps0)
 *       [audio low-pass] -> fir_filter_fff(1, taps1) [300 Hz high-pass] -> rational_resampler_fff(interp, decim, taps2)
 *       -> out (float, nout[r] per row: depends on the resampling ratio and on what the squelch gated away).
 * One chain = `rows` channels with identical parameters, each with its own streaming state (filter histories,
 * IIR / squelch state, resampler phase), so any split of the streams into calls gives the same samples.
 * Input: rows of n complex64 at iq + row_map[r] * in_stride (row_map NULL = identity; host array) - e.g. rows of a
 * device-resident rcb_pfb_process IQ output, so channelised samples never leave the GPU before the symbol filter.
 * Output: row r at out + r * out_stride (floats), nout[r] items (host int array of `rows`). */
enum { RCB_POST_P25_C4FM = 1, RCB_POST_ANALOG_FM = 2 };
typedef struct rcb_post_cfg {
    int kind;
    int rows;
    float gain;                        /* quadrature demod gain */
    const float* taps0; int ntaps0;    /* P25: channel prefilter       ANALOG: audio low-pass  */
    const float* taps1; int ntaps1;    /* P25: symbol filter           ANALOG: high-pass       */
    const float* taps2; int ntaps2;    /*                              ANALOG: resampler prototype */
    int interp, decim;                 /*                              ANALOG: resampler ratio (reduced) */
    double squelch_db, squelch_alpha;  /*                              ANALOG */
    int squelch_gate;
    double deemph_b0, deemph_b1, deemph_a1;
    int probe_len; float probe_scale;  /* P25: AFC probe (0 = none) */
} rcb_post_cfg;
int rcb_post_open(rcb_t* h, const rcb_post_cfg* cfg, int* chain_id);
int rcb_post_process(rcb_t* h, int chain_id, const void* iq, size_t n, size_t in_stride, const int* row_map,
                     int in_mem, void* out, size_t out_stride, int out_mem, int* nout, float* probe);
int rcb_post_close(rcb_t* h, int chain_id);

/* ---- K3: streaming windowed FFT + log-power accumulation ---------------------------------------
 * Replaces stream_to_vector -> fft.fft_vcc(L, True, window, True) -> complex_to_mag_squared ->
 * nlog10_ff(1, L, 1) -> moving_average_ff(avg, 1, ...)             fft_vector.py:37-60.
 * Frames are consecutive length-L blocks; one output vector per block of `avg` consecutive frames:
 * S[k] = sum_f (log10(max(|fftshift(FFT(x_f*w))[k]|^2, 1e-18)) + 1).  The vector the reference
 * writes (frames 900..999 of 1000, avg 100) is output block 9.  nsamples must be a multiple of L;
 * partial avg blocks are carried to the next call.  *nvec = vectors written to out_sums. */
int rcb_fft_config(rcb_t* h, int length, const float* window, int avg_frames);
int rcb_fft_reset(rcb_t* h);
int rcb_fft_process(rcb_t* h, const void* iq, size_t nsamples, int in_mem, void* out_sums,
                    size_t cap_vectors, int out_mem, size_t* nvec);
/* Which kernels run rcb_fft_process.
 * 0 (default): 16384-point frames - fft_vector.py:32, the reference's own scan length - run on fft_frame_kernel: the whole
 *    frame lives in one SM's shared memory (bulk-copied in, both FFT passes in place, log-power summed per group of
 *    about a dozen frames, 2 launches per call); every other length runs the column pass / row pass / fold pipeline per
 *    L2-resident sub-batch on two streams.
 * 1: ONE persistent launch per call for every length (task queue over column and row tiles of all frames, scratch ring
 *    kept in L2, block sums accumulated in frame order by the row tiles); needs frames of many tiles (2^18, 2^20 points)
 *    to be competitive.
 * 2: the column / row / fold pipeline for every length.
 * The block sum is associated in frame order by the tiled pipelines and in (frame-in-group, group) order by the
 * frame-resident kernel - both independent of how the stream is split into calls - so a switch restarts the current
 * averaging block (like rcb_fft_reset). */
int rcb_fft_set_pipeline(rcb_t* h, int persistent);

#ifdef __cplusplus
}
#endif
#endif /* B200CHAN_H */
