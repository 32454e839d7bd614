/*
 * pdt.h — C-ABI of the B200-native IQ demodulation chain (batch / context API).
 *
 * This is the THROUGHPUT boundary: many independent IQ captures, device-resident, one fused sm_100a
 * kernel per batch.  It replaces, for a whole capture at a time, the per-chunk loop body of the
 * reference drivers:
 *     POESTIPdemod/main.c:373-482   (StaticGain -> CarrierTrackPLL -> LowPassFilterInterp -> NormalizingAGC
 *                                    -> GardenerClockRecovery -> ManchesterDecode -> ByteSyncOnSyncword)
 *     ARGOSdemod/main.c:250-300     (… -> LowPassFilter -> NormalizingAGC -> Squelch -> … -> FindSyncWords)
 * The drop-in LEGACY boundary (the reference's own function signatures, one call per stage per chunk)
 * is declared in pdt_legacy.h; both live in the same shared objects:
 *     libpdt_f32.so  — DECIMAL_TYPE float  (POESTIPdemod/config.h:4)
 *     libpdt_f64.so  — DECIMAL_TYPE double (ARGOSdemod/config.h:4)
 * Plain C types only; no torch / CUDA types in any signature (streams and device pointers are void*).
 *
 * Error behaviour: functions returning int give 0 on success, a negative PDT_E* code otherwise and
 * record a message retrievable with pdt_last_error().  There is NO CPU fallback: without a usable CUDA
 * device every compute entry point fails with PDT_ENODEV.
 */
#ifndef PDT_H
#define PDT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDT_OK        0
#define PDT_ENODEV   -1   /* no CUDA device / driver */
#define PDT_ECUDA    -2   /* a CUDA call or kernel failed */
#define PDT_EINVAL   -3   /* bad argument */
#define PDT_ENOMEM   -4

#define PDT_MODE_POES  0
#define PDT_MODE_ARGOS 1

#define PDT_FRAME_MAX_BYTES 104

/* PDT_ENGINE_EXACT: one CTA per capture, every recurrence strictly serial (any mode / precision).
 * PDT_ENGINE_TILED: float POES chain only; recurrences parallelised in time with warm-up tiles that are accepted
 *                   only when bit-identical to the serial continuation (re-run otherwise) -> same results.
 * PDT_ENGINE_AUTO : TILED where it applies, else EXACT. */
#define PDT_ENGINE_AUTO  0
#define PDT_ENGINE_EXACT 1
#define PDT_ENGINE_TILED 2

/* All DSP constants of the reference drivers, in the units the reference uses (loop gains in rad/s,
 * scaled by 2π/Fs at the call site in double and narrowed to DECIMAL_TYPE: POESTIPdemod/main.c:413,429). */
typedef struct pdt_params {
    int      mode;               /* PDT_MODE_POES | PDT_MODE_ARGOS */
    double   sample_rate;        /* header.sample_rate (main.c:346) */
    uint32_t chunk;              /* DEFAULT_CHUNKSIZE: 10000 (POES main.c:30) / 2400 (ARGOS main.c:27).
                                    Part of the numerical behaviour (Gardner works in chunk-relative floats). */
    int      interp;             /* L = rint(150000/Fs) (main.c:347); ARGOS 1 */
    int      taps;               /* N = 26·L (main.c:348); ARGOS 50 */
    int      force_min_interp1;  /* declared deviation for Fs >= 300 ksps where the reference emits nothing */
    double   max_carrier_dev;    /* 4500 / 550 Hz */
    double   pll_acq_gain, pll_track_gain, pll_lock_alpha;   /* rad/s */
    double   pll_lock_thresh;
    double   agc_attack, agc_decay;                          /* rad/s */
    double   lpf_fc;             /* 11000 / 700 Hz */
    double   baud;               /* symbol (chip) rate: 16640.3 / 800 */
    double   gardner_err_lim, gardner_gain;                  /* 0.1, 3.0 */
    double   manchester_resync;  /* 1.0 (POES main.c:445 passes the literal) / 0.5 */
    double   squelch_thresh;     /* ARGOS 0.15; unused for POES */
    double   norm_factor;        /* 0 = StaticGain of the first chunk (main.c:384-389), else override (-n) */
    char     sync_word[32];      /* ASCII '0'/'1' */
    int      sync_len;           /* 19 / 13 */
    /* engine selection (not in the reference; results are identical, only the speed differs) */
    int      engine;             /* PDT_ENGINE_AUTO | PDT_ENGINE_EXACT | PDT_ENGINE_TILED */
    uint32_t pll_warm, pll_tile; /* tiled engine: PLL warm-up / tile length in samples (0 = derived from the loop bandwidth) */
    uint32_t agc_min_tile;       /* tiled engine: smallest AGC tile in interpolated samples (0 = default) */
    uint32_t acq_first;          /* tiled engine: samples covered by the first acquisition pass; captures that have not
                                    latched by then continue on a second stream (0 = default 131072) */
    /* stages the reference exports next to the path but calls from neither driver */
    int      sync_generic;       /* 1: the parameterised sync of common/ByteSync.c:16-144 instead of the application copy:
                                    a frame ends when its byte index exceeds sync_frame_len, bitIdx = sync_start_bit after a
                                    sync word ("3 for POES, 0 for ARGOS", :135), inverse search on, literal ED E2 in front */
    int      sync_frame_len, sync_start_bit;   /* frameLength (<= 103), startBit (0…7) */
    int      clock_recovery;     /* PDT_CLOCK_GARDNER (default) | PDT_CLOCK_MM: common/MMClockRecovery.c:5-84 in place of the
                                    Gardner loop (the call both drivers keep commented out: POESTIPdemod/main.c:435,
                                    ARGOSdemod/main.c:277).  Runs on the exact engine. */
    double   mm_step_range, mm_gain;           /* stepRange, kp of MMClockRecovery (the commented call sites pass 3, 0.15) */
} pdt_params;
#define PDT_CLOCK_GARDNER 0
#define PDT_CLOCK_MM      1

/* One decoded minor frame (POES, 104 bytes incl. the literal ED E2) or packet (ARGOS, 7 bytes). */
typedef struct pdt_frame {
    uint64_t sample_index;   /* absolute index, in interpolated samples, of the Gardner pick whose bit completed the sync word */
    uint32_t bit_index;      /* absolute index of that bit in the capture's bit stream */
    uint8_t  inverse;        /* 1 = matched the inverted sync word (POESTIPdemod/ByteSync.c:127) */
    uint8_t  n_bytes;        /* bytes valid in `bytes` (104 / 7 when complete, fewer for a trailing partial frame) */
    uint8_t  complete;
    uint8_t  pad;
    uint8_t  bytes[PDT_FRAME_MAX_BYTES];
} pdt_frame;                 /* 120 bytes */

/* Post-checks of one decoded POES minor frame, computed on the device right behind the frame shifter
 * (standalone_matlab/Functionized/checkParity.m:20-90, daytimeDecode.m:4): the step after the hot path. */
typedef struct pdt_frame_quality {
    uint16_t counter;        /* 9-bit minor-frame counter: (byte4 & 1) << 8 | byte5, 0…319 */
    uint8_t  spacecraft;     /* byte 2 (8 = NOAA-15, 13 = NOAA-18, 15 = NOAA-19) */
    uint8_t  parity_ok;      /* the five even-parity bits of word 103 agree with words 2-18 / 19-35 / 36-52 / 53-69 / 70-86 */
    uint8_t  parity_bits;    /* which of the five groups failed (bit 4 = first group … bit 0 = last) */
    uint8_t  continuous;     /* counter == previous complete frame's counter + 1 (mod 320); 1 for the first frame */
    uint8_t  valid;          /* frame complete (104 bytes) — the other fields are meaningful only then */
    uint8_t  pad;
} pdt_frame_quality;         /* 8 bytes */

/* Per-capture summary (what the reference prints on its progress line, main.c:461-481). */
typedef struct pdt_capture_stats {
    uint64_t n_samples, n_symbols, n_bits;
    uint32_t n_frames;       /* sync words accepted (== frames incl. a trailing partial one) */
    int32_t  locked;         /* PLL latch fired */
    uint64_t lock_sample;    /* absolute input-sample index of the latch */
    double   lock_freq_hz;   /* d_freq·Fs/2π at the latch (CarrierTrackingPLL.c:269) */
    double   norm_factor;    /* StaticGain result */
    double   avg_phase;      /* last CarrierTrackPLL return value */
    double   final_phase, final_freq, final_gain, final_next;   /* loop states at the end (re-stitch / debugging) */
    int32_t  prelocked;      /* stream segments only (pdt_demod_segments_device): 0 = the reference's acquisition sweep ran
                                (`locked`/`lock_sample`/`lock_freq_hz` are a real latch of CarrierTrackingPLL.c:266-274);
                                1 = started in TRACK mode from the FFT carrier estimate, no latch ever fired (`locked` is
                                then 1 by construction and lock_freq_hz is the estimate); 2 = the estimate was rejected
                                (spectral peak below PDT_PRELOCK_MIN_SNR times the mean) and the segment fell back to
                                the reference's sweep from zero */
    float    prelock_snr;    /* peak / mean of |DFT|^2 over the searched band of the carrier estimate (0 when not estimated) */
} pdt_capture_stats;
#define PDT_PRELOCK_MIN_SNR 16.0f

/* Optional per-capture trace taps (device pointers, any may be NULL).  REAL = float in libpdt_f32, double in f64. */
typedef struct pdt_traces {
    void     *pll_phase, *pll_freq, *pll_out;   /* REAL[n]      d_phase / d_freq BEFORE the update, and realDataOut */
    void     *lock;                             /* REAL[n]      d_locksig stream */
    void     *lpf, *agc;                        /* REAL[n·L]    FIR output, AGC (+squelch) output */
    void     *sym, *gardner_err;                /* REAL[cap]    symbol values, clamped Gardner error */
    uint64_t *gardner_idx;                      /* u64[cap]     absolute interp-sample index picked */
    uint8_t  *bits;                             /* u8[cap]      ASCII '0'/'1' */
    uint64_t  cap;
} pdt_traces;

typedef struct pdt_ctx pdt_ctx;

const char *pdt_version(void);
const char *pdt_last_error(void);
int         pdt_real_size(void);                       /* 4 (libpdt_f32) or 8 (libpdt_f64) */
int         pdt_device_count(void);                    /* <=0: no usable device */
int         pdt_set_device(int ordinal);

/* Reference defaults for a sample rate (mirrors the #defines at POESTIPdemod/main.c:30-104, ARGOSdemod/main.c:27-65). */
int         pdt_params_default(pdt_params *p, int mode, double sample_rate);

/* A context owns device workspaces for `max_captures` captures of up to `max_samples` IQ samples each and
 * `max_frames` frame slots per capture. */
pdt_ctx    *pdt_create(const pdt_params *p, uint32_t max_captures, uint64_t max_samples, uint32_t max_frames);
/* A context processes ONE batch at a time (its workspaces and internal streams belong to that batch until the results
 * have been fetched).  To keep several batches in flight — the serial acquisition tail of one batch then overlaps the bulk
 * kernels of the next — create one context per batch in flight and use them in rotation, each on its own stream
 * (bench.py does this with three). */
void        pdt_destroy(pdt_ctx *ctx);
int         pdt_get_params(const pdt_ctx *ctx, pdt_params *out);
int         pdt_get_taps(const pdt_ctx *ctx, void *h_out /* REAL[taps] */);

/* Demodulate a batch that is ALREADY in device memory.
 *   d_iq        REAL[2·n] interleaved I,Q per capture; capture c starts at d_iq + 2·c·stride_samples
 *               (pcm16 != 0: int16_t[2·n] raw PCM as in a WAV data chunk, normalised /32768 in-kernel, wave.c:141-166)
 *   n_samples   host array [n_captures] (NULL: every capture has stride_samples samples).  Entries may EXCEED
 *               stride_samples: the captures then overlap in memory (segments of one stream, see pdt_stream_plan);
 *               the input is only ever read.
 *   traces      host array [n_captures] of device trace taps, or NULL
 *   stream      cudaStream_t (NULL = default stream).  Asynchronous: results are valid after the stream is synchronised.
 * Results stay on the device until pdt_fetch(). */
int         pdt_demod_device(pdt_ctx *ctx, const void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                             const uint64_t *n_samples, const pdt_traces *traces, void *stream);

/* Same from HOST buffers (pinned or pageable): H2D of the samples, the kernels, and D2H of stats+frames, synchronous. */
int         pdt_demod_host(pdt_ctx *ctx, const void *h_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                           const uint64_t *n_samples, pdt_capture_stats *stats_out, pdt_frame *frames_out /* [n_captures·max_frames] */);

/* Asynchronous form: enqueues the chunked H2D (on an internal copy stream) and the kernels (on `stream`) and returns;
 * collect the results with pdt_fetch(ctx, …, stream).  Two or three contexts used in rotation keep the PCIe link and the
 * GPU busy at the same time (bench.py e2e).  The host buffer must stay valid (and should be pinned) until the fetch. */
int         pdt_demod_host_async(pdt_ctx *ctx, const void *h_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples,
                                 const uint64_t *n_samples, void *stream);

/* Copy results of the last pdt_demod_device() to the host (synchronises `stream`). */
int         pdt_fetch(pdt_ctx *ctx, uint32_t n_captures, pdt_capture_stats *stats_out, pdt_frame *frames_out, void *stream);

/* Frame post-checks of the last batch (POES): runs k_frame_checks on `stream` over the device frame table and copies
 * the [n_captures·max_frames] quality table to the host (synchronises `stream`). */
int         pdt_frame_checks(pdt_ctx *ctx, uint32_t n_captures, pdt_frame_quality *quality_out, void *stream);

/* ---- one long stream cut into overlapping segments (BASELINE configs[4]) ------------------------
 * The reference processes a recording strictly serially (POESTIPdemod/main.c:373-482).  A long stream is cut into
 * segments that start every `segment` samples and are lead + segment + tail samples long; each one is demodulated as an
 * independent capture of a batch (re-acquiring carrier, gain, symbol clock and frame sync inside its lead), and the frame
 * tables are stitched: segment s keeps the frames whose sync word completed in ITS window
 *     [s·segment + lead, (s+1)·segment + lead)      (segment 0: from 0; the last segment: to the end of the stream),
 * `tail` (> one minor frame) lets the last owned frame complete.  Segment 0 is bit-identical to the serial chain by
 * construction; later segments decode the same bits once locked but are not guaranteed bit-identical around a seam, so
 * the acceptance test is frame-counter continuity across the seams plus equality with the serial result on a prefix
 * (SURVEY §8d "C5").  Segments are independent units: they shard across GPUs like captures do (rank r takes a
 * contiguous range of segments; the only exchange is the gather of the frame tables before the stitch). */
/*
 * Segments behind the first do not repeat the reference's acquisition sweep from zero (CarrierTrackingPLL.c:115-262; a
 * serial, seconds-long search that the serial chain runs exactly once per recording): pdt_demod_segments_device starts
 * every capture with index >= n_serial directly in TRACK mode from a carrier estimate over its first 1024·D samples —
 * the same guess + warm-up with which the tiled engine starts the PLL tiles inside a capture.  Captures below n_serial
 * (the stream's first segment, on the rank that holds it) run the reference chain unchanged. */
typedef struct pdt_stream_plan {
    uint64_t total_samples, segment, lead, tail;
    uint32_t n_segments;
    uint32_t interp;             /* samples of pdt_frame.sample_index per input sample (max(params.interp, 1)) */
    uint32_t seam_tol;           /* ownership tolerance in interpolated samples (two symbols): a sync position is not
                                    bit-identical between neighbouring segments (own chunk grid, chunk-relative float
                                    Gardner state), so a frame within seam_tol of a window edge is claimed by BOTH
                                    neighbours and the later copy is dropped by position (see pdt_stream_stitch) */
    uint32_t pad;
} pdt_stream_plan;
/* lead / tail 0 = defaults (0.3 s and 0.13 s + 4096 samples of signal).  Captures handed to pdt_demod_segments_device():
 * d_iq = stream base (+ first·segment), stride_samples = segment, n_samples[s] = pdt_stream_segment_length(plan, s). */
int         pdt_stream_plan_make(pdt_stream_plan *plan, const pdt_params *p, uint64_t total_samples, uint64_t segment,
                                 uint64_t lead, uint64_t tail);
int         pdt_demod_segments_device(pdt_ctx *ctx, const void *d_iq, int pcm16, uint32_t n_segments, uint64_t stride_samples,
                                      const uint64_t *n_samples, uint32_t n_serial, void *stream);
uint64_t    pdt_stream_segment_length(const pdt_stream_plan *plan, uint32_t s);
/* Stitch the tables of segments [first, first + n) (stats[n], frames[n·max_frames]) into out[]: owned frames only, in
 * stream order, sample_index rewritten to the stream-global interpolated-sample index (bit_index is left segment-local).
 * Ownership is robust against the +-1-symbol position jitter between neighbouring segments: segment s claims
 * [lo - seam_tol, hi + seam_tol) and a frame whose global position lies within 2·seam_tol of the frame emitted just
 * before it is the same minor frame seen twice and is skipped (minor frames are 1664 symbols apart).  The outer edges of
 * the range [first, first + n) stay exact, so ranges stitched separately (one per rank) concatenate to a partition of the
 * stream; stitching the gathered tables in ONE call (stream.gather_and_stitch) is the form that is jitter-proof at every seam.
 * Equality with the serial chain holds from the point where the serial chain itself has locked: segments start in track
 * mode whether or not the reference's sweep would have latched by then (pdt_capture_stats.prelocked).
 * Returns the number of frames written, or <0. */
long        pdt_stream_stitch(const pdt_stream_plan *plan, uint32_t first, uint32_t n, const pdt_capture_stats *stats,
                              const pdt_frame *frames, uint32_t max_frames, pdt_frame *out, uint32_t out_cap);
/* Post-checks of a stitched list (parity word 103, counter, continuity ACROSS the seams): the host twin of
 * pdt_frame_checks, whose device table is per segment. */
int         pdt_stream_frame_checks(const pdt_frame *frames, uint32_t n_frames, pdt_frame_quality *quality_out);
/* Synthetic stream slice: samples [start, start + n) of ONE seeded POES stream of total_samples (Doppler crossing from
 * f0 to -f0 over the whole stream), written at d_iq. */
int         pdt_synth_poes_stream_device(void *d_iq, int pcm16, uint64_t start_sample, uint64_t n_samples,
                                         uint64_t total_samples, double sample_rate, uint64_t seed, void *stream);

/* ---- live mode: bounded-latency streaming ---------------------------------------------------------
 * The reference's sound-card drivers (POESTIPdemodPortAudio/main.c:324-401, ARGOSdemodPortAudio/main.c:290-329) read a
 * chunk from the audio device and run the stage functions on it, for ever; the stages carry their state in statics.  Live
 * mode is that loop for up to `max_captures` streams at once: every push hands the NEXT `n` samples of every stream to the
 * chain, which continues exactly where the previous push stopped — PLL, FIR history, AGC gain, Gardner position (and the
 * chunk buffer it looks back into), Manchester phase, the frame being shifted in.  A push is what one iteration of the
 * reference's loop is: the result equals the reference fed with the same sequence of chunk lengths (pushes longer than
 * params.chunk are cut into chunks of that length).  Latency of a push = one chunk through the serial chain (a few ms).
 *   pdt_live_begin        (re)start: all streams back to the reference's initial state
 *   pdt_live_push_device  enqueue one push (d_iq: stream s at d_iq + 2·s·stride_samples, n <= stride_samples samples each)
 *   pdt_live_push_host    same from host buffers, then pdt_fetch: synchronous
 * Results: pdt_capture_stats are running totals (n_samples, n_symbols, n_bits, n_frames since pdt_live_begin); the frame
 * table is a RING — frame number k of a stream (0-based, in order of its sync word) is slot k mod max_frames; a frame whose
 * `complete` is 0 is still being shifted in and will be finished by a later push.  sample_index is absolute in the stream. */
int         pdt_live_begin(pdt_ctx *ctx);
int         pdt_live_push_device(pdt_ctx *ctx, const void *d_iq, int pcm16, uint32_t n_streams, uint64_t stride_samples,
                                 uint64_t n, void *stream);
int         pdt_live_push_host(pdt_ctx *ctx, const void *h_iq, int pcm16, uint32_t n_streams, uint64_t stride_samples,
                               uint64_t n, pdt_capture_stats *stats_out, pdt_frame *frames_out);

/* Device addresses of the result tables of the last batch (for NCCL gathers without a host bounce). */
int         pdt_result_tables(pdt_ctx *ctx, void **d_stats, void **d_frames, uint32_t *max_frames);

/* Render frames exactly like the reference's output file ("%.5f ED E2 XX …\n", POESTIPdemod/ByteSync.c:96-101,62,69).
 * The time column emulates wave.c's float-accumulated axis.  Returns bytes written (excluding NUL) or <0. */
long        pdt_format_frames(const pdt_ctx *ctx, const pdt_frame *frames, uint32_t n_frames, char *buf, size_t cap);

/* Engine that the context resolved to (PDT_ENGINE_EXACT / PDT_ENGINE_TILED) and, for the tiled engine, the
 * speculation counters of the last batch (valid after the stream is synchronised):
 *   out[0] PLL tiles re-run, out[1] AGC tiles re-run, out[2] acquisition restarts, out[3] PLL tiles per capture (max). */
int         pdt_engine(const pdt_ctx *ctx);
int         pdt_tiled_counters(pdt_ctx *ctx, uint32_t out[4], void *stream);

/* Capture groups of the tiled engine: a batch is cut into up to `max_groups` groups of >= 64 captures that run the kernel
 * sequence on internal streams (one group's latency-bound kernels overlap another's streaming kernels).  0 = default (3).
 * Many small batches in flight (strong scaling: a fixed batch sharded over several GPUs) do better with 1: the overlap
 * then comes from the other batches and the process stays within the hardware work queues. */
int         pdt_set_groups(pdt_ctx *ctx, int max_groups);

/* Per-kernel device times of the tiled engine: when enabled, a CUDA event is recorded on the launch stream after
 * every kernel of a batch; pdt_kernel_times() synchronises and returns how many kernels the last batch launched,
 * their names (static strings) and durations in milliseconds. */
int         pdt_set_profiling(pdt_ctx *ctx, int enable);
int         pdt_kernel_times(pdt_ctx *ctx, const char **names, float *ms, int cap);
/* pdt_set_profiling(ctx, 2): timeline mode — the batch runs exactly as in production (capture groups on internal
 * streams) with a timing event after every kernel of every stream; pdt_timeline() returns for each of them its name,
 * its stream (capture group 0…, 99 = the slow-capture stream) and its END time in ms since the batch was forked. */
int         pdt_timeline(pdt_ctx *ctx, const char **names, int *groups, float *end_ms, int cap);

/* Debug: cycle accounting of the acquisition pipeline's second pass, summed over captures since the last reset:
 * [0] steps, [1] step cycles, [2] core-lane busy, [3] EMA-lane busy, [4] helper busy, [5] decision phase, [6] epochs. */
int         pdt_debug_acq_prof(uint64_t out[8], int reset);
/* cycle accounting of the exact engine (pdt_chain_kernel.cuh g_chain_prof; thread 0 of every CTA, summed over captures since the
 * last reset): [0] StaticGain [1] PLL [2] its P phase alone [3] C || E [4] H || P [5] C busy [6] E busy [7] PLL blocks [8] contradicted
 * blocks [9] FIR [10] AGC [11] clock recovery + bits [12] whole capture [13] samples [14] PLL control [15] its work alone [16] lock-EMA thread busy.  Diagnostics only. */
int         pdt_debug_chain_prof(uint64_t out[20], int reset);

/* Number of kernels launched by this library since load (bench.py reports it as gpu_launches). */
uint64_t    pdt_launch_count(void);

/* Seeded synthetic POES-TIP capture generator ON THE DEVICE (bench workloads; SURVEY §8d signal model).
 * Writes REAL[2·n] (or int16 when pcm16) per capture at d_iq + 2·c·stride.  Per-capture SNR/Doppler are derived from seed+c. */
int         pdt_synth_poes_device(void *d_iq, int pcm16, uint32_t n_captures, uint64_t stride_samples, uint64_t n_samples,
                                  double sample_rate, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDT_H */
