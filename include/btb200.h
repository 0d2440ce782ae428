/*
 * btb200.h -- C ABI of the B200-native gr-bluetooth multi-channel receive path.
 *
 * This is the drop-in boundary: everything the reference does between
 * "GNU Radio hands work() a window of IQ" and "ac()/aa() get a symbol pointer"
 * runs behind these entry points, in hand-written sm_100a kernels.  Plain
 * pointers and sizes only.  No CPU fallback: every entry point fails with
 * BTB200_ERR_NO_DEVICE / BTB200_ERR_CUDA when no usable GPU is present.
 *
 * What each entry point replaces in the reference (paths under
 * /root/reference):
 *
 *   btb200_create        gr::bluetooth::multi_block::multi_block()
 *                        lib/multi_block.cc:40-120, set_channels() :306-342,
 *                        set_symbol_history() :299-303 (tap design, channel
 *                        plan, DDC objects, history) -- reached through
 *                        multi_sniffer::make() lib/multi_sniffer_impl.cc:42-48,
 *                        multi_LAP::make() lib/multi_LAP_impl.cc:40-43,
 *                        multi_hopper::make() lib/multi_hopper_impl.cc:36-40.
 *   btb200_process       the body of multi_sniffer_impl::work()
 *                        lib/multi_sniffer_impl.cc:82-166 for n_slots
 *                        consecutive calls: channel_samples()
 *                        lib/multi_block.cc:180-228, check_snr() :253-296,
 *                        channel_symbols() :230-251 (demod :158-168, mm_cr
 *                        :128-155, slicer :171-178), classic_packet::sniff_ac
 *                        lib/packet_impl.cc:247-268 + check_ac :471-510 +
 *                        acgen :309-364, le_packet::sniff_aa :1452-1527.
 *                        Each returned hit is one ac()/aa() invocation
 *                        (lib/multi_sniffer_impl.cc:117,139) with its
 *                        arguments.
 *   btb200_process_device same, input already resident in HBM.
 *   btb200_get/set_mm_state  the three floats of M&M state that persist
 *                        across work() calls: include/gr_bluetooth/multi_block.h:86-93.
 *   btb200_get_stage     no reference counterpart: debug taps for parity tests.
 */
#ifndef BTB200_H
#define BTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTB200_ABI_VERSION 1

#if defined(__GNUC__)
#define BTB200_API __attribute__((visibility("default")))
#else
#define BTB200_API
#endif

/* error codes (negative); the reference abort()s or throws instead
 * (lib/multi_sniffer_impl.cc:36-40) */
enum {
  BTB200_OK              = 0,
  BTB200_ERR_ARG         = -1,   /* bad argument / configuration */
  BTB200_ERR_NO_DEVICE   = -2,   /* no CUDA device: there is NO CPU fallback */
  BTB200_ERR_CUDA        = -3,   /* CUDA runtime error, see btb200_last_error() */
  BTB200_ERR_NOMEM       = -4,
  BTB200_ERR_SHORT_INPUT = -5,   /* n_samples < (n_slots-1)*S + H */
  BTB200_ERR_TOO_MANY    = -6,   /* n_slots > max_slots_per_call */
  BTB200_ERR_OVERFLOW    = -7    /* reserved for callers: a hit or symbol buffer was too small (see btb200_hits.overflow) */
};

/* how clock-recovery / rotator state is carried from one channel-window to the next */
enum {
  /* reference behaviour: one M&M state shared by all channels and all work()
   * calls (include/gr_bluetooth/multi_block.h:86-93), DDC rotators free-running.
   * Serial by construction; bit-exact with the reference on whole files. */
  BTB200_MM_CHAINED   = 0,
  /* every channel-window starts from the constructor state
   * (lib/multi_block.cc:91-98), rotators restart at phase 1.  Windows become
   * pure functions of their samples: parallel, shardable, bit-exact with the
   * oracle run in the same mode. */
  BTB200_MM_STATELESS = 1
};

/* when the SNR squelch (lib/multi_block.cc:253-296) is evaluated in stateless mode */
enum {
  BTB200_SQUELCH_DEFAULT = 0,    /* lazy */
  /* eager: noise DDC + energies for every channel-window before demodulation, like the reference */
  BTB200_SQUELCH_EAGER   = 1,
  /* lazy: every window is demodulated and searched (no state is shared between windows, so a
   * squelched window cannot influence another); the exact squelch arithmetic -- 20001-tap noise
   * DDC at 100 Msps -- runs only for windows that produced a hit, and hits from windows the
   * reference would have squelched are dropped.  Same hit list, bit for bit. */
  BTB200_SQUELCH_LAZY    = 2
};

/* how the off-channel (noise) energy behind the squelch / printed snr is obtained (lazy squelch) */
enum {
  /* exact: the reference's 20001-tap noise DDC in the oracle's operation order, for every hit window;
   * hit.snr is bit-identical to the reference's double */
  BTB200_SNR_EXACT        = 0,
  /* fast, guarded: a polyphase + DFT estimate (fp32, FMA, ~1e-4 dB from the exact value) decides the
   * squelch and provides hit.snr whenever it is farther than the guard band (5e-3 dB) from the
   * squelch threshold and from every %.1f rounding boundary; otherwise the exact value is computed.
   * Same hit list and same printed snr=%.1f as BTB200_SNR_EXACT; hit.flags bit1 marks estimated snr. */
  BTB200_SNR_FAST_GUARDED = 1
};

/* how much of a window's symbol stream is produced before the access-code search (lazy squelch) */
enum {
  /* lazy tail (default where the geometry allows it): only lags 0..624 are ever searched, i.e. 697 of the
   * ~3742 symbols of a window, and the clock-recovery loop advances at most 5 samples per step, so every
   * window is demodulated up to symbol 704 only; windows with a hit are resumed from the saved loop state
   * to produce the rest (bit-identical: same loop, same state).  Symbol counts / bit streams of windows
   * without a hit then stop at 704 (btb200_get_stage). */
  BTB200_TAIL_LAZY = 0,
  BTB200_TAIL_FULL = 1           /* every window to the end */
};

/* how the per-channel DDCs (lib/multi_block.cc:180-228, 329-341) are evaluated */
enum {
  /* exact: every channel's band-pass FIR in direct form, one rounding per operation, sums in the reference's order:
   * DDC outputs, demod floats, soft symbols and bit streams equal the reference's bit for bit (either mm_mode) */
  BTB200_DDC_EXACT     = 0,
  /* polyphase (throughput mode, BTB200_MM_STATELESS only): ONE real-tap polyphase bank of fs / 1 MHz branches + DFT at
   * the channel bins for all channels, FM demod and window energy fused into its epilogue, off-channel energy from the
   * polyphase noise estimator.  Mathematically the same filters; floats agree with the reference's to a TOLERANCE
   * (demod floats: median 1e-5, 99.9 % below 5e-3 of the +-4 range at 100 Msps -- the reference's own taps carry fp32
   * phase errors of that size; snr within 2e-3 dB).  Because the Mueller & Mueller loop random-walks in noise, bit
   * streams of noise-only stretches and marginal detections differ from the reference's (measured: ~98 % of the hit
   * list in common, identical set of LAPs, same recall against ground truth); hit.flags bit2 marks such hits.
   * Needs an even integer number of samples per MHz. */
  BTB200_DDC_POLYPHASE = 1
};

enum {
  BTB200_SEARCH_BR = 1,          /* classic_packet::sniff_ac */
  BTB200_SEARCH_LE = 2,          /* le_packet::sniff_aa */
  /* with BTB200_SEARCH_BR: the access-code test of libbtbb's btbb_find_ac -- what multi_LAP and multi_UAP call
   * (lib/multi_LAP_impl.cc:93 with LAP_ANY and max_ac_errs 1, lib/multi_UAP_impl.cc:95 with the piconet's LAP and 2) --
   * instead of sniff_ac's: syndrome decoding of the (64,30) code of the sync word with up to `max_ac_errors` corrected
   * bits, or the Hamming distance to a given LAP's sync word; parameters in btb200_config.bch.  libbtbb is an external
   * library outside the reference tree: restated from its published algorithm, PARITY UNPINNED.
   * hit.lap is the corrected LAP, hit.ac_errors the corrected-bit count, hit.offset the preamble's position. */
  BTB200_SEARCH_BR_BCH = 4
};

/* which block's window geometry / search */
enum {
  BTB200_BLOCK_SNIFFER = 0,      /* history += 3125 symbols (multi_sniffer_impl.cc:61) */
  BTB200_BLOCK_LAP     = 1       /* history += 68 symbols   (multi_LAP_impl.cc:54)     */
};

typedef struct btb200_config {
  uint32_t abi_version;          /* BTB200_ABI_VERSION */
  double   sample_rate;          /* make() arg 1 */
  double   center_freq;          /* make() arg 2 */
  double   squelch_threshold;    /* make() arg 3, dB */
  uint32_t extra_history_symbols;/* 3125 or 68 */
  int32_t  mm_mode;              /* BTB200_MM_* */
  int32_t  search;               /* BTB200_SEARCH_* bit mask */
  int32_t  device;               /* CUDA device ordinal */
  uint32_t max_slots_per_call;   /* sizes device buffers; 0 = default */
  uint32_t keep_stages;          /* 1: keep demod/soft-symbol buffers for btb200_get_stage */
  uint32_t squelch_mode;         /* BTB200_SQUELCH_* (stateless mode only; chained is always eager) */
  uint32_t snr_mode;             /* BTB200_SNR_* (lazy squelch only) */
  uint32_t tail_mode;            /* BTB200_TAIL_* (lazy squelch only) */
  uint32_t ddc_mode;             /* BTB200_DDC_* */
  uint32_t bch;                  /* BTB200_SEARCH_BR_BCH: BTB200_BCH_ANY(max_ac_errors) or BTB200_BCH_LAP(lap, max_ac_errors),
                                  * max_ac_errors 0..2; ignored (write 0) otherwise */
} btb200_config;
#define BTB200_BCH_ANY(max_err)      ((((uint32_t)(max_err)) & 7u) << 28)
#define BTB200_BCH_LAP(lap, max_err) ((((uint32_t)(lap)) & 0xffffffu) | (1u << 24) | ((((uint32_t)(max_err)) & 7u) << 28))

/* derived constants (lib/multi_block.cc:56-119, 299-342) */
typedef struct btb200_info {
  int32_t samples_per_slot;      /* S */
  int32_t history;               /* H = window length in samples */
  int32_t decimation;            /* D */
  int32_t chan_taps, noise_taps; /* Nc, Nn */
  int32_t first_channel_sample, first_noise_sample;
  int32_t channel_low, channel_high, n_channels;
  int32_t ddc_out_per_window;    /* 7494 @100 Msps */
  int32_t noise_out_per_window;  /* 850 */
  float   demod_gain;
  float   omega_mid;
  uint32_t max_slots_per_call;
  int32_t sm_count;
} btb200_info;

/* one ac()/aa() invocation of the reference (multi_sniffer_impl.cc:117,139) */
typedef struct btb200_hit {
  uint32_t slot;                 /* work() call index == clkn (multi_sniffer_impl.cc:173) */
  uint16_t channel;              /* classic channel 0..78 */
  uint16_t kind;                 /* 0: BR access code, 1: LE access address */
  int32_t  offset;               /* symbol index in the window where the packet starts */
  int32_t  n_symbols;            /* the "len - i" argument of ac()/aa() */
  uint32_t lap;                  /* BR: LAP (symbols 38..61); LE: AA (symbols 8..39) */
  uint32_t flags;                /* bit0: snr within 1e-6 dB of the threshold; bit1: snr from the fast noise estimate;
                                  * bit2: found by the polyphase (tolerance-mode) front end */
  double   snr;                  /* dB, the value ac() prints */
  uint64_t sym_offset;           /* into btb200_hits.symbols */
  uint32_t sym_count;            /* min(n_symbols, 3125) symbols copied (what classic_packet::make keeps,
                                  * lib/packet_impl.cc:52-58), one per byte, air order; 0 when the arena was full */
  uint32_t ac_errors;            /* BR: symbols among the first 68 that differ from the access code of `lap` (check_ac's
                                  * count, 0..6); LE: 0 */
} btb200_hit;

typedef struct btb200_hits {
  btb200_hit *hits;              /* caller-allocated */
  uint32_t    cap;
  uint32_t    count;             /* out; ordered by (slot, channel, kind, offset) = reference visiting order */
  uint32_t    overflow;          /* out: hits dropped because cap was too small */
  uint8_t    *symbols;           /* caller-allocated arena, may be NULL (no symbols wanted).  Borrowed symbols: pass
                                  * symbols = NULL and symbols_cap = UINT64_MAX; collect() then sets `symbols` to the
                                  * context's own pinned arena (no copy; valid until the next submit on this context)
                                  * and sym_offset of each hit points into it */
  uint64_t    symbols_cap;
  uint64_t    symbols_used;      /* out */
} btb200_hits;

typedef struct btb200_ctx btb200_ctx;

/* Exactness contract of BTB200_DDC_EXACT (what "bit-exact with the reference" rests on):
 *  - every float operation of the sample path rounds once, sums run in ascending index order (the reference built
 *    with -ffp-contract=off; GNU Radio's VOLK kernels would order sums differently on another machine);
 *  - demod_out[0] of a window, which the reference never writes (lib/multi_block.cc:164) yet reads, is 0.0f;
 *  - the interpolator index imu = rint(mu * 128) is clamped to 0..128 where GNU Radio's mmse_fir_interpolator_ff
 *    throws; mu stays inside [0, 1) by construction, so the clamp never acts on finite input;
 *  - snr = 10 log10(on / off) is evaluated by the host's libm on the device's exact fp64 energies.
 * Rates: 625 * fs / 1 MHz must be a multiple of the decimation (int)(fs / 1 MHz) / 2 (all windows share one decimation
 * grid): true for the even integer Msps rates, not for e.g. 5, 7, 9, 13 Msps, which the reference's block accepts --
 * btb200_create returns BTB200_ERR_ARG for those (btb200_last_error(NULL) says why). */
BTB200_API int  btb200_create(const btb200_config *cfg, btb200_ctx **out);
BTB200_API void btb200_destroy(btb200_ctx *ctx);
BTB200_API int  btb200_get_info(const btb200_ctx *ctx, btb200_info *out);

/*
 * Process n_slots consecutive work() calls.  iq points at the first sample of
 * the first call's window -- exactly what GNU Radio passes as input_items[0]
 * to a sync_block with history H -- and must hold at least
 * (n_slots-1)*S + H complex64 samples (interleaved re,im).  first_slot is the
 * call index of the first window (the reference's d_cumulative_count / S).
 * Host memory (pageable or pinned) for btb200_process, device memory for
 * btb200_process_device.
 */
BTB200_API int  btb200_process(btb200_ctx *ctx, const float *iq, size_t n_samples,
                    uint64_t first_slot, uint32_t n_slots, btb200_hits *out);
BTB200_API int  btb200_process_device(btb200_ctx *ctx, const float *d_iq, size_t n_samples,
                           uint64_t first_slot, uint32_t n_slots, btb200_hits *out);

/*
 * One multi_hopper work() call (lib/multi_hopper_impl.cc:76-209), BTB200_MM_CHAINED only: process the
 * classic channels first_channel .. first_channel+n_channels-1 of ONE window in ascending order on the
 * shared clock-recovery state -- channel_samples, check_snr, channel_symbols and a single sniff_ac over
 * min(n_symbols-68, 625) lags per channel -- and stop after the first channel whose packet carries
 * stop_lap and has a packet header (the reference's `break`, multi_hopper_impl.cc:109-133; pass
 * 0xffffffff for "never", as hopalong() does with its single channel).  Channels that are not reached keep
 * their rotator and clock-recovery state untouched.  res[i] describes channel first_channel+i; symbols
 * receives, per channel with an access code, the symbols from the access code on (what the reference
 * hands to classic_packet::make).
 */
typedef struct btb200_chan_result {
  int32_t  channel;              /* classic channel number */
  int32_t  processed;            /* 0: not reached (after the break) */
  int32_t  pass;                 /* squelch decision */
  int32_t  n_symbols;            /* symbols produced by clock recovery */
  int32_t  ac_index;             /* first access code, -1 if none */
  uint32_t lap;
  double   snr;
  uint64_t sym_offset;           /* into the symbols arena */
  uint32_t sym_count;            /* min(n_symbols - ac_index, 3125) */
  uint32_t reserved;
} btb200_chan_result;
BTB200_API int btb200_process_channels(btb200_ctx *ctx, const float *iq, size_t n_samples, uint64_t slot,
                                       int32_t first_channel, int32_t n_channels, uint32_t stop_lap,
                                       btb200_chan_result *res, uint8_t *symbols, size_t symbols_cap);

/* split form for overlap / benchmarking.  btb200_process == submit + collect.
 *   submit         copies the input (own copy stream) and enqueues the batch's kernels on the device's compute
 *                  stream, which all contexts of a device share: batches execute in the order they were enqueued;
 *   collect_begin  (optional) waits for the batch's access-code search, fetches the hit list and enqueues the
 *                  deferred work (exact noise FIR + energies of the hit windows) WITHOUT waiting for it;
 *   collect        = collect_begin if it was not called, then waits and writes the hits.
 * A streaming caller with contexts A, B keeps the device busy and every copy hidden with
 *   submit(A,0) submit(B,1) | begin(A) submit(A',2) collect(A) | begin(B) submit(B',3) collect(B) | ...
 * (A' = a third context, or A again after collect(A) -- bench.py's end-to-end loop). */
BTB200_API int  btb200_submit(btb200_ctx *ctx, const float *iq, int iq_on_device, size_t n_samples,
                   uint64_t first_slot, uint32_t n_slots);
/* Window mask for the NEXT submit/process of a BTB200_MM_STATELESS context: mask[slot_in_batch * n_channels + chan_index]
 * != 0 selects the channel-windows to demodulate and search; the others are skipped as if squelched.  This is the batched
 * form of multi_hopper's hop-along phase (lib/multi_hopper_impl.cc:152-209: once CLK1-27 is known the hop channel of
 * every slot is known in advance, one channel per slot).  The mask is consumed by that one call. */
BTB200_API int  btb200_set_window_mask(btb200_ctx *ctx, const uint8_t *mask, uint32_t n_slots);

/* int16 input: interleaved (re, im) int16 pairs, what the reference's flowgraph feeds through
 * interleaved_short_to_complex in front of the block (apps/btrx:141-159).  Same batches as btb200_submit /
 * btb200_process with half the host-to-device bytes; the conversion to complex64 is exact, so results are identical
 * to feeding the converted samples. */
BTB200_API int  btb200_submit_i16(btb200_ctx *ctx, const int16_t *iq, int iq_on_device, size_t n_samples,
                       uint64_t first_slot, uint32_t n_slots);
BTB200_API int  btb200_process_i16(btb200_ctx *ctx, const int16_t *iq, size_t n_samples,
                        uint64_t first_slot, uint32_t n_slots, btb200_hits *out);
BTB200_API int  btb200_collect_begin(btb200_ctx *ctx);
BTB200_API int  btb200_collect(btb200_ctx *ctx, btb200_hits *out);

/* pinned host buffers for the H2D path */
BTB200_API int  btb200_host_alloc(void **ptr, size_t bytes);
BTB200_API void btb200_host_free(void *ptr);

/* M&M state carried between calls in BTB200_MM_CHAINED: {mu, omega, last_sample} */
BTB200_API int  btb200_get_mm_state(const btb200_ctx *ctx, float mm[3]);
BTB200_API int  btb200_set_mm_state(btb200_ctx *ctx, const float mm[3]);
/* restart the stream (slot counter, rotators, M&M) */
BTB200_API int  btb200_reset(btb200_ctx *ctx);

/* debug taps of the LAST processed batch (parity tests).  slot_in_batch and
 * chan_index (channel - channel_low) select a channel-window; dst receives up
 * to cap bytes; returns bytes written or a negative error. */
enum {
  BTB200_STAGE_ENERGY   = 1,     /* f64  on-channel mean |y|^2               */
  BTB200_STAGE_NOISE    = 2,     /* f64  off-channel mean |y|^2              */
  BTB200_STAGE_SNR      = 3,     /* f64  10*log10(on/off) (host libm)        */
  BTB200_STAGE_PASS     = 4,     /* i32  squelch decision                    */
  BTB200_STAGE_NSYM     = 5,     /* i32  symbols produced                    */
  BTB200_STAGE_BITS     = 6,     /* u8[nsym] sliced symbols                  */
  BTB200_STAGE_DDC      = 7,     /* c64[ddc_out_per_window] rotated DDC out  */
  BTB200_STAGE_DEMOD    = 8,     /* f32[ddc_out-1] (needs keep_stages)       */
  BTB200_STAGE_SOFT     = 9,     /* f32[nsym]      (needs keep_stages)       */
  BTB200_STAGE_NOISE_FAST = 10,  /* f64  fast estimate of the off-channel energy (BTB200_SNR_FAST_GUARDED) */
  BTB200_STAGE_CHAN_TAPS  = 20,  /* c64[Nc] reversed band-pass taps of chan_index */
  BTB200_STAGE_NOISE_TAPS = 21,  /* c64[Nn]                                  */
  BTB200_STAGE_MMSE_TABLE = 22,  /* f32[129*8]                               */
  BTB200_STAGE_ATAN_TABLE = 23,  /* f32[257]                                 */
  BTB200_STAGE_AC_LUT     = 24   /* u64[3*256+1] affine sync-word tables + constant */
};
BTB200_API int64_t btb200_get_stage(btb200_ctx *ctx, int stage, uint32_t slot_in_batch,
                         uint32_t chan_index, void *dst, size_t cap);

/* Debug entry for known-answer tests of the access-code search kernel alone (classic_packet::sniff_ac
 * lib/packet_impl.cc:247-268 + check_ac :471-510 + acgen :309-364, as run by the loop of
 * lib/multi_sniffer_impl.cc:107-126): the caller's symbol stream (one symbol per byte, air order) is cut into windows
 * that start every `stride` symbols and hold stride + 72 symbols; each window goes through the SAME kernel the receive
 * path uses (lags 0 .. min(len - 68, 625) - 1, first hit, skip 68, search on).  Hits come back with slot = window index,
 * channel = 0, offset = lag inside the window, lap, n_symbols = window length - offset; no symbols, snr = 0.
 * stride <= 625.  BR search only. */
BTB200_API int  btb200_search_bits(btb200_ctx *ctx, const uint8_t *symbols, size_t n_symbols, uint32_t stride,
                        btb200_hits *out);

/* Hop reversal, candidate search (lib/piconet_impl.cc:96-129 walks a 2^27-entry table the reference has to generate
 * first, :214-255): every CLK1-27 value = clock6 (mod 64) whose basic-hop channel -- folded by the aliasing receiver when
 * `aliased` -- equals first_channel, ascending, for address28 = (UAP << 24 | LAP) & 0xfffffff.  The hop selection kernel
 * is evaluated per clock value on the GPU, no table.  *count receives the number of candidates (out holds min(count, cap)). */
/* the hop selection kernel itself, evaluated on the host (same source as the device code): channel 0..78 at CLK1-27 = clock */
BTB200_API int  btb200_hop_select(uint32_t address28, int afh, uint32_t clock);
BTB200_API int  btb200_hop_candidates(int device, uint32_t address28, int afh, int aliased, uint32_t clock6, int first_channel,
                           uint32_t *out, uint32_t cap, uint32_t *count);

/* device-side stopwatch on the context's compute stream (CUDA events): start() records now; stop() records, waits and
 * returns the milliseconds in between -- brackets any sequence of submit/collect calls of the contexts of one device */
BTB200_API int  btb200_timer_start(btb200_ctx *ctx);
BTB200_API int  btb200_timer_stop(btb200_ctx *ctx, float *ms);

/* timing of the last batch, milliseconds, CUDA events on the ctx stream:
 * [0] H2D copy, [1] channel FIR, [2] noise FIR, [3] energy/squelch,
 * [4] demod+clock recovery, [5] access-code search, [6] D2H + host, [7] total device.
 * BTB200_DDC_POLYPHASE: [1] channelizer (with demod and energies), [2] noise estimate, [3] resume of the clock recovery
 * of the windows with hits, [4] clock recovery of the searchable prefix, [5] search, [6] hit list + gather + copies */
BTB200_API int  btb200_last_timing(const btb200_ctx *ctx, float ms[8]);
/* number of kernel launches issued by this ctx so far */
BTB200_API uint64_t btb200_launch_count(const btb200_ctx *ctx);

BTB200_API const char *btb200_strerror(int err);
BTB200_API const char *btb200_last_error(const btb200_ctx *ctx);   /* detail of the last CUDA failure */
BTB200_API const char *btb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BTB200_H */
