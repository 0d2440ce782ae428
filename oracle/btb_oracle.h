/*
 * oracle/btb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of gr-bluetooth's multi-channel receive hot path (SURVEY.md
 * section 8a rows a1-a15).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  Pinned against the verbatim reference build
 * (oracle/_ref/btref, which reproduces the survey's stdout digests) by
 * tests/test_oracle_ref.py and against the committed goldens in tests/golden/.
 */
#ifndef BTB_ORACLE_H
#define BTB_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double fs, fc, snr_db;
  int    extra_symbols;      /* 3125 (sniffer/hopper) or 68 (multi_LAP) */
  int    S;                  /* samples per slot                      multi_block.cc:56-58 */
  int    H;                  /* history() = window length in samples  multi_block.cc:99-119,299-303 */
  int    D;                  /* DDC decimation                        multi_block.cc:82 */
  int    Nc, Nn;             /* channel / noise prototype tap counts  multi_block.cc:65-79 */
  int    fcs, fns;           /* first channel / noise sample          multi_block.cc:104-114 */
  int    ch_lo, ch_hi, nch;  /* classic channels covered              multi_block.cc:306-342 */
  int    n_ddc;              /* DDC outputs per window                multi_block.cc:194-200 */
  int    n_noise;            /* noise DDC outputs per window          multi_block.cc:269 */
  float  demod_gain;         /* multi_block.cc:88 */
  float  omega_mid;          /* samples per symbol after decimation   multi_block.cc:83,94,96 */
} btbo_info;

typedef struct btbo_plan btbo_plan;
typedef struct btbo_state btbo_state;

btbo_plan  *btbo_plan_create(double fs, double fc, double snr_db, int extra_symbols);
void        btbo_plan_free(btbo_plan *);
void        btbo_plan_info(const btbo_plan *, btbo_info *out);
/* tables (pointers stay valid for the plan's lifetime) */
const float *btbo_chan_proto(const btbo_plan *);             /* Nc floats */
const float *btbo_noise_proto(const btbo_plan *);            /* Nn floats */
const float *btbo_chan_rtaps(const btbo_plan *, int chi);    /* Nc complex (re,im), reversed */
const float *btbo_noise_rtaps(const btbo_plan *, int chi);   /* Nn complex, reversed */
void         btbo_rot_incr(const btbo_plan *, int chi, int noise, float out[2]);
const float *btbo_mmse_table(const btbo_plan *);             /* 129*8 floats */
const float *btbo_atan_table(const btbo_plan *);             /* 257 floats */

btbo_state *btbo_state_create(const btbo_plan *);
void        btbo_state_free(btbo_state *);
void        btbo_state_get_mm(const btbo_state *, float mm[3]);   /* mu, omega, last_sample */
void        btbo_state_set_mm(btbo_state *, const float mm[3]);

/* ---- stage functions (each restates one reference function) ------------- */
/* packet_impl.cc:309-364 acgen: 72 air-order symbols (0/1) for a LAP */
void btbo_acgen_bits(uint32_t lap, uint8_t ac[72]);
/* same, packed MSB-first into the 9 bytes the reference's acgen() returns */
void btbo_acgen_bytes(uint32_t lap, uint8_t out[9]);
/* packet_impl.cc:471-510 */
int  btbo_check_ac(const uint8_t *stream, uint32_t lap);
/* libbtbb-style access-code test of one lag / search over a stream (btbb_find_ac as multi_LAP_impl.cc:93 and
 * multi_UAP_impl.cc:95 call it; restated from libbtbb's published algorithm by brute force -- PARITY UNPINNED).
 * lap = 0xffffffff: LAP_ANY. */
int  btbo_bch_lag(const uint8_t *stream, int max_ac_errors, uint32_t lap, uint32_t *lap_out, int *n_err);
int  btbo_find_ac_bch(const uint8_t *stream, int search_length, uint32_t lap, int max_ac_errors, uint32_t *lap_out, int *n_err);
/* packet_impl.cc:247-268; returns lag or -1 */
int  btbo_sniff_ac(const uint8_t *stream, int stream_length);
/* packet_impl.cc:1452-1527; returns lag or -1 */
int  btbo_sniff_aa(const uint8_t *stream, int stream_length, double freq);
/* the LUTs, regenerated in closed form; which: 0 classic
 * PREAMBLE(32) 1 BARKER(128) 2 le PREAMBLE(512) 3-6 le AA byte 0-3 (256)
 * 7 ACCESS_HDR_LSB 8 ACCESS_HDR_MSB 9 DATA_HDR_LSB 10 DATA_HDR_MSB (256) */
int  btbo_lut(int which, uint8_t *dst, int cap);
int  btbo_le_index(double freq);                                  /* packet_impl.cc:1285-1314 */
/* multi_block.cc:158-168 (out[0] pinned to 0.0f) ; n = n_ddc-1 outputs */
void btbo_demod(const btbo_plan *, const float *ddc_out /*complex*/, float *out, int n);
/* multi_block.cc:128-155 ; mm = {mu, omega, last}; returns symbols produced */
int  btbo_mm_cr(const btbo_plan *, float mm[3], const float *in, int nin, float *out, int nout);

/* ---- one work() call (multi_sniffer_impl.cc:82-166) --------------------- */
typedef struct {
  int32_t  slot;        /* work() call index == clkn */
  int16_t  channel;     /* classic channel number 0..78 */
  int16_t  kind;        /* 0 = BR access code (sniff_ac), 1 = LE (sniff_aa) */
  int32_t  offset;      /* index into the window's symbol array where the packet starts */
  int32_t  len;         /* "len - i" handed to ac()/aa() */
  uint32_t lap;         /* BR: LAP from symbols 38..61; LE: AA from symbols 8..39 (not de-whitened) */
  double   snr;
} btbo_hit;

typedef struct {
  /* per channel, index chi = channel - ch_lo; any pointer may be NULL */
  double  *energy;      /* [nch] */
  double  *noise;       /* [nch] */
  double  *snr;         /* [nch] */
  int32_t *pass;        /* [nch] */
  int32_t *nsym;        /* [nch] symbols produced (0 if squelched) */
  uint8_t *bits;        /* [nch][H] */
  float   *ddc;         /* [nch][n_ddc] complex */
  float   *demod;       /* [nch][n_ddc-1] */
  float   *soft;        /* [nch][n_ddc-1] */
} btbo_debug;

/* window = H complex samples (interleaved f32). flags bit0: stateless (M&M and
 * rotator reset per channel-window).  Appends to hits[*nhits], cap = hits_cap.
 * Returns 0, or -1 on hit overflow. */
int btbo_window(const btbo_plan *, btbo_state *, const float *window, int slot, int flags,
                btbo_hit *hits, int hits_cap, int *nhits, const btbo_debug *dbg);

/* ---- one multi_hopper work() call (multi_hopper_impl.cc:76-209) ---------------------------------
 * Processes the listed channel indices IN ORDER with the chained state (channel_samples, check_snr,
 * channel_symbols, ONE sniff_ac over min(nsym-68, 625) lags), and stops after the first channel whose
 * packet carries stop_lap and has a header (the reference's `break`, multi_hopper_impl.cc:109-133);
 * stop_lap = 0xffffffff never stops (hopalong passes a single channel).  Channels after the stop are
 * not touched at all (rotators and clock recovery keep their state).  symbols: [n][H] bytes. */
typedef struct {
  int32_t chi;          /* channel index */
  int32_t processed;    /* 0: not reached (after the break) */
  int32_t pass;         /* squelch */
  int32_t nsym;
  int32_t ac_index;     /* first access code, -1 if none */
  uint32_t lap;
  double  snr;
} btbo_chan_result;
int btbo_window_list(const btbo_plan *, btbo_state *, const float *window, const int32_t *chis, int n,
                     uint32_t stop_lap, btbo_chan_result *res, uint8_t *symbols);
/* classic_packet_impl::header_present, packet_impl.cc:1205-1242 */
int btbo_header_present(const uint8_t *symbols, int length);

/* Scheduler emulation over a sample array: calls
 * first_call .. first_call+num_calls-1 of a stream of n_total samples, of which
 * iq holds samples [iq_first, iq_first+iq_n) (zeros elsewhere).
 * bits_out (optional): [num_calls][nch][bits_stride] + nsym_out [num_calls][nch].
 * With flags bit0 (stateless) and threads > 1 the calls run under OpenMP. */
int btbo_run(const btbo_plan *, btbo_state *, const float *iq, long iq_first, long iq_n,
             long first_call, long num_calls, int flags, int threads,
             btbo_hit *hits, int hits_cap, int *nhits,
             uint8_t *bits_out, int bits_stride, int32_t *nsym_out,
             double *energy_out, double *noise_out);
#ifdef __cplusplus
}
#endif
#endif
