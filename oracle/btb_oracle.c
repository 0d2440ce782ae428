/*
 * oracle/btb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of gr-bluetooth's multi-channel receive hot path.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference).  GNU Radio arithmetic (third-party, absent) comes from
 * oracle/gr_arith.h.  Float rules: binary32, one rounding per operation
 * (-ffp-contract=off), ascending-index sums.
 *
 * PINNING: tests/test_oracle_ref.py runs this file against oracle/_ref/btref
 * (the reference's own lib/ *.cc compiled verbatim) on the four bundled
 * samples/ *.cfile and on channel37.dem: identical energies, bit streams and hit
 * lists; tests/test_oracle_golden.py checks it against the committed fixtures.
 */
#include "btb_oracle.h"
#include "gr_arith.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

/* include/gr_bluetooth/multi_block.h:47-60 */
#define SYMBOL_RATE 1000000
#define AC_SYMBOLS 68            /* SYMBOLS_PER_BASIC_RATE_SHORTENED_ACCESS_CODE */
#define LE_AA_SYMBOLS 40         /* SYMBOLS_PER_LOW_ENERGY_PREAMBLE_AA */
#define SLOT_SYMBOLS 625
#define BASE_FREQUENCY 2402000000.0
#define CHANNEL_WIDTH 1000000.0

struct btbo_plan {
  btbo_info info;
  float *chan_proto, *noise_proto;
  gra_fxlat *chan_ddc, *noise_ddc;        /* [nch]; state fields unused here (see btbo_state) */
  float mmse[GRA_MMSE_NSTEPS + 1][GRA_MMSE_NTAPS];
  float atan_t[257];
  /* M&M constants, multi_block.cc:91-96 */
  float gain_mu, gain_omega, omega_rel_limit, mu0;
};

struct btbo_state {
  float mu, omega, last_sample;           /* include/gr_bluetooth/multi_block.h:86-93 */
  gra_c32 *chan_phase, *noise_phase;      /* rotator state per DDC object (A.3) */
  unsigned *chan_counter, *noise_counter;
  int nch;
};

/* ------------------------------------------------------------------------- */
/* multi_block::multi_block + set_channels + set_symbol_history               */
/* lib/multi_block.cc:40-120, 299-342                                         */
btbo_plan *btbo_plan_create(double fs, double fc, double snr_db, int extra_symbols)
{
  btbo_plan *p = (btbo_plan *)calloc(1, sizeof *p);
  btbo_info *I = &p->info;
  I->fs = fs; I->fc = fc; I->snr_db = snr_db; I->extra_symbols = extra_symbols;
  double sps = fs / SYMBOL_RATE;                         /* :56 */
  double samples_per_slot = (int)SLOT_SYMBOLS * sps;     /* :58 (int cast binds to the constant) */
  I->S = (int)samples_per_slot;
  int history_required = (int)1 * samples_per_slot;      /* :59 */

  I->Nc = gra_lowpass_ntaps(fs, 300000);                 /* :63-69 */
  p->chan_proto = (float *)malloc(sizeof(float) * (size_t)I->Nc);
  gra_lowpass(1, fs, 500000, 300000, p->chan_proto);
  I->Nn = gra_lowpass_ntaps(fs, 10000);                  /* :71-79 */
  p->noise_proto = (float *)malloc(sizeof(float) * (size_t)I->Nn);
  gra_lowpass(1, fs, 22500, 10000, p->noise_proto);

  I->D = (int)sps / 2;                                   /* :82 */
  double chan_sps = sps / I->D;                          /* :83 */

  /* set_channels(), :306-342 */
  double center = (fc - BASE_FREQUENCY) / CHANNEL_WIDTH;
  double bw = fs / CHANNEL_WIDTH;
  double low_edge = center - bw / 2, high_edge = center + bw / 2;
  double min_w = 0.9;
  int lo = (int)(low_edge + min_w / 2 + 1);
  if (lo < 0) lo = 0;
  int hi = (int)(high_edge - min_w / 2);
  if (hi > 78) hi = 78;
  I->ch_lo = lo; I->ch_hi = hi; I->nch = hi - lo + 1;
  if (I->nch < 0) I->nch = 0;
  p->chan_ddc = (gra_fxlat *)calloc((size_t)(I->nch > 0 ? I->nch : 1), sizeof(gra_fxlat));
  p->noise_ddc = (gra_fxlat *)calloc((size_t)(I->nch > 0 ? I->nch : 1), sizeof(gra_fxlat));
  for (int ch = lo; ch <= hi; ch++) {
    double freq = BASE_FREQUENCY + ch * CHANNEL_WIDTH;   /* channel_abs_freq, :352-355 */
    gra_fxlat_init(&p->chan_ddc[ch - lo], I->D, p->chan_proto, I->Nc, freq - fc, fs);
    gra_fxlat_init(&p->noise_ddc[ch - lo], I->D, p->noise_proto, I->Nn, freq + 790000.0 - fc, fs);
  }

  I->demod_gain = (float)(chan_sps / M_PI_2);            /* :88 */
  p->gain_mu = 0.175f;                                   /* :91-96 */
  p->mu0 = 0.32f;
  p->omega_rel_limit = 0.005f;
  I->omega_mid = (float)chan_sps;
  p->gain_omega = (float)(.25 * p->gain_mu * p->gain_mu);   /* float*float promoted: .25*g*g in double */
  gra_mmse_table(p->mmse);
  gra_atan_table(p->atan_t);

  /* history, :99-119 */
  int channel_history = (int)(I->Nc + I->D * GRA_MMSE_NTAPS);
  int noise_history = I->Nn;
  if (channel_history > noise_history) {
    history_required += channel_history;
    I->fcs = 0; I->fns = channel_history - noise_history;
  } else {
    history_required += noise_history;
    I->fns = 0; I->fcs = noise_history - channel_history;
  }
  /* set_symbol_history(extra), :299-303 */
  I->H = (int)(history_required + extra_symbols * sps);

  /* channel_samples, :194-200 */
  int ddc_samples = I->H - (I->Nc - 1) - I->fcs;
  I->n_ddc = gra_fxlat_ninput_to_noutput(&p->chan_ddc[0], ddc_samples);
  /* check_snr, :269 */
  I->n_noise = gra_fxlat_ninput_to_noutput(&p->noise_ddc[0], (int)samples_per_slot);
  return p;
}

void btbo_plan_free(btbo_plan *p)
{
  if (!p) return;
  for (int i = 0; i < p->info.nch; i++) { gra_fxlat_free(&p->chan_ddc[i]); gra_fxlat_free(&p->noise_ddc[i]); }
  free(p->chan_ddc); free(p->noise_ddc); free(p->chan_proto); free(p->noise_proto); free(p);
}

void btbo_plan_info(const btbo_plan *p, btbo_info *out) { *out = p->info; }
const float *btbo_chan_proto(const btbo_plan *p) { return p->chan_proto; }
const float *btbo_noise_proto(const btbo_plan *p) { return p->noise_proto; }
const float *btbo_chan_rtaps(const btbo_plan *p, int chi) { return (const float *)p->chan_ddc[chi].rtaps; }
const float *btbo_noise_rtaps(const btbo_plan *p, int chi) { return (const float *)p->noise_ddc[chi].rtaps; }
void btbo_rot_incr(const btbo_plan *p, int chi, int noise, float out[2])
{
  const gra_fxlat *f = noise ? &p->noise_ddc[chi] : &p->chan_ddc[chi];
  out[0] = f->incr.re; out[1] = f->incr.im;
}
const float *btbo_mmse_table(const btbo_plan *p) { return &p->mmse[0][0]; }
const float *btbo_atan_table(const btbo_plan *p) { return p->atan_t; }

btbo_state *btbo_state_create(const btbo_plan *p)
{
  btbo_state *s = (btbo_state *)calloc(1, sizeof *s);
  int n = p->info.nch > 0 ? p->info.nch : 1;
  s->nch = p->info.nch;
  s->mu = p->mu0; s->omega = p->info.omega_mid; s->last_sample = 0;   /* multi_block.cc:92,94,98 */
  s->chan_phase = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)n);
  s->noise_phase = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)n);
  s->chan_counter = (unsigned *)calloc((size_t)n, sizeof(unsigned));
  s->noise_counter = (unsigned *)calloc((size_t)n, sizeof(unsigned));
  for (int i = 0; i < n; i++) {
    s->chan_phase[i].re = 1; s->chan_phase[i].im = 0;
    s->noise_phase[i].re = 1; s->noise_phase[i].im = 0;
  }
  return s;
}
void btbo_state_free(btbo_state *s)
{
  if (!s) return;
  free(s->chan_phase); free(s->noise_phase); free(s->chan_counter); free(s->noise_counter); free(s);
}
void btbo_state_get_mm(const btbo_state *s, float mm[3]) { mm[0] = s->mu; mm[1] = s->omega; mm[2] = s->last_sample; }
void btbo_state_set_mm(btbo_state *s, const float mm[3]) { s->mu = mm[0]; s->omega = mm[1]; s->last_sample = mm[2]; }

/* ------------------------------------------------------------------------- */
/* Access-code generation, lib/packet_impl.cc:278-364                         */

/* PN sequence of acgen (:315) laid over the 72 access-code positions, MSB first */
static const uint8_t PN_BYTES[9] = { 0x03, 0xF2, 0xA3, 0x3D, 0xD6, 0x9B, 0x12, 0x1C, 0x10 };
/* generator polynomial of the (64,30) expurgated BCH code (:318) */
static const uint8_t BCH_G[35] = { 1,0,0,1,0,1,0,1,1,0,1,1,1,1,0,0,1,0,0,0,1,1,1,0,1,0,1,0,0,0,0,1,1,0,1 };

static inline int pn_bit(int pos) { return (PN_BYTES[pos >> 3] >> (7 - (pos & 7))) & 1; }

void btbo_acgen_bits(uint32_t lap, uint8_t ac[72])
{
  uint8_t info[30], cw[34];
  int a23 = (lap >> 23) & 1;
  /* positions 38..61 carry a0..a23 (air order = LSB first), :322-327 */
  for (int i = 0; i < 24; i++) info[i] = (lap >> i) & 1;
  /* 6-bit Barker extension chosen by the LAP MSB, :329-334 */
  static const uint8_t bark1[6] = { 1, 1, 0, 0, 1, 0 }, bark0[6] = { 0, 0, 1, 1, 0, 1 };
  for (int i = 0; i < 6; i++) info[24 + i] = a23 ? bark1[i] : bark0[i];
  /* scramble the information bits with the PN, :336-344 */
  uint8_t data[30];
  for (int i = 0; i < 30; i++) data[i] = info[i] ^ pn_bit(38 + i);
  /* systematic encoding = polynomial division, lfsr() :278-306 with length 64, k 30 */
  memset(cw, 0, sizeof cw);
  for (int i = 29; i >= 0; i--) {
    uint8_t fb = data[i] ^ cw[33];
    for (int j = 33; j > 0; j--) cw[j] = cw[j - 1] ^ (BCH_G[j] & fb);
    cw[0] = BCH_G[0] & fb;
  }
  /* parity into positions 4..37, de-scrambled by the same PN, :349-357 */
  for (int i = 0; i < 34; i++) ac[4 + i] = cw[i] ^ pn_bit(4 + i);
  for (int i = 0; i < 30; i++) ac[38 + i] = info[i];
  /* preamble 1010 / 0101 by the first sync-word bit, :359-363 */
  if (ac[4]) { ac[0] = 1; ac[1] = 0; ac[2] = 1; ac[3] = 0; }
  else       { ac[0] = 0; ac[1] = 1; ac[2] = 0; ac[3] = 1; }
  /* trailer follows the Barker code, :329-334 (0x2a / 0xd5 low nibbles) */
  if (a23) { ac[68] = 1; ac[69] = 0; ac[70] = 1; ac[71] = 0; }
  else     { ac[68] = 0; ac[69] = 1; ac[70] = 0; ac[71] = 1; }
}

void btbo_acgen_bytes(uint32_t lap, uint8_t out[9])
{
  uint8_t ac[72];
  btbo_acgen_bits(lap, ac);
  for (int b = 0; b < 9; b++) {
    uint8_t v = 0;
    for (int i = 0; i < 8; i++) v = (uint8_t)((v << 1) | ac[8 * b + i]);   /* convert_to_grformat is MSB first, :93-101 */
    out[b] = v;
  }
}

/* lib/packet_impl.cc:471-510: regenerate the AC for the received LAP, count
 * mismatches over the first 68 symbols, reject at the 7th. */
int btbo_check_ac(const uint8_t *stream, uint32_t lap)
{
  uint8_t ac[72];
  int biterrors = 0;
  btbo_acgen_bits(lap, ac);
  for (int i = 0; i < AC_SYMBOLS; i++) {
    if (ac[i] != stream[i]) biterrors++;
    if (biterrors >= 7) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------- */
/* libbtbb-style access-code search: what multi_LAP / multi_UAP call through  */
/* btbb_find_ac (lib/multi_LAP_impl.cc:93, lib/multi_UAP_impl.cc:95).         */
/*                                                                             */
/* PARITY UNPINNED.  libbtbb is an external, un-vendored dependency            */
/* (cmake/Modules/FindBTBB.cmake; no version pinned by the reference, the call */
/* sites use the API level btbb_init(max_ac_errors) / btbb_find_ac(stream,     */
/* search_length, lap, max_ac_errors, &pkt) of the 2014/2015 releases).  This  */
/* restates its published algorithm (bluetooth_packet.c) by BRUTE FORCE over   */
/* the code the reference's own acgen defines, independent of the product's    */
/* syndrome tables:                                                            */
/*   given LAP: Hamming distance of the 64 sync-word symbols to acgen(LAP)'s   */
/*              sync word <= max_ac_errors;                                    */
/*   LAP_ANY:   the 7 top symbols (LAP MSB + Barker) are replaced by the nearer */
/*              of their two valid patterns; the word is accepted if flipping  */
/*              at most max_ac_errors of sync-word bits 0..57 makes it a code   */
/*              word (parity regenerated from the information bits agrees);    */
/*              LAP and error count come from the corrected word (Barker       */
/*              corrections are not counted).                                  */
/* The 68-symbol window is aligned like sniff_ac's (symbols 4..67 = sync word). */

/* parity symbols (ac positions 4..37) of 30 information bits, acgen's encoder (:336-357) with the information given */
static void bch_parity(const uint8_t info[30], uint8_t par[34])
{
  uint8_t cw[34];
  memset(cw, 0, sizeof cw);
  for (int i = 29; i >= 0; i--) {
    uint8_t fb = (uint8_t)((info[i] ^ pn_bit(38 + i)) ^ cw[33]);
    for (int j = 33; j > 0; j--) cw[j] = cw[j - 1] ^ (BCH_G[j] & fb);
    cw[0] = BCH_G[0] & fb;
  }
  for (int i = 0; i < 34; i++) par[i] = cw[i] ^ pn_bit(4 + i);
}

static int bch_is_codeword(const uint8_t sw[64])
{
  uint8_t par[34];
  bch_parity(sw + 34, par);
  return memcmp(par, sw, 34) == 0;
}

/* stream: >= 68 symbols at the lag under test.  lap = 0xffffffff: LAP_ANY.  Returns 1 and the LAP / corrected-bit
 * count when libbtbb's test accepts the lag. */
int btbo_bch_lag(const uint8_t *stream, int max_ac_errors, uint32_t lap, uint32_t *lap_out, int *n_err)
{
  uint8_t sw[64];
  for (int i = 0; i < 64; i++) sw[i] = stream[4 + i] & 1;
  if (lap != 0xffffffffu) {
    uint8_t ac[72];
    int d = 0;
    btbo_acgen_bits(lap, ac);
    for (int i = 0; i < 64; i++) d += ac[4 + i] != sw[i];
    *lap_out = lap; *n_err = d;
    return d <= max_ac_errors;
  }
  /* LAP MSB + Barker sequence: the nearer of {0,0,0,1,1,0,1} (a23 = 0) and its complement */
  static const uint8_t top0[7] = { 0, 0, 0, 1, 1, 0, 1 };
  int d0 = 0;
  for (int i = 0; i < 7; i++) d0 += sw[57 + i] != top0[i];
  for (int i = 0; i < 7; i++) sw[57 + i] = (d0 < 7 - d0) ? top0[i] : (uint8_t)!top0[i];
  for (int w = 0; w <= max_ac_errors && w <= 2; w++) {
    if (w == 0) { if (bch_is_codeword(sw)) goto found; continue; }
    for (int i = 0; i < 58; i++) {
      sw[i] ^= 1;
      if (w == 1) { if (bch_is_codeword(sw)) { *n_err = 1; goto found_e; } }
      else
        for (int j = i + 1; j < 58; j++) {
          sw[j] ^= 1;
          if (bch_is_codeword(sw)) { *n_err = 2; goto found_e; }
          sw[j] ^= 1;
        }
      sw[i] ^= 1;
    }
  }
  return 0;
found:
  *n_err = 0;
found_e:
  *lap_out = 0;
  for (int i = 0; i < 24; i++) *lap_out |= (uint32_t)sw[34 + i] << i;
  return 1;
}

/* btbb_find_ac over a symbol stream: first accepted lag < search_length, or -1 */
int btbo_find_ac_bch(const uint8_t *stream, int search_length, uint32_t lap, int max_ac_errors, uint32_t *lap_out, int *n_err)
{
  for (int count = 0; count < search_length; count++)
    if (btbo_bch_lag(stream + count, max_ac_errors, lap, lap_out, n_err)) return count;
  return -1;
}

static inline uint32_t air_to_host(const uint8_t *air, int bits)   /* :104-136 */
{
  uint32_t v = 0;
  for (int i = 0; i < bits; i++) v |= (uint32_t)(air[i] & 1) << i;
  return v;
}

static inline int pop8(unsigned v) { return __builtin_popcount(v); }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* Closed forms of the reference's LUTs; verified entry
 * by entry against `btref tables` in tests/test_oracle_ref.py. */
static int lut_classic_preamble(unsigned v) { return imin(pop8(v ^ 0x0A), pop8(v ^ 0x15)); }   /* packet_impl.cc:191-193 */
static int lut_barker(unsigned v)           { return imin(pop8(v ^ 39), pop8(v ^ 88)); }       /* :195-200 */
static int lut_le_preamble(unsigned v)      { return imin(pop8(v ^ 0x0AA), pop8(v ^ 0x155)); } /* :1316-1325 */
static const uint8_t LE_ADV_AA[4] = { 0xD6, 0xBE, 0x89, 0x8E };                                /* 0x8E89BED6 LSB first */
static int nearest(unsigned v, int (*valid)(unsigned))
{
  int best = 9;
  for (unsigned c = 0; c < 256; c++) if (valid(c)) best = imin(best, pop8(v ^ c));
  return best;
}
static int v_acc_lsb(unsigned c)  { return ((c & 0x3f) <= 6) && (((c >> 6) == 0) || ((c >> 6) == 3)); }
static int v_acc_msb(unsigned c)  { return c >= 0x06 && c <= 0x24; }
static int v_data_lsb(unsigned c) { return c < 0x20 && (c & 3) != 0; }
static int v_data_msb(unsigned c) { return c <= 0x1F; }

int btbo_lut(int which, uint8_t *dst, int cap)
{
  int n = (which == 0) ? 32 : (which == 1) ? 128 : (which == 2) ? 512 : 256;
  if (cap < n) return -1;
  for (int v = 0; v < n; v++) {
    int d;
    switch (which) {
    case 0: d = lut_classic_preamble((unsigned)v); break;
    case 1: d = lut_barker((unsigned)v); break;
    case 2: d = lut_le_preamble((unsigned)v); break;
    case 3: case 4: case 5: case 6: d = pop8((unsigned)v ^ LE_ADV_AA[which - 3]); break;
    case 7: d = nearest((unsigned)v, v_acc_lsb); break;
    case 8: d = nearest((unsigned)v, v_acc_msb); break;
    case 9: d = nearest((unsigned)v, v_data_lsb); break;
    case 10: d = nearest((unsigned)v, v_data_msb); break;
    default: return -1;
    }
    dst[v] = (uint8_t)d;
  }
  return n;
}

/* lib/packet_impl.cc:247-268 */
int btbo_sniff_ac(const uint8_t *stream, int stream_length)
{
  const int max_distance = 2;
  for (int count = 0; count < stream_length; count++) {
    const uint8_t *sym = stream + count;
    unsigned preamble = air_to_host(sym, 5);
    unsigned barker = air_to_host(sym + 61, 7);
    if (lut_classic_preamble(preamble) + lut_barker(barker) <= max_distance) {
      uint32_t lap = air_to_host(sym + 38, 24);
      if (btbo_check_ac(sym, lap)) return count;
    }
  }
  return -1;
}

/* whitening sequence shared by BR and LE, regenerated from its LFSR
 * (x^7 + x^4 + 1); equals packet::WHITENING_DATA (packet_impl.cc:84-90),
 * checked against `btref tables`. */
static uint8_t g_white[127];
static int g_white_ready = 0;
static void white_init(void)
{
  if (g_white_ready) return;
  /* register bits r0..r6, output r6... derived so that the 127-cycle matches the table start 1,1,1,0,0,0,1 */
  uint8_t r[7] = { 1, 1, 1, 1, 1, 1, 1 };
  for (int i = 0; i < 127; i++) {
    g_white[i] = r[6];
    uint8_t fb = r[6];
    r[6] = r[5]; r[5] = r[4]; r[4] = r[3] ^ fb; r[3] = r[2]; r[2] = r[1]; r[1] = r[0]; r[0] = fb;
  }
  g_white_ready = 1;
}

/* le_packet::INDICES (packet_impl.cc:1446-1450): position in the whitening
 * cycle at which the LFSR state equals (1, channel index bits) */
static uint8_t g_le_idx[40];
static int g_le_idx_ready = 0;
static void le_idx_init(void)
{
  if (g_le_idx_ready) return;
  white_init();
  /* state at cycle position i is the next 7 outputs reversed; the LE rule
   * seeds position0 = 1, positions 1..6 = channel index MSB..LSB.  Find, for
   * each index, the cycle offset whose upcoming outputs correspond to it. */
  for (int idx = 0; idx < 40; idx++) {
    uint8_t r[7];
    r[0] = 1;
    for (int b = 0; b < 6; b++) r[1 + b] = (idx >> (5 - b)) & 1;
    /* run this register and match its first 7 outputs against the cycle */
    uint8_t outs[7], q[7];
    memcpy(q, r, 7);
    for (int i = 0; i < 7; i++) {
      outs[i] = q[6];
      uint8_t fb = q[6];
      q[6] = q[5]; q[5] = q[4]; q[4] = q[3] ^ fb; q[3] = q[2]; q[2] = q[1]; q[1] = q[0]; q[0] = fb;
    }
    for (int off = 0; off < 127; off++) {
      int ok = 1;
      for (int i = 0; i < 7 && ok; i++) ok = (g_white[(off + i) % 127] == outs[i]);
      if (ok) { g_le_idx[idx] = (uint8_t)off; break; }
    }
  }
  g_le_idx_ready = 1;
}

/* freq2chan / chan2index / freq2index, lib/packet_impl.cc:1285-1314 */
int btbo_le_index(double freq)
{
  int chan = -1;
  if (freq >= 2402000000.0 && freq <= 2480000000.0)
    if (fmod(freq, 2000000.0) < 5000.0) chan = (int)((freq - 2402000000.0) / 2000000.0);
  if (chan < 0 || chan > 39) return -1;
  if (chan == 0) return 37;
  if (chan == 12) return 38;
  if (chan == 39) return 39;
  return (chan < 12) ? chan - 1 : chan - 2;
}

/* lib/packet_impl.cc:1452-1527 (the diagnostic printf at :1500-1512 is not restated) */
int btbo_sniff_aa(const uint8_t *stream, int stream_length, double freq)
{
  int index = btbo_le_index(freq);
  if (index < 0) return -1;
  le_idx_init();
  int adv = index >= 37;
  for (int count = 0; count < stream_length; count++) {
    const uint8_t *sym = stream + count;
    unsigned preamble = air_to_host(sym, 9);
    uint8_t hbuf[16];
    unsigned wi = g_le_idx[index];
    for (int hi = 0; hi < 16; hi++, wi = (wi + 1) % 127) hbuf[hi] = (sym[hi + 40] ^ g_white[wi]) & 1;
    unsigned h_lsb = air_to_host(hbuf, 8), h_msb = air_to_host(hbuf + 8, 8);
    int distance = lut_le_preamble(preamble);
    if (adv) distance += nearest(h_lsb, v_acc_lsb) + nearest(h_msb, v_acc_msb);
    else     distance += nearest(h_lsb, v_data_lsb) + nearest(h_msb, v_data_msb);
    int max_distance = 0;
    if (adv) {
      for (int k = 0; k < 4; k++) distance += pop8(air_to_host(sym + 8 + 8 * k, 8) ^ LE_ADV_AA[k]);
      max_distance += 2;
    }
    if (distance <= max_distance) return count;
  }
  return -1;
}

/* ------------------------------------------------------------------------- */
/* demod / mm_cr / slicer, lib/multi_block.cc:122-178                         */

void btbo_demod(const btbo_plan *p, const float *ddc_out, float *out, int n)
{
  const gra_c32 *x = (const gra_c32 *)ddc_out;
  if (n > 0) out[0] = 0.0f;                 /* never written by the reference (:164); pinned to 0 */
  for (int i = 1; i < n; i++) {
    /* in[i] * conj(in[i-1]) */
    float a = x[i].re, b = x[i].im, c = x[i - 1].re, d = -x[i - 1].im;
    float pr = a * c - b * d;
    float pi = a * d + b * c;
    out[i] = p->info.demod_gain * gra_fast_atan2f(p->atan_t, pi, pr);
  }
}

static inline float slice_pm1(float x) { return (x < 0) ? -1.0f : 1.0f; }   /* :122-125 */

int btbo_mm_cr(const btbo_plan *p, float mm[3], const float *in, int nin, float *out, int nout)
{
  unsigned ii = 0;
  int oo = 0;
  unsigned ni = (unsigned)(nin - GRA_MMSE_NTAPS);       /* :133 */
  float mu = mm[0], omega = mm[1], last = mm[2];
  const float omega_mid = p->info.omega_mid;
  while (oo < nout && ii < ni) {
    int bad = 0;
    out[oo] = gra_mmse_interpolate(p->mmse, &in[ii], mu, &bad);
    float mm_val = slice_pm1(last) * out[oo] - slice_pm1(out[oo]) * last;
    last = out[oo];
    omega = omega + p->gain_omega * mm_val;
    omega = omega_mid + gra_branchless_clip(omega - omega_mid, p->omega_rel_limit);
    mu = mu + (omega + p->gain_mu * mm_val);
    double fl = floor((double)mu);
    ii += (unsigned)(int)fl;
    mu = (float)((double)mu - fl);
    oo++;
  }
  mm[0] = mu; mm[1] = omega; mm[2] = last;
  return oo;
}

/* ------------------------------------------------------------------------- */
/* one work() call, lib/multi_sniffer_impl.cc:82-166                           */

int btbo_window(const btbo_plan *p, btbo_state *st, const float *window, int slot, int flags,
                btbo_hit *hits, int hits_cap, int *nhits, const btbo_debug *dbg)
{
  const btbo_info *I = &p->info;
  const gra_c32 *in = (const gra_c32 *)window;
  const int stateless = flags & 1;
  int rc = 0;
  gra_c32 *ddc_out = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)(I->n_ddc + 1));
  gra_c32 *nz_out = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)(I->n_noise + 1));
  float *demod_out = (float *)malloc(sizeof(float) * (size_t)(I->n_ddc + 1));
  float *cr_out = (float *)malloc(sizeof(float) * (size_t)(I->n_ddc + 1));
  uint8_t *symbols = (uint8_t *)malloc((size_t)I->H + 8);

  for (int chi = 0; chi < I->nch; chi++) {
    int ch = I->ch_lo + chi;
    double freq = BASE_FREQUENCY + ch * CHANNEL_WIDTH;

    /* channel_samples, lib/multi_block.cc:180-228 */
    gra_fxlat cf = p->chan_ddc[chi];
    if (stateless) { cf.phase.re = 1; cf.phase.im = 0; cf.counter = 0; }
    else { cf.phase = st->chan_phase[chi]; cf.counter = st->chan_counter[chi]; }
    gra_fxlat_work(&cf, I->n_ddc, in + I->fcs, ddc_out);
    if (!stateless) { st->chan_phase[chi] = cf.phase; st->chan_counter[chi] = cf.counter; }
    double energy = 0.0;
    for (int i = 0; i < I->n_ddc; i++) energy += gra_mag2(ddc_out[i]);
    energy /= I->n_ddc;

    /* check_snr, lib/multi_block.cc:253-296 */
    gra_fxlat nf = p->noise_ddc[chi];
    if (stateless) { nf.phase.re = 1; nf.phase.im = 0; nf.counter = 0; }
    else { nf.phase = st->noise_phase[chi]; nf.counter = st->noise_counter[chi]; }
    gra_fxlat_work(&nf, I->n_noise, in + I->fns, nz_out);
    if (!stateless) { st->noise_phase[chi] = nf.phase; st->noise_counter[chi] = nf.counter; }
    double off = 0.0;
    for (int i = 0; i < I->n_noise; i++) off += gra_mag2(nz_out[i]);
    off /= I->n_noise;
    double snr = 10.0 * log10(energy / off);
    int pass = (snr >= I->snr_db);

    if (dbg) {
      if (dbg->energy) dbg->energy[chi] = energy;
      if (dbg->noise) dbg->noise[chi] = off;
      if (dbg->snr) dbg->snr[chi] = snr;
      if (dbg->pass) dbg->pass[chi] = pass;
      if (dbg->nsym) dbg->nsym[chi] = 0;
      if (dbg->ddc) memcpy(dbg->ddc + (size_t)chi * 2 * I->n_ddc, ddc_out, sizeof(gra_c32) * (size_t)I->n_ddc);
    }
    if (!pass) continue;

    /* channel_symbols, lib/multi_block.cc:230-251 */
    int n_demod = I->n_ddc - 1;
    btbo_demod(p, (const float *)ddc_out, demod_out, n_demod);
    float mm[3];
    if (stateless) { mm[0] = p->mu0; mm[1] = I->omega_mid; mm[2] = 0; }
    else btbo_state_get_mm(st, mm);
    int len = btbo_mm_cr(p, mm, demod_out, n_demod, cr_out, n_demod);
    if (!stateless) btbo_state_set_mm(st, mm);
    for (int i = 0; i < len; i++) symbols[i] = (cr_out[i] < 0) ? 0 : 1;      /* slicer :171-178 */
    for (int i = len; i < I->H + 8; i++) symbols[i] = 0;
    if (dbg) {
      if (dbg->nsym) dbg->nsym[chi] = len;
      if (dbg->bits) memcpy(dbg->bits + (size_t)chi * I->H, symbols, (size_t)(len < I->H ? len : I->H));
      if (dbg->demod) memcpy(dbg->demod + (size_t)chi * n_demod, demod_out, sizeof(float) * (size_t)n_demod);
      if (dbg->soft) memcpy(dbg->soft + (size_t)chi * n_demod, cr_out, sizeof(float) * (size_t)len);
    }

    /* BR search, lib/multi_sniffer_impl.cc:107-128 */
    {
      const uint8_t *symp = symbols;
      int limit = ((len - AC_SYMBOLS) < SLOT_SYMBOLS) ? (len - AC_SYMBOLS) : SLOT_SYMBOLS;
      while (limit >= 0) {
        int i = btbo_sniff_ac(symp, limit);
        if (i < 0) break;
        int step = i + AC_SYMBOLS;
        if (*nhits >= hits_cap) { rc = -1; goto done; }
        btbo_hit *h = &hits[(*nhits)++];
        h->slot = slot; h->channel = (int16_t)ch; h->kind = 0;
        h->offset = (int32_t)(symp + i - symbols); h->len = len - i;
        h->lap = air_to_host(symp + i + 38, 24); h->snr = snr;
        len -= step;
        symp += step;
        limit -= step;
      }
    }
    /* LE search, lib/multi_sniffer_impl.cc:130-148 (len stays decremented) */
    {
      const uint8_t *symp = symbols;
      int limit = ((len - AC_SYMBOLS) < SLOT_SYMBOLS) ? (len - AC_SYMBOLS) : SLOT_SYMBOLS;
      while (limit >= 0) {
        int i = btbo_sniff_aa(symp, limit, freq);
        if (i < 0) break;
        int step = i + LE_AA_SYMBOLS;
        if (*nhits >= hits_cap) { rc = -1; goto done; }
        btbo_hit *h = &hits[(*nhits)++];
        h->slot = slot; h->channel = (int16_t)ch; h->kind = 1;
        h->offset = (int32_t)(symp + i - symbols); h->len = len - i;
        h->lap = air_to_host(symp + i + 8, 32); h->snr = snr;
        len -= step;
        symp += step;
        limit -= step;
      }
    }
  }
done:
  free(ddc_out); free(nz_out); free(demod_out); free(cr_out); free(symbols);
  return rc;
}

/* classic_packet_impl::header_present, lib/packet_impl.cc:1205-1242 */
int btbo_header_present(const uint8_t *sym, int length)
{
  if (length < 126) return 0;
  const uint8_t *s = sym + 67;
  int be = 0;
  const uint8_t msb = s[0];
  be += s[1] ^ !msb;
  be += s[2] ^ msb;
  be += s[3] ^ !msb;
  be += s[4] ^ msb;
  s += 5;
  for (int a = 0; a < 54; a += 3)
    be += ((s[a] ^ s[a + 1]) | (s[a + 1] ^ s[a + 2]) | (s[a + 2] ^ s[a]));
  return be < 5;
}

/* the channel loop of multi_hopper_impl::work / hopalong, lib/multi_hopper_impl.cc:93-137, 152-209 */
int btbo_window_list(const btbo_plan *p, btbo_state *st, const float *window, const int32_t *chis, int n,
                     uint32_t stop_lap, btbo_chan_result *res, uint8_t *symbols_out)
{
  const btbo_info *I = &p->info;
  const gra_c32 *in = (const gra_c32 *)window;
  gra_c32 *ddc_out = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)(I->n_ddc + 1));
  gra_c32 *nz_out = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)(I->n_noise + 1));
  float *demod_out = (float *)malloc(sizeof(float) * (size_t)(I->n_ddc + 1));
  float *cr_out = (float *)malloc(sizeof(float) * (size_t)(I->n_ddc + 1));
  int stopped = 0;
  for (int q = 0; q < n; q++) {
    btbo_chan_result *r = &res[q];
    const int chi = chis[q];
    uint8_t *symbols = symbols_out + (size_t)q * I->H;
    memset(r, 0, sizeof *r);
    r->chi = chi; r->ac_index = -1;
    if (stopped) continue;
    r->processed = 1;
    gra_fxlat cf = p->chan_ddc[chi];
    cf.phase = st->chan_phase[chi]; cf.counter = st->chan_counter[chi];
    gra_fxlat_work(&cf, I->n_ddc, in + I->fcs, ddc_out);
    st->chan_phase[chi] = cf.phase; st->chan_counter[chi] = cf.counter;
    double energy = 0.0;
    for (int i = 0; i < I->n_ddc; i++) energy += gra_mag2(ddc_out[i]);
    energy /= I->n_ddc;
    gra_fxlat nf = p->noise_ddc[chi];
    nf.phase = st->noise_phase[chi]; nf.counter = st->noise_counter[chi];
    gra_fxlat_work(&nf, I->n_noise, in + I->fns, nz_out);
    st->noise_phase[chi] = nf.phase; st->noise_counter[chi] = nf.counter;
    double off = 0.0;
    for (int i = 0; i < I->n_noise; i++) off += gra_mag2(nz_out[i]);
    off /= I->n_noise;
    r->snr = 10.0 * log10(energy / off);
    r->pass = (r->snr >= I->snr_db);
    if (!r->pass) continue;
    const int n_demod = I->n_ddc - 1;
    btbo_demod(p, (const float *)ddc_out, demod_out, n_demod);
    float mm[3];
    btbo_state_get_mm(st, mm);
    const int len = btbo_mm_cr(p, mm, demod_out, n_demod, cr_out, n_demod);
    btbo_state_set_mm(st, mm);
    memset(symbols, 0, (size_t)I->H);
    for (int i = 0; i < len; i++) symbols[i] = (cr_out[i] < 0) ? 0 : 1;
    r->nsym = len;
    if (len >= AC_SYMBOLS) {
      const int latest = ((len - AC_SYMBOLS) < SLOT_SYMBOLS) ? (len - AC_SYMBOLS) : SLOT_SYMBOLS;
      const int ac = btbo_sniff_ac(symbols, latest);
      r->ac_index = ac;
      if (ac >= 0) {
        r->lap = air_to_host(symbols + ac + 38, 24);
        if (r->lap == stop_lap && btbo_header_present(symbols + ac, (len - ac) < 3125 ? (len - ac) : 3125)) stopped = 1;
      }
    }
  }
  free(ddc_out); free(nz_out); free(demod_out); free(cr_out);
  return 0;
}

typedef struct {
  const btbo_plan *p; const float *iq; long iq_first, iq_n, first_call, num_calls;
  int flags, per_cap; btbo_hit *tmp; int *cnt;
  uint8_t *bits_out; int bits_stride; int32_t *nsym_out; double *energy_out, *noise_out;
  long next; int rc;
} run_ctx;

static void fill_window(const btbo_info *I, float *win, const float *iq, long iq_first, long iq_n, long k)
{
  long w0 = k * (long)I->S - (I->H - 1);
  for (long j = 0; j < I->H; j++) {
    long s = w0 + j - iq_first;
    if (w0 + j < 0 || s < 0 || s >= iq_n) { win[2 * j] = 0; win[2 * j + 1] = 0; }
    else { win[2 * j] = iq[2 * s]; win[2 * j + 1] = iq[2 * s + 1]; }
  }
}

static void *run_worker(void *arg)
{
  run_ctx *R = (run_ctx *)arg;
  const btbo_info *I = &R->p->info;
  float *win = (float *)malloc(sizeof(float) * 2 * (size_t)I->H);
  uint8_t *bits = R->bits_out ? (uint8_t *)malloc((size_t)I->nch * I->H) : NULL;
  int32_t *nsym_l = (int32_t *)malloc(sizeof(int32_t) * (size_t)I->nch);
  for (;;) {
    long c = __atomic_fetch_add(&R->next, 1, __ATOMIC_RELAXED);
    if (c >= R->num_calls) break;
    long k = R->first_call + c;
    fill_window(I, win, R->iq, R->iq_first, R->iq_n, k);
    btbo_debug dbg;
    memset(&dbg, 0, sizeof dbg);
    dbg.bits = bits;
    dbg.nsym = nsym_l;
    dbg.energy = R->energy_out ? R->energy_out + c * I->nch : NULL;
    dbg.noise = R->noise_out ? R->noise_out + c * I->nch : NULL;
    int n = 0;
    if (btbo_window(R->p, NULL, win, (int)k, R->flags, R->tmp + (size_t)c * R->per_cap, R->per_cap, &n, &dbg) != 0)
      __atomic_store_n(&R->rc, -1, __ATOMIC_RELAXED);
    R->cnt[c] = n;
    if (R->nsym_out) memcpy(R->nsym_out + c * I->nch, nsym_l, sizeof(int32_t) * (size_t)I->nch);
    if (R->bits_out)
      for (int chi = 0; chi < I->nch; chi++) {
        int m = nsym_l[chi];
        if (m > R->bits_stride) m = R->bits_stride;
        memcpy(R->bits_out + ((size_t)c * I->nch + chi) * R->bits_stride, bits + (size_t)chi * I->H, (size_t)m);
      }
  }
  free(win); free(bits); free(nsym_l);
  return NULL;
}

/* Scheduler emulation, SURVEY.md 3.4 / Appendix A.1 */
int btbo_run(const btbo_plan *p, btbo_state *st, const float *iq, long iq_first, long iq_n,
             long first_call, long num_calls, int flags, int threads,
             btbo_hit *hits, int hits_cap, int *nhits,
             uint8_t *bits_out, int bits_stride, int32_t *nsym_out,
             double *energy_out, double *noise_out)
{
  const btbo_info *I = &p->info;
  const int stateless = flags & 1;
  int rc = 0;
  *nhits = 0;
  if (!stateless || threads <= 1) {
    float *win = (float *)malloc(sizeof(float) * 2 * (size_t)I->H);
    uint8_t *bits = bits_out ? (uint8_t *)malloc((size_t)I->nch * I->H) : NULL;
    for (long c = 0; c < num_calls && rc == 0; c++) {
      long k = first_call + c;
      fill_window(I, win, iq, iq_first, iq_n, k);
      btbo_debug dbg;
      memset(&dbg, 0, sizeof dbg);
      dbg.bits = bits;
      dbg.nsym = nsym_out ? nsym_out + c * I->nch : NULL;
      dbg.energy = energy_out ? energy_out + c * I->nch : NULL;
      dbg.noise = noise_out ? noise_out + c * I->nch : NULL;
      rc = btbo_window(p, st, win, (int)k, flags, hits, hits_cap, nhits, &dbg);
      if (bits_out)
        for (int chi = 0; chi < I->nch; chi++) {
          int n = dbg.nsym ? dbg.nsym[chi] : 0;
          if (n > bits_stride) n = bits_stride;
          memcpy(bits_out + ((size_t)c * I->nch + chi) * bits_stride, bits + (size_t)chi * I->H, (size_t)n);
        }
    }
    free(win); free(bits);
    return rc;
  }
  /* stateless calls are independent: run them on a pthread pool, merge hits in call order */
  {
    const int per_cap = 64;
    run_ctx R;
    memset(&R, 0, sizeof R);
    R.p = p; R.iq = iq; R.iq_first = iq_first; R.iq_n = iq_n; R.first_call = first_call;
    R.num_calls = num_calls; R.flags = flags; R.per_cap = per_cap;
    R.tmp = (btbo_hit *)malloc(sizeof(btbo_hit) * (size_t)per_cap * (size_t)num_calls);
    R.cnt = (int *)calloc((size_t)num_calls, sizeof(int));
    R.bits_out = bits_out; R.bits_stride = bits_stride; R.nsym_out = nsym_out;
    R.energy_out = energy_out; R.noise_out = noise_out;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, run_worker, &R);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    rc = R.rc;
    for (long c = 0; c < num_calls; c++)
      for (int j = 0; j < R.cnt[c]; j++) {
        if (*nhits >= hits_cap) { rc = -1; break; }
        hits[(*nhits)++] = R.tmp[(size_t)c * per_cap + j];
      }
    free(R.tmp); free(R.cnt);
  }
  return rc;
}
