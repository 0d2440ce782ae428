/*
 * oracle/shim/btref_hooks.h -- TEST INFRASTRUCTURE.
 * Observation/intervention points of the GNU Radio stand-in.  The reference's
 * sources are compiled verbatim; everything the tests want to see of its
 * internals (DDC outputs, energies, demod floats, soft symbols) is observed at
 * the GNU Radio boundary the reference calls through.
 */
#ifndef BTREF_HOOKS_H
#define BTREF_HOOKS_H
#include <stdio.h>
#include <stdint.h>

namespace gr { namespace filter { class freq_xlating_fir_filter_ccf; } }

struct btref_hooks {
  int   call_index;                 /* work() call number, set by the driver */
  int   stateless;                  /* 1: rotator reset at every DDC work(); driver resets M&M */
  void (*on_channel_ddc)(void *user);   /* called at the start of every CHANNEL ddc work() */
  void *user;
  FILE *dump;                       /* binary record stream, NULL = no capture */
  int   heavy_from, heavy_to;       /* call range [from,to) in which heavy records are written */
  /* internal */
  int   cur_ddc_id;                 /* id of the DDC that ran last */
  int   cur_chan_ddc_id;            /* id of the channel DDC that ran last */
  int   cur_chan_nout;
};
extern btref_hooks g_btref;

/* record types (all little endian): u32 type, u32 call, i32 id, u32 n, payload */
enum {
  BTREF_REC_DDC    = 1,   /* heavy: n complex64 DDC outputs (after rotator), id = ddc id */
  BTREF_REC_ENERGY = 2,   /* light: one f64 = mean |y|^2 (double accumulate), id = ddc id, n = count */
  BTREF_REC_BITS   = 3,   /* light: n symbols, payload = n bytes 0/1 (sign of interpolate()), id = chan ddc id */
  BTREF_REC_SOFT   = 4,   /* heavy: n f32 interpolate() outputs, id = chan ddc id */
  BTREF_REC_DEMOD  = 5,   /* heavy: n f32 demod_out[] values (index 0 is the unwritten slot) */
  BTREF_REC_MU     = 6    /* heavy: n f32 mu values passed to interpolate() */
};

void btref_flush_symbols(void);     /* driver calls this after every work() */
#endif
