// rx_bodies.cuh -- per-thread bodies of the baseline ("v1") kernels.
//
// Each body computes what ONE thread of the corresponding kernel computes, from
// explicit indices, with no shared memory or warp intrinsics, so that the test
// harness can run them on the CPU (tests/emul) and so that the tuned kernels in
// rx_kernels.cu have a simple bit-exact yard-stick on the GPU.
//
// Device data layout (one batch of B slots, all row-major):
//   x      c32 [(B-1)*S + H]              input IQ, x[0] = first sample of window 0
//   Y      c32 [G][nch]     G=(B-1)*gps+n_ddc   de-duplicated channel FIR outputs
//                                          BEFORE the rotator; window b output i
//                                          is Y[b*gps + i]   (gps = S/D)
//   Nz     c32 [B][n_noise][nch]          noise FIR outputs before the rotator
//   phc    c32 [Bp][n_ddc][nch]           rotator phase multiplying channel output i
//   phn    c32 [Bp][n_noise][nch]         (Bp = 1 in stateless mode: same table every window)
//   energy,noise  f64 [B][nch]
//   dem    f32 [B][nch][n_dem_pad]        demod floats (index 0 = 0.0f)
//   bits   u32 [B][nch][bw]               sliced symbols, bit (i&31) of word i>>5
//   nsym   i32 [B][nch]
#pragma once
#include "rx_math.cuh"

namespace btb200 {

struct Geom {
  int S, H, D, Nc, Nn, fcs, fns, nch, n_ddc, n_noise, n_dem, gps;
  int n_dem_pad;      // row pitch of dem[]
  int bw;             // words per bits row
  int ch_lo;
  float demod_gain;
  MmConst mm;
  float mu0;
  double squelch_db;
  int search;         // BTB200_SEARCH_* mask
  int stateless;
  int early;          // lazy tail: clock recovery stops at sym_target symbols, demod computed for i < ne_dem
  int ne_dem, sym_target;
  // polyphase mode: the demod floats live on the GLOBAL decimation grid, dem[g][c] with g = b * gps + i (windows are
  // views, index 0 of a window reads as 0.0f like the reference's never-written demod_out[0]); else [b][n_dem_pad][c]
  int dem_grid;
  int dem_rows;       // rows of nch floats between consecutive windows: gps (grid) or n_dem_pad
};

// ---- channel FIR: Y[g][c] = sum_k x[fcs + g*D + k] * rt[c][k]  (k ascending)
// restates freq_xlating_fir_filter_ccf::work's dot product (A.3) for
// multi_block::channel_samples (lib/multi_block.cc:180-204)
BTB_HD c32 chan_fir_point(const Geom &G, const c32 *__restrict__ x, const c32 *__restrict__ rt_c, long g)
{
  const c32 *xi = x + G.fcs + g * (long)G.D;
  float ar = 0.0f, ai = 0.0f;
  for (int k = 0; k < G.Nc; k++) cmac(ar, ai, xi[k].re, xi[k].im, rt_c[k].re, rt_c[k].im);
  return c32{ar, ai};
}

// ---- noise FIR (multi_block::check_snr, lib/multi_block.cc:253-275)
BTB_HD c32 noise_fir_point(const Geom &G, const c32 *__restrict__ x, const c32 *__restrict__ rt_c, int b, int j)
{
  const c32 *xi = x + (long)b * G.S + G.fns + (long)j * G.D;
  float ar = 0.0f, ai = 0.0f;
  for (int k = 0; k < G.Nn; k++) cmac(ar, ai, xi[k].re, xi[k].im, rt_c[k].re, rt_c[k].im);
  return c32{ar, ai};
}

// ---- energies of one channel-window (lib/multi_block.cc:206-218, 277-287):
// rotate, |.|^2 in fp32, accumulate in fp64 in index order, divide by count.
BTB_HD void window_energy(const Geom &G, const c32 *__restrict__ Y, const c32 *__restrict__ Nz,
                          const c32 *__restrict__ phc, const c32 *__restrict__ phn,
                          int b, int c, int bp, double *e_on, double *e_off)
{
  double e = 0.0;
  const c32 *y = Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = phc + ((long)bp * G.n_ddc) * G.nch + c;
  for (int i = 0; i < G.n_ddc; i++) e += mag2(crot(y[(long)i * G.nch], p[(long)i * G.nch]));
  *e_on = e / G.n_ddc;
  double n = 0.0;
  const c32 *z = Nz + ((long)b * G.n_noise) * G.nch + c;
  const c32 *q = phn + ((long)bp * G.n_noise) * G.nch + c;
  for (int j = 0; j < G.n_noise; j++) n += mag2(crot(z[(long)j * G.nch], q[(long)j * G.nch]));
  *e_off = n / G.n_noise;
}

// ---- demod of one point i (1 <= i < n_dem); index 0 is the never-written slot = 0
BTB_HD float window_demod_point(const Geom &G, const c32 *__restrict__ Y, const c32 *__restrict__ phc,
                                const float *__restrict__ atan_tab, int b, int c, int bp, int i)
{
  if (i == 0) return 0.0f;
  const c32 *y = Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = phc + ((long)bp * G.n_ddc) * G.nch + c;
  const c32 cur = crot(y[(long)i * G.nch], p[(long)i * G.nch]);
  const c32 prev = crot(y[(long)(i - 1) * G.nch], p[(long)(i - 1) * G.nch]);
  return demod_point(atan_tab, G.demod_gain, cur, prev);
}

// ---- M&M clock recovery + slicer over one demod row (lib/multi_block.cc:128-155, 171-178)
// writes packed bits, optional soft symbols; returns symbol count.
BTB_HD int window_mm(const Geom &G, const float *__restrict__ mmse, const float *__restrict__ dem_row,
                     MmState &st, uint32_t *__restrict__ bits_row, float *__restrict__ soft_row)
{
  unsigned ii = 0;
  int oo = 0;
  const unsigned ni = (unsigned)(G.n_dem - 8);
  uint32_t word = 0;
  while (oo < G.n_dem && ii < ni) {
    const float out = mmse_interp(mmse, dem_row + ii, st.mu);
    if (soft_row) soft_row[oo] = out;
    if (!(out < 0)) word |= 1u << (oo & 31);
    if ((oo & 31) == 31) { bits_row[oo >> 5] = word; word = 0; }
    ii += (unsigned)mm_update(G.mm, st, out);
    oo++;
  }
  if (oo & 31) bits_row[oo >> 5] = word;
  for (int w = (oo + 31) >> 5; w < G.bw; w++) bits_row[w] = 0;
  return oo;
}

// 64 symbols starting at lag (bit i = symbol lag+i) and the next 8
BTB_HD void bits_window(const uint32_t *__restrict__ row, int lag, uint64_t *lo, uint32_t *hi)
{
  const int w = lag >> 5, s = lag & 31;
  const uint64_t a = row[w] | ((uint64_t)row[w + 1] << 32);
  const uint64_t b = row[w + 2] | ((uint64_t)row[w + 3] << 32);
  *lo = s ? (a >> s) | (b << (64 - s)) : a;
  *hi = (uint32_t)(b >> s) & 0xff;
}

BTB_HD int row_bit(const uint32_t *__restrict__ row, int i) { return (row[i >> 5] >> (i & 31)) & 1; }

// classic_packet_impl::header_present (lib/packet_impl.cc:1205-1242) on a packed symbol row:
// the packet starts at symbol `start` and has `length` symbols
BTB_HD int header_present_bits(const uint32_t *__restrict__ row, int start, int length)
{
  if (length < 126) return 0;
  const int s = start + 67;
  int be = 0;
  const int msb = row_bit(row, s);
  be += row_bit(row, s + 1) ^ !msb;
  be += row_bit(row, s + 2) ^ msb;
  be += row_bit(row, s + 3) ^ !msb;
  be += row_bit(row, s + 4) ^ msb;
  for (int a = 0; a < 54; a += 3) {
    const int x = row_bit(row, s + 5 + a), y = row_bit(row, s + 6 + a), z = row_bit(row, s + 7 + a);
    be += ((x ^ y) | (y ^ z) | (z ^ x));
  }
  return be < 5;
}

struct DevHit {
  int32_t  b;           // slot in batch
  int16_t  chi;         // channel index
  int16_t  kind;
  int32_t  offset;
  int32_t  n_symbols;
  uint32_t lap;
  uint32_t sym_count;
  uint64_t sym_offset;
};

// ---- the search loops of multi_sniffer_impl::work (lib/multi_sniffer_impl.cc:107-148)
// over one channel-window; emit() is called for every ac()/aa() invocation.
template <class Emit>
BTB_HD void window_search(const Geom &G, const uint64_t *__restrict__ ac_lut, const uint8_t *__restrict__ le_hdr_lut,
                          const uint32_t *__restrict__ row, int nsym, int le_idx, uint32_t le_white, Emit emit,
                          const BchDev *bch = nullptr)
{
  const bool use_bch = (G.search & 4) && bch != nullptr;     // BTB200_SEARCH_BR_BCH: libbtbb-style test instead of sniff_ac's
  int len = nsym;
  if (G.search & 1) {
    const int limit0 = (len - 68 < 625) ? len - 68 : 625;   // absolute end of the BR search
    int start = 0;                                           // absolute lag where the next sniff_ac starts
    while (limit0 - start >= 0) {
      int found = -1;
      uint32_t lap = 0;
      for (int lag = start; lag < limit0; lag++) {
        uint64_t lo; uint32_t hi;
        bits_window(row, lag, &lo, &hi);
        int ne = 0;
        if (use_bch ? br_lag_test_bch(*bch, lo, hi, &lap, &ne) : br_lag_test(ac_lut, lo, hi, &lap)) {
          found = lag;
          if (use_bch) lap = (lap & 0xffffff) | ((uint32_t)ne << 24);
          break;
        }
      }
      if (found < 0) break;
      if (!use_bch) {
        uint64_t lo; uint32_t hi;
        bits_window(row, found, &lo, &hi);
        lap |= (uint32_t)br_lag_errors(ac_lut, lo, hi) << 24;      // bits 24..31 of a BR hit's lap field: symbol errors of the access code
      }
      emit(0, found, nsym - found, lap);
      start = found + 68;
    }
    len = nsym - start;                                      // len after the BR steps
  }
  if ((G.search & 2) && le_idx >= 0) {
    const int limit0 = (len - 68 < 625) ? len - 68 : 625;
    int start = 0;
    const int len_le = len;
    while (limit0 - start >= 0) {
      int found = -1;
      for (int lag = start; lag < limit0; lag++) {
        uint64_t lo; uint32_t hi;
        bits_window(row, lag, &lo, &hi);
        if (le_lag_test(le_hdr_lut, lo, le_white, le_idx >= 37)) { found = lag; break; }
      }
      if (found < 0) break;
      uint64_t lo; uint32_t hi;
      bits_window(row, found, &lo, &hi);
      emit(1, found, len_le - found, (uint32_t)(lo >> 8));
      start = found + 40;
    }
  }
}

}  // namespace btb200
