// tests/emul/emul.cpp -- TEST HARNESS, not product code.
//
// Runs the per-thread bodies of the baseline kernels (gr-bluetooth_b200/csrc/
// rx_bodies.cuh, rx_math.cuh) and the host design step (plan.cpp) on the CPU, in
// the same order the CUDA launch code drives them, so the kernel logic can be
// checked against the oracle in the CPU-only test tier.  Never linked into
// libbtb200.so; the product has no CPU path.
#include "../../gr-bluetooth_b200/csrc/plan.hpp"
#include "../../gr-bluetooth_b200/csrc/rx_bodies.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace btb200;

extern "C" {

struct emul_hit { int32_t slot; int16_t channel; int16_t kind; int32_t offset; int32_t n_symbols; uint32_t lap; double snr; };

struct emul_out {
  double *energy, *noise;      // [B][nch]
  int32_t *nsym;               // [B][nch]
  uint8_t *bits;               // [B][nch][stride]
  int32_t bits_stride;
  emul_hit *hits; int32_t hits_cap; int32_t nhits;
  float mm[3];                 // chained state in/out
};

int emul_plan_info(double fs, double fc, double snr, int extra, int32_t out[16])
{
  Plan P;
  if (P.design(fs, fc, snr, extra)) return -1;
  int32_t v[16] = {P.S, P.H, P.D, P.Nc, P.Nn, P.fcs, P.fns, P.ch_lo, P.ch_hi, P.nch, P.n_ddc, P.n_noise, P.n_dem, P.grid_per_slot, 0, 0};
  std::memcpy(out, v, sizeof v);
  return 0;
}

// tables for direct comparison with the oracle's
int emul_plan_tables(double fs, double fc, int extra, int chi, float *chan_rtaps, float *noise_rtaps,
                     float *mmse, float *atan_tab, uint64_t *ac_lut, float *incr4)
{
  Plan P;
  if (P.design(fs, fc, 10.0, extra)) return -1;
  std::memcpy(chan_rtaps, &P.chan_rtaps[(size_t)chi * P.Nc], (size_t)P.Nc * 8);
  std::memcpy(noise_rtaps, &P.noise_rtaps[(size_t)chi * P.Nn], (size_t)P.Nn * 8);
  std::memcpy(mmse, P.mmse.data(), P.mmse.size() * 4);
  std::memcpy(atan_tab, P.atan_tab.data(), 257 * 4);
  std::memcpy(ac_lut, P.ac_lut.data(), 769 * 8);
  incr4[0] = P.chan_incr[chi].re; incr4[1] = P.chan_incr[chi].im;
  incr4[2] = P.noise_incr[chi].re; incr4[3] = P.noise_incr[chi].im;
  return 0;
}

uint64_t emul_sync_word(uint32_t lap) { return sync_word(lap); }

// x: (B-1)*S+H complex samples; rot_state (chained): rotators continue from rot_calls previous calls
int emul_run(double fs, double fc, double snr_db, int extra, int stateless, int search,
             const float *xf, int B, int first_slot, emul_out *o)
{
  Plan P;
  if (P.design(fs, fc, snr_db, extra)) return -1;
  Geom G{};
  G.S = P.S; G.H = P.H; G.D = P.D; G.Nc = P.Nc; G.Nn = P.Nn; G.fcs = P.fcs; G.fns = P.fns;
  G.nch = P.nch; G.n_ddc = P.n_ddc; G.n_noise = P.n_noise; G.n_dem = P.n_dem; G.gps = P.grid_per_slot;
  G.n_dem_pad = (P.n_dem + 3) & ~3;
  G.bw = (P.n_dem + 31) / 32 + 4;
  G.ch_lo = P.ch_lo; G.demod_gain = P.demod_gain;
  G.mm = MmConst{P.gain_mu, P.gain_omega, P.omega_mid, P.omega_lim};
  G.mu0 = P.mu0; G.squelch_db = P.squelch_db; G.search = search; G.stateless = stateless;
  const c32 *x = reinterpret_cast<const c32 *>(xf);
  const int nch = P.nch;
  const long Gtot = (long)(B - 1) * G.gps + G.n_ddc;
  std::vector<c32> Y((size_t)Gtot * nch), Nz((size_t)B * G.n_noise * nch);
  for (long g = 0; g < Gtot; g++)
    for (int c = 0; c < nch; c++)
      Y[(size_t)g * nch + c] = chan_fir_point(G, x, reinterpret_cast<const c32 *>(&P.chan_rtaps[(size_t)c * P.Nc]), g);
  for (int b = 0; b < B; b++)
    for (int j = 0; j < G.n_noise; j++)
      for (int c = 0; c < nch; c++)
        Nz[((size_t)b * G.n_noise + j) * nch + c] =
            noise_fir_point(G, x, reinterpret_cast<const c32 *>(&P.noise_rtaps[(size_t)c * P.Nn]), b, j);
  // rotator tables
  const int Bp = stateless ? 1 : B;
  std::vector<c32> phc((size_t)Bp * G.n_ddc * nch), phn((size_t)Bp * G.n_noise * nch);
  std::vector<Rotator> rc(nch), rn(nch);
  for (int c = 0; c < nch; c++) { rc[c].incr = P.chan_incr[c]; rn[c].incr = P.noise_incr[c]; }
  if (!stateless) {
    // advance the free-running rotators over the first_slot calls before this batch
    std::vector<cf32> scratch((size_t)std::max(G.n_ddc, G.n_noise));
    for (int k = 0; k < first_slot; k++)
      for (int c = 0; c < nch; c++) { rc[c].generate(scratch.data(), G.n_ddc, 1); rn[c].generate(scratch.data(), G.n_noise, 1); }
  }
  for (int b = 0; b < Bp; b++)
    for (int c = 0; c < nch; c++) {
      rc[c].generate(reinterpret_cast<cf32 *>(&phc[((size_t)b * G.n_ddc) * nch + c]), G.n_ddc, nch);
      rn[c].generate(reinterpret_cast<cf32 *>(&phn[((size_t)b * G.n_noise) * nch + c]), G.n_noise, nch);
    }
  uint8_t hdr[4 * 256];
  for (int w = 0; w < 4; w++) for (int v = 0; v < 256; v++) hdr[w * 256 + v] = (uint8_t)le_hdr_dist((uint32_t)v, w);
  std::vector<float> dem((size_t)G.n_dem_pad);
  std::vector<uint32_t> row((size_t)G.bw);
  MmState chained{o->mm[0], o->mm[1], o->mm[2]};
  o->nhits = 0;
  for (int b = 0; b < B; b++)
    for (int c = 0; c < nch; c++) {
      const size_t bc = (size_t)b * nch + c;
      double on, off;
      window_energy(G, Y.data(), Nz.data(), phc.data(), phn.data(), b, c, stateless ? 0 : b, &on, &off);
      o->energy[bc] = on; o->noise[bc] = off;
      const double snr = 10.0 * std::log10(on / off);
      o->nsym[bc] = 0;
      if (!(snr >= snr_db)) continue;
      for (int i = 0; i < G.n_dem; i++)
        dem[i] = window_demod_point(G, Y.data(), phc.data(), P.atan_tab.data(), b, c, stateless ? 0 : b, i);
      MmState st = stateless ? MmState{G.mu0, G.mm.omega_mid, 0.0f} : chained;
      const int nsym = window_mm(G, P.mmse.data(), dem.data(), st, row.data(), nullptr);
      if (!stateless) chained = st;
      o->nsym[bc] = nsym;
      if (o->bits)
        for (int i = 0; i < nsym && i < o->bits_stride; i++)
          o->bits[bc * o->bits_stride + i] = (row[i >> 5] >> (i & 31)) & 1;
      uint32_t white = 0;
      for (int i = 0; i < 16; i++) white |= (uint32_t)P.le_white16[(size_t)c * 16 + i] << i;
      auto emit = [&](int kind, int offset, int n_symbols, uint32_t lap) {
        if (o->nhits < o->hits_cap) {
          emul_hit &h = o->hits[o->nhits];
          h.slot = first_slot + b; h.channel = (int16_t)(P.ch_lo + c); h.kind = (int16_t)kind;
          h.offset = offset; h.n_symbols = n_symbols; h.lap = kind == 0 ? (lap & 0xffffff) : lap; h.snr = snr;
        }
        o->nhits++;
      };
      window_search(G, P.ac_lut.data(), hdr, row.data(), nsym, P.le_index[c], white, emit);
    }
  o->mm[0] = chained.mu; o->mm[1] = chained.omega; o->mm[2] = chained.last;
  return 0;
}

// Polyphase channelizer of the throughput mode, driven by the PRODUCT's host tables (PfbDesign, plan.cpp) in the
// stage order of the kernel (rx_pfb.cu): branch sums -> N1-point DFTs -> N2-point DFTs at the channel bins.
// Double precision, plain loops: checks the tables and the index algebra, not the kernel's thread mapping.
// x: n_x complex samples (x[0] = first sample of window 0, no pre-rotation applied); Z: [n_grid][nch] complex doubles;
// kappa: [nch] complex floats; dims: M, D, Q, N1, N2, tps, CPC, ncol, span, nfull, rem, a0
int emul_pfb(double fs, double fc, int extra, const float *xf, long n_x, int n_grid, double *Z, float *kappa, double *phi_out,
             int32_t dims[12])
{
  Plan P;
  if (P.design(fs, fc, 10.0, extra)) return -1;
  PfbDesign F;
  if (F.design(P, 63, 64, 5)) return -2;
  const int M = F.M, N1 = F.N1, N2 = F.N2;
  int32_t d[12] = {F.M, F.D, F.Q, F.N1, F.N2, F.tps, F.CPC, F.ncol, F.span, F.nfull, F.rem, F.a0};
  std::memcpy(dims, d, sizeof d);
  *phi_out = F.phi;
  for (int c = 0; c < P.nch; c++) { kappa[2 * c] = F.kappa[F.chan_col[c]].re; kappa[2 * c + 1] = F.kappa[F.chan_col[c]].im; }
  std::vector<double> ur(M), ui(M), vr((size_t)N1 * N2), vi((size_t)N1 * N2);
  for (int g = 0; g < n_grid; g++) {
    const long n0 = (long)P.fcs + (long)g * F.D;
    for (int r = 0; r < M; r++) {
      double ar = 0, ai = 0;
      for (int q = 0; q < F.Q; q++) {
        const long n = n0 + r + (long)M * q;
        if (n >= n_x) continue;
        // x'[n] = x[n] e^{-j 2 pi phi n / M}
        const double ang = -2.0 * M_PI * F.phi * (double)n / M;
        const double xr = xf[2 * n], xi = xf[2 * n + 1], cs = std::cos(ang), sn = std::sin(ang);
        const double h = F.hq[(size_t)q * M + r];
        ar += (xr * cs - xi * sn) * h; ai += (xr * sn + xi * cs) * h;
      }
      ur[r] = ar; ui[r] = ai;
    }
    for (int k1 = 0; k1 < N1; k1++)
      for (int n2 = 0; n2 < N2; n2++) {
        double ar = 0, ai = 0;
        for (int n1 = 0; n1 < N1; n1++) {
          const int r = (N2 * n1 + N1 * n2) % M;
          const double ang = -2.0 * M_PI * (double)((n1 * k1) % N1) / N1;
          const double cs = std::cos(ang), sn = std::sin(ang);
          ar += ur[r] * cs - ui[r] * sn; ai += ur[r] * sn + ui[r] * cs;
        }
        vr[(size_t)k1 * N2 + n2] = ar; vi[(size_t)k1 * N2 + n2] = ai;
      }
    for (int c = 0; c < P.nch; c++) {
      const int col = F.chan_col[c], k1 = col / F.CPC;
      if (F.col_chan[col] != c) return -3;
      double ar = 0, ai = 0;
      for (int n2 = 0; n2 < N2; n2++) {
        const cf32 w = F.WB[(size_t)n2 * F.ncol + col];
        const double a = vr[(size_t)k1 * N2 + n2], b = vi[(size_t)k1 * N2 + n2];
        ar += a * w.re - b * w.im; ai += a * w.im + b * w.re;
      }
      Z[((size_t)g * P.nch + c) * 2] = ar; Z[((size_t)g * P.nch + c) * 2 + 1] = ai;
    }
  }
  return 0;
}

// quadrature weights of the sub-sampled off-channel energy sum (plan.cpp: nest_quadrature); w: room for (N-1)/s+1+n_extra floats
double emul_nest_quadrature(int N, int s, int n_extra, int n_free, double omega_max, float *w, int32_t *n_used)
{
  std::vector<float> v;
  const double res = nest_quadrature(N, s, n_extra, n_free, omega_max, v);
  if (res < 0) return res;
  *n_used = (int32_t)v.size();
  std::memcpy(w, v.data(), v.size() * sizeof(float));
  return res;
}

// libbtbb-style access-code test of the product (rx_math.cuh: br_lag_test_bch with the tables of plan.cpp) on EVERY lag
// of a symbol stream (one symbol per byte): lags/laps/errs of the accepted lags, up to cap; returns their number.
// lap = 0xffffffff: LAP_ANY.
int emul_bch_scan(const uint8_t *symbols, long n, int max_err, uint32_t lap, int32_t *lags, uint32_t *laps, int32_t *errs, int cap)
{
  BchTables T;
  if (T.build(max_err)) return -1;
  BchDev B;
  B.par = T.par.data(); B.syn = T.syn.data(); B.err = T.err.data(); B.n = (int)T.syn.size(); B.max_err = max_err;
  B.lap = lap; B.target = lap != 0xffffffffu ? sync_word(lap) : 0;
  const long nw = (n + 31) / 32 + 4;
  std::vector<uint32_t> row((size_t)nw, 0u);
  for (long i = 0; i < n; i++) row[(size_t)(i >> 5)] |= (uint32_t)(symbols[i] & 1) << (i & 31);
  int found = 0;
  for (long lag = 0; lag + 68 <= n; lag++) {
    uint64_t lo; uint32_t hi;
    bits_window(row.data(), (int)lag, &lo, &hi);
    uint32_t l = 0; int e = 0;
    if (br_lag_test_bch(B, lo, hi, &l, &e)) {
      if (found < cap) { lags[found] = (int32_t)lag; laps[found] = l; errs[found] = e; }
      found++;
    }
  }
  return found;
}

// Off-channel energy estimate of one window, the way rx_nest.cu computes it, driven by the PRODUCT's host tables
// (PfbDesign::design_noise, nest_quadrature) in double precision with plain loops: pre-rotation, fold * M virtual
// branches over the flat tap array, fold, N1-point and N2-point DFTs at the channel bins, weighted |Z|^2.
// fold = 0: every output (weights 1).  esum[c] = the estimate of sum_j |y_j|^2 of window b (x[0] = first sample of
// window 0).  Checks the tables and the index algebra of the folded mode, not the kernel's thread mapping.
int emul_nest(double fs, double fc, int extra, const float *xf, long n_x, int b, int fold, double *esum)
{
  Plan P;
  if (P.design(fs, fc, 10.0, extra)) return -1;
  PfbDesign F;
  if (F.design_noise(P, 5, 16)) return -2;
  const int M = F.M, N1 = F.N1, N2 = F.N2;
  const int stride = fold ? 2 * fold : 1, MV = fold ? fold * M : M;
  std::vector<float> w;
  if (fold) {
    if (nest_quadrature(P.n_noise, stride, 2, 12, 2.0 * M_PI * 90e3 * P.D / P.fs, w) < 0) return -3;
  } else w.assign((size_t)P.n_noise, 1.0f);
  // outputs of a sequence with period `step` samples between them
  const long step = (long)stride * P.D;
  if (fold && step != MV) return -4;
  for (int c = 0; c < P.nch; c++) esum[c] = 0.0;
  std::vector<double> vr((size_t)MV), vi((size_t)MV), ur((size_t)M), ui((size_t)M), tr((size_t)N1 * N2), ti((size_t)N1 * N2);
  for (size_t u = 0; u < w.size(); u++) {
    const long n0 = (long)b * P.S + P.fns + (long)u * step;
    for (int r = 0; r < MV; r++) {
      double ar = 0, ai = 0;
      for (long k = r; k < P.Nn; k += MV) {                       // flat tap array: virtual branch r holds h'[r + MV q]
        const long n = n0 + k;
        if (n >= n_x) return -5;
        const double ang = -2.0 * M_PI * F.phi * (double)n / M;
        const double xr = xf[2 * n], xi = xf[2 * n + 1], cs = std::cos(ang), sn = std::sin(ang);
        const double h = F.hq[(size_t)k];
        ar += (xr * cs - xi * sn) * h; ai += (xr * sn + xi * cs) * h;
      }
      vr[r] = ar; vi[r] = ai;
    }
    for (int r = 0; r < M; r++) {
      double ar = 0, ai = 0;
      for (int f = 0; f * M + r < MV; f++) { ar += vr[(size_t)f * M + r]; ai += vi[(size_t)f * M + r]; }
      ur[r] = ar; ui[r] = ai;
    }
    for (int k1 = 0; k1 < N1; k1++)
      for (int n2 = 0; n2 < N2; n2++) {
        double ar = 0, ai = 0;
        for (int n1 = 0; n1 < N1; n1++) {
          const int r = (N2 * n1 + N1 * n2) % M;
          const double ang = -2.0 * M_PI * (double)((n1 * k1) % N1) / N1;
          const double cs = std::cos(ang), sn = std::sin(ang);
          ar += ur[r] * cs - ui[r] * sn; ai += ur[r] * sn + ui[r] * cs;
        }
        tr[(size_t)k1 * N2 + n2] = ar; ti[(size_t)k1 * N2 + n2] = ai;
      }
    for (int c = 0; c < P.nch; c++) {
      const int col = F.chan_col[c], k1 = col / F.CPC;
      double ar = 0, ai = 0;
      for (int n2 = 0; n2 < N2; n2++) {
        const cf32 ww = F.WB[(size_t)n2 * F.ncol + col];
        const double a = tr[(size_t)k1 * N2 + n2], bb = ti[(size_t)k1 * N2 + n2];
        ar += a * ww.re - bb * ww.im; ai += a * ww.im + bb * ww.re;
      }
      esum[c] += (double)w[u] * (ar * ar + ai * ai);
    }
  }
  return 0;
}

}  // extern "C"
