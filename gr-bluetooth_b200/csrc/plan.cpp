// plan.cpp -- see plan.hpp.  Host code, compiled with -ffp-contract=off.
#include "plan.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace btb200 {

namespace {

constexpr double kSymbolRate = 1e6;        // include/gr_bluetooth/multi_block.h:47
constexpr int    kSlotSymbols = 625;       // :56
constexpr double kBaseFreq = 2402000000.0; // :60
constexpr double kChanWidth = 1000000.0;   // :63
constexpr int    kMmseTaps = 8, kMmseSteps = 128;

// gr::filter::firdes::low_pass(gain=1, fs, fc, tw, WIN_HANN)  (Appendix A.2)
std::vector<float> hann_lowpass(double fs, double cutoff, double tw)
{
  int n = (int)(44.0 * fs / (22.0 * tw));
  if (!(n & 1)) n++;
  std::vector<float> win(n), taps(n);
  const int M = (n - 1) / 2;
  const float span = (float)(n - 1);
  for (int i = 0; i < n; i++) win[i] = (float)(0.5 - 0.5 * std::cos((2 * M_PI * i) / span));
  const double w0 = 2 * M_PI * cutoff / fs;
  for (int k = -M; k <= M; k++) {
    double ideal = (k == 0) ? w0 / M_PI : std::sin(k * w0) / (k * M_PI);
    taps[k + M] = (float)(ideal * win[k + M]);
  }
  double dc = taps[M];
  for (int k = 1; k <= M; k++) dc += 2 * taps[k + M];
  const double g = 1.0 / dc;
  for (auto &t : taps) t = (float)(t * g);
  return taps;
}

// freq_xlating_fir_filter_ccf tap translation + rotator increment (Appendix A.3)
void translate(const std::vector<float> &proto, double f_off, double fs, int decim,
               cf32 *rtaps, cf32 *incr)
{
  const int n = (int)proto.size();
  const float theta = (float)(2 * M_PI * f_off / fs);
  for (int i = 0; i < n; i++) {
    const float ang = i * theta;
    cf32 t{proto[i] * std::cos(ang), proto[i] * std::sin(ang)};
    rtaps[n - 1 - i] = t;
  }
  const float a = -theta * decim;
  cf32 e{std::cos(a), std::sin(a)};
  const float mag = std::hypot(e.re, e.im);
  *incr = {e.re / mag, e.im / mag};
}

// mmse_fir_interpolator_ff tap table (Appendix A.5): least-squares fractional
// delay against an ideal |f|<=0.25 low-pass, rounded through "%.5e".
std::vector<float> mmse_table()
{
  std::vector<float> out((kMmseSteps + 1) * kMmseTaps);
  auto g = [](double d) { return d == 0.0 ? 0.5 : std::sin(2 * M_PI * 0.25 * d) / (M_PI * d); };
  for (int s = 0; s <= kMmseSteps; s++) {
    double A[8][9];
    const double tau = -(double)s / kMmseSteps;
    for (int i = 0; i < 8; i++) {
      for (int j = 0; j < 8; j++) A[i][j] = g((double)(i - 4) - (double)(j - 4));
      A[i][8] = g((double)(i - 4) - tau);
    }
    for (int c = 0; c < 8; c++) {
      int piv = c;
      for (int r = c + 1; r < 8; r++) if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
      if (piv != c) for (int k = 0; k < 9; k++) std::swap(A[c][k], A[piv][k]);
      for (int r = c + 1; r < 8; r++) {
        const double m = A[r][c] / A[c][c];
        for (int k = c; k < 9; k++) A[r][k] -= m * A[c][k];
      }
    }
    double h[8];
    for (int r = 7; r >= 0; r--) {
      double acc = A[r][8];
      for (int k = r + 1; k < 8; k++) acc -= A[r][k] * h[k];
      h[r] = acc / A[r][r];
    }
    for (int i = 0; i < 8; i++) {
      char buf[40];
      double v = std::fabs(h[i]) < 5e-10 ? 0.0 : h[i];
      std::snprintf(buf, sizeof buf, "%.5e", v);
      out[s * kMmseTaps + i] = std::strtof(buf, nullptr);
    }
  }
  return out;
}

// whitening sequence (x^7+x^4+1), one period; packet::WHITENING_DATA lib/packet_impl.cc:84-90
void whitening(uint8_t w[127])
{
  uint8_t r[7] = {1, 1, 1, 1, 1, 1, 1};
  for (int i = 0; i < 127; i++) {
    w[i] = r[6];
    const uint8_t fb = r[6];
    r[6] = r[5]; r[5] = r[4]; r[4] = r[3] ^ fb; r[3] = r[2]; r[2] = r[1]; r[1] = r[0]; r[0] = fb;
  }
}

// le_packet::freq2index, lib/packet_impl.cc:1285-1314
int le_index_of(double freq)
{
  if (!(freq >= 2402000000.0 && freq <= 2480000000.0)) return -1;
  if (!(std::fmod(freq, 2000000.0) < 5000.0)) return -1;
  const int chan = (int)((freq - 2402000000.0) / 2000000.0);
  if (chan < 0 || chan > 39) return -1;
  if (chan == 0) return 37;
  if (chan == 12) return 38;
  if (chan == 39) return 39;
  return chan < 12 ? chan - 1 : chan - 2;
}

}  // namespace

void Rotator::generate(cf32 *dst, int n, int stride)
{
  for (int i = 0; i < n; i++) {
    dst[(size_t)i * stride] = phase;
    counter++;
    cf32 p{phase.re * incr.re - phase.im * incr.im, phase.re * incr.im + phase.im * incr.re};
    if ((counter % 512) == 0) {
      const float mag = std::hypot(p.re, p.im);
      p = {p.re / mag, p.im / mag};
    }
    phase = p;
  }
}

namespace {
int igcd(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// common part: factorisation, channel -> column maps, DFT matrix, kappa, for channel offsets f0 + c MHz
int pfb_common(PfbDesign &F, const Plan &P, double f0, int cols_per_thread)
{
  const double Md = P.fs / kSymbolRate;
  F.M = (int)std::llround(Md);
  if (std::fabs(Md - F.M) > 1e-9 || F.M < 2 || (F.M & 1)) return -1;
  F.D = P.D;
  if (2 * F.D != F.M || P.grid_per_slot <= 0) return -1;
  const int M = F.M;
  // M = N1 * N2 with gcd(N1, N2) = 1 and N1 in {4, 2, 1}
  F.N1 = 1;
  if (M % 4 == 0 && igcd(4, M / 4) == 1) F.N1 = 4;
  else if (M % 2 == 0 && igcd(2, M / 2) == 1) F.N1 = 2;
  F.N2 = M / F.N1;
  const int N1 = F.N1, N2 = F.N2;
  // channel c sits a_c + phi MHz from the centre frequency
  F.a0 = (int)std::floor(f0 + 1e-9);
  F.phi = f0 - F.a0;
  if (std::fabs(F.phi) < 1e-9) F.phi = 0;
  F.n2_of_rho.assign(N2, 0);
  for (int n2 = 0; n2 < N2; n2++) F.n2_of_rho[(N1 * n2) % N2] = n2;
  // columns: channels grouped by residue class k1 = bin mod N1, every class padded to CPC columns
  std::vector<std::vector<int>> cls(N1);
  for (int c = 0; c < P.nch; c++) cls[(((F.a0 + c) % M) + M) % M % N1].push_back(c);
  size_t mx = 0;
  for (auto &v : cls) mx = std::max(mx, v.size());
  F.CPC = (int)((mx + cols_per_thread - 1) / cols_per_thread) * cols_per_thread;
  if (F.CPC == 0) return -1;
  F.ncol = N1 * F.CPC;
  F.col_chan.assign(F.ncol, -1);
  F.chan_col.assign(P.nch, -1);
  F.WB.assign((size_t)N2 * F.ncol, cf32{0.0f, 0.0f});
  F.kappa.assign(F.ncol, cf32{1.0f, 0.0f});
  for (int k1 = 0; k1 < N1; k1++)
    for (size_t i = 0; i < cls[k1].size(); i++) {
      const int c = cls[k1][i], col = k1 * F.CPC + (int)i;
      F.col_chan[col] = c;
      F.chan_col[c] = col;
      const int a = F.a0 + c;
      const int k = ((a % M) + M) % M, k2 = k % N2;
      for (int n2 = 0; n2 < N2; n2++) {
        const double ang = -2.0 * M_PI * (double)((n2 * k2) % N2) / N2;
        F.WB[(size_t)n2 * F.ncol + col] = cf32{(float)std::cos(ang), (float)std::sin(ang)};
      }
      const double ak = -2.0 * M_PI * (double)((((long)a * F.D) % M + M) % M) / M;
      F.kappa[col] = cf32{(float)std::cos(ak), (float)std::sin(ak)};
    }
  return 0;
}
}  // namespace

int PfbDesign::design(const Plan &P, int tile_points, int tile_computed, int cols_per_thread)
{
  if (pfb_common(*this, P, (kBaseFreq + P.ch_lo * kChanWidth - P.fc) / 1e6, cols_per_thread)) return -1;
  Q = (P.Nc + M - 1) / M;
  if (Q > 8) return -1;
  Q = Q <= 7 ? 7 : 8;                                    // the kernel is instantiated for 7 and 8 taps per branch
  q_rows = Q;
  nfull = P.n_ddc / P.grid_per_slot;
  rem = P.n_ddc % P.grid_per_slot;
  tps = (P.grid_per_slot + tile_points - 1) / tile_points;
  span = (tile_computed - 1) * D + Q * M;
  // h'[k] = h[Nc-1-k] (the reversed taps multiply x[n0 + k]); branch r holds h'[r + M q]
  hq.assign((size_t)Q * M, 0.0f);
  for (int k = 0; k < P.Nc; k++) hq[(size_t)(k / M) * M + (k % M)] = P.chan_proto[P.Nc - 1 - k];
  return 0;
}

int PfbDesign::design_noise(const Plan &P, int cols_per_thread, int row_pad)
{
  if (pfb_common(*this, P, (kBaseFreq + P.ch_lo * kChanWidth + 790000.0 - P.fc) / 1e6, cols_per_thread)) return -1;
  Q = (P.Nn + M - 1) / M;
  q_rows = ((Q + row_pad - 1) / row_pad + 1) * row_pad;
  hq.assign((size_t)q_rows * M, 0.0f);
  for (int k = 0; k < P.Nn; k++) hq[(size_t)(k / M) * M + (k % M)] = P.noise_proto[P.Nn - 1 - k];
  return 0;
}

double nest_quadrature(int N, int s, int n_extra, int n_free, double omega_max, std::vector<float> &w)
{
  if (N < 16 || s < 1 || n_extra < 0 || n_free < 1 || !(omega_max > 0)) return -1.0;
  const int n_used = (N - 1) / s + 1 + n_extra;
  std::vector<int> idx;                                   // the free weights: n_free at either end (all, if they overlap)
  for (int m = 0; m < n_used; m++)
    if (m < n_free || m >= n_used - n_free) idx.push_back(m);
  const int nu = (int)idx.size(), nw = 400, rows = 2 * nw + nu;
  const long double reg = 1e-4L;                          // sqrt of the Tikhonov weight
  typedef long double ld;
  std::vector<ld> A((size_t)rows * nu, 0.0L), b((size_t)rows, 0.0L);
  std::vector<ld> Tre(nw), Tim(nw);
  for (int k = 0; k < nw; k++) {
    const ld om = (ld)omega_max * k / (nw - 1);
    // target sum_{j<N} e^{i om j} minus the contribution of the pinned weights
    ld tr = 0, ti = 0;
    for (int j = 0; j < N; j++) { tr += cosl(om * j); ti += sinl(om * j); }
    Tre[k] = tr; Tim[k] = ti;
    ld rr = tr, ri = ti;
    for (int m = 0; m < n_used; m++) { rr -= (ld)s * cosl(om * s * m); ri -= (ld)s * sinl(om * s * m); }
    for (int u = 0; u < nu; u++) {
      A[(size_t)(2 * k) * nu + u] = cosl(om * s * idx[u]);
      A[(size_t)(2 * k + 1) * nu + u] = sinl(om * s * idx[u]);
    }
    b[2 * k] = rr; b[2 * k + 1] = ri;
  }
  for (int u = 0; u < nu; u++) A[(size_t)(2 * nw + u) * nu + u] = reg;
  // Householder QR, column by column; b is transformed along
  for (int c = 0; c < nu; c++) {
    ld norm = 0;
    for (int r = c; r < rows; r++) norm += A[(size_t)r * nu + c] * A[(size_t)r * nu + c];
    norm = sqrtl(norm);
    if (norm == 0) return -2.0;
    const ld alpha = A[(size_t)c * nu + c] > 0 ? -norm : norm;
    std::vector<ld> v((size_t)rows, 0.0L);
    for (int r = c; r < rows; r++) v[r] = A[(size_t)r * nu + c];
    v[c] -= alpha;
    ld vv = 0;
    for (int r = c; r < rows; r++) vv += v[r] * v[r];
    if (vv == 0) continue;
    for (int cc = c; cc < nu; cc++) {
      ld dot = 0;
      for (int r = c; r < rows; r++) dot += v[r] * A[(size_t)r * nu + cc];
      const ld f = 2 * dot / vv;
      for (int r = c; r < rows; r++) A[(size_t)r * nu + cc] -= f * v[r];
    }
    ld dot = 0;
    for (int r = c; r < rows; r++) dot += v[r] * b[r];
    const ld f = 2 * dot / vv;
    for (int r = c; r < rows; r++) b[r] -= f * v[r];
  }
  std::vector<ld> dw((size_t)nu, 0.0L);
  for (int c = nu - 1; c >= 0; c--) {
    ld acc = b[c];
    for (int cc = c + 1; cc < nu; cc++) acc -= A[(size_t)c * nu + cc] * dw[cc];
    dw[c] = acc / A[(size_t)c * nu + c];
  }
  w.assign((size_t)n_used, (float)s);
  for (int u = 0; u < nu; u++) w[(size_t)idx[u]] = (float)((ld)s + dw[u]);
  // residual with the weights as stored (float)
  double worst = 0;
  for (int k = 0; k < nw; k++) {
    const ld om = (ld)omega_max * k / (nw - 1);
    ld rr = -Tre[k], ri = -Tim[k];
    for (int m = 0; m < n_used; m++) { rr += (ld)w[m] * cosl(om * s * m); ri += (ld)w[m] * sinl(om * s * m); }
    const double e = (double)sqrtl(rr * rr + ri * ri);
    if (e > worst) worst = e;
  }
  return worst;
}

// classic_packet::acgen, lib/packet_impl.cc:309-364: (64,30) BCH sync word.
// the (64,30) code word of 30 information bits (bit k = sync-word bit 34 + k: 24 LAP bits, then the 6 Barker bits) with
// the PN overlay applied: bits 0..33 parity, 34..63 information
uint64_t sync_from_info(uint32_t info30)
{
  // PN overlay p[0..63] for sync-word bit i (= access-code position 4+i)
  static const uint8_t pn_bytes[9] = {0x03, 0xF2, 0xA3, 0x3D, 0xD6, 0x9B, 0x12, 0x1C, 0x10};
  auto pn = [&](int pos) { return (pn_bytes[pos >> 3] >> (7 - (pos & 7))) & 1; };
  static const uint8_t gen[35] = {1,0,0,1,0,1,0,1,1,0,1,1,1,1,0,0,1,0,0,0,1,1,1,0,1,0,1,0,0,0,0,1,1,0,1};
  uint8_t info[30];
  for (int i = 0; i < 30; i++) info[i] = (info30 >> i) & 1;
  uint8_t reg[34] = {0};
  for (int i = 29; i >= 0; i--) {
    const uint8_t fb = (uint8_t)((info[i] ^ pn(38 + i)) ^ reg[33]);
    for (int j = 33; j > 0; j--) reg[j] = reg[j - 1] ^ (gen[j] & fb);
    reg[0] = gen[0] & fb;
  }
  uint64_t w = 0;
  for (int i = 0; i < 34; i++) w |= (uint64_t)(reg[i] ^ pn(4 + i)) << i;
  for (int i = 0; i < 30; i++) w |= (uint64_t)info[i] << (34 + i);
  return w;
}

uint64_t sync_word(uint32_t lap)
{
  // information bits 24..29: the Barker sequence selected by the LAP's most significant bit
  const int msb = (lap >> 23) & 1;
  const uint32_t bark = msb ? 0x13u : 0x2Cu;      // bits 0..5 = {1,1,0,0,1,0} / {0,0,1,1,0,1}
  return sync_from_info((lap & 0xffffffu) | (bark << 24));
}

int BchTables::build(int max_errors)
{
  if (max_errors < 0 || max_errors > 2) return -1;
  max_err = max_errors;
  const uint64_t c0 = sync_from_info(0);
  par.assign(4 * 256 + 1, 0);
  for (int byte = 0; byte < 4; byte++)
    for (int v = 0; v < 256; v++) {
      const uint32_t info = (uint32_t)v << (8 * byte);
      if (info >> 30) continue;
      par[(size_t)byte * 256 + v] = (sync_from_info(info) ^ c0) & kBchParityMask;
    }
  par[1024] = c0 & kBchParityMask;
  // every error pattern of weight 1..max_err on sync-word bits 0..57 with its syndrome (received parity ^ parity of the
  // received information bits), sorted by syndrome; distinct because the code's minimum distance is 14
  auto syndrome_of = [&](uint64_t e) { return (e ^ sync_from_info((uint32_t)(e >> 34)) ^ c0) & kBchParityMask; };
  std::vector<std::pair<uint64_t, uint64_t>> tab;
  if (max_err >= 1)
    for (int i = 0; i < 58; i++) tab.push_back({syndrome_of(1ull << i), 1ull << i});
  if (max_err >= 2)
    for (int i = 0; i < 58; i++)
      for (int j = i + 1; j < 58; j++) tab.push_back({syndrome_of((1ull << i) | (1ull << j)), (1ull << i) | (1ull << j)});
  std::sort(tab.begin(), tab.end());
  syn.clear(); err.clear();
  for (size_t i = 0; i < tab.size(); i++) {
    if (tab[i].first == 0 || (i && tab[i].first == tab[i - 1].first)) return -2;
    syn.push_back(tab[i].first); err.push_back(tab[i].second);
  }
  return 0;
}

int Plan::design(double fs_, double fc_, double squelch, int extra)
{
  fs = fs_; fc = fc_; squelch_db = squelch; extra_symbols = extra;
  if (!(fs >= 2e6) || !(fc > 0) || extra < 0) return -1;
  const double sps = fs / kSymbolRate;                       // multi_block.cc:56
  const double slot = (int)kSlotSymbols * sps;               // :58
  S = (int)slot;
  int hist = (int)(1 * slot);                                // :59
  chan_proto = hann_lowpass(fs, 500000, 300000);             // :63-69
  noise_proto = hann_lowpass(fs, 22500, 10000);              // :71-79
  Nc = (int)chan_proto.size();
  Nn = (int)noise_proto.size();
  D = (int)sps / 2;                                          // :82
  if (D < 1) return -1;
  const double csps = sps / D;                               // :83
  // windows must share one decimation grid: S = 625 sps must be a multiple of D = (int)sps / 2.  True for the even
  // integer Msps rates (every BASELINE configuration); NOT for e.g. 5, 7, 9 or 13 Msps, which the reference's block
  // accepts -- those rates are rejected here (documented in include/btb200.h, btb200_create)
  if (S % D != 0) return -2;
  grid_per_slot = S / D;

  // set_channels(), :306-342
  const double center = (fc - kBaseFreq) / kChanWidth;
  const double bw = fs / kChanWidth;
  int lo = (int)(center - bw / 2 + 0.9 / 2 + 1);
  if (lo < 0) lo = 0;
  int hi = (int)(center + bw / 2 - 0.9 / 2);
  if (hi > 78) hi = 78;
  ch_lo = lo; ch_hi = hi; nch = hi - lo + 1;
  if (nch <= 0) return -1;

  chan_rtaps.resize((size_t)nch * Nc);
  noise_rtaps.resize((size_t)nch * Nn);
  chan_incr.resize(nch);
  noise_incr.resize(nch);
  for (int c = 0; c < nch; c++) {
    const double f = kBaseFreq + (lo + c) * kChanWidth;
    translate(chan_proto, f - fc, fs, D, &chan_rtaps[(size_t)c * Nc], &chan_incr[c]);
    translate(noise_proto, f + 790000.0 - fc, fs, D, &noise_rtaps[(size_t)c * Nn], &noise_incr[c]);
  }

  demod_gain = (float)(csps / M_PI_2);                       // :88
  gain_mu = 0.175f; mu0 = 0.32f; omega_lim = 0.005f;         // :91-93
  omega_mid = (float)csps;                                   // :94,96
  gain_omega = (float)(.25 * gain_mu * gain_mu);             // :95

  const int chist = Nc + D * kMmseTaps;                      // :101-103
  if (chist > Nn) { hist += chist; fcs = 0; fns = chist - Nn; }
  else            { hist += Nn; fns = 0; fcs = Nn - chist; }
  H = (int)(hist + extra * sps);                             // :299-303

  int avail = H - (Nc - 1) - fcs;                            // :194
  int nd = avail - Nc + 1; if (nd < 0) nd = 0;
  n_ddc = nd / D;                                            // :200
  int nn = (int)slot - Nn + 1; if (nn < 0) nn = 0;
  n_noise = nn / D;                                          // :269
  n_dem = n_ddc - 1;
  if (n_ddc < 16 || n_noise < 1) return -1;

  mmse = mmse_table();
  atan_tab.resize(257);
  for (int i = 0; i < 256; i++) atan_tab[i] = (float)std::atan(i / 255.0);
  atan_tab[256] = atan_tab[255];

  // acgen is affine over GF(2) in the LAP: sync(lap) = C ^ T0[b0] ^ T1[b1] ^ T2[b2]
  ac_lut.assign(3 * 256 + 1, 0);
  const uint64_t c0 = sync_word(0);
  for (int byte = 0; byte < 3; byte++)
    for (int v = 0; v < 256; v++)
      ac_lut[byte * 256 + v] = sync_word((uint32_t)v << (8 * byte)) ^ c0;
  ac_lut[768] = c0;

  // LE per-channel whitening of the 16 header bits (packet_impl.cc:1446-1483)
  uint8_t w[127];
  whitening(w);
  le_white16.assign((size_t)nch * 16, 0);
  le_index.assign(nch, -1);
  for (int c = 0; c < nch; c++) {
    const int idx = le_index_of(kBaseFreq + (lo + c) * kChanWidth);
    le_index[c] = (int8_t)idx;
    if (idx < 0) continue;
    // whitening register seeded with (1, index MSB..LSB); find its place in the cycle
    uint8_t r[7];
    r[0] = 1;
    for (int b = 0; b < 6; b++) r[1 + b] = (idx >> (5 - b)) & 1;
    uint8_t first7[7], q[7];
    std::memcpy(q, r, 7);
    for (int i = 0; i < 7; i++) {
      first7[i] = q[6];
      const uint8_t fb = q[6];
      q[6] = q[5]; q[5] = q[4]; q[4] = q[3] ^ fb; q[3] = q[2]; q[2] = q[1]; q[1] = q[0]; q[0] = fb;
    }
    int off = 0;
    for (; off < 127; off++) {
      bool ok = true;
      for (int i = 0; i < 7 && ok; i++) ok = w[(off + i) % 127] == first7[i];
      if (ok) break;
    }
    for (int i = 0; i < 16; i++) le_white16[(size_t)c * 16 + i] = w[(off + i) % 127];
  }
  return 0;
}

}  // namespace btb200
