// plan.hpp -- host-side design step of the B200 receive path.
//
// Everything multi_block's constructor derives from (sample_rate, center_freq)
// -- lib/multi_block.cc:40-120, 299-342 of the reference -- plus the GNU Radio
// 3.7 design arithmetic it calls (firdes::low_pass, freq_xlating_fir_filter_ccf
// tap rotation and rotator, mmse_fir_interpolator_ff taps, fast_atan2f table;
// SURVEY.md Appendix A).  Computed once on the host in the same precision and
// operation order as the reference so the uploaded tables are bit-identical.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace btb200 {

struct cf32 { float re, im; };

struct Plan {
  // make() arguments
  double fs = 0, fc = 0, squelch_db = 0;
  int extra_symbols = 3125;
  // derived geometry
  int S = 0;              // samples per slot
  int H = 0;              // history = window length
  int D = 1;              // DDC decimation
  int Nc = 0, Nn = 0;     // prototype lengths
  int fcs = 0, fns = 0;   // first channel / noise sample inside the window
  int ch_lo = 0, ch_hi = -1, nch = 0;
  int n_ddc = 0;          // channel DDC outputs per window
  int n_noise = 0;        // noise DDC outputs per window
  int n_dem = 0;          // demod outputs per window (n_ddc - 1)
  int grid_per_slot = 0;  // S / D: decimated grid points per slot
  float demod_gain = 0;
  // M&M constants (lib/multi_block.cc:91-96)
  float gain_mu = 0, gain_omega = 0, omega_mid = 0, omega_lim = 0, mu0 = 0;
  // tables
  std::vector<float> chan_proto, noise_proto;
  std::vector<cf32>  chan_rtaps;    // [nch][Nc], reversed: rtaps[k] multiplies x[n0+k]
  std::vector<cf32>  noise_rtaps;   // [nch][Nn]
  std::vector<cf32>  chan_incr, noise_incr;   // rotator increments [nch]
  std::vector<float> mmse;          // [129][8]
  std::vector<float> atan_tab;      // [257]
  std::vector<uint64_t> ac_lut;     // [3][256] affine sync-word tables, then constant
  std::vector<uint8_t>  le_white16; // [nch][16] whitening bits for the LE header, per channel
  std::vector<int8_t>   le_index;   // [nch] LE channel index or -1

  int design(double fs, double fc, double squelch_db, int extra_symbols);   // 0 or negative error
};

// Host tables of the polyphase channelizer (throughput mode, rx_pfb.cu): the per-channel DDCs of
// lib/multi_block.cc:329-341 as one real-tap bank of M = fs / 1 MHz branches + a Good-Thomas DFT (M = N1 * N2)
// evaluated at the channel bins.  See rx_pfb.cu for the derivation.
struct PfbDesign {
  int M = 0, D = 0, Q = 0, N1 = 1, N2 = 1;
  int nfull = 0, rem = 0, tps = 0, CPC = 0, ncol = 0, span = 0;
  int a0 = 0;                        // integer MHz offset of the lowest channel
  double phi = 0;                    // fractional MHz offset common to all channels
  std::vector<float> hq;             // [Q][M]
  std::vector<int> n2_of_rho;        // [N2]
  std::vector<cf32> WB;              // [N2][ncol]
  std::vector<int> col_chan;         // [ncol] channel index or -1
  std::vector<int> chan_col;         // [nch]
  std::vector<cf32> kappa;           // [ncol]
  // 0, or -1 when the configuration does not fit the model (non-integer or odd samples per MHz, > 8 taps per branch)
  int design(const Plan &P, int tile_points, int tile_computed, int cols_per_thread);
  // Same structure for the off-channel ("noise") DDCs of check_snr (lib/multi_block.cc:253-296): prototype
  // noise_proto (Nn taps, Q = ceil(Nn / M) taps per branch, rows padded with zeros to a multiple of `row_pad` plus
  // `row_pad` more), channel offsets + 790 kHz.  The pre-rotation by phi is applied to the input by a separate pass.
  int design_noise(const Plan &P, int cols_per_thread, int row_pad);
  int q_rows = 0;                    // rows of hq (noise variant: padded)
};

// Quadrature weights of the sub-sampled off-channel energy sum (rx_nest.cu, throughput mode).  check_snr sums |y_j|^2
// over the N outputs j < N of the noise DDC (lib/multi_block.cc:279-284); |y|^2 is band-limited far below the
// reference's output rate (the Hann low-pass of 22.5 kHz + 10 kHz transition leaves y within ~+-45 kHz at 2 Msps), so
//     sum_{j<N} g(j)  ~=  sum_{m<n_used} w[m] g(s m),     n_used = (N-1)/s + 1 + n_extra
// for every g band-limited to |omega| <= omega_max (radians per output).  Interior weights are s; the n_free weights
// at either end are the least-squares solution over a grid of complex exponentials in the band (Householder QR,
// Tikhonov-regularised towards s).  Returns the largest residual |sum_m w e^{i w s m} - sum_j e^{i w j}| over the
// grid (the worst-case error for a unit-amplitude in-band tone; the sum itself is N), or a negative value on error.
double nest_quadrature(int N, int s, int n_extra, int n_free, double omega_max, std::vector<float> &w);

// Free-running rotator of one DDC object (GNU Radio's gr::blocks::rotator):
// phase multiplies output i, then advances; renormalised every 512 outputs.
struct Rotator {
  cf32 phase{1.0f, 0.0f};
  cf32 incr{1.0f, 0.0f};
  unsigned counter = 0;
  void reset() { phase = {1.0f, 0.0f}; counter = 0; }
  // writes the n phases that multiply the next n outputs, advancing the state
  void generate(cf32 *dst, int n, int stride);
};

// 64-bit sync word of a LAP, bit i = access-code symbol 4+i
// (restates classic_packet::acgen, lib/packet_impl.cc:309-364)
uint64_t sync_word(uint32_t lap);
// the same code word generator for arbitrary information bits (bit k = sync-word bit 34 + k)
uint64_t sync_from_info(uint32_t info30);

// Tables of the libbtbb-style access-code search (BTB200_SEARCH_BR_BCH; what multi_LAP / multi_UAP call through
// btbb_find_ac, lib/multi_LAP_impl.cc:93, lib/multi_UAP_impl.cc:95): syndrome decoding of the (64,30) code the sync
// word is built on.  par: parity of the information bits as four byte tables + constant (the generator is affine);
// syn/err: the syndromes of every error pattern of weight <= max_err on sync-word bits 0..57, sorted, with the patterns.
constexpr uint64_t kBchParityMask = (1ull << 34) - 1;
struct BchTables {
  int max_err = 0;
  std::vector<uint64_t> par;   // [4][256] + constant
  std::vector<uint64_t> syn;   // sorted
  std::vector<uint64_t> err;   // pattern of syn[i]
  int build(int max_errors);   // 0, or negative (max_errors outside 0..2)
};

}  // namespace btb200
