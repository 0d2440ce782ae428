/*
 * oracle/shim/gr_shim.cc -- TEST INFRASTRUCTURE.
 * Out-of-line parts of the GNU Radio stand-in (arithmetic lives in
 * oracle/gr_arith.h) plus the capture hooks.
 */
#include <gnuradio/sync_block.h>
#include <gnuradio/math.h>
#include <gnuradio/filter/firdes.h>
#include <gnuradio/filter/freq_xlating_fir_filter_ccf.h>
#include <gnuradio/filter/mmse_fir_interpolator_ff.h>
#include <gnuradio/blocks/complex_to_mag_squared.h>
#include "btref_hooks.h"
#include <vector>

btref_hooks g_btref = { 0, 0, NULL, NULL, NULL, 0, 0, -1, -1, 0 };

static float g_atan_table[257];
static float g_mmse[GRA_MMSE_NSTEPS + 1][GRA_MMSE_NTAPS];
static bool  g_tables_ready = false;
static int   g_next_ddc_id = 0;

static std::vector<float> g_soft, g_mu;
static const float *g_demod_base = NULL;

static void tables_init()
{
  if (g_tables_ready) return;
  gra_atan_table(g_atan_table);
  gra_mmse_table(g_mmse);
  g_tables_ready = true;
}

static void rec_header(uint32_t type, int32_t id, uint32_t n)
{
  uint32_t h[4] = { type, (uint32_t)g_btref.call_index, (uint32_t)id, n };
  fwrite(h, sizeof h, 1, g_btref.dump);
}

static bool heavy_on()
{
  return g_btref.dump && g_btref.call_index >= g_btref.heavy_from &&
         g_btref.call_index < g_btref.heavy_to;
}

void btref_flush_symbols(void)
{
  if (!g_soft.empty() && g_btref.dump) {
    uint32_t n = (uint32_t)g_soft.size();
    rec_header(BTREF_REC_BITS, g_btref.cur_chan_ddc_id, n);
    std::vector<unsigned char> bits(n);
    for (uint32_t i = 0; i < n; i++) bits[i] = (g_soft[i] < 0) ? 0 : 1;   /* multi_block.cc:171-178 */
    fwrite(bits.data(), 1, n, g_btref.dump);
    if (heavy_on()) {
      rec_header(BTREF_REC_SOFT, g_btref.cur_chan_ddc_id, n);
      fwrite(g_soft.data(), sizeof(float), n, g_btref.dump);
      rec_header(BTREF_REC_MU, g_btref.cur_chan_ddc_id, n);
      fwrite(g_mu.data(), sizeof(float), n, g_btref.dump);
    }
  }
  g_soft.clear();
  g_mu.clear();
  g_demod_base = NULL;
}

namespace gr {

float fast_atan2f(float y, float x)
{
  tables_init();
  return gra_fast_atan2f(g_atan_table, y, x);
}

namespace filter {

freq_xlating_fir_filter_ccf::sptr
freq_xlating_fir_filter_ccf::make(int decimation, const std::vector<float> &taps,
                                  double center_freq, double fs)
{
  return sptr(new freq_xlating_fir_filter_ccf(decimation, taps, center_freq, fs));
}

freq_xlating_fir_filter_ccf::freq_xlating_fir_filter_ccf(int decimation,
    const std::vector<float> &taps, double center_freq, double fs)
{
  gra_fxlat_init(&d_f, decimation, taps.data(), (int)taps.size(), center_freq, fs);
  d_id = g_next_ddc_id++;
  d_center_freq = center_freq;
}

freq_xlating_fir_filter_ccf::~freq_xlating_fir_filter_ccf() { gra_fxlat_free(&d_f); }

int freq_xlating_fir_filter_ccf::work(int noutput_items, gr_vector_const_void_star &in,
                                      gr_vector_void_star &out)
{
  bool is_channel = (d_id % 2) == 0;     /* set_channels(): channel ddc then noise ddc per channel */
  if (is_channel) {
    btref_flush_symbols();
    g_btref.cur_chan_ddc_id = d_id;
    g_btref.cur_chan_nout = noutput_items;
    if (g_btref.on_channel_ddc) g_btref.on_channel_ddc(g_btref.user);
  }
  g_btref.cur_ddc_id = d_id;
  if (g_btref.stateless) {
    d_f.phase.re = 1.0f; d_f.phase.im = 0.0f; d_f.counter = 0;
  }
  int n = gra_fxlat_work(&d_f, noutput_items, (const gra_c32 *)in[0], (gra_c32 *)out[0]);
  if (heavy_on()) {
    rec_header(BTREF_REC_DDC, d_id, (uint32_t)n);
    fwrite(out[0], sizeof(gra_c32), (size_t)n, g_btref.dump);
  }
  return n;
}

mmse_fir_interpolator_ff::mmse_fir_interpolator_ff() { tables_init(); }

float mmse_fir_interpolator_ff::interpolate(const float input[], float mu) const
{
  if (g_soft.empty()) {
    /* first interpolation of a channel-window: input == demod_out (ii = 0, multi_block.cc:139) */
    g_demod_base = input;
    if (heavy_on()) {
      uint32_t n = (uint32_t)(g_btref.cur_chan_nout - 1);
      rec_header(BTREF_REC_DEMOD, g_btref.cur_chan_ddc_id, n);
      fwrite(input, sizeof(float), n, g_btref.dump);
    }
  }
  int bad = 0;
  float r = gra_mmse_interpolate(g_mmse, input, mu, &bad);
  if (bad) { fprintf(stderr, "shim: interpolate imu out of range (mu=%g)\n", mu); abort(); }
  if (g_btref.dump) { g_soft.push_back(r); g_mu.push_back(mu); }
  return r;
}

} // namespace filter

namespace blocks {

int complex_to_mag_squared::work(int noutput_items, gr_vector_const_void_star &in,
                                 gr_vector_void_star &out)
{
  const gra_c32 *x = (const gra_c32 *)in[0];
  float *y = (float *)out[0];
  double energy = 0.0;
  for (int i = 0; i < noutput_items; i++) {
    y[i] = gra_mag2(x[i]);
    energy += y[i];                       /* same accumulate as multi_block.cc:214-218 */
  }
  if (g_btref.dump) {
    energy /= noutput_items;
    rec_header(BTREF_REC_ENERGY, g_btref.cur_ddc_id, (uint32_t)noutput_items);
    fwrite(&energy, sizeof energy, 1, g_btref.dump);
  }
  return noutput_items;
}

} // namespace blocks
} // namespace gr
