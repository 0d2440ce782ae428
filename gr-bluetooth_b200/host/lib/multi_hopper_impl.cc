// multi_hopper_impl.cc -- the target-LAP follower block (lib/multi_hopper_impl.cc:76-209 of the reference).
//
// One slot per work() call, always on the shared (chained) clock-recovery state: which channels a call
// demodulates depends on the piconet state the previous call left behind -- every channel until CLK1-27
// is known (ending early at the first packet of the target LAP that has a header), afterwards only the
// predicted hop channel.  The channel loop runs on the GPU (btb200_process_channels), the piconet logic
// (UAP/CLK1-6 from headers, hop reversal for CLK1-27, packet decode) in bt_host.{h,cc}.
#include "multi_hopper_impl.h"
#include "btb200.h"
#include <cstdio>
#include <stdexcept>
#include <string>

namespace gr {
namespace bluetooth {

multi_hopper::sptr multi_hopper::make(double sample_rate, double center_freq, double squelch_threshold, int LAP,
                                      bool aliased, bool tun)
{
  return gnuradio::get_initial_sptr(new multi_hopper_impl(sample_rate, center_freq, squelch_threshold, LAP, aliased, tun));
}

multi_hopper_impl::multi_hopper_impl(double sample_rate, double center_freq, double squelch_threshold, int LAP,
                                     bool aliased, bool tun)
    : gr::sync_block("bluetooth multi hopper block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      multi_block(sample_rate, center_freq, squelch_threshold, 3125, BTB200_SEARCH_BR, /*force_chained=*/true)
{
  if (aliased) throw std::runtime_error("multi_hopper: the aliased receiver mode is not supported on the B200 path");
  const int lo = (int)((d_low_freq - 2402000000.0) / 1e6), hi = (int)((d_high_freq - 2402000000.0) / 1e6);
  d_host.reset(new btb200_host::HopperHost((uint32_t)LAP, aliased, lo, hi));
  if (tun) d_host->set_tun_fd(btb200_host::open_tun_output());     /* lib/multi_hopper_impl.cc:56-64 */
  d_res.resize((size_t)(hi - lo + 1));
  d_symbols.resize((size_t)(hi - lo + 1) * 3125);
}

multi_hopper_impl::~multi_hopper_impl() {}

int multi_hopper_impl::work(int, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  const int S = (int)d_samples_per_slot;
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  const auto pl = d_host->plan(clkn);
  if (pl.n_channels > 0) {
    int rc = btb200_process_channels(d_ctx, reinterpret_cast<const float *>(input_items[0]), history(), clkn,
                                     pl.first_channel, pl.n_channels, pl.stop_lap, d_res.data(), d_symbols.data(),
                                     d_symbols.size());
    if (rc != BTB200_OK)
      throw std::runtime_error(std::string("btb200_process_channels: ") + btb200_strerror(rc) + " (" + btb200_last_error(d_ctx) + ")");
    for (int q = 0; q < pl.n_channels; q++) {
      const btb200_chan_result &r = d_res[(size_t)q];
      if (!r.processed) break;
      if (r.ac_index < 0) continue;
      const char *sp = reinterpret_cast<const char *>(d_symbols.data() + r.sym_offset);
      if (pl.hopalong) d_host->hop_packet(pl, sp, (int)r.sym_count);
      else if (d_host->scan_packet(clkn, r.channel, sp, (int)r.sym_count)) break;
    }
  }
  d_cumulative_count += (uint64_t)S;
  return S;
}

}  // namespace bluetooth
}  // namespace gr
