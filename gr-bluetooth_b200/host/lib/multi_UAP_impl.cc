// multi_UAP_impl.cc -- the UAP-discovery block on the B200 path.
//
// The reference's work() (lib/multi_UAP_impl.cc:67-124) is the hopper's scan loop with one difference: per slot it
// walks the channels in order on the shared clock-recovery state, looks for the TARGET piconet's access code and,
// at the first such packet that carries a header, feeds it to the UAP/CLK1-6 discovery and leaves the channel
// loop (`break`, :112); once the UAP is determined the process ends (`exit(0)`, :110).  The channel loop with its
// early exit runs on the GPU (btb200_process_channels, chained state); the discovery is the native
// Piconet::uap_from_header of bt_host.cc (the same arithmetic as basic_rate_piconet_impl::UAP_from_header,
// lib/piconet_impl.cc:433-517).
//
// Differences from the reference, stated: it calls libbtbb (btbb_find_ac for the piconet's LAP with up to 2 bit errors,
// btbb_uap_from_header) -- an external library that is not part of the reference tree, so parity there is unpinned.
// The search is the same test restated (BTB200_SEARCH_BR_BCH with the LAP: Hamming distance of the 64 sync-word
// symbols to the LAP's sync word <= 2; BTB200_AC_SEARCH=sniff_ac selects sniff_ac's rule), the discovery is the
// native one, and the packets get their real CLKN.  Where the reference calls exit(0), work() returns -1
// (gr::block::WORK_DONE).
#include "multi_UAP_impl.h"
#include "btb200.h"
#include <stdexcept>
#include <string>

namespace gr {
namespace bluetooth {

multi_UAP::sptr multi_UAP::make(double sample_rate, double center_freq, double squelch_threshold, int LAP)
{
  return gnuradio::get_initial_sptr(new multi_UAP_impl(sample_rate, center_freq, squelch_threshold, LAP));
}

multi_UAP_impl::multi_UAP_impl(double sample_rate, double center_freq, double squelch_threshold, int LAP)
    : gr::sync_block("bluetooth multi UAP block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      multi_block(sample_rate, center_freq, squelch_threshold, 3125, BTB200_SEARCH_BR | BTB200_SEARCH_BR_BCH, /*force_chained=*/true,
                  BTB200_BCH_LAP((unsigned)LAP, 2))
{
  const int lo = (int)((d_low_freq - 2402000000.0) / 1e6), hi = (int)((d_high_freq - 2402000000.0) / 1e6);
  d_host.reset(new btb200_host::UapHost((uint32_t)LAP, lo, hi));
  d_res.resize((size_t)(hi - lo + 1));
  d_symbols.resize((size_t)(hi - lo + 1) * 3125);
}

multi_UAP_impl::~multi_UAP_impl() {}

int multi_UAP_impl::work(int, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  if (d_host->done()) return -1;                     /* WORK_DONE: the reference has exited by now */
  const int S = (int)d_samples_per_slot;
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  const int n = d_host->ch_hi() - d_host->ch_lo() + 1;
  int rc = btb200_process_channels(d_ctx, reinterpret_cast<const float *>(input_items[0]), history(), clkn,
                                   d_host->ch_lo(), n, d_host->lap(), d_res.data(), d_symbols.data(), d_symbols.size());
  if (rc != BTB200_OK)
    throw std::runtime_error(std::string("btb200_process_channels: ") + btb200_strerror(rc) + " (" + btb200_last_error(d_ctx) + ")");
  for (int q = 0; q < n; q++) {
    const btb200_chan_result &r = d_res[(size_t)q];
    if (!r.processed) break;
    if (r.ac_index < 0) continue;
    const char *sp = reinterpret_cast<const char *>(d_symbols.data() + r.sym_offset);
    if (d_host->packet(clkn, r.channel, sp, (int)r.sym_count)) break;
  }
  d_cumulative_count += (uint64_t)S;
  return d_host->done() ? -1 : S;
}

}  // namespace bluetooth
}  // namespace gr
