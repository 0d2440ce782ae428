// multi_LAP_impl.cc -- LAP printer block (lib/multi_LAP_impl.cc:65-114 of the reference).
// Window geometry: history + 68 symbols (multi_LAP_impl.cc:54).  The reference searches with libbtbb's
// btbb_find_ac(symbols, latest_ac, LAP_ANY, max_ac_errs = 1, &pkt) (:74, :93) -- an external library that is not part
// of the reference tree.  This block runs the same test on the GPU (BTB200_SEARCH_BR_BCH, csrc/rx_math.cuh: Barker
// correction, syndrome decoding of the sync word's (64,30) code with one corrected bit), restated from libbtbb's
// published algorithm: PARITY UNPINNED (no libbtbb to diff against).  One report per channel-window, like
// the reference's single call; `err` is the number of corrected bits (btbb_packet_get_ac_errors).
// BTB200_AC_SEARCH=sniff_ac selects classic_packet::sniff_ac semantics instead (err = check_ac's count).
#include "multi_LAP_impl.h"
#include "btb200.h"
#include <cstdio>

namespace gr {
namespace bluetooth {

multi_LAP::sptr multi_LAP::make(double sample_rate, double center_freq, double squelch_threshold)
{
  return gnuradio::get_initial_sptr(new multi_LAP_impl(sample_rate, center_freq, squelch_threshold));
}

multi_LAP_impl::multi_LAP_impl(double sample_rate, double center_freq, double squelch_threshold)
    : gr::sync_block("bluetooth multi LAP block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      multi_block(sample_rate, center_freq, squelch_threshold, 68, BTB200_SEARCH_BR | BTB200_SEARCH_BR_BCH, false, BTB200_BCH_ANY(1))
{
}

multi_LAP_impl::~multi_LAP_impl() {}

int multi_LAP_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  return process_windows(noutput_items, input_items);
}

void multi_LAP_impl::handle_hit(const btb200_hit &hit, const char *, int, double)
{
  if ((int)hit.slot == d_last_slot && (int)hit.channel == d_last_channel) return;   // one report per channel-window
  d_last_slot = (int)hit.slot;
  d_last_channel = (int)hit.channel;
  std::printf("GOT PACKET: ch=%d, LAP=%06x, err=%u at time slot %d\n", hit.channel, hit.lap, hit.ac_errors, (int)hit.slot);
}

}  // namespace bluetooth
}  // namespace gr
