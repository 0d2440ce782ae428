// multi_sniffer_impl.cc -- the all-piconet sniffer block on the B200 path.
//
// work() (lib/multi_sniffer_impl.cc:82-166 of the reference) is one call into the C ABI for a
// batch of slots; every returned hit then runs through the reference's per-packet call chain
// ac()/aa() -> id / discover / decode / recall / fhs (lib/multi_sniffer_impl.cc:169-365),
// restated here over the native packet layer of bt_host.{h,cc}.  The text printed is the
// reference's, line for line (tests: btrx_b200 stdout digests == the reference's).
#include "multi_sniffer_impl.h"
#include "btb200.h"
#include <cstdio>
#include <cstring>

namespace gr {
namespace bluetooth {

multi_sniffer::sptr multi_sniffer::make(double sample_rate, double center_freq, double squelch_threshold, bool tun)
{
  return gnuradio::get_initial_sptr(new multi_sniffer_impl(sample_rate, center_freq, squelch_threshold, tun));
}

multi_sniffer_impl::multi_sniffer_impl(double sample_rate, double center_freq, double squelch_threshold, bool tun)
    : gr::sync_block("bluetooth multi sniffer block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      multi_block(sample_rate, center_freq, squelch_threshold, 3125, BTB200_SEARCH_BR | BTB200_SEARCH_LE),
      d_tun(tun)
{
  if (d_tun) {
    /* Tun interface (lib/multi_sniffer_impl.cc:63-72): a failure only warns */
    const int fd = btb200_host::open_tun_output();
    if (fd < 0) d_tun = false;
    d_host.set_tun_fd(fd);
  }
}

multi_sniffer_impl::~multi_sniffer_impl() {}

int multi_sniffer_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  return process_windows(noutput_items, input_items);
}

void multi_sniffer_impl::handle_hit(const btb200_hit &hit, const char *symbols, int n_symbols, double freq)
{
  // the ABI hands over min(len, 3125) symbols -- all that classic_packet::make / le_packet::make keep
  const int len = hit.n_symbols < n_symbols ? hit.n_symbols : n_symbols;
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);   // :173
  if (hit.kind == 0) d_host.ac(symbols, len, clkn, freq, hit.snr);
  else d_host.aa(symbols, len, clkn, freq, hit.snr);
}

}  // namespace bluetooth
}  // namespace gr
