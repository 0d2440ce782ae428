// multi_sniffer_impl.cc -- the all-piconet sniffer block on the B200 path.
//
// work() (lib/multi_sniffer_impl.cc:82-166 of the reference) is one call into the C ABI; the
// per-packet handlers print what the reference's ac()/aa() print for the fields the hot path
// produces.  The host packet layer behind them (header decode, UAP/CLK discovery:
// lib/packet_impl.cc:512-1275, lib/piconet_impl.cc) is the next row of SURVEY.md 8f and is not
// part of this round: a BR packet is classified ID / has-header exactly like
// classic_packet_impl::header_present() (lib/packet_impl.cc:1205-1242) and reported as such.
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
  if (d_tun)
    std::fprintf(stderr, "warning: the TUN/Wireshark interface (lib/tun.cc) is not part of the B200 path, disabling it\n");
}

multi_sniffer_impl::~multi_sniffer_impl() {}

int multi_sniffer_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  return process_windows(noutput_items, input_items);
}

void multi_sniffer_impl::handle_hit(const btb200_hit &hit, const char *symbols, int n_symbols, double freq)
{
  if (hit.kind == 0) ac(symbols, hit.n_symbols, n_symbols, freq, hit.snr, hit.lap);
  else aa(symbols, hit.n_symbols, n_symbols, freq, hit.snr);
}

// classic_packet_impl::header_present(), lib/packet_impl.cc:1205-1242
static bool header_present(const char *sym, int length)
{
  if (length < 126) return false;
  const char *s = sym + 67;
  int be = 0;
  const char msb = s[0];
  be += s[1] ^ !msb;
  be += s[2] ^ msb;
  be += s[3] ^ !msb;
  be += s[4] ^ msb;
  s += 5;
  for (int a = 0; a < 54; a += 3)
    be += ((s[a] ^ s[a + 1]) | (s[a + 1] ^ s[a + 2]) | (s[a + 2] ^ s[a]));
  return be < 5;   // ID_THRESHOLD
}

void multi_sniffer_impl::ac(const char *symbols, int len, int sym_avail, double freq, double snr, uint32_t lap)
{
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  const int channel = (int)((freq - 2402000000.0) / 1000000.0);
  std::printf("time %6d, snr=%.1f, channel %2d, LAP %06x ", clkn, snr, channel, lap);
  const int n = len < sym_avail ? len : sym_avail;
  if (header_present(symbols, n > 3125 ? 3125 : n))
    std::printf("header\n");      // reference: discover()/decode() output follows here
  else
    std::printf("ID\n");
}

void multi_sniffer_impl::aa(const char *symbols, int len, int sym_avail, double freq, double snr)
{
  (void)len;
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  uint32_t aa = 0;
  for (int i = 0; i < 32 && 8 + i < sym_avail; i++) aa |= (uint32_t)(symbols[8 + i] & 1) << i;
  std::printf("time %6d, snr=%.1f, BTLE AA=%08x freq=%.0f\n", clkn, snr, aa, freq / 1e6);
}

}  // namespace bluetooth
}  // namespace gr
