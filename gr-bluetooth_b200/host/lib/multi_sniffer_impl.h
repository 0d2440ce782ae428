#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_SNIFFER_IMPL_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_SNIFFER_IMPL_H
#include "gr_bluetooth/multi_sniffer.h"
#include "bt_host.h"
#include <map>
#include <memory>

namespace gr {
namespace bluetooth {

class multi_sniffer_impl : virtual public multi_sniffer {
 private:
  bool d_tun;
  btb200_host::SnifferHost d_host;       // the reference's ac()/aa() call chain (lib/multi_sniffer_impl.cc:169-365)
  void handle_hit(const btb200_hit &hit, const char *symbols, int n_symbols, double freq);

 public:
  multi_sniffer_impl(double sample_rate, double center_freq, double squelch_threshold, bool tun);
  ~multi_sniffer_impl();
  int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}  // namespace bluetooth
}  // namespace gr
#endif
