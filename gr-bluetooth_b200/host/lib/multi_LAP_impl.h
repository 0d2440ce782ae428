#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_LAP_IMPL_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_LAP_IMPL_H
#include "gr_bluetooth/multi_LAP.h"

namespace gr {
namespace bluetooth {

class multi_LAP_impl : virtual public multi_LAP {
 private:
  void handle_hit(const btb200_hit &hit, const char *symbols, int n_symbols, double freq);
  int d_last_slot = -1, d_last_channel = -1;

 public:
  multi_LAP_impl(double sample_rate, double center_freq, double squelch_threshold);
  ~multi_LAP_impl();
  int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}  // namespace bluetooth
}  // namespace gr
#endif
