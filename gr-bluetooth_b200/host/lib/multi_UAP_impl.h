#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_UAP_IMPL_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_UAP_IMPL_H
#include "gr_bluetooth/multi_UAP.h"
#include "bt_host.h"
#include <memory>
#include <vector>

struct btb200_chan_result;

namespace gr {
namespace bluetooth {

class multi_UAP_impl : virtual public multi_UAP {
 private:
  std::unique_ptr<btb200_host::UapHost> d_host;
  std::vector<btb200_chan_result> d_res;
  std::vector<uint8_t> d_symbols;
  void handle_hit(const btb200_hit &, const char *, int, double) {}

 public:
  multi_UAP_impl(double sample_rate, double center_freq, double squelch_threshold, int LAP);
  ~multi_UAP_impl();
  int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}  // namespace bluetooth
}  // namespace gr
#endif
