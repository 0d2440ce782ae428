#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_HOPPER_IMPL_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_HOPPER_IMPL_H
#include "gr_bluetooth/multi_hopper.h"
#include "bt_host.h"
#include <memory>
#include <vector>

struct btb200_chan_result;

namespace gr {
namespace bluetooth {

class multi_hopper_impl : virtual public multi_hopper {
 private:
  std::unique_ptr<btb200_host::HopperHost> d_host;
  std::vector<btb200_chan_result> d_res;
  std::vector<uint8_t> d_symbols;
  // throughput variant of the hop-along phase (BTB200_MM_MODE=stateless): a second, stateless context processes
  // batches of slots with one masked channel per slot
  btb200_ctx *d_hop_ctx = nullptr;
  std::vector<uint8_t> d_mask;
  std::vector<btb200_hit> d_hits;
  bool d_batched = false;
  void handle_hit(const btb200_hit &, const char *, int, double) {}
  int hopalong_batch(int noutput_items, gr_vector_const_void_star &input_items);

 public:
  multi_hopper_impl(double sample_rate, double center_freq, double squelch_threshold, int LAP, bool aliased, bool tun);
  ~multi_hopper_impl();
  int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}  // namespace bluetooth
}  // namespace gr
#endif
