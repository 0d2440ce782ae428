// gr::bluetooth::multi_block -- B200 edition.
//
// Same role and public surface as the reference's base class
// (include/gr_bluetooth/multi_block.h:40-170 there): a gr::sync_block with one complex
// input and no outputs whose work() consumes one 625 us slot per window.  The arithmetic the
// reference does inline (channel_samples / check_snr / channel_symbols and the access-code
// search) happens behind the C ABI of include/btb200.h, in CUDA kernels; this class keeps the
// scheduler-facing contract (history(), work() return value, slot counter) and turns the
// returned hits into the ac()/aa() callbacks of the derived blocks.
#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_BLOCK_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_BLOCK_H

#include <gr_bluetooth/api.h>
#include <gnuradio/sync_block.h>
#include <stdint.h>
#include <vector>

struct btb200_ctx;
struct btb200_hit;

namespace gr {
namespace bluetooth {

class GR_BLUETOOTH_API multi_block : virtual public gr::sync_block {
 protected:
  multi_block() {}   // to allow for pure virtual
  // extra_symbols: 3125 for sniffer/hopper, 68 for multi_LAP (set_symbol_history in the reference)
  // bch: btb200_config.bch when search_mask carries BTB200_SEARCH_BR_BCH (the libbtbb-style access-code test)
  multi_block(double sample_rate, double center_freq, double squelch_threshold, int extra_symbols,
              int search_mask, bool force_chained = false, unsigned bch = 0);

  static const int SYMBOLS_PER_BASIC_RATE_SLOT = 625;

  double d_sample_rate = 0, d_center_freq = 0, d_target_snr = 0;
  double d_samples_per_slot = 0;
  double d_low_freq = 0, d_high_freq = 0;
  uint64_t d_cumulative_count = 0;       // samples elapsed, as in the reference
  btb200_ctx *d_ctx = nullptr;
  unsigned d_batch_slots = 1;            // windows handed to the GPU per work() call when available
  bool d_stateless = false;
  double d_device_ms = 0;                // device time of the batches processed so far (CUDA events, btb200_last_timing)

  // one callback per detected packet, in the reference's visiting order
  virtual void handle_hit(const btb200_hit &hit, const char *symbols, int n_symbols, double freq) = 0;

  // the body shared by every block's work(): process as many complete windows as the scheduler
  // supplied (at least one), call handle_hit() for each ac()/aa() event, return items consumed
  int process_windows(int noutput_items, gr_vector_const_void_star &input_items);

 public:
  virtual ~multi_block();
  // environment knobs (the make() signatures stay the reference's):
  //   BTB200_MM_MODE=chained|stateless   BTB200_DDC=exact|polyphase   BTB200_BATCH_SLOTS=n   BTB200_DEVICE=k
  unsigned batch_slots() const { return d_batch_slots; }
  double samples_per_slot() const { return d_samples_per_slot; }
  double device_ms() const { return d_device_ms; }
  virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                   gr_vector_void_star &output_items) = 0;
};

}  // namespace bluetooth
}  // namespace gr
#endif
