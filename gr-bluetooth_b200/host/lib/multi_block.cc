// multi_block.cc -- scheduler-facing half of the blocks; see include/gr_bluetooth/multi_block.h.
#include "gr_bluetooth/multi_block.h"
#include "btb200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace gr {
namespace bluetooth {

multi_block::multi_block(double sample_rate, double center_freq, double squelch_threshold,
                         int extra_symbols, int search_mask, bool force_chained)
    : gr::sync_block("bluetooth multi block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0))
{
  d_sample_rate = sample_rate;
  d_center_freq = center_freq;
  d_target_snr = squelch_threshold;
  btb200_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = BTB200_ABI_VERSION;
  cfg.sample_rate = sample_rate;
  cfg.center_freq = center_freq;
  cfg.squelch_threshold = squelch_threshold;
  cfg.extra_history_symbols = (uint32_t)extra_symbols;
  cfg.search = search_mask;
  // reference semantics by default: one clock-recovery state shared by all channel-windows
  const char *mm = std::getenv("BTB200_MM_MODE");
  cfg.mm_mode = (!force_chained && mm && std::string(mm) == "stateless") ? BTB200_MM_STATELESS : BTB200_MM_CHAINED;
  const char *bs = std::getenv("BTB200_BATCH_SLOTS");
  d_batch_slots = bs ? (unsigned)std::atoi(bs) : 16u;
  if (d_batch_slots < 1) d_batch_slots = 1;
  cfg.max_slots_per_call = d_batch_slots;
  const char *dv = std::getenv("BTB200_DEVICE");
  cfg.device = dv ? std::atoi(dv) : 0;
  int rc = btb200_create(&cfg, &d_ctx);
  if (rc != BTB200_OK)
    throw std::runtime_error(std::string("btb200_create: ") + btb200_strerror(rc) + " (" + btb200_last_error(nullptr) + ")");
  btb200_info info;
  btb200_get_info(d_ctx, &info);
  d_samples_per_slot = info.samples_per_slot;
  d_low_freq = 2402000000.0 + 1e6 * info.channel_low;
  d_high_freq = 2402000000.0 + 1e6 * info.channel_high;
  // the reference's constructor line (lib/multi_block.cc:116), printed before the symbol history is added
  const int chist = info.chan_taps + info.decimation * 8;
  const int base = info.samples_per_slot + (chist > info.noise_taps ? chist : info.noise_taps);
  std::printf("history set to %d samples: channel=%d, noise=%d\n", base, chist, info.noise_taps);
  set_history((unsigned)info.history);
}

multi_block::~multi_block() { btb200_destroy(d_ctx); }

int multi_block::process_windows(int noutput_items, gr_vector_const_void_star &input_items)
{
  const int S = (int)d_samples_per_slot;
  const int H = (int)history();
  // the scheduler guarantees noutput_items + H - 1 input items; windows fully contained in them:
  int n = 1 + (noutput_items > 0 ? (noutput_items - 1) / S : 0);
  if (n > (int)d_batch_slots) n = (int)d_batch_slots;
  static thread_local std::vector<btb200_hit> hits;
  static thread_local std::vector<uint8_t> symbols;
  hits.resize(4096);
  symbols.resize(4096u * 3125u / 4);
  btb200_hits out;
  std::memset(&out, 0, sizeof out);
  out.hits = hits.data();
  out.cap = (uint32_t)hits.size();
  out.symbols = symbols.data();
  out.symbols_cap = symbols.size();
  const uint64_t first_slot = d_cumulative_count / (uint64_t)S;
  int rc = btb200_process(d_ctx, reinterpret_cast<const float *>(input_items[0]),
                          (size_t)(n - 1) * S + H, first_slot, (uint32_t)n, &out);
  if (rc != BTB200_OK)
    throw std::runtime_error(std::string("btb200_process: ") + btb200_strerror(rc) + " (" + btb200_last_error(d_ctx) + ")");
  uint32_t cur_slot = (uint32_t)first_slot;
  for (uint32_t i = 0; i < out.count; i++) {
    const btb200_hit &h = out.hits[i];
    // keep d_cumulative_count where the reference has it while ac()/aa() run (clkn, multi_sniffer_impl.cc:173)
    d_cumulative_count += (uint64_t)(h.slot - cur_slot) * S;
    cur_slot = h.slot;
    handle_hit(h, reinterpret_cast<const char *>(symbols.data() + h.sym_offset), (int)h.sym_count,
               2402000000.0 + 1e6 * h.channel);
  }
  d_cumulative_count += (uint64_t)(first_slot + n - cur_slot) * S;
  return n * S;
}

}  // namespace bluetooth
}  // namespace gr
