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
                         int extra_symbols, int search_mask, bool force_chained, unsigned bch)
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
  cfg.bch = bch;
  // BTB200_AC_SEARCH=sniff_ac: the blocks the reference builds on libbtbb (multi_LAP, multi_UAP) fall back to
  // classic_packet::sniff_ac semantics (what this port used before the libbtbb-style test existed)
  if (const char *acs = std::getenv("BTB200_AC_SEARCH"))
    if (std::string(acs) == "sniff_ac") { cfg.search &= ~BTB200_SEARCH_BR_BCH; cfg.bch = 0; }
  // reference semantics by default: one clock-recovery state shared by all channel-windows
  const char *mm = std::getenv("BTB200_MM_MODE");
  cfg.mm_mode = (!force_chained && mm && std::string(mm) == "stateless") ? BTB200_MM_STATELESS : BTB200_MM_CHAINED;
  d_stateless = cfg.mm_mode == BTB200_MM_STATELESS;
  // BTB200_DDC=polyphase selects the throughput front end (stateless mode only; tolerance-level floats)
  const char *dd = std::getenv("BTB200_DDC");
  if (d_stateless && dd && std::string(dd) == "polyphase") cfg.ddc_mode = BTB200_DDC_POLYPHASE;
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
  // every hit may carry 3125 symbols (classic_packet::make keeps MAX_SYMBOLS, lib/packet_impl.cc:52-58); the hit
  // buffer grows and the batch is re-run when dense traffic overflows it -- a silently dropped or symbol-less hit
  // would print a wrong line
  if (hits.size() < 1024) hits.resize(1024);
  symbols.resize(hits.size() * 3125u);
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
  { float tm[8]; if (btb200_last_timing(d_ctx, tm) == BTB200_OK) d_device_ms += tm[7]; }
  if (out.overflow) {
    // more hits than the buffer holds: this only happens in stateless (batch) mode, where the batch can simply be
    // processed again; the chained stream state has already advanced and cannot be replayed
    if (hits.size() >= (1u << 20) || !d_stateless)
      throw std::runtime_error("btb200_process: hit buffer overflow (" + std::to_string(out.overflow) + " hits dropped)");
    hits.resize(hits.size() * 4);
    return process_windows(noutput_items, input_items);
  }
  for (uint32_t i = 0; i < out.count; i++) {
    const int expect = out.hits[i].n_symbols < 3125 ? out.hits[i].n_symbols : 3125;
    if ((int)out.hits[i].sym_count < (expect > 0 ? expect : 0))
      throw std::runtime_error("btb200_process: symbol arena exhausted");
  }
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
