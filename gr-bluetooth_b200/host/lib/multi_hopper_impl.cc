// multi_hopper_impl.cc -- the target-LAP follower block (lib/multi_hopper_impl.cc:76-209 of the reference).
//
// One slot per work() call, always on the shared (chained) clock-recovery state: which channels a call
// demodulates depends on the piconet state the previous call left behind -- every channel until CLK1-27
// is known (ending early at the first packet of the target LAP that has a header), afterwards only the
// predicted hop channel.  The channel loop runs on the GPU (btb200_process_channels), the piconet logic
// (UAP/CLK1-6 from headers, hop reversal for CLK1-27, packet decode) in bt_host.{h,cc}.
#include "multi_hopper_impl.h"
#include "btb200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
// candidate search of the hop reversal on the GPU (btb200_hop_candidates); declining lets the host logic do it
int g_hop_device = 0;
bool gpu_hop_candidates(uint32_t address28, bool afh, bool aliased, uint32_t clock6, int first_channel, std::vector<uint32_t> &out)
{
  uint32_t n = 0;
  out.resize(1u << 16);
  int rc = btb200_hop_candidates(g_hop_device, address28, afh, aliased, clock6, first_channel, out.data(), (uint32_t)out.size(), &n);
  if (rc == BTB200_OK && n > out.size()) {
    out.resize(n);
    rc = btb200_hop_candidates(g_hop_device, address28, afh, aliased, clock6, first_channel, out.data(), (uint32_t)out.size(), &n);
  }
  if (rc != BTB200_OK) return false;
  out.resize(n);
  return true;
}
}  // namespace

namespace gr {
namespace bluetooth {

multi_hopper::sptr multi_hopper::make(double sample_rate, double center_freq, double squelch_threshold, int LAP,
                                      bool aliased, bool tun)
{
  return gnuradio::get_initial_sptr(new multi_hopper_impl(sample_rate, center_freq, squelch_threshold, LAP, aliased, tun));
}

multi_hopper_impl::multi_hopper_impl(double sample_rate, double center_freq, double squelch_threshold, int LAP,
                                     bool aliased, bool tun)
    : gr::sync_block("bluetooth multi hopper block", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                     gr::io_signature::make(0, 0, 0)),
      multi_block(sample_rate, center_freq, squelch_threshold, 3125, BTB200_SEARCH_BR, /*force_chained=*/true)
{
  const int lo = (int)((d_low_freq - 2402000000.0) / 1e6), hi = (int)((d_high_freq - 2402000000.0) / 1e6);
  d_host.reset(new btb200_host::HopperHost((uint32_t)LAP, aliased, lo, hi));
  if (tun) d_host->set_tun_fd(btb200_host::open_tun_output());     /* lib/multi_hopper_impl.cc:56-64 */
  d_res.resize((size_t)(hi - lo + 1));
  d_symbols.resize((size_t)(hi - lo + 1) * 3125);
  { const char *dv0 = std::getenv("BTB200_DEVICE"); g_hop_device = dv0 ? std::atoi(dv0) : 0; }
  btb200_host::Piconet::s_candidate_fn = gpu_hop_candidates;
  // BTB200_MM_MODE=stateless: once CLK1-27 is known, follow the piconet in BATCHES of slots.  The hop channel of
  // every future slot is known then (lib/multi_hopper_impl.cc:152-166: clock = clkn + offset, channel = hop(clock)), so
  // the slots no longer have to be visited one work() call at a time; each window starts from the constructor's
  // clock-recovery state instead of the chained one (the throughput semantics of the sniffer block).
  const char *mm = std::getenv("BTB200_MM_MODE");
  if (mm && std::string(mm) == "stateless") {
    btb200_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = BTB200_ABI_VERSION;
    cfg.sample_rate = sample_rate;
    cfg.center_freq = center_freq;
    cfg.squelch_threshold = squelch_threshold;
    cfg.extra_history_symbols = 3125;
    cfg.search = BTB200_SEARCH_BR;
    cfg.mm_mode = BTB200_MM_STATELESS;
    cfg.max_slots_per_call = d_batch_slots;
    const char *dv = std::getenv("BTB200_DEVICE");
    cfg.device = dv ? std::atoi(dv) : 0;
    const char *dd = std::getenv("BTB200_DDC");
    if (dd && std::string(dd) == "polyphase") cfg.ddc_mode = BTB200_DDC_POLYPHASE;
    int rc = btb200_create(&cfg, &d_hop_ctx);
    if (rc != BTB200_OK)
      throw std::runtime_error(std::string("btb200_create (hop-along): ") + btb200_strerror(rc) + " (" + btb200_last_error(nullptr) + ")");
    d_batched = true;
    d_mask.resize((size_t)d_batch_slots * (size_t)(hi - lo + 1));
    d_hits.resize(4096);
  }
}

multi_hopper_impl::~multi_hopper_impl() { if (d_hop_ctx) btb200_destroy(d_hop_ctx); }

// Batched form of the block (BTB200_MM_MODE=stateless): as many whole windows as the scheduler supplied go through the
// stateless receive path at once.  While the clock is unknown every channel is demodulated and searched (the scan of
// lib/multi_hopper_impl.cc:93-137); from the slot on which CLK1-27 is acquired -- it can happen in the middle of a batch --
// only the predicted hop channel counts (hopalong(), :152-209), and batches that start with the clock known mask
// everything else out before the GPU sees it.  Per (slot, channel) the reference runs ONE sniff_ac: the first access code.
int multi_hopper_impl::hopalong_batch(int noutput_items, gr_vector_const_void_star &input_items)
{
  const int S = (int)d_samples_per_slot, H = (int)history();
  int n = 1 + (noutput_items > 0 ? (noutput_items - 1) / S : 0);
  if (n > (int)d_batch_slots) n = (int)d_batch_slots;
  const uint32_t clkn0 = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  const int lo = (int)((d_low_freq - 2402000000.0) / 1e6), nch = (int)d_res.size();
  const bool masked = d_host->plan(clkn0).hopalong;
  bool any = !masked;
  if (masked) {
    std::fill(d_mask.begin(), d_mask.begin() + (size_t)n * nch, 0);
    for (int k = 0; k < n; k++) {
      const auto pl = d_host->plan(clkn0 + (uint32_t)k);
      if (pl.n_channels == 1) { d_mask[(size_t)k * nch + (pl.first_channel - lo)] = 1; any = true; }
    }
  }
  if (any) {
    for (;;) {
      btb200_hits out;
      std::memset(&out, 0, sizeof out);
      out.hits = d_hits.data();
      out.cap = (uint32_t)d_hits.size();
      out.symbols = nullptr;
      out.symbols_cap = UINT64_MAX;                      // borrowed symbols: no second host copy
      int rc = masked ? btb200_set_window_mask(d_hop_ctx, d_mask.data(), (uint32_t)n) : BTB200_OK;
      if (rc == BTB200_OK)
        rc = btb200_process(d_hop_ctx, reinterpret_cast<const float *>(input_items[0]), (size_t)(n - 1) * S + H, clkn0, (uint32_t)n, &out);
      if (rc != BTB200_OK)
        throw std::runtime_error(std::string("btb200_process (hopper batch): ") + btb200_strerror(rc) + " (" + btb200_last_error(d_hop_ctx) + ")");
      { float tm[8]; if (btb200_last_timing(d_hop_ctx, tm) == BTB200_OK) d_device_ms += tm[7]; }
      if (out.overflow) { d_hits.resize(d_hits.size() * 4); continue; }
      uint32_t last_slot = 0xffffffffu, done_slot = 0xffffffffu;
      int last_chan = -1;
      for (uint32_t i = 0; i < out.count; i++) {
        const btb200_hit &h = out.hits[i];
        if (h.kind != 0) continue;
        if (h.slot == last_slot && h.channel == last_chan) continue;       // first access code of a channel-window only
        last_slot = h.slot; last_chan = h.channel;
        if (h.slot == done_slot) continue;                                   // the reference left this slot's channel loop
        const char *sp = reinterpret_cast<const char *>(out.symbols + h.sym_offset);
        const auto pl = d_host->plan(h.slot);
        if (pl.hopalong) {
          if (pl.n_channels == 1 && (int)h.channel == pl.first_channel) {      // one channel per slot in this phase
            d_host->hop_packet(pl, sp, (int)h.sym_count);
            done_slot = h.slot;
          }
        } else if (d_host->scan_packet(h.slot, h.channel, sp, (int)h.sym_count)) {
          done_slot = h.slot;                                                // `break`, multi_hopper_impl.cc:129,133
        }
      }
      break;
    }
  }
  d_cumulative_count += (uint64_t)n * (uint64_t)S;
  return n * S;
}

int multi_hopper_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
  const int S = (int)d_samples_per_slot;
  const uint32_t clkn = (uint32_t)((int)(d_cumulative_count / d_samples_per_slot) & 0x7ffffff);
  const auto pl = d_host->plan(clkn);
  if (d_batched) return hopalong_batch(noutput_items, input_items);
  if (pl.n_channels > 0) {
    int rc = btb200_process_channels(d_ctx, reinterpret_cast<const float *>(input_items[0]), history(), clkn,
                                     pl.first_channel, pl.n_channels, pl.stop_lap, d_res.data(), d_symbols.data(),
                                     d_symbols.size());
    if (rc != BTB200_OK)
      throw std::runtime_error(std::string("btb200_process_channels: ") + btb200_strerror(rc) + " (" + btb200_last_error(d_ctx) + ")");
    for (int q = 0; q < pl.n_channels; q++) {
      const btb200_chan_result &r = d_res[(size_t)q];
      if (!r.processed) break;
      if (r.ac_index < 0) continue;
      const char *sp = reinterpret_cast<const char *>(d_symbols.data() + r.sym_offset);
      if (pl.hopalong) d_host->hop_packet(pl, sp, (int)r.sym_count);
      else if (d_host->scan_packet(clkn, r.channel, sp, (int)r.sym_count)) break;
    }
  }
  d_cumulative_count += (uint64_t)S;
  return S;
}

}  // namespace bluetooth
}  // namespace gr
