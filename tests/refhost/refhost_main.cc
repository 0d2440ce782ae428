// tests/refhost/refhost_main.cc -- TEST HARNESS, not product code.
//
// Proves that the C ABI carries everything the reference's host packet layer needs: the GPU
// front end (libbtb200.so, chained mode) feeds every hit into the REFERENCE's own
// multi_sniffer_impl::ac()/aa() (lib/multi_sniffer_impl.cc:169-365, packet_impl.cc,
// piconet_impl.cc compiled verbatim in oracle/_ref/obj) and the resulting stdout must be
// byte-identical to the reference running alone.  Built only where /root/reference exists
// (tests/refhost/Makefile); the binary travels to the GPU box.
// system and shim headers first, so that the access hack below only touches the reference headers
#include <iostream>
#include <sstream>
#include <map>
#include <string>
#include <vector>
#include <gnuradio/sync_block.h>
#include <gnuradio/io_signature.h>
#include <gnuradio/filter/freq_xlating_fir_filter_ccf.h>
#include <gnuradio/filter/mmse_fir_interpolator_ff.h>
#include "tun.h"
#define private public
#define protected public
#include "multi_sniffer_impl.h"
#undef private
#undef protected
#include "btb200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace gr::bluetooth;

int main(int argc, char **argv)
{
  if (argc < 4) { std::fprintf(stderr, "usage: refhost FS FC FILE.cfile [batch]\n"); return 2; }
  const double fs = std::atof(argv[1]), fc = std::atof(argv[2]);
  const uint32_t batch = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 64u;
  multi_sniffer_impl ref(fs, fc, 10.0, false);           // prints the "history set to" line itself
  btb200_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = BTB200_ABI_VERSION;
  cfg.sample_rate = fs; cfg.center_freq = fc; cfg.squelch_threshold = 10.0;
  cfg.extra_history_symbols = 3125; cfg.mm_mode = BTB200_MM_CHAINED;
  cfg.search = BTB200_SEARCH_BR | BTB200_SEARCH_LE; cfg.max_slots_per_call = batch;
  btb200_ctx *ctx = nullptr;
  int rc = btb200_create(&cfg, &ctx);
  if (rc) { std::fprintf(stderr, "btb200_create: %s (%s)\n", btb200_strerror(rc), btb200_last_error(nullptr)); return 1; }
  btb200_info info;
  btb200_get_info(ctx, &info);
  const long S = info.samples_per_slot, H = info.history;
  if ((long)ref.history() != H) { std::fprintf(stderr, "history mismatch %u vs %ld\n", ref.history(), H); return 1; }
  FILE *f = std::fopen(argv[3], "rb");
  if (!f) { std::perror(argv[3]); return 2; }
  std::fseek(f, 0, SEEK_END);
  const long total = std::ftell(f) / 8;
  std::fseek(f, 0, SEEK_SET);
  std::vector<float> buf((size_t)(H - 1 + total) * 2, 0.0f);
  if (std::fread(&buf[(size_t)(H - 1) * 2], 8, (size_t)total, f) != (size_t)total) return 2;
  std::fclose(f);
  const long ncalls = (total + S - 1) / S;
  std::vector<btb200_hit> hits(8192);
  std::vector<uint8_t> syms(8192u * 3200u);
  for (long k = 0; k < ncalls; k += batch) {
    const uint32_t n = (uint32_t)((ncalls - k) < (long)batch ? (ncalls - k) : (long)batch);
    btb200_hits out;
    std::memset(&out, 0, sizeof out);
    out.hits = hits.data(); out.cap = (uint32_t)hits.size();
    out.symbols = syms.data(); out.symbols_cap = syms.size();
    rc = btb200_process(ctx, &buf[(size_t)k * S * 2], (size_t)(n - 1) * S + H, (uint64_t)k, n, &out);
    if (rc) { std::fprintf(stderr, "btb200_process: %s (%s)\n", btb200_strerror(rc), btb200_last_error(ctx)); return 1; }
    if (out.overflow) { std::fprintf(stderr, "hit overflow\n"); return 1; }
    for (uint32_t i = 0; i < out.count; i++) {
      const btb200_hit &h = out.hits[i];
      ref.d_cumulative_count = (uint64_t)h.slot * (uint64_t)S;            // clkn source, multi_sniffer_impl.cc:173
      char *sp = reinterpret_cast<char *>(syms.data() + h.sym_offset);
      const double freq = 2402000000.0 + 1e6 * h.channel;
      if (h.kind == 0) ref.ac(sp, h.n_symbols, freq, h.snr);
      else ref.aa(sp, h.n_symbols, freq, h.snr);
    }
  }
  std::fflush(stdout);
  btb200_destroy(ctx);
  return 0;
}
