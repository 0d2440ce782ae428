/*
 * oracle/btref_main.cc -- TEST INFRASTRUCTURE.
 *
 * Driver for the VERBATIM reference build (oracle/_ref/btref): the reference's
 * lib/multi_block.cc, multi_sniffer_impl.cc, multi_hopper_impl.cc,
 * packet_impl.cc, piconet_impl.cc, tun.cc are compiled unmodified from
 * /root/reference against oracle/shim (GNU Radio stand-in).  This file only
 * emulates the GNU Radio scheduler contract for a sync_block with history
 * (SURVEY.md 3.4 / Appendix A.1): buf = zeros(H-1) ++ samples; call k gets
 * in[0] = &buf[k*S] and consumes S = samples per slot.
 *
 * Sub-commands:
 *   sniff   --fs F --fc F --snr F --in FILE [--i16] [--first-call K] [--num-calls N]
 *           [--stateless] [--dump FILE] [--heavy A:B] [--quiet] [--tun-out FILE]
 *   hop     same + --lap HEX [--aliased]
 *   acgen   HEXLAP...                 print the 9 access-code bytes (packet_impl.cc:309)
 *   sniffdem FILE                     loop sniff_ac over a 1-bit-per-byte symbol file
 *   tables                            print the reference's detection LUTs
 */
/* --tun-out needs the blocks' private TAP descriptor; the reference headers stay untouched, the keyword is
 * redefined around their inclusion only (standard and shim headers first, so nothing else is affected) */
#include <map>
#include <string>
#include <vector>
#include <iostream>
#include <memory>
#include <gnuradio/sync_block.h>
#include <gnuradio/io_signature.h>
#include <fcntl.h>
#define private protected
#include "multi_sniffer_impl.h"
#include "multi_hopper_impl.h"
#undef private
#include "btref_hooks.h"
#include <string>
#include <vector>
#include <chrono>

using namespace gr::bluetooth;

static gr::io_signature::sptr sig() { return gr::io_signature::make(1, 1, 8); }

/* Derived only to reach multi_block's protected M&M state for --stateless. */
struct probe_sniffer : public multi_sniffer_impl {
  probe_sniffer(double fs, double fc, double snr)
    : multi_block(fs, fc, snr), gr::sync_block("probe", sig(), sig()),
      multi_sniffer_impl(fs, fc, snr, false) {}
  void reset_mm() { d_mu = 0.32; d_omega = d_omega_mid; d_last_sample = 0; }   /* multi_block.cc:91-98 */
  void tun_to(int fd) { d_tun = true; d_tunfd = fd; }      /* frames of the Wireshark interface -> fd */
  int S() const { return (int)d_samples_per_slot; }
  void info() const {
    fprintf(stderr, "btref: S=%d D=%d Nc=%zu Nn=%zu fcs=%d fns=%d low=%.0f high=%.0f gain=%.9g\n",
            (int)d_samples_per_slot, d_ddc_decimation_rate, d_channel_filter.size(),
            d_noise_filter.size(), d_first_channel_sample, d_first_noise_sample,
            d_low_freq, d_high_freq, (double)d_demod_gain);
  }
  const std::vector<float> &chan_taps() const { return d_channel_filter; }
  const std::vector<float> &noise_taps() const { return d_noise_filter; }
};

struct probe_hopper : public multi_hopper_impl {
  probe_hopper(double fs, double fc, double snr, int lap, bool aliased)
    : multi_block(fs, fc, snr), gr::sync_block("probe", sig(), sig()),
      multi_hopper_impl(fs, fc, snr, lap, aliased, false) {}
  void reset_mm() { d_mu = 0.32; d_omega = d_omega_mid; d_last_sample = 0; }
  void tun_to(int fd) { d_tun = true; d_tunfd = fd; }
  int S() const { return (int)d_samples_per_slot; }
};

static void hook_sniffer(void *u) { ((probe_sniffer *)u)->reset_mm(); }
static void hook_hopper(void *u) { ((probe_hopper *)u)->reset_mm(); }

struct args {
  double fs = 2e6, fc = 2476e6, snr = 10;
  std::string in, dump, tun_out;
  bool i16 = false, stateless = false, aliased = false, timing = false;
  long first_call = 0, num_calls = -1;
  int heavy_a = 0, heavy_b = 0;
  int lap = 0;
};

static bool load_samples(const args &a, long first_sample, long count, std::vector<gr_complex> &dst, long &total)
{
  FILE *f = fopen(a.in.c_str(), "rb");
  if (!f) { perror(a.in.c_str()); return false; }
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  size_t esz = a.i16 ? 4 : 8;
  total = bytes / (long)esz;
  /* dst[j] = sample first_sample + j, zeros outside [0,total) */
  dst.assign((size_t)count, gr_complex(0, 0));
  long lo = first_sample < 0 ? 0 : first_sample;
  long hi = first_sample + count; if (hi > total) hi = total;
  if (hi > lo) {
    fseek(f, lo * (long)esz, SEEK_SET);
    if (a.i16) {
      std::vector<int16_t> tmp((size_t)(hi - lo) * 2);
      if (fread(tmp.data(), 4, (size_t)(hi - lo), f) != (size_t)(hi - lo)) { fclose(f); return false; }
      for (long j = 0; j < hi - lo; j++)
        dst[(size_t)(lo - first_sample + j)] = gr_complex(tmp[2 * j], tmp[2 * j + 1]);
    } else {
      if (fread(&dst[(size_t)(lo - first_sample)], 8, (size_t)(hi - lo), f) != (size_t)(hi - lo)) { fclose(f); return false; }
    }
  }
  fclose(f);
  return true;
}

template <class B>
static int run_block(B &blk, const args &a)
{
  const long S = blk.S();
  const long H = blk.history();
  FILE *f = fopen(a.in.c_str(), "rb");
  if (!f) { perror(a.in.c_str()); return 2; }
  fseek(f, 0, SEEK_END);
  long total = ftell(f) / (a.i16 ? 4 : 8);
  fclose(f);
  long ncalls_all = (total + S - 1) / S;               /* k*S < N */
  long k0 = a.first_call;
  long k1 = (a.num_calls < 0) ? ncalls_all : k0 + a.num_calls;
  if (k1 > ncalls_all) k1 = ncalls_all;
  if (k1 <= k0) return 0;
  /* call k sees absolute samples [k*S-(H-1), k*S] */
  long first_sample = k0 * S - (H - 1);
  long count = (k1 - 1 - k0) * S + H;
  std::vector<gr_complex> buf;
  if (!load_samples(a, first_sample, count, buf, total)) return 2;

  if (!a.dump.empty()) {
    g_btref.dump = fopen(a.dump.c_str(), "wb");
    if (!g_btref.dump) { perror(a.dump.c_str()); return 2; }
    g_btref.heavy_from = a.heavy_a;
    g_btref.heavy_to = a.heavy_b;
  }
  g_btref.stateless = a.stateless ? 1 : 0;

  gr_vector_const_void_star in(1);
  gr_vector_void_star out;
  auto t0 = std::chrono::steady_clock::now();
  for (long k = k0; k < k1; k++) {
    g_btref.call_index = (int)k;
    in[0] = &buf[(size_t)((k - k0) * S)];
    int r = blk.work(32768, in, out);
    btref_flush_symbols();
    if (r != S) { fprintf(stderr, "btref: work returned %d != %ld\n", r, S); return 3; }
  }
  auto t1 = std::chrono::steady_clock::now();
  fflush(stdout);
  if (a.timing) {
    double s = std::chrono::duration<double>(t1 - t0).count();
    fprintf(stderr, "btref-timing: calls=%ld samples=%ld seconds=%.6f\n", k1 - k0, (k1 - k0) * S, s);
  }
  if (g_btref.dump) fclose(g_btref.dump);
  return 0;
}

static void print_table(const char *name, const uint8_t *t, int n)
{
  printf("%s %d", name, n);
  for (int i = 0; i < n; i++) printf(" %d", t[i]);
  printf("\n");
}

int main(int argc, char **argv)
{
  if (argc < 2) { fprintf(stderr, "usage: btref sniff|hop|acgen|sniffdem|tables ...\n"); return 1; }
  std::string cmd = argv[1];
  args a;
  std::vector<std::string> pos;
  for (int i = 2; i < argc; i++) {
    std::string s = argv[i];
    auto next = [&]() -> const char * { return (i + 1 < argc) ? argv[++i] : ""; };
    if (s == "--fs") a.fs = atof(next());
    else if (s == "--fc") a.fc = atof(next());
    else if (s == "--snr") a.snr = atof(next());
    else if (s == "--in") a.in = next();
    else if (s == "--dump") a.dump = next();
    else if (s == "--tun-out") a.tun_out = next();
    else if (s == "--i16") a.i16 = true;
    else if (s == "--stateless") a.stateless = true;
    else if (s == "--aliased") a.aliased = true;
    else if (s == "--timing") a.timing = true;
    else if (s == "--first-call") a.first_call = atol(next());
    else if (s == "--num-calls") a.num_calls = atol(next());
    else if (s == "--lap") a.lap = (int)strtol(next(), NULL, 16);
    else if (s == "--heavy") { sscanf(next(), "%d:%d", &a.heavy_a, &a.heavy_b); }
    else pos.push_back(s);
  }

  if (cmd == "sniff") {
    probe_sniffer blk(a.fs, a.fc, a.snr);
    blk.info();
    if (a.stateless) { g_btref.on_channel_ddc = hook_sniffer; g_btref.user = &blk; }
    if (!a.tun_out.empty()) blk.tun_to(open(a.tun_out.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644));
    return run_block(blk, a);
  }
  if (cmd == "hop") {
    probe_hopper blk(a.fs, a.fc, a.snr, a.lap, a.aliased);
    if (a.stateless) { g_btref.on_channel_ddc = hook_hopper; g_btref.user = &blk; }
    if (!a.tun_out.empty()) blk.tun_to(open(a.tun_out.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644));
    return run_block(blk, a);
  }
  if (cmd == "design") {
    /* print the derived constants; with --dump also the two prototype tap sets (f32) */
    probe_sniffer blk(a.fs, a.fc, a.snr);
    blk.info();
    printf("H %u\n", blk.history());
    if (!a.dump.empty()) {
      FILE *f = fopen(a.dump.c_str(), "wb");
      fwrite(blk.chan_taps().data(), 4, blk.chan_taps().size(), f);
      fwrite(blk.noise_taps().data(), 4, blk.noise_taps().size(), f);
      fclose(f);
    }
    return 0;
  }
  if (cmd == "acgen") {
    for (auto &p : pos) {
      int lap = (int)strtol(p.c_str(), NULL, 16);
      uint8_t *ac = classic_packet::acgen(lap);
      printf("%06x ", lap);
      for (int i = 0; i < 9; i++) printf("%02x", ac[i]);
      printf("\n");
      free(ac);
    }
    return 0;
  }
  if (cmd == "sniffdem") {
    /* loop sniff_ac over a symbol file, skipping 68 symbols after every hit
     * (the same stepping as multi_sniffer_impl.cc:112-126 without the slot limit) */
    FILE *f = fopen(pos.at(0).c_str(), "rb");
    if (!f) { perror(pos[0].c_str()); return 2; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<char> sym((size_t)n + 80, 0);
    if (fread(sym.data(), 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    long pos_ = 0, hits = 0;
    while (pos_ + 72 < n) {
      int i = classic_packet::sniff_ac(&sym[(size_t)pos_], (int)(n - 72 - pos_));
      if (i < 0) break;
      uint32_t lap = packet::air_to_host32(&sym[(size_t)(pos_ + i + 38)], 24);
      printf("%ld %06x\n", pos_ + i, lap);
      hits++;
      pos_ += i + 68;
    }
    fprintf(stderr, "btref: %ld symbols, %ld hits\n", n, hits);
    return 0;
  }
  if (cmd == "tables") {
    print_table("classic.PREAMBLE_DISTANCE", classic_packet::PREAMBLE_DISTANCE, 32);
    print_table("classic.BARKER_DISTANCE", classic_packet::BARKER_DISTANCE, 128);
    print_table("classic.INDICES", classic_packet::INDICES, 64);
    print_table("packet.WHITENING_DATA", packet::WHITENING_DATA, 127);
    print_table("le.INDICES", le_packet::INDICES, 40);
    print_table("le.PREAMBLE_DISTANCE", le_packet::PREAMBLE_DISTANCE, 512);
    print_table("le.ACCESS_ADDRESS_DISTANCE_0", le_packet::ACCESS_ADDRESS_DISTANCE_0, 256);
    print_table("le.ACCESS_ADDRESS_DISTANCE_1", le_packet::ACCESS_ADDRESS_DISTANCE_1, 256);
    print_table("le.ACCESS_ADDRESS_DISTANCE_2", le_packet::ACCESS_ADDRESS_DISTANCE_2, 256);
    print_table("le.ACCESS_ADDRESS_DISTANCE_3", le_packet::ACCESS_ADDRESS_DISTANCE_3, 256);
    print_table("le.ACCESS_HEADER_DISTANCE_LSB", le_packet::ACCESS_HEADER_DISTANCE_LSB, 256);
    print_table("le.ACCESS_HEADER_DISTANCE_MSB", le_packet::ACCESS_HEADER_DISTANCE_MSB, 256);
    print_table("le.DATA_HEADER_DISTANCE_LSB", le_packet::DATA_HEADER_DISTANCE_LSB, 256);
    print_table("le.DATA_HEADER_DISTANCE_MSB", le_packet::DATA_HEADER_DISTANCE_MSB, 256);
    for (int ch = 0; ch < 40; ch++) printf("le.chan2index %d %d\n", ch, le_packet::chan2index(ch));
    return 0;
  }
  fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 1;
}
