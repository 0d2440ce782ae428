// tests/hostlayer/hopper_main.cc -- TEST HARNESS.
// The product's native hopper logic (HopperHost + Piconet hop reversal, bt_host.cc) driven by the
// ORACLE front end (btbo_window_list): plays multi_hopper_impl::work() over a capture and prints what
// the block prints; the CPU test compares it with the reference's digest.
#include "../../gr-bluetooth_b200/host/lib/bt_host.h"
extern "C" {
#include "../../oracle/btb_oracle.h"
}
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
  if (argc < 5) { std::fprintf(stderr, "usage: hopper FS FC LAPHEX FILE.cfile [TUNFRAMES.out | uap]\n"); return 2; }
  const double fs = std::atof(argv[1]), fc = std::atof(argv[2]);
  const uint32_t lap = (uint32_t)std::strtoul(argv[3], nullptr, 16);
  btbo_plan *P = btbo_plan_create(fs, fc, 10.0, 3125);
  btbo_state *st = btbo_state_create(P);
  btbo_info I;
  btbo_plan_info(P, &I);
  FILE *f = std::fopen(argv[4], "rb");
  if (!f) return 2;
  std::fseek(f, 0, SEEK_END);
  const long total = std::ftell(f) / 8;
  std::fseek(f, 0, SEEK_SET);
  std::vector<float> buf((size_t)(I.H - 1 + total) * 2, 0.0f);
  if (std::fread(&buf[(size_t)(I.H - 1) * 2], 8, (size_t)total, f) != (size_t)total) return 2;
  std::fclose(f);
  const int chist = I.Nc + I.D * 8;
  std::printf("history set to %d samples: channel=%d, noise=%d\n", I.S + (chist > I.Nn ? chist : I.Nn), chist, I.Nn);
  const bool aliased = argc > 5 && std::string(argv[5]) == "aliased";
  btb200_host::HopperHost host(lap, aliased, I.ch_lo, I.ch_hi);
  if (argc > 5 && std::string(argv[5]) != "uap" && !aliased) host.set_tun_fd(open(argv[5], O_WRONLY | O_CREAT | O_TRUNC, 0644));
  std::vector<btbo_chan_result> res((size_t)I.nch);
  std::vector<uint8_t> sym((size_t)I.nch * I.H);
  std::vector<int32_t> chis((size_t)I.nch);
  const long ncalls = (total + I.S - 1) / I.S;
  if (argc > 5 && std::string(argv[5]) == "uap") {
    // multi_UAP: the same channel loop, packets of the target piconet feed the UAP discovery until it is known
    btb200_host::UapHost uh(lap, I.ch_lo, I.ch_hi);
    for (long k = 0; k < ncalls && !uh.done(); k++) {
      for (int q = 0; q < I.nch; q++) chis[(size_t)q] = q;
      btbo_window_list(P, st, &buf[(size_t)k * I.S * 2], chis.data(), I.nch, lap, res.data(), sym.data());
      for (int q = 0; q < I.nch; q++) {
        const btbo_chan_result &r = res[(size_t)q];
        if (!r.processed) break;
        if (r.ac_index < 0) continue;
        if (uh.packet((uint32_t)k, I.ch_lo + q, reinterpret_cast<const char *>(&sym[(size_t)q * I.H + r.ac_index]), r.nsym - r.ac_index)) break;
      }
    }
    std::printf("multi_UAP %s: UAP 0x%02x\n", uh.done() ? "done" : "not found", uh.uap());
    return 0;
  }
  for (long k = 0; k < ncalls; k++) {
    const auto pl = host.plan((uint32_t)k);
    if (pl.n_channels == 0) continue;
    for (int q = 0; q < pl.n_channels; q++) chis[(size_t)q] = pl.first_channel + q - I.ch_lo;
    btbo_window_list(P, st, &buf[(size_t)k * I.S * 2], chis.data(), pl.n_channels, pl.stop_lap, res.data(), sym.data());
    for (int q = 0; q < pl.n_channels; q++) {
      const btbo_chan_result &r = res[(size_t)q];
      if (!r.processed) break;
      if (r.ac_index < 0) continue;
      const char *sp = reinterpret_cast<const char *>(&sym[(size_t)q * I.H + r.ac_index]);
      if (pl.hopalong) host.hop_packet(pl, sp, r.nsym - r.ac_index);
      else if (host.scan_packet((uint32_t)k, pl.first_channel + q, sp, r.nsym - r.ac_index)) break;
    }
  }
  return 0;
}
