// btrx_b200 -- command-line receiver with the flag surface of the reference's apps/btrx
// (Python 2 + GNU Radio there; options at apps/btrx:23-58).  File input only: it plays the
// role of the GNU Radio scheduler for one block -- zero history in front of the stream, then
// work() until the file is consumed.
//   -f/--freq Hz   -r/--rate sps   -i/--input-file FILE   -S (all-piconet sniffer, default)
//   -L (LAP printer)   -l HEXLAP [-p] (follow one piconet: multi_hopper)   -s/--snr dB   -N/--nsamples n   -2/--input-shorts
// and, not in the reference: --tile N (play the file N times back to back, as one stream; BASELINE configs[3]),
// --stats (one JSON line on stderr: samples, seconds, Msamples/s through work())
#include "gr_bluetooth/multi_sniffer.h"
#include "gr_bluetooth/multi_LAP.h"
#include "gr_bluetooth/multi_hopper.h"
#include "gr_bluetooth/multi_UAP.h"
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

static double parse_eng(const char *s)
{
  char *end = nullptr;
  double v = std::strtod(s, &end);
  if (end && (*end == 'M' || *end == 'm')) v *= 1e6;
  else if (end && (*end == 'k' || *end == 'K')) v *= 1e3;
  else if (end && (*end == 'G' || *end == 'g')) v *= 1e9;
  return v;
}

int main(int argc, char **argv)
{
  double freq = 2476e6, rate = 2e6, snr = 10;
  std::string in;
  long nsamples = -1;
  bool shorts = false, hop_mode = false, tun = false, have_lap = false, sniff = false;
  int target_lap = 0;
  long tile = 1;
  bool stats = false, aliased = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "-f" || a == "--freq") freq = parse_eng(next());
    else if (a == "-r" || a == "--rate") rate = parse_eng(next());
    else if (a == "-i" || a == "--input-file") in = next();
    else if (a == "-s" || a == "--snr") snr = std::atof(next());
    else if (a == "-N" || a == "--nsamples") nsamples = (long)parse_eng(next());
    else if (a == "-2" || a == "--input-shorts") shorts = true;
    else if (a == "-S" || a == "--sniff-all") sniff = true;
    else if (a == "-L" || a == "--lap-printer") sniff = false;
    else if (a == "-l" || a == "--lap") { target_lap = (int)std::strtol(next(), nullptr, 16); have_lap = true; }     // apps/btrx:45
    else if (a == "-p" || a == "--hop") hop_mode = true;
    else if (a == "--tile") tile = std::atol(next());
    else if (a == "--stats") stats = true;
    else if (a == "--aliased") aliased = true;                      // apps/btrx:37-38: aliasing receiver (USRP2 firmware)
    else if (a == "-w" || a == "--wireshark") tun = true;           // apps/btrx:57-58; BTB200_TUN_FILE redirects the frames to a file
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (in.empty()) { std::fprintf(stderr, "usage: btrx_b200 -f FREQ -r RATE -i FILE [-S | -L | -l LAP [-p]] [-w] [-s SNR] [-N n] [-2]\n"); return 2; }
  FILE *f = std::fopen(in.c_str(), "rb");
  if (!f) { std::perror(in.c_str()); return 2; }
  std::fseek(f, 0, SEEK_END);
  long total = std::ftell(f) / (shorts ? 4 : 8);
  std::fseek(f, 0, SEEK_SET);
  if (nsamples >= 0 && nsamples < total) total = nsamples;

  boost::shared_ptr<gr::bluetooth::multi_block> blk;
  try {
    // mode selection of apps/btrx:140-159: -S sniffer; no LAP: LAP printer; LAP + -p: hopper; LAP alone: UAP discovery
    if (sniff) blk = gr::bluetooth::multi_sniffer::make(rate, freq, snr, tun);
    else if (have_lap && hop_mode) blk = gr::bluetooth::multi_hopper::make(rate, freq, snr, target_lap, aliased, tun);
    else if (have_lap) blk = gr::bluetooth::multi_UAP::make(rate, freq, snr, target_lap);
    else blk = gr::bluetooth::multi_LAP::make(rate, freq, snr);      // print the LAP of every frame detected (also -L)
  } catch (const std::exception &e) {
    std::fprintf(stderr, "btrx_b200: %s\n", e.what());
    return 1;
  }
  const long H = blk->history(), S = (long)blk->samples_per_slot();
  if (tile < 1) tile = 1;
  const long file_total = total;
  total *= tile;
  std::vector<gr_complex> buf((size_t)(H - 1 + total));
  if (shorts) {
    std::vector<short> tmp((size_t)file_total * 2);
    if (std::fread(tmp.data(), 4, (size_t)file_total, f) != (size_t)file_total) return 2;
    for (long i = 0; i < file_total; i++) buf[(size_t)(H - 1 + i)] = gr_complex(tmp[2 * i], tmp[2 * i + 1]);
  } else if (std::fread(&buf[(size_t)(H - 1)], 8, (size_t)file_total, f) != (size_t)file_total) return 2;
  std::fclose(f);
  for (long t = 1; t < tile; t++)
    std::memcpy(&buf[(size_t)(H - 1 + t * file_total)], &buf[(size_t)(H - 1)], (size_t)file_total * sizeof(gr_complex));

  gr_vector_const_void_star inv(1);
  gr_vector_void_star outv;
  long consumed = 0;                    // new samples consumed so far
  const auto t_start = std::chrono::steady_clock::now();
  const long ncalls = (total + S - 1) / S;
  long k = 0;
  while (k < ncalls) {
    // offer the block as many whole slots as it batches; the last window must end inside the stream
    long n = (long)blk->batch_slots();
    if (k + n > ncalls) n = ncalls - k;
    inv[0] = &buf[(size_t)consumed];
    const int got = blk->work((int)((n - 1) * S + 1), inv, outv);
    if (got < 0) break;                 // WORK_DONE (multi_UAP: the UAP has been determined)
    consumed += got;
    k += got / S;
  }
  std::fflush(stdout);
  if (stats) {
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::fprintf(stderr, "{\"samples\": %ld, \"seconds\": %.6f, \"msamples_per_s\": %.3f, \"work_calls_slots\": %ld, \"device_ms\": %.3f}\n",
                 consumed, sec, consumed / sec / 1e6, k, blk->device_ms());
  }
  return 0;
}
