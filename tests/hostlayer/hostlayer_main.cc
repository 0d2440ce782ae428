// tests/hostlayer/hostlayer_main.cc -- TEST HARNESS.
// Feeds a list of hits (from any source -- the CPU test tier uses the oracle) into the product's
// native host packet layer (gr-bluetooth_b200/host/lib/bt_host.cc, SnifferHost) and prints what
// the sniffer block would print.  Record format: u32 slot, i32 kind, f64 freq, f64 snr, i32 len,
// then len symbol bytes.  Optional second argument: file that receives the TAP frames.
#include "../../gr-bluetooth_b200/host/lib/bt_host.h"
#include <cstdio>
#include <fcntl.h>
#include <vector>

int main(int argc, char **argv)
{
  if (argc < 2) return 2;
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  btb200_host::SnifferHost host;
  if (argc > 2) host.set_tun_fd(open(argv[2], O_WRONLY | O_CREAT | O_TRUNC, 0644));   // Wireshark frames -> file
  for (;;) {
    uint32_t slot; int32_t kind, len; double freq, snr;
    if (std::fread(&slot, 4, 1, f) != 1) break;
    if (std::fread(&kind, 4, 1, f) != 1 || std::fread(&freq, 8, 1, f) != 1 || std::fread(&snr, 8, 1, f) != 1 ||
        std::fread(&len, 4, 1, f) != 1) return 3;
    std::vector<char> sym((size_t)len + 1);
    if (len && std::fread(sym.data(), 1, (size_t)len, f) != (size_t)len) return 3;
    if (kind == 0) host.ac(sym.data(), len, slot & 0x7ffffff, freq, snr);
    else host.aa(sym.data(), len, slot & 0x7ffffff, freq, snr);
  }
  std::fclose(f);
  return 0;
}
