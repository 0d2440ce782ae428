// bt_host.cc -- see bt_host.h.
#include <fcntl.h>
#include <linux/if.h>
#include <linux/if_tun.h>
#include <sys/ioctl.h>
#include <unistd.h>
#include "bt_host.h"
#include "../../csrc/hop_select.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace btb200_host {

// ---------------------------------------------------------------------------------------------
// whitening: one LFSR (x^7 + x^4 + 1) for BR and LE
namespace {
struct WhiteTables {
  uint8_t seq[127];
  int classic[64];
  int le[40];
  WhiteTables()
  {
    uint8_t r[7] = {1, 1, 1, 1, 1, 1, 1};
    for (int i = 0; i < 127; i++) { seq[i] = r[6]; step(r); }
    // BR: register = (CLK1..CLK6, 1)        (Core spec vol 2 part B 7.2)
    for (int c = 0; c < 64; c++) {
      uint8_t q[7];
      for (int i = 0; i < 6; i++) q[i] = (c >> i) & 1;
      q[6] = 1;
      classic[c] = locate(q);
    }
    // LE: register = (1, channel index MSB..LSB)   (vol 6 part B 3.2)
    for (int c = 0; c < 40; c++) {
      uint8_t q[7];
      q[0] = 1;
      for (int i = 0; i < 6; i++) q[1 + i] = (c >> (5 - i)) & 1;
      le[c] = locate(q);
    }
  }
  static void step(uint8_t *r)
  {
    const uint8_t fb = r[6];
    r[6] = r[5]; r[5] = r[4]; r[4] = r[3] ^ fb; r[3] = r[2]; r[2] = r[1]; r[1] = r[0]; r[0] = fb;
  }
  int locate(const uint8_t *reg) const
  {
    uint8_t q[7], outs[7];
    std::memcpy(q, reg, 7);
    for (int i = 0; i < 7; i++) { outs[i] = q[6]; step(q); }
    for (int off = 0; off < 127; off++) {
      bool ok = true;
      for (int i = 0; i < 7 && ok; i++) ok = seq[(off + i) % 127] == outs[i];
      if (ok) return off;
    }
    return 0;
  }
};
const WhiteTables &tables() { static const WhiteTables t; return t; }

inline uint8_t reverse8(uint8_t b)                 // packet::reverse, packet_impl.cc:77-82
{
  uint8_t r = 0;
  for (int i = 0; i < 8; i++) r |= ((b >> i) & 1) << (7 - i);
  return r;
}

// parity of a systematic cyclic code by polynomial division (classic_packet::lfsr, packet_impl.cc:278-306)
void lfsr_parity(const char *data, int length, int k, const uint8_t *g, uint8_t *cw)
{
  const int n = length - k;
  std::memset(cw, 0, (size_t)n);
  for (int i = k - 1; i >= 0; i--) {
    const uint8_t fb = (uint8_t)((data[i] & 1) ^ cw[n - 1]);
    for (int j = n - 1; j > 0; j--) cw[j] = cw[j - 1] ^ (g[j] & fb);
    cw[0] = g[0] & fb;
  }
}
}  // namespace

const uint8_t *whitening_sequence() { return tables().seq; }
int classic_whitening_index(int clk6) { return tables().classic[clk6 & 0x3f]; }
int le_whitening_index(int idx) { return tables().le[idx]; }

int le_freq_to_index(double freq)
{
  if (!(freq >= 2402000000.0 && freq <= 2480000000.0)) return -1;
  if (!(std::fmod(freq, 2000000.0) < 5000.0)) return -1;
  const int chan = (int)((freq - 2402000000.0) / 2000000.0);
  if (chan < 0 || chan > 39) return -1;
  if (chan == 0) return 37;
  if (chan == 12) return 38;
  if (chan == 39) return 39;
  return chan < 12 ? chan - 1 : chan - 2;
}

// ---------------------------------------------------------------------------------------------
ClassicPacket::ClassicPacket(const char *stream, int length, uint32_t clkn_, double freq)
    : clkn(clkn_), d_sym(MAX_SYMBOLS + 64, 0), d_payload(2744 + 64, 0)
{
  if (length > MAX_SYMBOLS) length = MAX_SYMBOLS;
  if (length < 0) length = 0;
  for (int i = 0; i < length; i++) d_sym[i] = stream[i] & 1;
  d_length = length;
  d_lap = air_to_host(&d_sym[38], 24);
  channel = (freq >= 2402000000.0 && freq <= 2480000000.0) ? (int)((freq - 2402000000.0) / 1000000.0) : -1;
  std::memset(d_packet_header, 0, sizeof d_packet_header);
  std::memset(d_payload_header, 0, sizeof d_payload_header);
}

bool ClassicPacket::header_present() const
{
  if (d_length < 126) return false;
  const char *s = &d_sym[67];
  int be = 0;
  const char msb = s[0];
  be += s[1] ^ !msb;
  be += s[2] ^ msb;
  be += s[3] ^ !msb;
  be += s[4] ^ msb;
  s += 5;
  for (int a = 0; a < 54; a += 3)
    be += ((s[a] ^ s[a + 1]) | (s[a + 1] ^ s[a + 2]) | (s[a + 2] ^ s[a]));
  return be < 5;                                   // ID_THRESHOLD, packet.h:185
}

bool ClassicPacket::unfec13(const char *in, char *out, int length)
{
  int be = 0;
  for (int i = 0; i < length; i++) {
    const int a = 3 * i, b = a + 1, c = a + 2;
    out[i] = (char)((in[a] & in[b]) | (in[b] & in[c]) | (in[c] & in[a]));
    be += ((in[a] ^ in[b]) | (in[b] ^ in[c]) | (in[c] ^ in[a]));
  }
  return be < (length / 4);
}

// (15,10) shortened Hamming blocks.  The reference's correction switch can never match (its
// syndrome word keeps the mismatch count in the upper bits), so a block decodes only when at most
// one of its five parity bits disagrees, and the whole call fails otherwise.
bool ClassicPacket::unfec23(const char *in, int length, std::vector<char> &out)
{
  static const uint8_t fecgen[6] = {1, 1, 0, 1, 0, 1};
  if (length % 10) length += 10 - (length % 10);
  const int blocks = length / 10;
  out.assign((size_t)length, 0);
  for (int b = 0; b < blocks; b++) {
    const char *blk = in + 15 * b;
    for (int i = 0; i < 10; i++) out[(size_t)(10 * b + i)] = blk[i];
    uint8_t cw[5];
    lfsr_parity(blk, 15, 10, fecgen, cw);
    int diff = 0;
    for (int i = 0; i < 5; i++) diff += (cw[i] != (uint8_t)(blk[10 + i] & 1));
    if (diff > 1) return false;
  }
  return true;
}

void ClassicPacket::unwhiten(const char *in, char *out, int clock, int length, int skip) const
{
  const uint8_t *w = whitening_sequence();
  int index = (classic_whitening_index(clock) + skip) % 127;
  for (int i = 0; i < length; i++) {
    out[i] = (char)((in[i] ^ w[index]) & 1);       // d_whitened is always true for sniffed packets
    index = (index + 1) % 127;
  }
}

uint16_t ClassicPacket::crcgen(const char *payload, int length, int uap)
{
  uint16_t reg = (uint16_t)((reverse8((uint8_t)uap) << 8) & 0xff00);
  for (int i = 0; i < length; i++) {
    const char byte = payload[i];
    reg = (uint16_t)((reg >> 1) | (((reg & 0x0001) ^ (byte & 0x01)) << 15));
    reg ^= ((reg & 0x8000) >> 5);
    reg ^= ((reg & 0x8000) >> 12);
  }
  return reg;
}

int ClassicPacket::uap_from_hec(uint16_t data, uint8_t hec)
{
  for (int i = 9; i >= 0; i--) {
    if (hec & 0x80) hec ^= 0x65;
    hec = (uint8_t)((hec << 1) | (((hec >> 7) ^ (data >> i)) & 0x01));
  }
  return reverse8(hec);
}

uint8_t ClassicPacket::try_clock(int clock)
{
  char header[18], unwhitened[18];
  if (!unfec13(&d_sym[72], header, 18)) return 0;
  unwhiten(header, unwhitened, clock, 18, 0);
  const uint16_t hdr_data = (uint16_t)air_to_host(unwhitened, 10);
  const uint8_t hec = (uint8_t)air_to_host(&unwhitened[10], 8);
  d_uap = (uint8_t)uap_from_hec(hdr_data, hec);
  d_type = (int)air_to_host(&unwhitened[3], 4);
  return d_uap;
}

void ClassicPacket::set_clock(uint32_t clock, bool have27)
{
  d_clock = have27 ? (clock & 0x7ffffff) : (clock & 0x3f);
  d_have_clk6 = true;
  d_have_clk27 = have27;
}

std::vector<uint8_t> ClassicPacket::tun_format() const
{
  std::vector<uint8_t> out((size_t)9 + (size_t)d_payload_length);
  for (int i = 0; i < 4; i++) out[(size_t)i] = (uint8_t)(d_clock >> (8 * i));
  out[4] = (uint8_t)channel;
  out[5] = (uint8_t)((d_have_clk27 ? 1 : 0) | ((d_have_nap ? 1 : 0) << 1));
  out[6] = (uint8_t)air_to_host(&d_packet_header[0], 7);      /* LT_ADDR and type */
  out[7] = (uint8_t)air_to_host(&d_packet_header[7], 3);      /* flags */
  out[8] = (uint8_t)air_to_host(&d_packet_header[10], 8);     /* HEC */
  for (int i = 0; i < d_payload_length; i++) out[(size_t)9 + (size_t)i] = (uint8_t)air_to_host(&d_payload[(size_t)i * 8], 8);
  return out;
}

bool ClassicPacket::payload_crc() const
{
  const uint16_t crc = crcgen(d_payload.data(), (d_payload_length - 2) * 8, d_uap);
  const uint16_t check = (uint16_t)air_to_host(&d_payload[(size_t)(d_payload_length - 2) * 8], 16);
  return crc == check;
}

int ClassicPacket::crc_check(int clock)
{
  int retval = 1;
  switch (d_type) {
    case 2: retval = fhs(clock); break;
    case 8: case 3: case 10: case 14: retval = DM(clock); break;
    case 4: case 11: case 15: retval = DH(clock); break;
    case 7: retval = EV3(clock); break;
    case 12: retval = EV4(clock); break;
    case 13: retval = EV5(clock); break;
    case 5: retval = HV(clock); break;
    default: break;
  }
  if (retval == 0 && (d_type != 2 && d_type != 3 && d_type != 5)) return 1;
  if (retval > 1 && (d_type == 7 || d_type == 13)) return 1;
  return retval;
}

int ClassicPacket::fhs(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  d_payload_length = 20;
  if (size < d_payload_length * 12) return 1;
  std::vector<char> corrected;
  if (!unfec23(stream, d_payload_length * 8, corrected)) return 0;
  unwhiten(corrected.data(), d_payload.data(), clock, d_payload_length * 8, 18);
  if (payload_crc()) return 1000;
  for (clock = 32; clock < 64; clock++) {
    unwhiten(corrected.data(), d_payload.data(), clock, d_payload_length * 8, 18);
    if (payload_crc()) return 1000;
  }
  return 0;
}

bool ClassicPacket::decode_payload_header(const char *stream, int clock, int header_bytes, int size, bool fec)
{
  std::vector<char> corrected;
  if (header_bytes == 2) {
    if (size < 16) return false;
    if (fec) {
      if (size < 30) return false;
      if (!unfec23(stream, 16, corrected)) return false;
      unwhiten(corrected.data(), d_payload_header, clock, 16, 18);
    } else {
      unwhiten(stream, d_payload_header, clock, 16, 18);
    }
    d_payload_length = (int)air_to_host(&d_payload_header[3], 10) + 4;
  } else {
    if (size < 8) return false;
    if (fec) {
      if (size < 15) return false;
      if (!unfec23(stream, 8, corrected)) return false;
      unwhiten(corrected.data(), d_payload_header, clock, 8, 18);
    } else {
      unwhiten(stream, d_payload_header, clock, 8, 18);
    }
    d_payload_length = (int)air_to_host(&d_payload_header[3], 5) + 3;
  }
  d_llid = (int)air_to_host(&d_payload_header[0], 2);
  d_flow = (int)air_to_host(&d_payload_header[2], 1);
  d_payload_header_length = header_bytes;
  return true;
}

int ClassicPacket::DM(int clock)
{
  int header_bytes = 2, max_length;
  const char *stream = &d_sym[126];
  int size = d_length - 126;
  switch (d_type) {
    case 8: stream += 80; size -= 80; header_bytes = 1; max_length = 12; break;
    case 3: header_bytes = 1; max_length = 20; break;
    case 10: max_length = 125; break;
    case 14: max_length = 228; break;
    default: return 0;
  }
  if (!decode_payload_header(stream, clock, header_bytes, size, true)) return 0;
  if (d_payload_length > max_length) return 1;
  const int bitlength = d_payload_length * 8;
  if (bitlength > size) return 1;
  std::vector<char> corrected;
  if (!unfec23(stream, bitlength, corrected)) return 0;
  unwhiten(corrected.data(), d_payload.data(), clock, bitlength, 18);
  if (payload_crc()) return 10;
  return 1;
}

int ClassicPacket::DH(int clock)
{
  int header_bytes = 2, max_length;
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  switch (d_type) {
    case 9: case 4: header_bytes = 1; max_length = 30; break;
    case 11: max_length = 187; break;
    case 15: max_length = 343; break;
    default: return 0;
  }
  if (!decode_payload_header(stream, clock, header_bytes, size, false)) return 0;
  if (d_payload_length > max_length) return 1;
  const int bitlength = d_payload_length * 8;
  if (bitlength > size) return 1;
  unwhiten(stream, d_payload.data(), clock, bitlength, 18);
  if (d_type == 9) return 1;
  if (payload_crc()) return 10;
  return 1;
}

int ClassicPacket::EV3(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  const int maxlength = 32;
  for (d_payload_length = 0; d_payload_length < maxlength; d_payload_length++) {
    const int bits = d_payload_length * 8;
    if ((bits + 8) > size) return 1;
    unwhiten(stream, &d_payload[(size_t)bits], clock, 8, 18 + bits);
    if ((d_payload_length > 2) && payload_crc()) return 10;
  }
  return 1;
}

int ClassicPacket::EV4(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  const int maxlength = 1470, minlength = 45;
  int syms = 0, bits = 0;
  d_payload_length = 1;
  std::vector<char> corrected;
  while (syms < maxlength) {
    if (syms + 15 > size) return 1;
    if (!unfec23(stream + syms, 10, corrected)) return (syms < minlength) ? 0 : 1;
    unwhiten(corrected.data(), &d_payload[(size_t)bits], clock, 10, 18 + bits);
    while (d_payload_length * 8 <= bits) {
      if (payload_crc()) return 10;
      d_payload_length++;
    }
    syms += 15;
    bits += 10;
  }
  return 1;
}

int ClassicPacket::EV5(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  const int maxlength = 182;
  for (d_payload_length = 0; d_payload_length < maxlength; d_payload_length++) {
    const int bits = d_payload_length * 8;
    if ((bits + 8) > size) return 1;
    unwhiten(stream, &d_payload[(size_t)bits], clock, 8, 18 + bits);
    if ((d_payload_length > 2) && payload_crc()) return 10;
  }
  return 1;
}

int ClassicPacket::HV(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126;
  if (size < 240) { d_payload_length = 0; return 1; }
  switch (d_type) {
    case 5: {
      char corrected[80];
      if (!unfec13(stream, corrected, 80)) return 0;
      d_payload_length = 10;
      unwhiten(corrected, d_payload.data(), clock, d_payload_length * 8, 18);
      break;
    }
    case 6: {
      std::vector<char> corrected;
      if (!unfec23(stream, 160, corrected)) return 0;
      d_payload_length = 20;
      unwhiten(corrected.data(), d_payload.data(), clock, d_payload_length * 8, 18);
      break;
    }
    case 7:
      d_payload_length = 30;
      unwhiten(stream, d_payload.data(), clock, d_payload_length * 8, 18);
      break;
  }
  return 1;
}

bool ClassicPacket::decode_header()
{
  char header[18];
  if (d_have_clk6 && unfec13(&d_sym[72], header, 18)) {
    unwhiten(header, d_packet_header, (int)d_clock, 18, 0);
    const uint16_t hdr_data = (uint16_t)air_to_host(d_packet_header, 10);
    const uint8_t hec = (uint8_t)air_to_host(&d_packet_header[10], 8);
    const uint8_t uap = (uint8_t)uap_from_hec(hdr_data, hec);
    if (uap == d_uap) {
      d_type = (int)air_to_host(&d_packet_header[3], 4);
      return true;
    }
    std::printf("bad HEC! %02x %02x %i ", uap, d_uap, (int)air_to_host(&d_packet_header[3], 4));
  }
  std::printf("failed to decode header\n");
  return false;
}

void ClassicPacket::decode_payload()
{
  d_payload_header_length = 0;
  const int clk = (int)d_clock;
  switch (d_type) {
    case 0: case 1: d_payload_length = 0; break;
    case 2: fhs(clk); break;
    case 3: DM(clk); break;
    case 4: DH(clk); break;
    case 5: HV(clk); break;
    case 6: HV(clk); break;
    case 7: if (EV3(clk) <= 1) HV(clk); break;
    case 8: DM(clk); break;
    case 9: DH(clk); break;
    case 10: DM(clk); break;
    case 11: DH(clk); break;
    case 12: EV4(clk); break;
    case 13: EV5(clk);        /* the reference falls through into DM5 here (packet_impl.cc:1146-1152) */
    case 14: DM(clk); break;
    case 15: DH(clk); break;
  }
  d_have_payload = true;
}

void ClassicPacket::decode()
{
  d_have_payload = false;
  if (decode_header()) decode_payload();
}

void ClassicPacket::print() const
{
  static const char *names[16] = {"NULL", "POLL", "FHS", "DM1", "DH1/2-DH1", "HV1", "HV2/2-EV3", "HV3/EV3/3-EV3",
                                  "DV/3-DH1", "AUX1", "DM3/2-DH3", "DH3/3-DH3", "EV4/2-EV5", "EV5/3-EV5",
                                  "DM5/2-DH5", "DH5/3-DH5"};
  if (!d_have_payload) return;
  std::printf("%s\n", names[d_type & 15]);
  if (d_payload_header_length > 0) {
    std::printf("  LLID: %d\n", d_llid);
    std::printf("  flow: %d\n", d_flow);
    std::printf("  payload length: %d\n", d_payload_length);
  }
}

// ---------------------------------------------------------------------------------------------
std::shared_ptr<ClassicPacket> Piconet::dequeue()
{
  if (d_queue.empty()) return nullptr;
  auto p = d_queue.front();
  d_queue.pop_front();
  return p;
}

// piconet_impl.cc:433-517: eliminate CLK1-6 candidates with the HEC (UAP consistency) and payload CRCs
bool Piconet::uap_from_header(ClassicPacket &pkt)
{
  int first_clock = 0, starting = 0, remaining = 0;
  const uint32_t clkn = pkt.clkn;
  if (!d_got_first_packet) d_first_pkt_time = clkn;
  if (d_packets_observed >= MAX_PATTERN_LENGTH) {
    std::printf("Oops. More hops than we can remember.\n");
    reset();
    return false;
  }
  d_pattern_indices[d_packets_observed] = (int)(clkn - d_first_pkt_time);
  d_pattern_channels[d_packets_observed] = (uint8_t)pkt.channel;
  d_packets_observed++;
  d_total_packets_observed++;
  for (int count = 0; count < 64; count++) {
    if (!d_got_first_packet || d_clock6_candidates[count] > -1) {
      const int clock = (int)(((uint32_t)count + clkn - d_first_pkt_time) % 64);
      starting++;
      const uint8_t uap = pkt.try_clock(clock);
      int retval = -1;
      if (!d_got_first_packet || uap == d_clock6_candidates[count]) retval = pkt.crc_check(clock);
      switch (retval) {
        case -1:
        case 0:
          d_clock6_candidates[count] = -1;
          break;
        case 1:
          d_clock6_candidates[count] = uap;
          first_clock = count;
          remaining++;
          break;
        default:
          std::printf("Correct CRC! UAP = 0x%x found after %d total packets.\n", uap, d_total_packets_observed);
          d_clk_offset = ((uint32_t)count - (d_first_pkt_time & 0x3f)) & 0x3f;
          d_uap = uap;
          d_have_clk6 = true;
          d_have_uap = true;
          d_total_packets_observed = 0;
          return true;
      }
    }
  }
  d_got_first_packet = true;
  std::printf("reduced from %d to %d CLK1-6 candidates\n", starting, remaining);
  if (remaining == 1) {
    d_clk_offset = ((uint32_t)first_clock - (d_first_pkt_time & 0x3f)) & 0x3f;
    d_uap = (uint8_t)d_clock6_candidates[first_clock];
    d_have_clk6 = true;
    d_have_uap = true;
    std::printf("We have a winner! UAP = 0x%x found after %d total packets.\n", d_uap, d_total_packets_observed);
    d_total_packets_observed = 0;
    return true;
  }
  if (remaining == 0) reset();
  return false;
}

void Piconet::reset()
{
  std::printf("no candidates remaining! starting over . . .\n");
  if (d_hop_reversal_inited) d_clock_candidates.clear();
  d_got_first_packet = false;
  d_packets_observed = 0;
  d_hop_reversal_inited = false;
  d_have_uap = false;
  d_have_clk6 = false;
  d_have_clk27 = false;
  /* two packets in a row on one channel hint at adaptive frequency hopping: try AFH next time */
  d_afh = d_looks_like_afh;
  d_looks_like_afh = false;
}

// hop selection kernel evaluated per clock value: csrc/hop_select.hpp (shared with the GPU candidate search)
int Piconet::hop_select(uint32_t addr, bool afh, uint32_t clock) { return btb200::hop_select(addr, afh, clock); }

Piconet::candidate_fn Piconet::s_candidate_fn = nullptr;

int Piconet::init_hop_reversal(bool aliased)
{
  std::printf("\nCalculating complete hopping sequence.\n");
  d_hop_addr = (((uint32_t)d_uap << 24) | d_lap) & 0xfffffffu;
  d_aliased = aliased;
  const uint32_t clock = (d_clk_offset + d_first_pkt_time) & 0x3f;
  // candidates: clock values with the known low bits whose hop lands on the first observed channel
  d_clock_candidates.clear();
  const int first_channel = d_pattern_channels[0];
  if (!(s_candidate_fn && s_candidate_fn(d_hop_addr, d_afh, d_aliased, clock, first_channel, d_clock_candidates))) {
    d_clock_candidates.clear();
    for (uint32_t c = clock; c < (uint32_t)SEQUENCE_LENGTH; c += 0x40) {
      const int hc = hop_select(d_hop_addr, d_afh, c);
      if ((d_aliased ? (int)aliased_channel((char)hc) : hc) == first_channel) d_clock_candidates.push_back(c);
    }
  }
  d_num_candidates = (int)d_clock_candidates.size();
  d_winnowed = 0;
  d_hop_reversal_inited = true;
  d_have_clk27 = false;
  std::printf("%d initial CLK1-27 candidates\n", d_num_candidates);
  return d_num_candidates;
}

int Piconet::winnow(int offset, char channel)
{
  int n = 0;
  for (int i = 0; i < d_num_candidates; i++) {
    const char s = hop((int)((d_clock_candidates[(size_t)i] + (uint32_t)offset) % (uint32_t)SEQUENCE_LENGTH));
    const char obs = d_aliased ? aliased_channel(s) : s;
    if (obs == channel) d_clock_candidates[(size_t)n++] = d_clock_candidates[(size_t)i];
  }
  d_num_candidates = n;
  if (n == 1) {
    d_clk_offset = (d_clock_candidates[0] - d_first_pkt_time) & 0x7ffffff;
    d_have_clk27 = true;
    std::printf("\nAcquired CLK1-27 offset = 0x%07x\n", d_clk_offset);
  } else if (n == 0) {
    reset();
  } else {
    std::printf("%d CLK1-27 candidates remaining\n", n);
  }
  return n;
}

int Piconet::winnow()
{
  int n = d_num_candidates;
  for (; d_winnowed < d_packets_observed; d_winnowed++) {
    const int index = d_pattern_indices[d_winnowed];
    const uint8_t channel = d_pattern_channels[d_winnowed];
    n = winnow(index, (char)channel);
    if (d_packets_observed > 0 && d_winnowed > 0) {
      const int last_index = d_pattern_indices[d_winnowed - 1];
      const uint8_t last_channel = d_pattern_channels[d_winnowed - 1];
      if (!d_looks_like_afh && (index == last_index + 1) && (channel == last_channel)) d_looks_like_afh = true;
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
void le_print(const char *stream, int available, double freq)
{
  const int index = le_freq_to_index(freq);
  const int MAXS = 8 * (1 + 4 + 39 + 3);           // LE_MAX_SYMBOLS, packet.h:283-285
  char link[8 * 47];
  for (int i = 0; i < MAXS; i++) link[i] = (i < available) ? (stream[i] & 1) : 0;
  const uint8_t *w = whitening_sequence();
  for (int i = 40, wi = le_whitening_index(index < 0 ? 0 : index); i < MAXS; i++, wi = (wi + 1) % 127) link[i] ^= w[wi];
  const uint32_t aa = air_to_host(&link[8], 32);
  const uint16_t header = (uint16_t)air_to_host(&link[40], 16);
  uint8_t pdu[48];
  int pi = 0;
  for (int i = 56; i + 8 < MAXS; pi++, i += 8) pdu[pi] = (uint8_t)air_to_host(&link[i], 8);
  if (index >= 37) {
    const int type = header & 0xf, txadd = (header >> 6) & 1, rxadd = (header >> 7) & 1;
    const unsigned len = (header >> 8) & 0x3f;
    std::printf("BTLE index=%02d, AA=%08x, PDUType=%d, TxAdd=%d, RxAdd=%d, Length=%d\n", index, aa, type, txadd, rxadd, len);
    switch (type) {
      case 0: case 2: case 4: case 6:
        std::printf("  AdvA=%02x%02x%02x%02x%02x%02x\n", pdu[0], pdu[1], pdu[2], pdu[3], pdu[4], pdu[5]);
        std::printf(type == 4 ? "\n  (char) ScanRspData=" : "\n  (char) AdvData=");
        for (unsigned i = 6; i < len; i++) {
          char c = (char)pdu[i];
          if ((c < ' ') || (c > '~')) c = '.';
          std::printf(" %c", c);
        }
        std::printf(type == 4 ? "\n  (byte) ScanRspData=" : "\n  (byte) AdvData=");
        for (unsigned i = 6; i < len; i++) std::printf("%02x", pdu[i]);
        std::printf("\n");
        break;
      case 1:
        std::printf("  AdvA=%02x%02x%02x%02x%02x%02x\n  InitA=%02x%02x%02x%02x%02x%02x\n", pdu[0], pdu[1], pdu[2], pdu[3],
                    pdu[4], pdu[5], pdu[6], pdu[7], pdu[8], pdu[9], pdu[10], pdu[11]);
        break;
      case 3:
        std::printf("  ScanA=%02x%02x%02x%02x%02x%02x\n  AdvA=%02x%02x%02x%02x%02x%02x\n", pdu[0], pdu[1], pdu[2], pdu[3],
                    pdu[4], pdu[5], pdu[6], pdu[7], pdu[8], pdu[9], pdu[10], pdu[11]);
        break;
      case 5: {
        std::printf("  InitA=%02x%02x%02x%02x%02x%02x\n  AdvA=%02x%02x%02x%02x%02x%02x\n", pdu[0], pdu[1], pdu[2], pdu[3],
                    pdu[4], pdu[5], pdu[6], pdu[7], pdu[8], pdu[9], pdu[10], pdu[11]);
        const uint32_t AA = pdu[12] | ((uint32_t)pdu[13] << 8) | ((uint32_t)pdu[14] << 16) | ((uint32_t)pdu[15] << 24);
        const uint32_t crcinit = pdu[16] | ((uint32_t)pdu[17] << 8) | ((uint32_t)pdu[18] << 16);
        const uint16_t winoff = (uint16_t)(pdu[20] | (pdu[21] << 8)), interval = (uint16_t)(pdu[22] | (pdu[23] << 8));
        const uint16_t latency = (uint16_t)(pdu[24] | (pdu[25] << 8)), timeout = (uint16_t)(pdu[26] | (pdu[27] << 8));
        const uint64_t chm = pdu[28] | ((uint64_t)pdu[29] << 8) | ((uint64_t)pdu[30] << 16) | ((uint64_t)pdu[31] << 24) |
                             ((uint64_t)pdu[32] << 32);
        std::printf("  AA=%08x, CRCInit=%06x, WinSize=%d, WinOffset=%d\n", AA, crcinit, pdu[19], winoff);
        std::printf("  Interval=%d, Latency=%d, Timeout=%d, ChM=%010lx, Hop=%d, SCA=%d\n", interval, latency, timeout,
                    (unsigned long)chm, pdu[33] & 0x1f, (pdu[33] >> 5) & 7);
        break;
      }
      default: break;
    }
  } else {
    std::printf("BTLE index=%02d, AA=%08x, LLID=%d, NESN=%d, SN=%d, MD=%d, Length=%d\n", index, aa, header & 3,
                (header >> 2) & 1, (header >> 3) & 1, (header >> 4) & 1, (header >> 8) & 0x1f);
  }
}

// ---------------------------------------------------------------------------------------------
/* handle AC, lib/multi_sniffer_impl.cc:169-206 */
void SnifferHost::ac(const char *symbols, int len, uint32_t clkn, double freq, double snr)
{
  auto pkt = std::make_shared<ClassicPacket>(symbols, len, clkn, freq);
  const uint32_t lap = pkt->lap();
  std::printf("time %6d, snr=%.1f, channel %2d, LAP %06x ", clkn, snr, pkt->channel, lap);
  if (pkt->header_present()) {
    auto &slot = d_piconets[(int)lap];
    if (!slot) slot = std::make_shared<Piconet>(lap);
    auto pn = slot;
    if (pn->have_clk6() && pn->have_uap()) decode(pkt, pn, true);
    else discover(pkt, pn);
    /* an inquiry response must not leave piconet state behind */
    if (lap == GIAC || lap == LIAC) d_piconets.erase((int)lap);
  } else {
    id(lap);
  }
}

/* handle AA, lib/multi_sniffer_impl.cc:208-227 */
void SnifferHost::aa(const char *symbols, int len, uint32_t clkn, double freq, double snr)
{
  std::printf("time %6d, snr=%.1f, ", clkn, snr);
  le_print(symbols, len, freq);
}

/* ID packet (no header), :229-236 */
void SnifferHost::id(uint32_t lap)
{
  std::printf("ID\n");
  write_frame(d_tunfd, nullptr, 0, 0, lap, TUN_ETHER_TYPE);
}

/* decode packets with headers, :238-281 */
void SnifferHost::decode(std::shared_ptr<ClassicPacket> pkt, std::shared_ptr<Piconet> pn, bool first_run)
{
  const uint32_t clock = pkt->clkn + pn->offset();
  pkt->set_clock(clock, pn->have_clk27());
  pkt->set_uap(pn->uap());
  pkt->decode();
  if (pkt->got_payload()) {
    pkt->print();
    if (d_tunfd >= 0) {
      /* destination address = NAP:UAP:LAP as far as known (:252-267) */
      uint64_t addr = ((uint32_t)pkt->uap() << 24) | pkt->lap();
      if (pn->have_nap()) {
        addr |= (uint64_t)pn->nap() << 32;
        pkt->set_nap(pn->nap());
      }
      const std::vector<uint8_t> data = pkt->tun_format();
      write_frame(d_tunfd, data.data(), (unsigned)data.size(), 0, addr, TUN_ETHER_TYPE);
    }
    if (pkt->type() == 2) fhs(pkt);
  } else if (first_run) {
    std::printf("lost clock!\n");
    pn->reset();
    discover(pkt, pn);           /* start rediscovery with this packet */
  } else {
    std::printf("Giving up on queued packet!\n");
  }
}

/* work on UAP/CLK1-6 discovery, :288-300 */
void SnifferHost::discover(std::shared_ptr<ClassicPacket> pkt, std::shared_ptr<Piconet> pn)
{
  std::printf("working on UAP/CLK1-6\n");
  pn->enqueue(pkt);
  if (pn->uap_from_header(*pkt)) recall(pn);
}

/* decode stored packets, :306-321 */
void SnifferHost::recall(std::shared_ptr<Piconet> pn)
{
  std::printf("Decoding queued packets\n");
  while (auto pkt = pn->dequeue()) {
    std::printf("time %6d, channel %2d, LAP %06x ", pkt->clkn, pkt->channel, pkt->lap());
    decode(pkt, pn, false);
  }
  std::printf("Finished decoding queued packets\n");
}

/* pull information out of an FHS packet, :326-369 */
void SnifferHost::fhs(std::shared_ptr<ClassicPacket> pkt)
{
  const uint32_t lap = pkt->lap_from_fhs();
  const uint8_t uap = pkt->uap_from_fhs();
  const uint16_t nap = pkt->nap_from_fhs();
  const uint32_t clk = pkt->clock_from_fhs() << 1;        /* units of 625 us */
  const uint32_t offset = (clk - pkt->clkn) & 0x7ffffff;
  std::printf("FHS contents: BD_ADDR ");
  std::printf("%2.2x:", (nap >> 8) & 0xff);
  std::printf("%2.2x:", nap & 0xff);
  std::printf("%2.2x:", uap);
  std::printf("%2.2x:", (lap >> 16) & 0xff);
  std::printf("%2.2x:", (lap >> 8) & 0xff);
  std::printf("%2.2x", lap & 0xff);
  std::printf(", CLK %07x\n", clk);
  auto &slot = d_piconets[(int)lap];
  if (!slot) slot = std::make_shared<Piconet>(lap);
  slot->set_uap(uap);
  slot->set_nap(nap);
  slot->set_offset(offset);
}


// ---------------------------------------------------------------------------------------------
HopperHost::SlotPlan HopperHost::plan(uint32_t clkn) const
{
  SlotPlan p;
  if (d_piconet.have_clk27()) {
    /* follow along on the predicted channel, multi_hopper_impl.cc:152-166 */
    p.hopalong = true;
    p.clock27 = (clkn + d_piconet.offset()) & 0x7ffffff;
    const int hopch = d_piconet.hop((int)p.clock27);
    p.obs_channel = d_aliased ? Piconet::aliased_channel((char)hopch) : hopch;
    if (p.obs_channel >= d_ch_lo && p.obs_channel <= d_ch_hi) { p.first_channel = hopch; p.n_channels = 1; }
  } else {
    p.first_channel = d_ch_lo;
    p.n_channels = d_ch_hi - d_ch_lo + 1;
    p.stop_lap = d_lap;
  }
  return p;
}

/* multi_hopper_impl.cc:107-135 */
bool HopperHost::scan_packet(uint32_t clkn, int channel, const char *symbols, int len)
{
  ClassicPacket pkt(symbols, len, clkn, 2402000000.0 + 1e6 * channel);
  if (!(pkt.lap() == d_lap && pkt.header_present())) return false;
  if (!d_piconet.have_clk6()) {
    /* working on CLK1-6/UAP discovery */
    d_piconet.uap_from_header(pkt);
    if (d_piconet.have_clk6()) {
      /* got CLK1-6/UAP, start working on CLK1-27 with the packets seen so far */
      d_piconet.init_hop_reversal(d_aliased);
      d_piconet.winnow();
    }
  } else {
    /* continue working on CLK1-27: timing of an additional packet */
    d_piconet.uap_from_header(pkt);
    if (d_piconet.have_clk6()) d_piconet.winnow();
  }
  return true;
}

/* multi_hopper_impl.cc:176-205 */
void HopperHost::hop_packet(const SlotPlan &p, const char *symbols, int len)
{
  ClassicPacket pkt(symbols, len, 0, 2402000000.0 + 1e6 * p.obs_channel);
  if (pkt.lap() != d_lap) return;
  std::printf("clock 0x%07x, channel %2d: ", p.clock27, pkt.channel);
  if (pkt.header_present()) {
    pkt.set_uap(d_piconet.uap());
    pkt.set_clock(p.clock27, true);
    pkt.decode();
    if (pkt.got_payload()) {
      pkt.print();
      if (d_tunfd >= 0) {
        /* the reference keeps this address in an int (:190-193): sign-extended when UAP >= 0x80 */
        const int addr = (int)(((uint32_t)pkt.uap() << 24) | pkt.lap());
        const std::vector<uint8_t> data = pkt.tun_format();
        write_frame(d_tunfd, data.data(), (unsigned)data.size(), 0, (uint64_t)(int64_t)addr, TUN_ETHER_TYPE);
      }
    }
  } else {
    std::printf("ID\n");
    if (d_tunfd >= 0) {
      const int addr = (int)(((uint32_t)d_piconet.uap() << 24) | pkt.lap());
      write_frame(d_tunfd, nullptr, 0, 0, (uint64_t)(int64_t)addr, TUN_ETHER_TYPE);
    }
  }
}

/* multi_UAP_impl.cc:100-113 */
bool UapHost::packet(uint32_t clkn, int channel, const char *symbols, int len)
{
  ClassicPacket pkt(symbols, len, clkn, 2402000000.0 + 1e6 * channel);
  if (!(pkt.lap() == d_lap && pkt.header_present())) return false;
  if (d_piconet.uap_from_header(pkt)) d_done = true;           /* the reference exits the process here */
  return true;
}

/* ---- Wireshark interface (lib/tun.cc) ---------------------------------------------------- */
int write_frame(int fd, const uint8_t *data, unsigned data_len, uint64_t src_addr, uint64_t dst_addr,
                unsigned short ether_type)
{
  if (fd < 0) return (int)data_len;
  const unsigned MTU = 1500, HDR = 14;
  uint8_t frame[MTU] = {0};
  for (int i = 0; i < 6; i++) {
    const int shift = 8 * (5 - i);
    frame[i] = (uint8_t)(dst_addr >> shift);
    frame[6 + i] = (uint8_t)(src_addr >> shift);
  }
  frame[12] = (uint8_t)(ether_type >> 8);
  frame[13] = (uint8_t)ether_type;
  const unsigned room = MTU - HDR, n = data_len < room ? data_len : room;
  if (n) std::memcpy(frame + HDR, data, n);
  if (::write(fd, frame, HDR + n) == -1) {
    std::perror("write");
    return -1;
  }
  return (int)data_len;
}

int open_tap(const char *name)
{
  const int fd = ::open("/dev/net/tun", O_RDWR);
  if (fd < 0) return -1;
  struct ifreq ifr;
  std::memset(&ifr, 0, sizeof ifr);
  ifr.ifr_flags = IFF_TAP | IFF_NO_PI;
  std::strncpy(ifr.ifr_name, name, IFNAMSIZ - 1);
  if (::ioctl(fd, TUNSETIFF, (void *)&ifr) < 0) { ::close(fd); return -1; }
  return fd;
}

int open_tun_output()
{
  int fd = -1;
  if (const char *path = std::getenv("BTB200_TUN_FILE")) fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  else fd = open_tap("btbb");                                   /* lib/multi_sniffer_impl.cc:64-66 */
  if (fd < 0) std::fprintf(stderr, "warning: was not able to open TUN device, disabling Wireshark interface\n");
  return fd;
}

}  // namespace btb200_host
