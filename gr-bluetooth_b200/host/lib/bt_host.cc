// bt_host.cc -- see bt_host.h.
#include <fcntl.h>
#include <linux/if.h>
#include <linux/if_tun.h>
#include <sys/ioctl.h>
#include <unistd.h>
#include "bt_host.h"
#include "../../csrc/hop_select.hpp"
#include <algorithm>
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

// A packet header follows the access code when the 4-symbol trailer alternates from the sync word's last bit and the
// 54 header symbols look like 18 bits sent three times each (1/3 FEC).  Both checks are Hamming distances of packed
// words: trailer ^ expected pattern, and (h ^ h>>1 | h>>1 ^ h>>2) sampled at every third position counts the triples
// whose three copies disagree.  Present when fewer than ID_THRESHOLD (5, packet.h:185) symbols are off
// (classic_packet_impl::header_present, packet_impl.cc:1205-1242).
bool ClassicPacket::header_present() const
{
  if (d_length < 126) return false;
  uint64_t trailer = 0, hdr = 0;
  for (int i = 0; i < 5; i++) trailer |= (uint64_t)(d_sym[67 + i] & 1) << i;
  for (int i = 0; i < 54; i++) hdr |= (uint64_t)(d_sym[72 + i] & 1) << i;
  const uint64_t want4 = (trailer & 1) ? 0xAull : 0x5ull;      // symbols 1..4 after a 1: 0,1,0,1; after a 0: 1,0,1,0
  int off = __builtin_popcountll(((trailer >> 1) & 0xF) ^ want4);
  uint64_t every_third = 0;
  for (int t = 0; t < 18; t++) every_third |= 1ull << (3 * t);
  const uint64_t d01 = hdr ^ (hdr >> 1), d12 = (hdr >> 1) ^ (hdr >> 2);
  off += __builtin_popcountll((d01 | d12) & every_third);
  return off < 5;
}

// 1/3 repetition code: majority vote; the block is accepted while fewer than a quarter of the triples disagree
// (classic_packet_impl::unfec13, packet_impl.cc:367-384)
bool ClassicPacket::unfec13(const char *in, char *out, int length)
{
  int split = 0;
  for (int i = 0; i < length; i++) {
    const int ones = (in[3 * i] & 1) + (in[3 * i + 1] & 1) + (in[3 * i + 2] & 1);
    out[i] = (char)(ones >> 1);
    split += (ones == 1 || ones == 2);
  }
  return split < length / 4;
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

// Payload CRC: CRC-16/CCITT, LSB first (polynomial 0x1021 reflected = 0x8408), register preset with the bit-reversed UAP
// in its upper byte (classic_packet_impl::crcgen, packet_impl.cc:529-548).  Whole bytes go through a 256-entry table.
uint16_t ClassicPacket::crcgen(const char *payload, int length, int uap)
{
  static const struct Table {
    uint16_t t[256];
    Table()
    {
      for (int v = 0; v < 256; v++) {
        uint16_t r = (uint16_t)v;
        for (int b = 0; b < 8; b++) r = (uint16_t)((r >> 1) ^ ((r & 1) ? 0x8408 : 0));
        t[v] = r;
      }
    }
  } crc;
  uint16_t reg = (uint16_t)(reverse8((uint8_t)uap) << 8);
  int i = 0;
  for (; i + 8 <= length; i += 8) {
    uint8_t byte = 0;
    for (int b = 0; b < 8; b++) byte |= (uint8_t)((payload[i + b] & 1) << b);
    reg = (uint16_t)((reg >> 8) ^ crc.t[(reg ^ byte) & 0xff]);
  }
  for (; i < length; i++) reg = (uint16_t)((reg >> 1) ^ (((reg ^ payload[i]) & 1) ? 0x8408 : 0));
  return reg;
}

// The HEC is an 8-bit LFSR (x^8 + x^7 + x^5 + x^2 + x + 1) preset with the UAP and clocked over the 10 header bits, so
// for a received (header data, HEC) pair the UAP that would have produced it is a GF(2)-LINEAR function of the 18 bits:
// two XOR tables, built once from the responses to the 18 unit vectors (classic_packet_impl::UAP_from_hec,
// packet_impl.cc:593-606 runs the register backwards for every call).
int ClassicPacket::uap_from_hec(uint16_t data, uint8_t hec)
{
  static const struct Tables {
    uint8_t by_hec[256], by_data[1024];
    static uint8_t unit(uint16_t data, uint8_t hec)
    {
      // one backward pass of the register for a single input pattern
      for (int i = 9; i >= 0; i--) {
        if (hec & 0x80) hec ^= 0x65;
        hec = (uint8_t)((hec << 1) | (((hec >> 7) ^ (data >> i)) & 1));
      }
      return reverse8(hec);
    }
    Tables()
    {
      uint8_t hb[8], db[10];
      for (int b = 0; b < 8; b++) hb[b] = unit(0, (uint8_t)(1u << b));
      for (int b = 0; b < 10; b++) db[b] = unit((uint16_t)(1u << b), 0);
      for (int v = 0; v < 256; v++) { uint8_t r = 0; for (int b = 0; b < 8; b++) if (v >> b & 1) r ^= hb[b]; by_hec[v] = r; }
      for (int v = 0; v < 1024; v++) { uint8_t r = 0; for (int b = 0; b < 10; b++) if (v >> b & 1) r ^= db[b]; by_data[v] = r; }
    }
  } lin;
  return lin.by_hec[hec] ^ lin.by_data[data & 0x3ff];
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

// ---- payload formats, one row per TYPE code of the packet header (Core spec vol 2 part B 6.5) ---------------------
// The reference spreads this over seven member functions with a switch on the type inside each (DM, DH, EV3, EV4,
// EV5, HV, fhs: packet_impl.cc:688-1042) and two more switches that pick among them (crc_check :609-668,
// decode_payload :1092-1158); here the differences are data.  Return values keep the reference's meaning:
// 0 = this cannot be such a packet, 1 = plausible, 10 = CRC correct, 1000 = FHS with correct CRC.
namespace {
enum class Body : uint8_t { None, Fhs, Acl, Scan, Ev4, Sco };
struct PayloadFormat {
  Body body;           // how the payload is parsed
  bool fec23;          // Acl: protected by the (15,10) code
  uint8_t hdr_bytes;   // Acl: payload header length
  uint16_t limit;      // Acl: largest payload (header + body + CRC) in bytes; Scan: bytes tried
  uint8_t skip;        // Acl: voice symbols in front of the data field (DV)
  bool crc;            // Acl: has a CRC (AUX1 has none)
  bool in_discovery;   // crc_check() looks at this type (the others always count as "plausible")
};
const PayloadFormat kFormat[16] = {
    /* 0 NULL  */ {Body::None, false, 0, 0, 0, false, false},
    /* 1 POLL  */ {Body::None, false, 0, 0, 0, false, false},
    /* 2 FHS   */ {Body::Fhs, true, 0, 20, 0, true, true},
    /* 3 DM1   */ {Body::Acl, true, 1, 20, 0, true, true},
    /* 4 DH1   */ {Body::Acl, false, 1, 30, 0, true, true},
    /* 5 HV1   */ {Body::Sco, false, 0, 10, 0, false, true},
    /* 6 HV2   */ {Body::Sco, false, 0, 20, 0, false, false},
    /* 7 HV3/EV3 */ {Body::Scan, false, 0, 32, 0, true, true},
    /* 8 DV    */ {Body::Acl, true, 1, 12, 80, true, true},
    /* 9 AUX1  */ {Body::Acl, false, 1, 30, 0, false, false},
    /* 10 DM3  */ {Body::Acl, true, 2, 125, 0, true, true},
    /* 11 DH3  */ {Body::Acl, false, 2, 187, 0, true, true},
    /* 12 EV4  */ {Body::Ev4, true, 0, 0, 0, true, true},
    /* 13 EV5  */ {Body::Scan, false, 0, 182, 0, true, true},
    /* 14 DM5  */ {Body::Acl, true, 2, 228, 0, true, true},
    /* 15 DH5  */ {Body::Acl, false, 2, 343, 0, true, true},
};
}  // namespace

// what a payload under the whitening of `clock` says about the CLK1-6 / UAP hypothesis (packet_impl.cc:609-668)
int ClassicPacket::crc_check(int clock)
{
  const PayloadFormat &f = kFormat[d_type & 15];
  if (!f.in_discovery) return 1;
  int verdict = 1;
  switch (f.body) {
    case Body::Fhs: verdict = fhs(clock); break;
    case Body::Acl: verdict = acl(clock); break;
    case Body::Scan: verdict = crc_scan(clock); break;
    case Body::Ev4: verdict = EV4(clock); break;
    case Body::Sco: verdict = sco(clock); break;
    default: break;
  }
  // only FHS, DM1 and HV1 are trusted to rule a clock OUT; a CRC hit on a guessed-length EV3/EV5 proves nothing
  if (verdict == 0 && !(d_type == 2 || d_type == 3 || d_type == 5)) return 1;
  if (verdict > 1 && (d_type == 7 || d_type == 13)) return 1;
  return verdict;
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

// ACL-style payload: payload header, body, CRC -- DM1/3/5, DH1/3/5, DV's data field, AUX1
// (classic_packet_impl::DM / DH, packet_impl.cc:772-870)
int ClassicPacket::acl(int clock)
{
  const PayloadFormat &f = kFormat[d_type & 15];
  if (f.body != Body::Acl) return 0;
  const char *stream = &d_sym[126 + f.skip];
  const int size = d_length - 126 - f.skip;
  if (!decode_payload_header(stream, clock, f.hdr_bytes, size, f.fec23)) return 0;
  if (d_payload_length > f.limit) return 1;
  const int bits = d_payload_length * 8;
  if (bits > size) return 1;
  if (f.fec23) {
    std::vector<char> corrected;
    if (!unfec23(stream, bits, corrected)) return 0;
    unwhiten(corrected.data(), d_payload.data(), clock, bits, 18);
  } else {
    unwhiten(stream, d_payload.data(), clock, bits, 18);
  }
  if (!f.crc) return 1;
  return payload_crc() ? 10 : 1;
}

// eSCO payloads of unknown length without FEC (EV3, EV5): grow the payload a byte at a time until a CRC fits
// (classic_packet_impl::EV3 / EV5, packet_impl.cc:872-899, 950-977)
int ClassicPacket::crc_scan(int clock)
{
  const char *stream = &d_sym[126];
  const int size = d_length - 126, limit = kFormat[d_type & 15].limit;
  for (d_payload_length = 0; d_payload_length < limit; d_payload_length++) {
    const int bits = d_payload_length * 8;
    if (bits + 8 > size) return 1;
    unwhiten(stream, &d_payload[(size_t)bits], clock, 8, 18 + bits);
    if (d_payload_length > 2 && payload_crc()) return 10;
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

// SCO voice: 240 symbols carrying 10, 20 or 30 bytes under 1/3, 2/3 or no FEC, no CRC (classic_packet_impl::HV,
// packet_impl.cc:979-1042; the reference also sends EV3 candidates here as HV3)
int ClassicPacket::sco(int clock)
{
  const char *stream = &d_sym[126];
  if (d_length - 126 < 240) { d_payload_length = 0; return 1; }
  if (d_type == 5) {
    char voice[80];
    if (!unfec13(stream, voice, 80)) return 0;
    d_payload_length = 10;
    unwhiten(voice, d_payload.data(), clock, 80, 18);
  } else if (d_type == 6) {
    std::vector<char> voice;
    if (!unfec23(stream, 160, voice)) return 0;
    d_payload_length = 20;
    unwhiten(voice.data(), d_payload.data(), clock, 160, 18);
  } else if (d_type == 7) {
    d_payload_length = 30;
    unwhiten(stream, d_payload.data(), clock, 240, 18);
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

// payload of a packet whose header decoded (classic_packet_impl::decode_payload, packet_impl.cc:1092-1158)
void ClassicPacket::decode_payload()
{
  d_payload_header_length = 0;
  const int clk = (int)d_clock;
  switch (kFormat[d_type & 15].body) {
    case Body::None: d_payload_length = 0; break;
    case Body::Fhs: fhs(clk); break;
    case Body::Acl: acl(clk); break;
    case Body::Sco: sco(clk); break;
    case Body::Ev4: EV4(clk); break;
    case Body::Scan:
      // type 7 is EV3 when a CRC fits and HV3 otherwise; type 13 is EV5 (the reference then falls into its DM5 case,
      // which does nothing for this type: packet_impl.cc:1146-1152)
      if (crc_scan(clk) <= 1 && d_type == 7) sco(clk);
      break;
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

// CLK1-6 / UAP discovery (basic_rate_piconet_impl::UAP_from_header, piconet_impl.cc:433-517).  Every packet with a
// header is tried under each of the up to 64 clock hypotheses still alive: the header's HEC yields the UAP that
// hypothesis implies, and the payload CRC under that UAP and whitening says whether the pair is impossible (drop
// it), possible (keep it) or certain (done).  A hypothesis survives only while it keeps implying the same UAP.
bool Piconet::uap_from_header(ClassicPacket &pkt)
{
  const uint32_t clkn = pkt.clkn;
  const bool first = !d_got_first_packet;
  if (first) d_first_pkt_time = clkn;
  if (d_packets_observed >= MAX_PATTERN_LENGTH) {
    std::printf("Oops. More hops than we can remember.\n");
    reset();
    return false;
  }
  // remember when and where it was seen: the hop reversal replays these observations
  d_pattern_indices[d_packets_observed] = (int)(clkn - d_first_pkt_time);
  d_pattern_channels[d_packets_observed] = (uint8_t)pkt.channel;
  d_packets_observed++;
  d_total_packets_observed++;

  const uint32_t elapsed = clkn - d_first_pkt_time;
  int tried = 0, alive = 0, survivor = 0;
  for (int h = 0; h < 64; h++) {
    if (!first && d_clock6_candidates[h] < 0) continue;               // hypothesis already ruled out
    tried++;
    const int clock = (int)(((uint32_t)h + elapsed) % 64);             // CLK1-6 of THIS packet under hypothesis h
    const uint8_t uap = pkt.try_clock(clock);
    const bool consistent = first || uap == d_clock6_candidates[h];
    const int verdict = consistent ? pkt.crc_check(clock) : -1;
    if (verdict > 1) {
      // a payload CRC that checks out under this whitening and UAP settles both at once
      std::printf("Correct CRC! UAP = 0x%x found after %d total packets.\n", uap, d_total_packets_observed);
      d_clk_offset = ((uint32_t)h - (d_first_pkt_time & 0x3f)) & 0x3f;
      d_uap = uap;
      d_have_clk6 = d_have_uap = true;
      d_total_packets_observed = 0;
      return true;
    }
    if (verdict == 1) { d_clock6_candidates[h] = uap; survivor = h; alive++; }
    else d_clock6_candidates[h] = -1;
  }
  d_got_first_packet = true;
  std::printf("reduced from %d to %d CLK1-6 candidates\n", tried, alive);
  if (alive == 1) {
    d_clk_offset = ((uint32_t)survivor - (d_first_pkt_time & 0x3f)) & 0x3f;
    d_uap = (uint8_t)d_clock6_candidates[survivor];
    d_have_clk6 = d_have_uap = true;
    std::printf("We have a winner! UAP = 0x%x found after %d total packets.\n", d_uap, d_total_packets_observed);
    d_total_packets_observed = 0;
    return true;
  }
  if (alive == 0) reset();
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
  // the reference picks its first candidates BEFORE it stores the new aliasing flag (piconet_impl.cc:118-124), i.e. with
  // the flag of the previous attempt (false the first time): kept, the printed candidate counts depend on it
  const bool aliased_for_candidates = d_aliased;
  const uint32_t clock = (d_clk_offset + d_first_pkt_time) & 0x3f;
  // candidates: clock values with the known low bits whose hop lands on the first observed channel
  d_clock_candidates.clear();
  const int first_channel = d_pattern_channels[0];
  if (!(s_candidate_fn && s_candidate_fn(d_hop_addr, d_afh, aliased_for_candidates, clock, first_channel, d_clock_candidates))) {
    d_clock_candidates.clear();
    for (uint32_t c = clock; c < (uint32_t)SEQUENCE_LENGTH; c += 0x40) {
      const int hc = hop_select(d_hop_addr, d_afh, c);
      if ((aliased_for_candidates ? (int)aliased_channel((char)hc) : hc) == first_channel) d_clock_candidates.push_back(c);
    }
  }
  d_aliased = aliased;
  d_num_candidates = (int)d_clock_candidates.size();
  d_winnowed = 0;
  d_hop_reversal_inited = true;
  d_have_clk27 = false;
  std::printf("%d initial CLK1-27 candidates\n", d_num_candidates);
  return d_num_candidates;
}

// keep the candidates whose hop `offset` slots after the first packet is the channel observed then
// (basic_rate_piconet_impl::winnow, piconet_impl.cc:303-343)
int Piconet::winnow(int offset, char channel)
{
  auto misses = [&](uint32_t cand) {
    const char h = hop((int)((cand + (uint32_t)offset) % (uint32_t)SEQUENCE_LENGTH));
    return (d_aliased ? aliased_channel(h) : h) != channel;
  };
  d_clock_candidates.resize((size_t)d_num_candidates);
  d_clock_candidates.erase(std::remove_if(d_clock_candidates.begin(), d_clock_candidates.end(), misses), d_clock_candidates.end());
  d_num_candidates = (int)d_clock_candidates.size();
  if (d_num_candidates == 1) {
    d_clk_offset = (d_clock_candidates[0] - d_first_pkt_time) & 0x7ffffff;
    d_have_clk27 = true;
    std::printf("\nAcquired CLK1-27 offset = 0x%07x\n", d_clk_offset);
  } else if (d_num_candidates == 0) {
    reset();
  } else {
    std::printf("%d CLK1-27 candidates remaining\n", d_num_candidates);
  }
  return d_num_candidates;
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
// BLE link-layer printout (le_packet_impl constructor + print, packet_impl.cc:1529-1646), driven by a description of
// each advertising PDU: which 6-byte device addresses lead the payload and what follows them.
namespace {
struct LePdu {
  const char *addr[2];      // names of the leading device addresses (nullptr: none)
  const char *data;         // name of the byte string after the first address (AdvData / ScanRspData), or nullptr
  bool connect;             // CONNECT_REQ: link parameters follow the two addresses
};
const LePdu kLePdu[8] = {
    /* 0 ADV_IND         */ {{"AdvA", nullptr}, "AdvData", false},
    /* 1 ADV_DIRECT_IND  */ {{"AdvA", "InitA"}, nullptr, false},
    /* 2 ADV_NONCONN_IND */ {{"AdvA", nullptr}, "AdvData", false},
    /* 3 SCAN_REQ        */ {{"ScanA", "AdvA"}, nullptr, false},
    /* 4 SCAN_RSP        */ {{"AdvA", nullptr}, "ScanRspData", false},
    /* 5 CONNECT_REQ     */ {{"InitA", "AdvA"}, nullptr, true},
    /* 6 ADV_SCAN_IND    */ {{"AdvA", nullptr}, "AdvData", false},
    /* 7                 */ {{nullptr, nullptr}, nullptr, false},
};

uint64_t le_field(const uint8_t *p, int bytes)          // little-endian multi-byte field
{
  uint64_t v = 0;
  for (int i = 0; i < bytes; i++) v |= (uint64_t)p[i] << (8 * i);
  return v;
}
}  // namespace

void le_print(const char *stream, int available, double freq)
{
  const int index = le_freq_to_index(freq);
  const int n_sym = 8 * (1 + 4 + 39 + 3);           // LE_MAX_SYMBOLS, packet.h:283-285
  // de-whiten everything after the access address, then read the fields LSB first
  char link[8 * 47];
  const uint8_t *w = whitening_sequence();
  int wi = le_whitening_index(index < 0 ? 0 : index);
  for (int i = 0; i < n_sym; i++) {
    link[i] = (i < available) ? (stream[i] & 1) : 0;
    if (i >= 40) { link[i] ^= w[wi]; wi = (wi + 1) % 127; }
  }
  const uint32_t aa = air_to_host(&link[8], 32);
  const unsigned header = air_to_host(&link[40], 16);
  uint8_t pdu[48];
  for (int b = 0, i = 56; i + 8 < n_sym; b++, i += 8) pdu[b] = (uint8_t)air_to_host(&link[i], 8);
  if (index < 37) {
    std::printf("BTLE index=%02d, AA=%08x, LLID=%d, NESN=%d, SN=%d, MD=%d, Length=%d\n", index, aa, header & 3,
                (header >> 2) & 1, (header >> 3) & 1, (header >> 4) & 1, (header >> 8) & 0x1f);
    return;
  }
  const unsigned type = header & 0xf, len = (header >> 8) & 0x3f;
  std::printf("BTLE index=%02d, AA=%08x, PDUType=%d, TxAdd=%d, RxAdd=%d, Length=%d\n", index, aa, type, (header >> 6) & 1,
              (header >> 7) & 1, len);
  if (type > 6) return;
  const LePdu &d = kLePdu[type];
  for (int a = 0; a < 2 && d.addr[a]; a++) {
    const uint8_t *p = pdu + 6 * a;
    std::printf("  %s=%02x%02x%02x%02x%02x%02x\n", d.addr[a], p[0], p[1], p[2], p[3], p[4], p[5]);
  }
  if (d.data) {
    std::printf("\n  (char) %s=", d.data);
    for (unsigned i = 6; i < len; i++) std::printf(" %c", (pdu[i] < ' ' || pdu[i] > '~') ? '.' : (char)pdu[i]);
    std::printf("\n  (byte) %s=", d.data);
    for (unsigned i = 6; i < len; i++) std::printf("%02x", pdu[i]);
    std::printf("\n");
  }
  if (d.connect) {
    const uint8_t *q = pdu + 12;      // AA(4) CRCInit(3) WinSize(1) WinOffset(2) Interval(2) Latency(2) Timeout(2) ChM(5) Hop/SCA(1)
    std::printf("  AA=%08x, CRCInit=%06x, WinSize=%d, WinOffset=%d\n", (unsigned)le_field(q, 4), (unsigned)le_field(q + 4, 3), q[7],
                (int)le_field(q + 8, 2));
    std::printf("  Interval=%d, Latency=%d, Timeout=%d, ChM=%010lx, Hop=%d, SCA=%d\n", (int)le_field(q + 10, 2),
                (int)le_field(q + 12, 2), (int)le_field(q + 14, 2), (unsigned long)le_field(q + 16, 5), q[21] & 0x1f, (q[21] >> 5) & 7);
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
    // the reference tests the OBSERVED channel against the band and then demodulates the hop channel itself
    // (multi_hopper_impl.cc:158-168); with the aliasing receiver the latter can lie outside the band, where the
    // reference dereferences a filter that does not exist -- such slots are skipped here
    if (p.obs_channel >= d_ch_lo && p.obs_channel <= d_ch_hi && hopch >= d_ch_lo && hopch <= d_ch_hi) { p.first_channel = hopch; p.n_channels = 1; }
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
