// bt_host.h -- host-side Bluetooth packet layer of the B200 blocks.
//
// What happens to a packet AFTER the GPU path has found it: header / payload decoding of
// basic-rate packets (whitening, FEC 1/3 and 2/3, HEC, CRC), UAP and CLK1-6 discovery per
// piconet, FHS parsing, and the BLE header printout.  Behavioural restatement of the
// reference's lib/packet_impl.cc:367-1275,1529-1665 and lib/piconet_impl.cc:33-60,370-547
// (each function cites its lines), written so that the blocks print exactly what the
// reference's ac()/aa() call chains print.  Runs on a handful of packets per second: plain C++.
#pragma once
#include <stdint.h>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace btb200_host {

// one period of the x^7+x^4+1 whitening sequence (packet::WHITENING_DATA, packet_impl.cc:84-90)
const uint8_t *whitening_sequence();
// position in that sequence for CLK1-6 (classic_packet::INDICES, packet_impl.cc:185-189)
int classic_whitening_index(int clk6);
// position for a BLE channel index (le_packet::INDICES, packet_impl.cc:1446-1450)
int le_whitening_index(int channel_index);
// le_packet::freq2index, packet_impl.cc:1285-1314
int le_freq_to_index(double freq);

inline uint32_t air_to_host(const char *air, int bits)          // packet_impl.cc:104-136
{
  uint32_t v = 0;
  for (int i = 0; i < bits; i++) v |= (uint32_t)(air[i] & 1) << i;
  return v;
}

class ClassicPacket {
 public:
  static const int MAX_SYMBOLS = 3125;                          // include/gr_bluetooth/packet.h:59
  ClassicPacket(const char *stream, int length, uint32_t clkn, double freq);   // packet_impl.cc:40-58,210-245
  uint32_t clkn;
  int channel;
  uint32_t lap() const { return d_lap; }
  bool header_present() const;                                  // :1205-1242
  uint8_t try_clock(int clock);                                 // :1045-1063
  int crc_check(int clock);                                     // :609-668
  void set_clock(uint32_t clock, bool have27);                  // :578-590
  void set_uap(uint8_t uap) { d_uap = uap; }
  uint8_t uap() const { return d_uap; }
  void set_nap(uint16_t nap) { d_nap = nap; d_have_nap = true; }   // :570-574
  int payload_length() const { return d_payload_length; }
  // 9 + payload_length bytes for the Wireshark interface: CLK (4, little endian), channel, flags, the packet
  // header squeezed into 3 bytes, then the payload bytes (packet_impl.cc:1175-1202)
  std::vector<uint8_t> tun_format() const;
  void decode();                                                // :173-179 (decode_header + decode_payload)
  bool got_payload() const { return d_have_payload; }
  int type() const { return d_type; }
  void print() const;                                           // :1161-1173
  // FHS fields (:1243-1275)
  uint32_t lap_from_fhs() const { return air_to_host(&d_payload[34], 24); }
  uint8_t uap_from_fhs() const { return (uint8_t)air_to_host(&d_payload[64], 8); }
  uint16_t nap_from_fhs() const { return (uint16_t)(uint8_t)air_to_host(&d_payload[72], 16); }   // air_to_host8(...,16) in the reference
  uint32_t clock_from_fhs() const { return air_to_host(&d_payload[115], 26); }

  static bool unfec13(const char *in, char *out, int length);   // :367-384
  static bool unfec23(const char *in, int length, std::vector<char> &out);   // :387-468
  static uint16_t crcgen(const char *payload, int length, int uap);          // :529-548
  static int uap_from_hec(uint16_t data, uint8_t hec);          // :593-606

 private:
  std::vector<char> d_sym;      // MAX_SYMBOLS entries, zero beyond d_length
  int d_length;
  uint32_t d_lap;
  uint8_t d_uap = 0;
  uint32_t d_clock = 0;
  bool d_have_clk6 = false, d_have_payload = false;
  bool d_have_clk27 = false, d_have_nap = false;
  uint16_t d_nap = 0;
  int d_type = 0;
  int d_payload_length = 0, d_payload_header_length = 0, d_llid = 0, d_flow = 0;
  char d_packet_header[18];
  char d_payload_header[16];
  std::vector<char> d_payload;  // 2744 bits worth of chars in the reference (one bit per char)

  void unwhiten(const char *in, char *out, int clock, int length, int skip) const;   // :513-526
  bool payload_crc() const;                                     // :671-680
  bool decode_header();                                         // :1066-1090
  void decode_payload();                                        // :1092-1158
  bool decode_payload_header(const char *stream, int clock, int header_bytes, int size, bool fec);   // :726-770
  // payload parsers, driven by the per-type format table in bt_host.cc
  int fhs(int clock);
  int acl(int clock);           // DM1/3/5, DH1/3/5, DV data field, AUX1
  int crc_scan(int clock);      // EV3, EV5: length found by trying CRCs
  int EV4(int clock);
  int sco(int clock);           // HV1/2/3
};

// UAP / CLK1-6 discovery state of one piconet (basic_rate_piconet_impl, piconet_impl.cc:63-80,370-547)
class Piconet {
 public:
  explicit Piconet(uint32_t lap) : d_lap(lap) {}
  void enqueue(std::shared_ptr<ClassicPacket> p) { d_queue.push_back(p); }
  std::shared_ptr<ClassicPacket> dequeue();
  bool uap_from_header(ClassicPacket &pkt);
  void reset();
  bool have_uap() const { return d_have_uap; }
  bool have_nap() const { return d_have_nap; }
  bool have_clk6() const { return d_have_clk6; }
  bool have_clk27() const { return d_have_clk27; }
  uint8_t uap() const { return d_uap; }
  uint16_t nap() const { return d_nap; }
  uint32_t offset() const { return d_clk_offset; }
  void set_uap(uint8_t u) { d_uap = u; d_have_uap = true; }
  void set_nap(uint16_t n) { d_nap = n; d_have_nap = true; }
  void set_offset(uint32_t o) { d_clk_offset = o; d_have_clk6 = true; d_have_clk27 = true; }

  // CLK1-27 discovery by hop reversal (piconet_impl.cc:96-368): the complete 2^27-entry hopping
  // sequence of (UAP, LAP) is generated, candidates = clock values whose hop matches the first
  // observed channel, then winnowed with every later observation.
  int init_hop_reversal(bool aliased);                          // :96-129
  int winnow();                                                 // :345-368
  // Channel of the basic hop sequence at CLK1-27 = clock.  The reference tabulates all 2^27 entries (128 MiB,
  // piconet_impl.cc:214-255) before it can look one up (:279-282); here the hop selection kernel (Core spec vol 2
  // part B 2.6) is evaluated on demand from the clock's bit fields -- same values, no table.
  char hop(int clock) const { return (char)hop_select(d_hop_addr, d_afh, (uint32_t)clock); }
  static int hop_select(uint32_t address28, bool afh, uint32_t clock);
  // candidate search of the hop reversal on another engine (the GPU kernel behind btb200_hop_candidates): fills `out`
  // with every clock = clock6 (mod 64) below 2^27 whose hop is first_channel, ascending; returns false to decline
  typedef bool (*candidate_fn)(uint32_t address28, bool afh, bool aliased, uint32_t clock6, int first_channel,
                               std::vector<uint32_t> &out);
  static candidate_fn s_candidate_fn;
  static char aliased_channel(char channel) { return (char)(((channel + 24) % 25) + 26); }   // :520-523

 private:
  static const int MAX_PATTERN_LENGTH = 1000;                   // lib/piconet_impl.h:45
  static const int SEQUENCE_LENGTH = 134217728;                 // include/gr_bluetooth/piconet.h:83
  static const int CHANNELS = 79, ALIASED_CHANNELS = 25;
  int d_pattern_indices[MAX_PATTERN_LENGTH];
  uint8_t d_pattern_channels[MAX_PATTERN_LENGTH];
  uint32_t d_hop_addr = 0;            // UAP/LAP bits the hop kernel uses
  std::vector<uint32_t> d_clock_candidates;
  int d_num_candidates = 0, d_winnowed = 0;
  bool d_hop_reversal_inited = false, d_aliased = false, d_afh = false, d_looks_like_afh = false;
  int winnow(int offset, char channel);                         // :303-343
  uint32_t d_lap;
  std::deque<std::shared_ptr<ClassicPacket>> d_queue;
  bool d_got_first_packet = false, d_have_uap = false, d_have_nap = false, d_have_clk6 = false, d_have_clk27 = false;
  int d_packets_observed = 0, d_total_packets_observed = 0;
  uint32_t d_first_pkt_time = 0, d_clk_offset = 0;
  uint8_t d_uap = 0;
  uint16_t d_nap = 0;
  int d_clock6_candidates[64];
};

// le_packet_impl: constructor + print (packet_impl.cc:1529-1646)
void le_print(const char *stream, int available, double freq);

// The per-packet call chain of the sniffer block: multi_sniffer_impl::ac / aa / id / decode /
// discover / recall / fhs, lib/multi_sniffer_impl.cc:169-365.  Independent of the GPU path so
// that it can be driven from any hit source (the CPU test tier drives it from the oracle).
// One frame on the TAP interface (lib/tun.cc:91-123): Ethernet header (destination = the six low-order bytes
// of dst_addr, big endian; source likewise; EtherType) followed by the data.  A negative fd is "no interface":
// nothing is written.  Returns data_len, or -1 when the write fails.
static const unsigned short TUN_ETHER_TYPE = 0xFFF0;              // lib/multi_sniffer_impl.h:52, multi_hopper_impl.h:62
int write_frame(int fd, const uint8_t *data, unsigned data_len, uint64_t src_addr, uint64_t dst_addr,
                unsigned short ether_type);
// open the TAP device `name` (lib/tun.cc:41-78: /dev/net/tun, IFF_TAP | IFF_NO_PI); -1 when that is not possible
int open_tap(const char *name);
// what the blocks do for tun = true: the file named by BTB200_TUN_FILE if set (frames are appended as written:
// capture/tests), else the TAP device "btbb" as in the reference; -1 and the reference's warning when neither opens
int open_tun_output();

class SnifferHost {
 public:
  // frames go to this descriptor (a TAP device, or any file/pipe for capture and tests); -1 = off
  void set_tun_fd(int fd) { d_tunfd = fd; }
  void ac(const char *symbols, int len, uint32_t clkn, double freq, double snr);
  void aa(const char *symbols, int len, uint32_t clkn, double freq, double snr);

 private:
  static const uint32_t GIAC = 0x9E8B33, LIAC = 0x9E8B00;      // lib/multi_sniffer_impl.h:41-42
  int d_tunfd = -1;
  std::map<int, std::shared_ptr<Piconet>> d_piconets;
  void id(uint32_t lap);
  void decode(std::shared_ptr<ClassicPacket> pkt, std::shared_ptr<Piconet> pn, bool first_run);
  void discover(std::shared_ptr<ClassicPacket> pkt, std::shared_ptr<Piconet> pn);
  void recall(std::shared_ptr<Piconet> pn);
  void fhs(std::shared_ptr<ClassicPacket> pkt);
};

// The per-slot logic of the hopper block: multi_hopper_impl::work / hopalong,
// lib/multi_hopper_impl.cc:76-209 -- scan every channel until CLK1-27 is known (UAP/CLK1-6 from
// packet headers, then hop reversal), afterwards follow the piconet on its predicted channel.
class HopperHost {
 public:
  HopperHost(uint32_t lap, bool aliased, int ch_lo, int ch_hi)
      : d_lap(lap), d_aliased(aliased), d_ch_lo(ch_lo), d_ch_hi(ch_hi), d_piconet(lap) {}
  struct SlotPlan {
    bool hopalong = false;
    uint32_t clock27 = 0;
    int first_channel = 0, n_channels = 0;     // classic channels to process, in ascending order
    int obs_channel = -1;                      // hopalong: channel the packet is reported on
    uint32_t stop_lap = 0xffffffffu;           // scan: a packet with this LAP and a header ends the channel loop
  };
  SlotPlan plan(uint32_t clkn) const;
  // scan phase: the first access code found on `channel`; returns true when the reference breaks out of the loop
  bool scan_packet(uint32_t clkn, int channel, const char *symbols, int len);
  // hopalong phase: the first access code found on the predicted channel
  void hop_packet(const SlotPlan &p, const char *symbols, int len);
  void set_tun_fd(int fd) { d_tunfd = fd; }

 private:
  uint32_t d_lap;
  bool d_aliased;
  int d_ch_lo, d_ch_hi;
  Piconet d_piconet;
  int d_tunfd = -1;
};

// The per-slot logic of the UAP block (multi_UAP_impl::work, lib/multi_UAP_impl.cc:67-124): packets of the target
// piconet that carry a header feed the UAP/CLK1-6 discovery until the UAP is known.
class UapHost {
 public:
  UapHost(uint32_t lap, int ch_lo, int ch_hi) : d_lap(lap), d_ch_lo(ch_lo), d_ch_hi(ch_hi), d_piconet(lap) {}
  // the first access code found on `channel`; true when the reference leaves the channel loop (target packet with a header)
  bool packet(uint32_t clkn, int channel, const char *symbols, int len);
  bool done() const { return d_done; }
  uint8_t uap() const { return d_piconet.uap(); }
  uint32_t lap() const { return d_lap; }
  int ch_lo() const { return d_ch_lo; }
  int ch_hi() const { return d_ch_hi; }

 private:
  uint32_t d_lap;
  int d_ch_lo, d_ch_hi;
  Piconet d_piconet;
  bool d_done = false;
};

}  // namespace btb200_host
