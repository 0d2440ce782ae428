// hop_select.hpp -- Bluetooth basic hop selection kernel evaluated per clock value (Core spec vol 2 part B 2.6.2-2.6.3).
//
// The reference fills a 2^27-entry table with five nested loops before it can look a hop up
// (lib/piconet_impl.cc:131-159, 165-199, 214-255: 128 MiB, about a second); entry `clock` of that table is a pure
// function of the clock's bit fields, computed here directly -- shared by the host piconet logic (host/lib/bt_host.cc)
// and the GPU candidate search (csrc/rx_hop.cu):
//   clock = [h:2 | i:5 | j:5 | k:9 | x:5 | t:1]   (t = CLK1: 0 master-to-slave, 1 slave-to-master slots)
//   A = addr[27:23] ^ i,  B = addr[22:19],  C = addr{8,6,4,2,0} ^ j,  D = addr[18:10] ^ k,  E = addr{13,11,9,7,5,3,1}
//   F = 16 * clock[26:6]  (the running offset the table generator adds every 64 entries)
//   Z = ((x + A) mod 32) ^ B  ->  PERM5(Z; P13-9 = C (t = 0) or ~C (t = 1), P8-0 = D)  ->  (perm + E + F + 32 t) mod 79
//   -> channel = (2 * that) mod 79  (even channels first, then the odd ones)
// addr = (UAP << 24 | LAP) & 0xfffffff.  With AFH the slave answers on the master's channel (t = 1 repeats t = 0).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BTB_HOP_HD __host__ __device__ __forceinline__
#else
#define BTB_HOP_HD inline
#endif

namespace btb200 {

// PERM5: 14 butterflies on a 5-bit word, control bit P13 first (Core spec figure 2.19); stage s swaps bits
// (kHopStageA >> 3 s) & 7 and (kHopStageB >> 3 s) & 7 when control bit 13 - s is set
constexpr uint64_t kHopStageA = 01u | (00u << 3) | (01u << 6) | (02u << 9) | (00ull << 12) | (01ull << 15) | (03ull << 18) | (00ull << 21) |
                                (01ull << 24) | (00ull << 27) | (03ull << 30) | (01ull << 33) | (02ull << 36) | (00ull << 39);
constexpr uint64_t kHopStageB = 02u | (03u << 3) | (03u << 6) | (04u << 9) | (03ull << 12) | (04ull << 15) | (04ull << 18) | (02ull << 21) |
                                (03ull << 24) | (04ull << 27) | (04ull << 30) | (02ull << 33) | (03ull << 36) | (01ull << 39);

BTB_HOP_HD uint32_t hop_perm5(uint32_t z, uint32_t ctl14)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int s = 0; s < 14; s++) {
    const uint32_t a = (uint32_t)(kHopStageA >> (3 * s)) & 7u, b = (uint32_t)(kHopStageB >> (3 * s)) & 7u;
    const uint32_t t = ((z >> a) ^ (z >> b)) & (ctl14 >> (13 - s)) & 1u;
    z ^= (t << a) | (t << b);
  }
  return z;
}

BTB_HOP_HD uint32_t hop_gather(uint32_t v, uint32_t first, uint32_t count)
{
  uint32_t out = 0;
  for (uint32_t i = 0; i < count; i++) out |= ((v >> (first + 2 * i)) & 1u) << i;     // every second bit
  return out;
}

BTB_HOP_HD int hop_select(uint32_t addr, bool afh, uint32_t clock)
{
  const uint32_t t = clock & 1u, x = (clock >> 1) & 31u, k = (clock >> 6) & 511u, j = (clock >> 15) & 31u, i = (clock >> 20) & 31u;
  const uint32_t A = ((addr >> 23) & 31u) ^ i, B = (addr >> 19) & 15u;
  const uint32_t C = hop_gather(addr, 0, 5) ^ j, D = ((addr >> 10) & 511u) ^ k, E = hop_gather(addr, 1, 7);
  const uint32_t F = (16u * (clock >> 6)) % 79u;
  const bool second = t && !afh;
  const uint32_t Z = ((x + A) & 31u) ^ B;
  const uint32_t ctl = (((second ? C ^ 31u : C) & 31u) << 9) | D;
  const uint32_t r = (hop_perm5(Z, ctl) + E + F + (second ? 32u : 0u)) % 79u;
  return (int)((2u * r) % 79u);
}

// the aliasing receiver folds the band onto 25 channels (lib/piconet_impl.cc:520-523)
BTB_HOP_HD int hop_aliased(int channel) { return ((channel + 24) % 25) + 26; }

}  // namespace btb200
