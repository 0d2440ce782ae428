// rx_math.cuh -- arithmetic of the receive path, shared by every kernel.
//
// Float contract (DESIGN.md "Exactness"): binary32, ONE rounding per operation,
// sums in ascending index order -- the same rule the oracle is built with
// (-ffp-contract=off).  This translation unit is compiled with --fmad=false so
// nvcc never contracts a*b+c; where an FMA is wanted it is written as fmaf().
//
// Functions are __host__ __device__ so the host-emulation harness in tests/
// (tests/emul) can run the very same bodies on the CPU without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BTB_HD __host__ __device__ __forceinline__
#else
#define BTB_HD inline
#endif

namespace btb200 {

struct alignas(8) c32 { float re, im; };

// x * t accumulated the way std::complex<float> does it with contraction off:
// (a*c - b*d, a*d + b*c), then acc += product.     [gr_arith.h: gra_dot_cc]
BTB_HD void cmac(float &ar, float &ai, float a, float b, float c, float d)
{
  const float pr = a * c - b * d;
  const float pi = a * d + b * c;
  ar = ar + pr;
  ai = ai + pi;
}

// rotator output z = acc * phase                     [GNU Radio rotator::rotate]
BTB_HD c32 crot(c32 v, c32 p)
{
  c32 z;
  z.re = v.re * p.re - v.im * p.im;
  z.im = v.re * p.im + v.im * p.re;
  return z;
}

BTB_HD float mag2(c32 z) { return z.re * z.re + z.im * z.im; }   // complex_to_mag_squared

// gr::fast_atan2f: 256-entry table, linear interpolation
BTB_HD float fast_atan2f(const float *__restrict__ T, float y, float x)
{
  const float ya = fabsf(y), xa = fabsf(x);
  if (!((ya > 0.0f) || (xa > 0.0f))) return 0.0f;
  const float z = (ya < xa) ? ya / xa : xa / ya;
  float base;
  if (z < 0.003921569f) {
    base = z;
  } else {
    float alpha = z * 255.0f;
    const int idx = ((int)alpha) & 0xff;
    alpha = alpha - (float)idx;
    base = T[idx];
    base = base + (T[idx + 1] - T[idx]) * alpha;
  }
  float angle;
  if (xa > ya) {
    if (x >= 0.0f) {
      angle = (y >= 0.0f) ? base : -base;
    } else {
      angle = (float)3.14159265358979323846;
      angle = (y >= 0.0f) ? angle - base : base - angle;
    }
  } else {
    if (y >= 0.0f) {
      angle = (float)1.57079632679489661923;
      angle = (x >= 0.0f) ? angle - base : angle + base;
    } else {
      angle = (float)-1.57079632679489661923;
      angle = (x >= 0.0f) ? angle + base : angle - base;
    }
  }
  return angle;
}

// multi_block::demod for one output: gain * atan2(imag, real) of cur * conj(prev)
// (lib/multi_block.cc:158-168)
BTB_HD float demod_point(const float *__restrict__ T, float gain, c32 cur, c32 prev)
{
  const float a = cur.re, b = cur.im, c = prev.re, d = -prev.im;
  const float pr = a * c - b * d;
  const float pi = a * d + b * c;
  return gain * fast_atan2f(T, pi, pr);
}

struct MmConst { float gain_mu, gain_omega, omega_mid, omega_lim; };
struct MmState { float mu, omega, last; };

// mmse_fir_interpolator_ff::interpolate: sum_k in[k] * taps[imu][7-k]
BTB_HD float mmse_interp(const float *__restrict__ tab, const float *__restrict__ in, float mu)
{
#if defined(__CUDA_ARCH__)
  int imu = __float2int_rn(mu * 128.0f);
#else
  int imu = (int)rint(mu * 128.0f);
#endif
  imu = imu < 0 ? 0 : (imu > 128 ? 128 : imu);     // GNU Radio throws outside 0..128; mu is in [0,1) here
  const float *t = tab + imu * 8;
  float acc = 0.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 8; k++) acc = acc + in[k] * t[7 - k];
  return acc;
}

// one Mueller & Mueller iteration after the interpolation (lib/multi_block.cc:139-152);
// returns the integer input advance
BTB_HD int mm_update(const MmConst &K, MmState &s, float out)
{
  const float sl = (s.last < 0) ? -1.0f : 1.0f;
  const float so = (out < 0) ? -1.0f : 1.0f;
  const float mm_val = sl * out - so * s.last;
  s.last = out;
  float om = s.omega + K.gain_omega * mm_val;
  // branchless_clip(om - mid, lim): 0.5 * (|x+c| - |x-c|)
  const float x = om - K.omega_mid;
  float x1 = fabsf(x + K.omega_lim);
  const float x2 = fabsf(x - K.omega_lim);
  x1 = x1 - x2;
  s.omega = K.omega_mid + 0.5f * x1;
  float mu = s.mu + (s.omega + K.gain_mu * mm_val);
#if defined(__CUDA_ARCH__)
  // floor and float->int without the conversion unit: for 0 <= mu < 2^23 adding 2^23 with round-toward-zero drops the
  // fraction exactly, the integer sits in the mantissa and subtracting 2^23 back is exact (same value as floorf)
  if (mu >= 0.0f && mu < 8388608.0f) {
    const float t = __fadd_rz(mu, 8388608.0f);
    s.mu = mu - (t - 8388608.0f);
    return __float_as_int(t) - 0x4B000000;
  }
#endif
  const float fl = floorf(mu);
  s.mu = mu - fl;
  return (int)fl;
}

// ---- access-code predicates (lib/packet_impl.cc:247-268, 471-510) ---------
BTB_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
BTB_HD int popc64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}
BTB_HD int imin2(int a, int b) { return a < b ? a : b; }

// w_lo: symbols lag+0..lag+63 (bit i = symbol lag+i), w_hi: symbols lag+64..lag+71 in bits 0..7.
// Returns 1 when sniff_ac would stop at this lag; *lap_out = LAP read from the stream.
BTB_HD int br_lag_test(const uint64_t *__restrict__ lut, uint64_t w_lo, uint32_t w_hi, uint32_t *lap_out)
{
  const uint32_t pre = (uint32_t)w_lo & 0x1f;                                       // symbols 0..4
  const uint32_t bark = (uint32_t)((w_lo >> 61) | ((uint64_t)w_hi << 3)) & 0x7f;     // symbols 61..67
  const int d = imin2(popc32(pre ^ 0x0A), popc32(pre ^ 0x15)) +
                imin2(popc32(bark ^ 39), popc32(bark ^ 88));
  if (d > 2) return 0;
  const uint32_t lap = (uint32_t)(w_lo >> 38) & 0xffffff;                            // symbols 38..61
  const uint64_t sync = lut[768] ^ lut[lap & 0xff] ^ lut[256 + ((lap >> 8) & 0xff)] ^ lut[512 + (lap >> 16)];
  // received symbols 4..67
  const uint64_t rx_sync = (w_lo >> 4) | ((uint64_t)(w_hi & 0xf) << 60);
  // expected preamble 0101 (symbols 0..3 = 0,1,0,1) when sync bit0 == 0, else 1010
  const uint32_t exp_pre = (sync & 1) ? 0x5u : 0xAu;      // bit i = symbol i: 1,0,1,0 -> 0b0101
  const int errs = popc64(rx_sync ^ sync) + popc32(((uint32_t)w_lo & 0xf) ^ exp_pre);
  *lap_out = lap;
  return errs < 7;
}

// check_ac's count for a lag that passed (lib/packet_impl.cc:471-510): symbols among the first 68 that differ from the
// access code regenerated for the LAP read from the stream; 0..6
BTB_HD int br_lag_errors(const uint64_t *__restrict__ lut, uint64_t w_lo, uint32_t w_hi)
{
  const uint32_t lap = (uint32_t)(w_lo >> 38) & 0xffffff;
  const uint64_t sync = lut[768] ^ lut[lap & 0xff] ^ lut[256 + ((lap >> 8) & 0xff)] ^ lut[512 + (lap >> 16)];
  const uint64_t rx_sync = (w_lo >> 4) | ((uint64_t)(w_hi & 0xf) << 60);
  const uint32_t exp_pre = (sync & 1) ? 0x5u : 0xAu;
  return popc64(rx_sync ^ sync) + popc32(((uint32_t)w_lo & 0xf) ^ exp_pre);
}

// libbtbb-style access-code test (BTB200_SEARCH_BR_BCH): what multi_LAP / multi_UAP get from btbb_find_ac
// (lib/multi_LAP_impl.cc:93, lib/multi_UAP_impl.cc:95).  libbtbb is an external library that is not part of the
// reference tree; this restates its published algorithm (bluetooth_packet.c, the API level with
// btbb_find_ac(stream, search_length, lap, max_ac_errors, &pkt) that the reference's call sites use) -- PARITY
// UNPINNED, see DESIGN.md:
//   * a given LAP: Hamming distance of the 64 received sync-word symbols to that LAP's sync word <= max_ac_errors;
//   * LAP_ANY: the 7 top symbols (LAP MSB + Barker sequence) are replaced by the nearer of their two valid patterns,
//     the syndrome of the (64,30) code word under the PN overlay is computed, a non-zero syndrome is looked up among
//     the error patterns of weight <= max_ac_errors on sync-word bits 0..57 and corrected; the LAP is read from the
//     corrected word, the error count is the weight of the pattern (Barker corrections are not counted).
// The window is aligned like sniff_ac's: symbols 4..67 of the lag are the sync word, so a hit's offset is the
// preamble's (libbtbb reports the sync word's position, 4 later, and also tries the first 4 positions of a stream).
struct BchDev {
  const uint64_t *par = nullptr;     // [4][256] + constant: parity bits of the information bits (plan.hpp BchTables)
  const uint64_t *syn = nullptr;     // [n] sorted syndromes of the correctable error patterns
  const uint64_t *err = nullptr;     // [n] the patterns
  int n = 0;
  int max_err = 0;
  uint32_t lap = 0xffffffffu;        // 0xffffffff: LAP_ANY
  uint64_t target = 0;               // sync word of `lap`
};

BTB_HD int br_lag_test_bch(const BchDev &B, uint64_t w_lo, uint32_t w_hi, uint32_t *lap_out, int *n_err)
{
  uint64_t sw = (w_lo >> 4) | ((uint64_t)(w_hi & 0xf) << 60);                      // symbols 4..67
  if (B.lap != 0xffffffffu) {
    const int e = popc64(sw ^ B.target);
    *lap_out = B.lap; *n_err = e;
    return e <= B.max_err;
  }
  const uint32_t top = (uint32_t)(sw >> 57);                                       // LAP MSB + Barker: 88 or 39
  const uint32_t fix = popc32(top ^ 88u) < popc32(top ^ 39u) ? 88u : 39u;          // complements: no ties
  sw = (sw & ((1ull << 57) - 1)) | ((uint64_t)fix << 57);
  const uint32_t info = (uint32_t)(sw >> 34);
  const uint64_t par = B.par[1024] ^ B.par[info & 0xff] ^ B.par[256 + ((info >> 8) & 0xff)] ^ B.par[512 + ((info >> 16) & 0xff)] ^
                       B.par[768 + (info >> 24)];
  const uint64_t syn = (sw ^ par) & ((1ull << 34) - 1);
  int ne = 0;
  if (syn) {
    int lo = 0, hi = B.n - 1, at = -1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const uint64_t v = B.syn[mid];
      if (v == syn) { at = mid; break; }
      if (v < syn) lo = mid + 1; else hi = mid - 1;
    }
    if (at < 0) return 0;
    const uint64_t e = B.err[at];
    sw ^= e;
    ne = popc64(e);
  }
  *lap_out = (uint32_t)(sw >> 34) & 0xffffff;
  *n_err = ne;
  return 1;
}

// nearest-valid-byte distances of the LE header tables (lib/packet_impl.cc:1327-1444 in closed form)
BTB_HD int le_hdr_dist(uint32_t v, int which)
{
  int best = 9;
  for (uint32_t c = 0; c < 256; c++) {
    bool ok;
    switch (which) {
      case 0: ok = ((c & 0x3f) <= 6) && (((c >> 6) == 0) || ((c >> 6) == 3)); break;   // adv header LSB
      case 1: ok = c >= 0x06 && c <= 0x24; break;                                       // adv header MSB
      case 2: ok = c < 0x20 && (c & 3) != 0; break;                                     // data header LSB
      default: ok = c <= 0x1F; break;                                                   // data header MSB
    }
    if (ok) best = imin2(best, popc32(v ^ c));
  }
  return best;
}

// le_packet::sniff_aa predicate for one lag.  hdr_lut: [4][256] bytes (adv lsb, adv msb, data lsb, data msb).
// w_lo: symbols lag..lag+63; whitening16: bit i = whitening bit for header symbol 40+i.
BTB_HD int le_lag_test(const uint8_t *__restrict__ hdr_lut, uint64_t w_lo, uint32_t whitening16, int adv)
{
  const uint32_t pre = (uint32_t)w_lo & 0x1ff;                       // 9 preamble symbols
  int dist = imin2(popc32(pre ^ 0x0AA), popc32(pre ^ 0x155));
  const uint32_t hdr = ((uint32_t)(w_lo >> 40) & 0xffff) ^ whitening16;
  const uint32_t hl = hdr & 0xff, hm = hdr >> 8;
  int maxd = 0;
  if (adv) {
    dist += hdr_lut[hl] + hdr_lut[256 + hm];
    const uint32_t aa = (uint32_t)(w_lo >> 8);
    dist += popc32(aa ^ 0x8E89BED6u);
    maxd = 2;
  } else {
    dist += hdr_lut[512 + hl] + hdr_lut[768 + hm];
  }
  return dist <= maxd;
}

}  // namespace btb200
