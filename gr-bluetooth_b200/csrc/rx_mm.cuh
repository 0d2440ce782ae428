// rx_mm.cuh -- the stateless Mueller & Mueller clock-recovery loop (lib/multi_block.cc:128-155 + slicer :171-178), one
// thread per channel-window, shared by the stand-alone kernel (rx_kernels.cu) and the noise estimator's launch
// (rx_nest.cu), whose first blocks resume the chains of the windows with hits on SMs of their own.
// Every float operation of the loop is written with explicit single-rounding arithmetic (mm_update, the 8-tap dot
// product in ascending order), so both translation units produce the same bits.
#pragma once
#include "rx_kernels.cuh"

namespace btb200 {

__device__ __forceinline__ void mm_cp_async4(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void mm_cp_async16(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void mm_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void mm_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void mm_cp_async_wait_prev() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// ld.shared with an explicit 32-bit address and an immediate offset
template <int IMM>
__device__ __forceinline__ float mm_lds(unsigned addr)
{
  float v;
  asm volatile("ld.volatile.shared.f32 %0, [%1+%2];\n" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}
// the 8-tap interpolation of one step, ascending order, one rounding per operation
// ld.shared.v4 with an explicit 32-bit address and an immediate offset
template <int IMM>
__device__ __forceinline__ float4 mm_lds4(unsigned addr)
{
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
  return v;
}
// The 8-tap interpolation of one step, ascending order, one rounding per operation.  RS = bytes between consecutive
// samples of a chain in the ring; ma = address of the 8 taps of the step's fraction, contiguous ([imu][8] table).
// A chain is ONE dependent instruction stream with at most a few warps per scheduler to hide behind, and the
// assembler sinks every load to its first use, where it costs its full latency -- sixteen times per step with
// scalar loads.  So: the loads are volatile (kept in program order), the eight samples go first (their address
// follows from the integer advance, before the rounded fraction is known), the taps arrive as two 16-byte loads,
// and the first multiply cannot be placed before the last load has been issued.
// byte permute with a selector the assembler cannot see through (0x3210 at run time: d = a); used to make a value wait
// for another one's arrival without changing it
__device__ __forceinline__ float mm_tie(float a, float other, unsigned sel)
{
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(__float_as_uint(a)), "r"(__float_as_uint(other)), "r"(sel));
  return __uint_as_float(d);
}
template <int RS>
__device__ __forceinline__ float mm_interp8(unsigned ra, unsigned ma, float acc, unsigned sel)
{
  const float x0 = mm_lds<0 * RS>(ra), x1 = mm_lds<1 * RS>(ra), x2 = mm_lds<2 * RS>(ra), x3 = mm_lds<3 * RS>(ra);
  const float x4 = mm_lds<4 * RS>(ra), x5 = mm_lds<5 * RS>(ra), x6 = mm_lds<6 * RS>(ra), x7 = mm_lds<7 * RS>(ra);
  float4 tb = mm_lds4<16>(ma);
  const float4 ta = mm_lds4<0>(ma);
  // the upper taps pass through a permute that also reads the lower ones: no multiply can be scheduled between the
  // two loads (where it would stall the second load behind the first one's latency); both are in flight together
  tb.x = mm_tie(tb.x, ta.w, sel); tb.y = mm_tie(tb.y, ta.w, sel); tb.z = mm_tie(tb.z, ta.w, sel); tb.w = mm_tie(tb.w, ta.w, sel);
  acc = acc + x0 * ta.x; acc = acc + x1 * ta.y; acc = acc + x2 * ta.z; acc = acc + x3 * ta.w;
  acc = acc + x4 * tb.x; acc = acc + x5 * tb.y; acc = acc + x6 * tb.z; acc = acc + x7 * tb.w;
  return acc;
}

struct MmSave { float mu, omega, last; unsigned ii; int oo; uint32_t word; };
// ring depth of the clock-recovery loop: a refill reaches ii+96 and the reader is at most 40 samples further
// when the next one is issued, so 128 rows never overwrite a live sample
constexpr int MM_RD = 128;
constexpr int MM_MAXADV = 5;         // |demod| <= gain * pi bounds the timing error term: at most 5 samples per symbol
constexpr int MM_PC = MM_RD + 12;   // channel-major source: floats per chain in the ring (8 mirrored rows + pad, 16-byte multiple)
constexpr size_t mm_smem_bytes(int blk) { return sizeof(float) * MM_PC * blk + sizeof(float) * 8 * 132; }

// mode 0: every window from the constructor state to the end (reference loop).
// mode 1 (lazy tail, first pass): every window, stop at G.sym_target symbols, demod floats exist for i < G.ne_dem
//         (enough for sym_target symbols at the loop's maximum advance); the loop state is saved.
// mode 2 (lazy tail, resume): LISTED windows continue from the saved state to the end.
// The whole block calls this (the interpolator table is staged by all threads); threads >= BLK leave after that.
// block = index of this group of BLK windows; mm_smem = mm_smem_bytes(BLK) bytes of shared memory.
//
// CM (mode 2 in the throughput mode): the demod floats come from the CHANNEL-MAJOR copy demC[c * pitchC + grid row]
// the channelizer writes next to the row-major one.  The listed windows are scattered over slots and channels, so
// from the row-major grid every sample of every chain is its own 32-byte sector and its own 4-byte copy (one
// load/store-unit pass per lane: 830 MB of sectors per 2 300 windows, and the unit the chains' own shared-memory
// loads queue behind); from the channel-major copy a lane fetches 16 bytes = 4 consecutive samples of its window
// per copy and uses every byte of every sector.  A copy lands as 16 contiguous bytes, so the ring is chain-major,
// [chain][MM_PC] (rows are bank-conflicted 4-way between lanes, which two warps per SM do not notice); rows are
// counted from the 16-byte boundary at or below the window's first grid row (`sh` rows earlier), the same for the
// whole warp, so the copies are aligned for every lane.
template <int BLK, bool CM = false>
__device__ __forceinline__ void mm_stateless_block(const Geom &G, const DevBatch &W, const float *__restrict__ mmse_g,
                                                   const float *__restrict__ demT, int mode, MmSave *__restrict__ save,
                                                   const int4 *__restrict__ list, int n_list, unsigned char *mm_smem, int block,
                                                   const float *__restrict__ demC = nullptr, long pitchC = 0)
{
  static_assert(BLK % 32 == 0, "whole warps");
  constexpr int RD = MM_RD;        // ring depth (demod samples per window)
  constexpr int PERIOD = 8;        // steps per block
  constexpr int NEED = MM_MAXADV * PERIOD + 8;   // a block reads samples below ii + NEED
  constexpr unsigned FULL = 0xffffffffu;
  // Every chain has its own column of the ring, but the ROWS move in step for the 32 chains of a warp: row i holds
  // sample i of each chain's window, the warp copies rows [pf, want) with want = (smallest ii of the warp) + RD, one
  // cp.async per row -- for windows of consecutive channels that is one or two cache lines per instruction and ONE
  // shared-memory wavefront (per-chain row schedules cost a wavefront per distinct row: 45 % of the shared-memory
  // pipe, the unit that bounds this loop).  The loop advances ii by at most MM_MAXADV samples per step (|demod| <=
  // gain pi bounds the timing error), so a block of PERIOD steps reads below ii + NEED; a chain that is further
  // ahead of the slowest one than the landed rows allow sits the block out (chains drift apart by a few tens of
  // samples over a window).  Two copy groups may be in flight.  Rows 0..7 are mirrored at RD..RD+7 so the 8 samples
  // of an interpolation are always 8 consecutive rows: one base address, immediate offsets.
  float (*ring)[BLK] = reinterpret_cast<float (*)[BLK]>(mm_smem);                        // [RD + 8][BLK]
  float *ringc = reinterpret_cast<float *>(mm_smem);                                     // CM: [BLK][MM_PC]
  constexpr int PC = MM_PC;
  // interpolator taps, [imu][8]: entry k of row imu = taps[imu][7 - k] multiplies sample k of the step (two 16-byte loads)
  float *s_mmse = reinterpret_cast<float *>(mm_smem + sizeof(float) * (CM ? PC : RD + 8) * BLK);     // [130][8]
  for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) s_mmse[(i & ~7) + 7 - (i & 7)] = mmse_g[i];
  if (threadIdx.x == 0) {
    s_mmse[129 * 8] = __int_as_float((int)(0x4B400000u << 5));                                         // spare row: see mmse_biased
    s_mmse[129 * 8 + 1] = __int_as_float(0x3210);                                                      // identity selector of mm_tie
  }
  __syncthreads();
  if (threadIdx.x >= BLK) return;                          // whole warps leave; the others stay complete to the end
  int idx = block * BLK + threadIdx.x;
  bool live;
  if (mode == 2) {
    if (n_list < 0) n_list = *W.tail.n_list;             // device-driven tail: the list was built on the device
    live = idx < n_list;
    if (live) { const int4 it = list[idx]; idx = it.x * G.nch + it.y; }
  } else {
    live = idx < W.B * G.nch;
    if (live && !W.pass[idx]) { W.nsym[idx] = 0; live = false; }
  }
  if (!live) idx = 0;
  const int b = idx / G.nch, c = idx - b * G.nch;
  // CM: rows count from the 16-byte boundary at or below the window's first grid row; gp is row 0 of that numbering
  const int sh = CM ? (int)(((long)b * G.dem_rows) & 3) : 0;
  const float *gp = CM ? demC + (long)c * pitchC + ((long)b * G.dem_rows - sh)
                       : demT + ((long)b * G.dem_rows) * G.nch + c;     // sample 0 of the window
  uint32_t *__restrict__ bits_row = W.bits + (long)idx * G.bw;
  float *soft_row = W.soft ? W.soft + (long)idx * G.n_dem_pad : nullptr;
  MmState st{G.mu0, G.mm.omega_mid, 0.0f};
  unsigned ii = 0;
  int oo = 0;
  uint32_t word = 0;
  if (mode == 2 && live) {
    const MmSave sv = save[idx];
    st = MmState{sv.mu, sv.omega, sv.last};
    ii = sv.ii; oo = sv.oo; word = sv.word;
  }
  ii += (unsigned)sh;                                            // CM: ii, ni and avail live in the shifted row numbering
  // demod floats that exist (CM: the warp copies whole 16-byte groups, up to the group that holds the last row of any lane)
  const int avail = CM ? ((G.n_dem + 3 + 3) & ~3) : (mode == 1) ? G.ne_dem : G.n_dem;
  const int oo_end = (mode == 1) ? G.sym_target : G.n_dem;
  const unsigned ni = (unsigned)((mode == 1 ? G.ne_dem : G.n_dem) - 8 + sh);
  const int tid = threadIdx.x;
  const int nch = G.nch;
  const unsigned ring_tid = CM ? (unsigned)__cvta_generic_to_shared(&ringc[tid * PC]) : (unsigned)__cvta_generic_to_shared(&ring[0][tid]);
  // table address biased by the exponent bits of the magic constant: row imu is at mmse_biased + 32 bits(1.5 2^23 + imu)
  // mod 2^32 (the bias is read back from shared memory so that ptxas cannot split it off again as an add per load)
  const unsigned mmse_biased = (unsigned)__cvta_generic_to_shared(&s_mmse[0]) - (unsigned)__float_as_int(s_mmse[129 * 8]);
  const unsigned tie_sel = (unsigned)__float_as_int(*(volatile float *)&s_mmse[129 * 8 + 1]);
  const MmConst K = G.mm;
  bool done = !live || !(oo < oo_end && ii < ni);
  // rows [pf - RD, pf) are in (or on their way to) the ring; pf is the same for the 32 chains of the warp
  int pf = (int)__reduce_min_sync(FULL, done ? 0x7fffffffu : ii);
  if (pf != 0x7fffffff) {
    if constexpr (CM) pf &= ~3; else gp += (long)pf * nch;
    auto refill = [&](int want) {
      if constexpr (CM) {
        for (; pf < want; pf += 4) {                          // pf and want are multiples of 4
          const int rr = pf & (RD - 1);
          if (live) {
            mm_cp_async16(&ringc[tid * PC + rr], gp + pf);
            if (rr < 8) mm_cp_async16(&ringc[tid * PC + RD + rr], gp + pf);
          }
        }
      } else {
        for (; pf < want; pf++, gp += nch) {
          const int rr = pf & (RD - 1);
          if (live) {
            mm_cp_async4(&ring[rr][tid], gp);
            if (rr < 8) mm_cp_async4(&ring[rr + RD][tid], gp);
          }
        }
      }
      mm_cp_async_commit();
    };
    {
      int want = pf + RD; if (want > avail) want = avail;
      refill(want);
      mm_cp_async_wait_all();
      if (!CM && G.dem_grid && live && ii == 0) ring[0][tid] = 0.0f;   // demod_out[0] of a window is never written by the reference
    }
    // reach_prev = first row NOT covered by the copies issued before the most recent group (landed after wait_group 1)
    int reach_prev = pf, reach_last = pf;
    while (true) {
      done = !live || !(oo < oo_end && ii < ni);
      if (__all_sync(FULL, done)) break;
      const int ii_min = (int)__reduce_min_sync(FULL, done ? 0x7fffffffu : ii);
      if (ii_min + NEED > reach_prev) { mm_cp_async_wait_all(); reach_prev = reach_last; }
      else mm_cp_async_wait_prev();
      const int landed = reach_prev;
      {
        int want = (CM ? (ii_min & ~3) : ii_min) + RD; if (want > avail) want = avail;      // rows from ii_min on stay
        refill(want);
        reach_prev = reach_last;
        reach_last = pf;
      }
      if (done || ((int)ii + NEED > landed && landed < avail)) continue;     // finished, or too far ahead: sit this block out
      if ((oo & 7) == 0 && oo + PERIOD <= oo_end && ii + (unsigned)(MM_MAXADV * PERIOD) < ni) {
        // Fast block: PERIOD steps of straight-line code.  No exit tests inside (the loop cannot end within them: the
        // advance per step is at most MM_MAXADV samples), the sliced bits collected in a byte, shared memory addressed
        // explicitly.  Same arithmetic as mm_update / mmse_interp, operation for operation and rounding for rounding:
        //   sl * out = +-out;  clip = 0.5 (|x + c| - |x - c|);
        //   floor(m) and (int) floor(m) from tz = RZ(m + 2^23): fl = tz - 2^23, advance = bits(tz) - bits(2^23);
        //   rint(128 (m - fl)) from ONE fused multiply-add: m - fl is exact and 1.5 2^23 - 128 fl = fma(tz, -128,
        //   2^30 + 1.5 2^23) is an exact integer below 2^24, so fma(m, 128, that) rounds 128 (m - fl) + 1.5 2^23
        //   once, like fadd(128 mu, 1.5 2^23) = the magic rint of the slow path.
        // A step whose m is outside [0, 2^23) (never, for finite input) flags the block, which is then redone by the
        // step-by-step loop below from the saved state.
        const MmState st0 = st;
        const unsigned ii0 = ii;
        // byte address of a chain's row = ring_tid + (row << SH): rows are BLK floats apart, or 1 float (CM)
        constexpr int SH = CM ? 2 : (BLK == 32 ? 7 : BLK == 64 ? 8 : BLK == 128 ? 9 : BLK == 256 ? 10 : -1);
        static_assert(SH > 0, "BLK must be 32, 64, 128 or 256");
        constexpr int RS = CM ? 4 : BLK * 4;
        constexpr unsigned RMASK = (unsigned)(RD - 1) << SH;
        unsigned iiw = ii << SH;
        unsigned tb = umin((unsigned)__float_as_int(__fmaf_rn(st.mu, 128.0f, 12582912.0f)), 0x4B400080u);
        unsigned byte = 0, bad = 0;
        float mu = st.mu, omega = st.omega, last = st.last;
  #pragma unroll
        for (int t = 0; t < PERIOD; t++) {
          const unsigned ma = mmse_biased + (tb << 5);
          const unsigned ra = ring_tid + (iiw & RMASK);
          const float out = mm_interp8<RS>(ra, ma, 0.0f, tie_sel);
          if (soft_row) soft_row[oo + t] = out;
          const bool neg = out < 0;
          if (!neg) byte |= 1u << t;
          const float a = (last < 0) ? -out : out;
          const float b = neg ? -last : last;
          const float mm_val = a - b;
          last = out;
          const float x = (omega + K.gain_omega * mm_val) - K.omega_mid;
          omega = K.omega_mid + 0.5f * (fabsf(x + K.omega_lim) - fabsf(x - K.omega_lim));
          const float m = mu + (omega + K.gain_mu * mm_val);
          const float tz = __fadd_rz(m, 8388608.0f);
          const unsigned tzb = (unsigned)__float_as_int(tz);
          tb = umin((unsigned)__float_as_int(__fmaf_rn(m, 128.0f, __fmaf_rn(tz, -128.0f, 1086324736.0f))), 0x4B400080u);
          iiw += SH >= 8 ? (tzb << SH) : ((tzb - 0x4B000000u) << SH);        // bits(2^23) << 8 is 0 mod 2^32
          bad |= tzb - 0x4B000000u;
          mu = m - (tz - 8388608.0f);
        }
        if (bad < 0x800000u) {
          st = MmState{mu, omega, last};
          ii = ii0 + ((iiw - (ii0 << SH)) >> SH);
          word |= byte << (oo & 31);
          oo += PERIOD;
          if ((oo & 31) == 0) { bits_row[(oo >> 5) - 1] = word; word = 0; }
          continue;
        }
        st = st0;
        ii = ii0;
      }
#pragma unroll 1
      for (int t = 0; t < PERIOD && oo < oo_end && ii < ni; t++) {
        // rint(mu * 128) without the conversion unit: 0 <= mu < 1, so adding 1.5 * 2^23 rounds to nearest-even at the
        // units place and leaves the integer in the mantissa (= __float2int_rn)
        int imu = __float_as_int(__fadd_rn(st.mu * 128.0f, 12582912.0f)) - 0x4B400000;
        imu = imu < 0 ? 0 : (imu > 128 ? 128 : imu);
        const float *rp = CM ? &ringc[tid * PC + (ii & (RD - 1))] : &ring[ii & (RD - 1)][tid];
        const float *mp = &s_mmse[imu * 8];
        float out = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) out = out + rp[k * (CM ? 1 : BLK)] * mp[k];
        if (soft_row) soft_row[oo] = out;
        if (!(out < 0)) word |= 1u << (oo & 31);
        if ((oo & 31) == 31) { bits_row[oo >> 5] = word; word = 0; }
        ii += (unsigned)mm_update(K, st, out);
        oo++;
      }
    }
    mm_cp_async_wait_all();
  }
  if (!live) return;
  if (mode == 1) save[idx] = MmSave{st.mu, st.omega, st.last, ii, oo, word};
  if (oo & 31) bits_row[oo >> 5] = word;
  for (int w = (oo + 31) >> 5; w < G.bw; w++) bits_row[w] = 0;
  W.nsym[idx] = oo;
}

}  // namespace btb200
