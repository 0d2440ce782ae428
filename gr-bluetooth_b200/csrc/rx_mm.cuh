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
__device__ __forceinline__ void mm_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void mm_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

struct MmSave { float mu, omega, last; unsigned ii; int oo; uint32_t word; };
// ring depth of the clock-recovery loop: a refill reaches ii+96 and the reader is at most 40 samples further
// when the next one is issued, so 128 rows never overwrite a live sample
constexpr int MM_RD = 128;
constexpr size_t mm_smem_bytes(int blk) { return sizeof(float) * (MM_RD + 8) * blk + sizeof(float) * 8 * 132; }

// mode 0: every window from the constructor state to the end (reference loop).
// mode 1 (lazy tail, first pass): every window, stop at G.sym_target symbols, demod floats exist for i < G.ne_dem
//         (enough for sym_target symbols at the loop's maximum advance); the loop state is saved.
// mode 2 (lazy tail, resume): LISTED windows continue from the saved state to the end.
// The whole block calls this (the interpolator table is staged by all threads); threads >= BLK leave after that.
// block = index of this group of BLK windows; mm_smem = mm_smem_bytes(BLK) bytes of shared memory.
template <int BLK>
__device__ __forceinline__ void mm_stateless_block(const Geom &G, const DevBatch &W, const float *__restrict__ mmse_g,
                                                   const float *__restrict__ demT, int mode, MmSave *__restrict__ save,
                                                   const int4 *__restrict__ list, int n_list, unsigned char *mm_smem, int block)
{

  constexpr int RD = MM_RD;        // ring depth (demod samples per window)
  constexpr int AHEAD = 88;        // refill target: ii + 8 + AHEAD
  constexpr int PERIOD = 8;        // steps between refills
  // The loop advances ii by at most 5 samples per step (|demod| <= gain*pi bounds mm_val), so the refill issued at
  // the start of a block of PERIOD steps (it reaches ii + 96) covers everything the NEXT block can read (< ii + 88)
  // and lands while this block runs; 128 rows never overwrite a live sample.  Rows 0..7 are mirrored at 128..135 so
  // the 8 samples of an interpolation are always 8 consecutive rows: one base address, immediate offsets.
  float (*ring)[BLK] = reinterpret_cast<float (*)[BLK]>(mm_smem);                        // [RD + 8][BLK]
  float (*s_mmse)[132] = reinterpret_cast<float (*)[132]>(mm_smem + sizeof(float) * (RD + 8) * BLK);   // [8][132]
  for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) s_mmse[7 - (i & 7)][i >> 3] = mmse_g[i];   // [k][imu] = taps[imu][7-k]
  __syncthreads();
  if (threadIdx.x >= BLK) return;
  int idx = block * BLK + threadIdx.x;
  if (mode == 2) {
    if (n_list < 0) n_list = *W.tail.n_list;             // device-driven tail: the list was built on the device
    if (idx >= n_list) return;
    const int4 it = list[idx];
    idx = it.x * G.nch + it.y;
  } else {
    if (idx >= W.B * G.nch) return;
    if (!W.pass[idx]) { W.nsym[idx] = 0; return; }
  }
  const int b = idx / G.nch, c = idx - b * G.nch;
  const float *gp = demT + ((long)b * G.dem_rows) * G.nch + c;     // next sample to prefetch
  uint32_t *__restrict__ bits_row = W.bits + (long)idx * G.bw;
  float *soft_row = W.soft ? W.soft + (long)idx * G.n_dem_pad : nullptr;
  MmState st{G.mu0, G.mm.omega_mid, 0.0f};
  unsigned ii = 0;
  int oo = 0;
  uint32_t word = 0;
  if (mode == 2) {
    const MmSave sv = save[idx];
    st = MmState{sv.mu, sv.omega, sv.last};
    ii = sv.ii; oo = sv.oo; word = sv.word;
  }
  const int avail = (mode == 1) ? G.ne_dem : G.n_dem;            // demod floats that exist
  const int oo_end = (mode == 1) ? G.sym_target : G.n_dem;
  const unsigned ni = (unsigned)((mode == 1 ? G.ne_dem : G.n_dem) - 8);
  int pf = (int)ii;                // samples [pf-RD, pf) are in (or on their way to) the ring
  const int tid = threadIdx.x;
  const int nch = G.nch;
  const MmConst K = G.mm;
  gp += (long)pf * nch;
  auto refill = [&](int want) {
    for (; pf < want; pf++, gp += nch) {
      const int rr = pf & (RD - 1);
      mm_cp_async4(&ring[rr][tid], gp);
      if (rr < 8) mm_cp_async4(&ring[rr + RD][tid], gp);
    }
    mm_cp_async_commit();
  };
  {
    int want = (int)ii + 8 + AHEAD; if (want > avail) want = avail;
    refill(want);
    mm_cp_async_wait_all();
    if (G.dem_grid && ii == 0) ring[0][tid] = 0.0f;      // demod_out[0] of a window is never written by the reference
  }
  while (oo < oo_end && ii < ni) {
    mm_cp_async_wait_all();              // the refill issued a block ago has long landed
    {
      int want = (int)ii + 8 + AHEAD; if (want > avail) want = avail;
      refill(want);
    }
#pragma unroll 1
    for (int t = 0; t < PERIOD && oo < oo_end && ii < ni; t++) {
      // rint(mu * 128) without the conversion unit: 0 <= mu < 1, so adding 1.5 * 2^23 rounds to nearest-even at the
      // units place and leaves the integer in the mantissa (= __float2int_rn)
      int imu = __float_as_int(__fadd_rn(st.mu * 128.0f, 12582912.0f)) - 0x4B400000;
      imu = imu < 0 ? 0 : (imu > 128 ? 128 : imu);
      const float *rp = &ring[ii & (RD - 1)][tid];
      const float *mp = &s_mmse[0][imu];
      float out = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; k++) out = out + rp[k * BLK] * mp[k * 132];
      if (soft_row) soft_row[oo] = out;
      if (!(out < 0)) word |= 1u << (oo & 31);
      if ((oo & 31) == 31) { bits_row[oo >> 5] = word; word = 0; }
      ii += (unsigned)mm_update(K, st, out);
      oo++;
    }
  }
  mm_cp_async_wait_all();
  if (mode == 1) save[idx] = MmSave{st.mu, st.omega, st.last, ii, oo, word};
  if (oo & 31) bits_row[oo >> 5] = word;
  for (int w = (oo + 31) >> 5; w < G.bw; w++) bits_row[w] = 0;
  W.nsym[idx] = oo;
}

}  // namespace btb200
