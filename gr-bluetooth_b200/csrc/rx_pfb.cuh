// rx_pfb.cuh -- polyphase (weighted overlap-add) channelizer of the throughput mode, see rx_pfb.cu.
#pragma once
#include <cuda_runtime.h>
#include "rx_math.cuh"

namespace btb200 {

constexpr int PFB_T = 63;          // grid points a tile owns
constexpr int PFB_TT = 64;         // grid points a tile computes (one more in front: the demod needs Z[g-1])
constexpr int PFB_NCOL = 5;        // channels per thread in the DFT stage

struct PfbPlan {
  int M = 0, D = 0, Q = 0;         // samples per MHz (= branches), decimation (2 D = M), taps per branch
  int N1 = 1, N2 = 1;              // M = N1 * N2, coprime (Good-Thomas); N1 in {1, 2, 4}
  int gps = 0, n_ddc = 0;          // grid points per slot, DDC outputs per window
  int nfull = 0, rem = 0;          // n_ddc = nfull * gps + rem
  int fcs = 0, nch = 0;
  int tps = 0;                     // tiles per slot segment
  int CPC = 0, ncol = 0;           // columns per residue class (multiple of PFB_NCOL), columns = N1 * CPC
  int span = 0;                    // samples staged per tile: (TT - 1) * D + Q * M
  float gain = 0;                  // demod gain (lib/multi_block.cc:88)
  float phi_step = 0;              // pre-rotation: x'[i] = x[i] e^{j pi phi_step i}, phi_step = -2 phi / M (0: none)
  // device tables: ONE blob in the kernel's shared-memory layout (pfb_pack_tables), fetched per tile by a bulk copy:
  //   WB [N2][ncol] c32 W_N2^{n2 k2(col)} | hq [Q][M] f32 h'[r + M q] | atan [257] | kappa [ncol] c32
  //   e^{-j 2 pi a_c D / M} | col_chan [ncol] i32 (-1: padding) | n2_of_rho [N2] i32
  const unsigned char *tables = nullptr;
  const int *chan_col = nullptr;   // [nch]   column of a channel
  // outputs
  float *dem = nullptr;            // [Gtot][nch]  demod floats on the global decimation grid
  float *E = nullptr;              // [segments * tps][ncol][2]  per-tile sums of |Z|^2 (all points / points below rem)
  // optional channel-major copy of the demod floats, demC[c * pitchC + g]: what the resume of the clock-recovery
  // chains of the windows with hits reads (rx_mm.cuh, CM) -- a window is a contiguous run of its channel's row there
  float *demC = nullptr;
  long pitchC = 0;                 // floats per channel row (multiple of 4, >= Gtot + 8)
  // optional second output: the input rotated by the noise DDCs' common fractional offset, xr[n] = x[n] phasor[n % period]
  // -- what the noise estimator (rx_nest.cu) reads.  Every tile writes the samples of its own grid points from the
  // span it has staged anyway (tile 0 also the fcs samples in front of the first grid point), which saves the
  // estimator's own pass over the input (one read of the batch).  Covers samples [0, fcs + Gtot D).
  c32 *xr = nullptr;
  const c32 *phasor = nullptr;     // [period]
  int period = 0;
};

size_t pfb_smem_bytes(const PfbPlan &P);
size_t pfb_table_bytes(const PfbPlan &P);
void pfb_pack_tables(const PfbPlan &P, const c32 *WB, const float *hq, const float *atan_tab, const c32 *kappa,
                     const int *col_chan, const int *n2_of_rho, unsigned char *blob);
int  pfb_setup(const PfbPlan &P);                                 // opt in to the dynamic shared memory; 0 or -1
long pfb_tiles(const PfbPlan &P, int B);
// input samples (from x[0]) the tiles below `tile_end` read
long pfb_samples(const PfbPlan &P, long tile_end);
// x: batch input (x[0] = first sample of window 0), n_samples valid samples; tiles [tile0, tile1)
void launch_pfb(const PfbPlan &P, const c32 *x, long n_samples, int B, long tile0, long tile1, cudaStream_t s);
// e_on[b][c] = mean |Z|^2 over window b (the on-channel energy of lib/multi_block.cc:206-218)
void launch_pfb_energy(const PfbPlan &P, int B, double *e_on, cudaStream_t s);

// int16 interleaved IQ -> complex64 (the reference's interleaved_short_to_complex in front of the block, apps/btrx:141-159)
void launch_i16_to_c32(const int16_t *src, c32 *dst, long n_samples, cudaStream_t s);

}  // namespace btb200
