// rx_nest.cuh -- off-channel ("noise") energy estimator of the throughput mode, see rx_nest.cu.
#pragma once
#include <cuda_runtime.h>
#include "rx_math.cuh"
#include "rx_kernels.cuh"

namespace btb200 {

constexpr int NEST_R = 16;           // outputs a thread slides through the taps
constexpr int NEST_K = 3;            // runs per parity per tile
constexpr int NEST_TO = NEST_R * NEST_K;   // outputs per parity per tile
constexpr int NEST_NCOL = 5;         // channels per thread in the DFT stage
constexpr int NEST_RUNS_V = 3;       // fold >= 2: runs of 16 outputs per tile (threads = NEST_RUNS_V * fold * M)

struct NestPlan {
  int M = 0, D = 0, Q = 0, q_rows = 0;   // branches, decimation (2 D = M), taps per branch, rows of hq (multiple of 16)
  int N1 = 1, N2 = 1, CPC = 0, ncol = 0, nch = 0;
  int S = 0, fns = 0, n_noise = 0;
  int tiles_per_slot = 0;
  int period = 0;
  // stride 1: every output of the window (two interleaved sequences, 2 x 48 per tile).  stride 2: the EVEN outputs only
  // (one sequence, 96 per tile) combined with `weights` -- |y|^2 is the square of a signal the 45 kHz noise filter
  // band-limits to +-65 kHz, sampled at 2 Msps by the reference, so its sum over the 850 outputs equals twice the sum
  // over the even ones plus Euler-Maclaurin end corrections (weights 2, with 1.3125, 2.25, 1.9375 at either end and a
  // closing 0.3125) to ~1e-4 relative -- measured <= 3.3e-4 with 25 dB bursts in half of the neighbouring channels,
  // 6e-5 at the benchmark's traffic.  Half the multiply-adds.
  int stride = 1;
  int n_used = 0;                    // stride >= 2: outputs of the sub-sampled sequence that carry weight
  const float *weights = nullptr;    // stride >= 2: [n_used]
  // fold F >= 2: every (2 F)-th output only (stride = 2 F; |y|^2 lives within +-90 kHz, the sub-sampled sequence still
  // runs at 2 Msps / (2 F)), weights from nest_quadrature() (plan.hpp: least squares over the band, end corrections
  // included).  Outputs F M samples apart are one tap apart in a bank of MV = F M "virtual" branches with
  // ceil(Nn / MV) taps each -- the SAME flat tap array h'[k] and the SAME contiguous input range read with row length
  // MV -- so the tap loop is the one of the even-output mode with M -> MV; the F virtual branches r + M f of a real
  // branch r are added up on the way into the N1-point DFTs.  1/F of the multiply-adds of the even-output mode.
  int fold = 1;
  // v2 (fold 2, M = 100 = 4 x 25): the same tiles cut into single runs of 16 outputs, 2 M threads and ~70 KB of shared
  // memory per block, THREE blocks per SM -- the pipeline fill, the per-tile DFT phases and the chunk barriers of one
  // block hide behind the tap loops of the other two (rx_nest.cu: k_nest2).  hq1 = the tap array as plain floats.
  int v2 = 0;
  const float *hq1 = nullptr;
  int q_rows_v = 0;                  // fold >= 2: rows of the tap array read with row length fold * M (multiple of 16, >= ceil(Nn / MV) + 16)
  const float2 *hq2 = nullptr;       // [q_rows][M]  (h, h) with h = h'[r + M q], zero padded: operands of the packed FMAs
  const int *n2_of_rho = nullptr;    // [N2]
  const c32 *WB = nullptr;           // [N2][ncol]
  const int *col_chan = nullptr;     // [ncol]
  const int *chan_col = nullptr;     // [nch]
  const c32 *phasor = nullptr;       // [period]  e^{-j 2 pi phi n / M}
  c32 *xr = nullptr;                 // pre-rotated input
  float *E2 = nullptr;               // [B][tiles_per_slot][ncol]  per-tile sums of |Z|^2
  double *esum = nullptr;            // [B][nch]  sum over the window's noise outputs
};

// The estimator's launch can carry the resume of the clock-recovery chains of the windows with hits (lazy tail,
// rx_mm.cuh) in its FIRST n_blocks blocks: each of them owns an SM for the duration (the estimator runs one block per
// SM), so the latency-bound chains are not starved of issue slots by compute-bound neighbours, while the other SMs
// work through the estimator's tiles.  n_blocks = 0: estimator only.
struct NestResume {
  Geom G;
  DevBatch W;
  const float *mmse = nullptr;
  const float *demT = nullptr;
  const float *demC = nullptr;       // channel-major copy of the demod floats (rx_pfb.cuh), or null: read demT
  long pitchC = 0;
  void *save = nullptr;
  int n_blocks = 0;                  // blocks of `blk` windows: enough for every window of the batch
  int blk = 64;                      // windows (chains) per resume block: 64, or 32 (channel-major source only)
  // v2 (several blocks per SM): a resume block that has windows raises sm_flag[its SM] while its chains run and the
  // estimator blocks that find the flag up wait (bounded) before they start -- the chains keep their SM to themselves
  int *sm_flag = nullptr;            // [256], zero before the launch; null: share the SM
};
constexpr int NEST_RESUME_BLK = 64;

size_t nest_smem_bytes(const NestPlan &P);
int  nest_setup(const NestPlan &P);      // 0, or -1 when the configuration is outside the kernel's limits
// xr[n] = x[n] * phasor[(n0 + n) % period] for n < n_samples
void launch_nest_prerot(const NestPlan &P, const c32 *x, long n_samples, cudaStream_t s);
// esum[b][c] = sum_j |noise DDC output j of window b, channel c|^2 (lib/multi_block.cc:253-287) from P.xr
// true when launch_nest can carry the resume for this plan (block size and shared memory suffice)
bool nest_can_resume(const NestPlan &P);
void launch_nest(const NestPlan &P, int B, cudaStream_t s, const NestResume *resume = nullptr);

}  // namespace btb200
