// rx_fast.cuh -- fast (tolerance-level) noise-floor estimator, see rx_fast.cu.
#pragma once
#include <cuda_runtime.h>
#include "rx_math.cuh"

namespace btb200 {

struct FastNoisePlan {
  int M = 0;            // samples per MHz = polyphase period (fs / 1 MHz)
  int C = 0;            // output classes: M / D
  int Q = 0;            // taps per branch: ceil(Nn / M)
  int period = 0;       // phasor table length
  int nchp = 0;         // channels padded to a multiple of 16
  // device tables
  const float *hpad = nullptr;      // [(Q + 16) * M] zero padded prototype, hpad[(q + 8) * M + r] = h'[r + M q]
  const c32 *phasor = nullptr;      // [period]  e^{-j 2 pi phi n / M}
  const c32 *twid = nullptr;        // [M][nchp] e^{-j 2 pi a_c r / M}
  c32 *U = nullptr;                 // [B][n_noise][M] branch sums
  double *esum = nullptr;           // [B][nch] sum_j |Y_c[j]|^2
};

// x: batch input; per slot b the noise DDC reads x[b*S + fns + j*D + k]
void launch_noise_fast(const FastNoisePlan &F, const c32 *x, int B, int S, int fns, int D, int n_noise, int nch,
                       cudaStream_t s);

}  // namespace btb200
