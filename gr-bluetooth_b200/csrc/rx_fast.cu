// rx_fast.cu -- polyphase + DFT estimate of the off-channel ("noise") energy of every
// channel-window, used by BTB200_SNR_FAST_GUARDED.  NOT on the bit-exact path: compiled with
// FMA contraction, arbitrary summation order, ~1e-6 relative error.  The exact value (20001-tap
// direct-form FIR in the oracle's order, rx_kernels.cu) is still computed whenever the estimate
// is within the guard band of a decision or print-rounding boundary (DESIGN.md "SNR modes").
//
// Math (lib/multi_block.cc:253-296 restated): the noise DDC of channel c is
//   y_c[j] = sum_k x[n_j + k] h'[k] e^{-j theta_c k}  (up to a unit-modulus factor),
//   theta_c = 2 pi (a_c + phi) / M,  a_c integer MHz offset, phi the common fractional MHz offset
//   (0.79 at fc = 2441 MHz), M = samples per MHz.  With x'[n] = x[n] e^{-j 2 pi phi n / M}:
//   |y_c[j]| = | sum_{r<M} e^{-j 2 pi a_c r / M} u_j[r] |,   u_j[r] = sum_q x'[n_j + r + M q] h'[r + M q]
// i.e. ONE real-tap polyphase bank (u, 20001 real-complex MACs per output time for all channels)
// followed by an M-point DFT evaluated at the channel bins.
#include "rx_fast.cuh"

namespace btb200 {

constexpr int PF_RUN = 8;      // consecutive same-class outputs per thread

// Branch sums.  Thread = (branch r, run of PF_RUN outputs of one class); consecutive threads =
// consecutive r, so input, phasor and tap loads are contiguous.  Sliding windows of PF_RUN taps
// keep every loaded input sample in use for PF_RUN real-complex MACs.
__global__ void __launch_bounds__(128) k_noise_poly(FastNoisePlan F, const c32 *__restrict__ x, int S, int fns, int D,
                                                    int n_noise)
{
  const int b = blockIdx.z, p = blockIdx.y;
  const int n_i = (n_noise - p + F.C - 1) / F.C;            // outputs of this class
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t % F.M, ib = t / F.M;
  const int i0 = ib * PF_RUN;
  if (i0 >= n_i) return;
  const long n0 = (long)b * S + fns + (long)p * D + r + (long)F.M * i0;     // sample of (output i0, tap q = 0)
  const c32 *xp = x + n0;
  int pidx = (int)(n0 % F.period);
  const int pstep = F.M % F.period;
  const float *hp = F.hpad + r;                              // hp[(q + 8) * M]
  float ar[PF_RUN], ai[PF_RUN], tw[PF_RUN];
#pragma unroll
  for (int o = 0; o < PF_RUN; o++) { ar[o] = 0.f; ai[o] = 0.f; }
  // step s loads sample X[i0 + s]; output o uses tap q = s - o.  tw[o] holds h'[r + M (s - o)].
#pragma unroll
  for (int o = 0; o < PF_RUN; o++) tw[o] = hp[(long)(8 - o) * F.M];          // s = 0: q = -o (zero padding for q < 0)
  const int steps = F.Q + PF_RUN - 1;
  for (int s = 0; s < steps; s++) {
    const c32 v = xp[(long)s * F.M];
    const c32 ph = F.phasor[pidx];
    pidx += pstep; if (pidx >= F.period) pidx -= F.period;
    const float xr = v.re * ph.re - v.im * ph.im;
    const float xi = v.re * ph.im + v.im * ph.re;
#pragma unroll
    for (int o = 0; o < PF_RUN; o++) { ar[o] = fmaf(xr, tw[o], ar[o]); ai[o] = fmaf(xi, tw[o], ai[o]); }
#pragma unroll
    for (int o = PF_RUN - 1; o > 0; o--) tw[o] = tw[o - 1];
    tw[0] = hp[(long)(s + 1 + 8) * F.M];
  }
#pragma unroll
  for (int o = 0; o < PF_RUN; o++) {
    const int j = p + F.C * (i0 + o);
    if (i0 + o < n_i) F.U[((long)b * n_noise + j) * F.M + r] = c32{ar[o], ai[o]};
  }
}

// DFT at the channel bins + |.|^2 accumulation.  Block = (slot b, tile of 32 output times);
// thread = (channel, group of 8 output times).
constexpr int DF_JT = 32, DF_JR = 8;
__global__ void k_noise_dft(FastNoisePlan F, int n_noise, int nch)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  c32 *Wt = reinterpret_cast<c32 *>(smem_raw);                 // [M][nchp]
  c32 *Us = Wt + (size_t)F.M * F.nchp;                         // [DF_JT][M]
  __shared__ float s_part[DF_JT / DF_JR][96];
  const int b = blockIdx.y, j0 = blockIdx.x * DF_JT;
  const int nthreads = blockDim.x;
  for (int i = threadIdx.x; i < F.M * F.nchp; i += nthreads) Wt[i] = F.twid[i];
  const int jn = (n_noise - j0) < DF_JT ? (n_noise - j0) : DF_JT;
  const c32 *Ug = F.U + ((long)b * n_noise + j0) * F.M;
  for (int i = threadIdx.x; i < DF_JT * F.M; i += nthreads) Us[i] = (i < jn * F.M) ? Ug[i] : c32{0.f, 0.f};
  __syncthreads();
  const int c = threadIdx.x % F.nchp, jq = threadIdx.x / F.nchp;
  float ar[DF_JR], ai[DF_JR];
#pragma unroll
  for (int k = 0; k < DF_JR; k++) { ar[k] = 0.f; ai[k] = 0.f; }
  const c32 *up = Us + (size_t)jq * DF_JR * F.M;
  for (int r = 0; r < F.M; r++) {
    const c32 w = Wt[r * F.nchp + c];
#pragma unroll
    for (int k = 0; k < DF_JR; k++) {
      const c32 u = up[k * F.M + r];
      ar[k] = fmaf(u.re, w.re, ar[k]); ar[k] = fmaf(-u.im, w.im, ar[k]);
      ai[k] = fmaf(u.re, w.im, ai[k]); ai[k] = fmaf(u.im, w.re, ai[k]);
    }
  }
  float e = 0.f;
#pragma unroll
  for (int k = 0; k < DF_JR; k++) e += ar[k] * ar[k] + ai[k] * ai[k];      // padded output times contribute 0
  s_part[jq][c] = e;
  __syncthreads();
  if (jq == 0 && c < nch) {
    double tot = 0.0;
    for (int q = 0; q < DF_JT / DF_JR; q++) tot += (double)s_part[q][c];
    atomicAdd(&F.esum[(long)b * nch + c], tot);
  }
}

void launch_noise_fast(const FastNoisePlan &F, const c32 *x, int B, int S, int fns, int D, int n_noise, int nch,
                       cudaStream_t s)
{
  cudaMemsetAsync(F.esum, 0, sizeof(double) * (size_t)B * nch, s);
  const int n_i = (n_noise + F.C - 1) / F.C;
  const int threads = F.M * ((n_i + PF_RUN - 1) / PF_RUN);
  dim3 g1((unsigned)((threads + 127) / 128), (unsigned)F.C, (unsigned)B);
  k_noise_poly<<<g1, 128, 0, s>>>(F, x, S, fns, D, n_noise);
  const int nthreads = F.nchp * (DF_JT / DF_JR);
  const size_t smem = sizeof(c32) * ((size_t)F.M * F.nchp + (size_t)DF_JT * F.M);
  static bool opted[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!opted[dev & 63]) { cudaFuncSetAttribute(k_noise_dft, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); opted[dev & 63] = true; }
  dim3 g2((unsigned)((n_noise + DF_JT - 1) / DF_JT), (unsigned)B);
  k_noise_dft<<<g2, nthreads, smem, s>>>(F, n_noise, nch);
}

}  // namespace btb200
