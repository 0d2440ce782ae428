// rx_nest.cu -- off-channel ("noise") energy of EVERY channel-window in the throughput mode: the reference's
// 20001-tap noise DDCs (check_snr, lib/multi_block.cc:253-296: one freq_xlating_fir_filter_ccf per channel at
// f_ch + 790 kHz, 850 outputs of the window's first slot, mean |y|^2) restated as ONE real-tap polyphase bank
// of M = fs / 1 MHz branches with Q = ceil(Nn / M) = 201 taps each, shared by all channels, followed by the same
// Good-Thomas DFT at the channel bins as rx_pfb.cu.  fp32 with FMA: a tolerance-level estimate (measured <= 7e-4 dB
// from the exact energy, tests/test_gpu_parity.py::test_fast_noise_estimate_accuracy); the exact value is
// rx_firdl.cu's job.
//
//   x'[n]   = x[n] e^{-j 2 pi phi n / M}              pre-rotation by the common fractional offset (own pass)
//   u_j[r]  = sum_{q<Q} x'[n_j + r + M q] h'[r + M q]    n_j = b S + fns + j D,  j < 850
//   |Y_c[j]| = | sum_r e^{-j 2 pi a_c r / M} u_j[r] |
//
// One block = one tile of 2 x 48 outputs (48 per parity of j: outputs two apart are exactly one tap apart, 2 D = M)
// of one slot.  Thread = (branch r, parity p, run k of 16 outputs): it slides its 16 accumulators through the 201
// taps -- per step ONE new input sample and ONE new tap feed 16 complex-by-real MACs (the last 16 taps sit in a
// statically rotated register window) -- so shared-memory traffic is 12 bytes per 32 FMA.  Inputs and taps stream
// through shared memory in chunks of 16 steps: a ring of 6 x 16 x M samples (both parities read the SAME contiguous
// sample range) and a double-buffered tap chunk, each filled by one TMA bulk copy (cp.async.bulk + mbarrier) issued
// a whole chunk of compute ahead.  Then, per tile: N1-point DFTs in place, N2-point DFTs at the channel bins (warp
// = 32 output groups x 4 channels), |Z|^2 summed over the tile's valid outputs.
#include "rx_nest.cuh"
#include "rx_tma.cuh"
#include "rx_packed.cuh"
#include "rx_mm.cuh"

namespace btb200 {

namespace {

constexpr int RING_MAX = 2 * NEST_K + 2;  // ring capacity in chunks of 16 steps: live chunks + one in flight (stride 1: K + 2, stride 2: 2 K + 2)
// compute chunk c reads ring chunks c .. c + live_ahead: stride 1: runs 0..K-1 plus the half-step offset of the odd
// sequence; stride 2: runs 0..2K-1, no offset
__host__ __device__ inline int live_ahead(const NestPlan &P) { return P.fold > 1 ? NEST_RUNS_V - 1 : P.stride == 2 ? 2 * NEST_K - 1 : NEST_K; }
__host__ __device__ inline int ring_chunks(const NestPlan &P) { return live_ahead(P) + 2; }
constexpr int CH = 16;               // steps per chunk

struct NestSmem { size_t ring, taps, wb, n2r, epart, bar, total; };

__host__ __device__ inline NestSmem nest_layout(const NestPlan &P)
{
  NestSmem L{};
  size_t o = 0;
  auto take = [&o](size_t bytes, size_t align) { o = (o + align - 1) / align * align; const size_t r = o; o += bytes; return r; };
  const int MV = P.fold * P.M;                                            // row length of the tap loop
  const int n_out = P.fold > 1 ? NEST_R * NEST_RUNS_V : 2 * NEST_TO;      // outputs per tile
  const size_t ring = (size_t)ring_chunks(P) * CH * MV * sizeof(c32);
  const size_t u = (size_t)n_out * (MV + 1) * sizeof(c32);                // U / V tile, row pitch MV + 1
  L.ring = take(ring > u ? ring : u, 128);
  L.taps = take((size_t)2 * CH * MV * sizeof(float2), 128);
  L.wb = take((size_t)P.N2 * P.ncol * sizeof(c32), 16);
  L.n2r = take((size_t)P.N2 * sizeof(int), 16);
  L.epart = take((size_t)32 * P.ncol * sizeof(float), 16);
  L.bar = take((RING_MAX + 2) * 8, 8);
  L.total = o;
  return L;
}

__global__ void k_nest_prerot(const c32 *__restrict__ x, c32 *__restrict__ xr, long n, const c32 *__restrict__ ph, int period)
{
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const c32 v = x[i], p = ph[(int)(i % period)];
  xr[i] = c32{v.re * p.re - v.im * p.im, v.re * p.im + v.im * p.re};
}

// MT: the number of branches as a compile-time constant (100 = the benchmark configuration: every shared-memory
// access of the tap loop then has an immediate offset), or 0 for "read it from the plan"
// FOLD: 1 = every output or the even ones (P.stride); F >= 2 = every (2 F)-th output through F M virtual branches
template <int N1, int MT, int FOLD>
__global__ void __launch_bounds__(2 * NEST_K * 100, 1) k_nest(NestPlan P, NestResume R)
{
  extern __shared__ __align__(128) unsigned char smem[];
  if ((int)blockIdx.x < R.n_blocks) {
    // the first blocks resume clock-recovery chains instead (one block per SM: the chains get the SM to themselves)
    if (R.demC)
      mm_stateless_block<NEST_RESUME_BLK, true>(R.G, R.W, R.mmse, R.demT, 2, reinterpret_cast<MmSave *>(R.save), R.W.tail.list, -1,
                                                smem, (int)blockIdx.x, R.demC, R.pitchC);
    else
      mm_stateless_block<NEST_RESUME_BLK>(R.G, R.W, R.mmse, R.demT, 2, reinterpret_cast<MmSave *>(R.save), R.W.tail.list, -1,
                                          smem, (int)blockIdx.x);
    return;
  }
  const int tile_index = (int)blockIdx.x - R.n_blocks;
  const NestSmem L = nest_layout(P);
  c32 *ring = reinterpret_cast<c32 *>(smem + L.ring);
  float2 *taps = reinterpret_cast<float2 *>(smem + L.taps);
  c32 *WBs = reinterpret_cast<c32 *>(smem + L.wb);
  int *n2r = reinterpret_cast<int *>(smem + L.n2r);
  float *epart = reinterpret_cast<float *>(smem + L.epart);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);          // [0..5] ring slots, [6..7] tap buffers

  constexpr bool FV = FOLD > 1;
  const int M = MT ? MT : P.M, N2 = MT ? MT / N1 : P.N2, ncol = P.ncol;
  const int MV = FOLD * M;                                             // row length of the tap loop (virtual branches)
  const int tid = threadIdx.x, nthr = blockDim.x;                      // nthr = 2 * NEST_K * M, or NEST_RUNS_V * MV
  const int r = tid % MV, pk = tid / MV, p = pk & 1, k = pk >> 1;
  const int b = tile_index / P.tiles_per_slot, tile = tile_index - b * P.tiles_per_slot;
  const int RING_CHUNKS = ring_chunks(P), AHEAD = live_ahead(P);
  const bool sub = FV || P.stride == 2;
  // stride 1: the tile is 48 outputs of each parity, run (p, k) = 16 outputs i0 + 16 k + o of parity p.
  // stride 2: the tile is 96 consecutive outputs of the even sequence, run g = 2 k + p... = 16 outputs i0 + 16 g + o.
  // fold F:   the tile is 16 NEST_RUNS_V consecutive outputs of the sub-sampled sequence, run = tid / MV.
  const int run = sub ? pk : k;                                        // run start, in units of 16 steps
  constexpr int N_OUT = FV ? NEST_R * NEST_RUNS_V : 2 * NEST_TO;       // outputs per tile
  const int i0 = tile * (sub ? N_OUT : NEST_TO);                       // first output index (in its sequence) of the tile
  const long n_base = (long)b * P.S + P.fns + (long)MV * i0;           // sample of ring step 0, branch 0, first sequence
  const int n_chunks = (FV ? P.q_rows_v : P.q_rows) / CH;              // compute chunks
  const int n_ring = n_chunks + AHEAD;                                 // ring chunks the tile reads
  const unsigned ring_bytes = (unsigned)(CH * MV * sizeof(c32)), tap_bytes = (unsigned)(CH * MV * sizeof(float2));

  if (tid == 0) {
    for (int i = 0; i < RING_MAX + 2; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto load_ring = [&](int m) {
    uint64_t *bb = &bar[m % RING_CHUNKS];
    mbar_expect_tx(bb, ring_bytes);
    tma_bulk_g2s(ring + (size_t)(m % RING_CHUNKS) * CH * MV, P.xr + n_base + (long)m * CH * MV, ring_bytes, bb);
  };
  auto load_taps = [&](int c) {
    uint64_t *bb = &bar[RING_MAX + (c & 1)];
    mbar_expect_tx(bb, tap_bytes);
    tma_bulk_g2s(taps + (size_t)(c & 1) * CH * MV, P.hq2 + (size_t)c * CH * MV, tap_bytes, bb);
  };
  if (tid == 0) {
    for (int m = 0; m < RING_CHUNKS && m < n_ring; m++) load_ring(m);
    load_taps(0);
    if (n_chunks > 1) load_taps(1);
  }
  for (int i = tid; i < N2 * ncol; i += nthr) WBs[i] = P.WB[i];
  for (int i = tid; i < N2; i += nthr) n2r[i] = P.n2_of_rho[i];

  // ---- 1. branch sums: 16 outputs i0 + 16 k + o of parity p, branch r
  // packed fp32 (FFMA2): accumulator = (re, im) of one output, input = (re, im) as loaded, tap = (h, h) as loaded.
  // A 3-register FFMA whose operands share a register bank issues at half rate; the packed form reads register
  // PAIRS, one of which (the input) is reused by all 16 MACs of a step, so it runs at the full fp32 rate.
  u64 acc[NEST_R], H[NEST_R];
#pragma unroll
  for (int o = 0; o < NEST_R; o++) { acc[o] = 0ull; H[o] = 0ull; }
  const int ring_samples = RING_CHUNKS * CH * MV;
  const u64 *ring64 = reinterpret_cast<const u64 *>(ring);
  for (int c = 0; c < n_chunks; c++) {
    if (c == 0) { for (int m = 0; m < AHEAD && m < n_ring; m++) mbar_wait(&bar[m], 0); }
    if (c + AHEAD < n_ring) mbar_wait(&bar[(c + AHEAD) % RING_CHUNKS], (unsigned)(((c + AHEAD) / RING_CHUNKS) & 1));
    mbar_wait(&bar[RING_MAX + (c & 1)], (unsigned)((c >> 1) & 1));
    int off = (int)(((long)MV * (NEST_R * run + CH * c) + (sub ? 0 : (long)(M / 2) * p) + r) % ring_samples);
    const u64 *tp = reinterpret_cast<const u64 *>(taps + (size_t)(c & 1) * CH * MV + r);
    if (off + CH * MV <= ring_samples) {
      // the 16 steps do not cross the end of the ring (4 chunks in 5): one base address, immediate offsets
      const u64 *xp = ring64 + off;
#pragma unroll
      for (int s = 0; s < CH; s++) {
        const u64 X = xp[s * MV];
        H[s] = tp[s * MV];                            // tap q = 16 c + s; H[i] holds the latest tap with q = i (mod 16)
#pragma unroll
        for (int o = 0; o < NEST_R; o++) acc[o] = pk_fma(X, H[(s - o) & (NEST_R - 1)], acc[o]);   // tap q = 16 c + s - o
      }
    } else {
#pragma unroll
      for (int s = 0; s < CH; s++) {
        const u64 X = ring64[off];
        off += MV; if (off >= ring_samples) off -= ring_samples;
        H[s] = tp[s * MV];
#pragma unroll
        for (int o = 0; o < NEST_R; o++) acc[o] = pk_fma(X, H[(s - o) & (NEST_R - 1)], acc[o]);
      }
    }
    __syncthreads();                                  // ring chunk c and tap chunk c are free
    if (tid == 0) {
      fence_proxy_async();
      if (c + RING_CHUNKS < n_ring) load_ring(c + RING_CHUNKS);
      if (c + 2 < n_chunks) load_taps(c + 2);
    }
  }
  // ---- the tile's branch sums to shared memory (the ring is dead): U[out][r], out = p * TO + 16 k + o
  const int UP = MV + 1;
  c32 *U = ring;
#pragma unroll
  for (int o = 0; o < NEST_R; o++)
    reinterpret_cast<u64 *>(U)[(size_t)((sub ? 0 : p * NEST_TO) + NEST_R * run + o) * UP + r] = acc[o];
  __syncthreads();

  // ---- 2. N1-point DFTs, in place: V[out][k1 * N2 + n2] = sum_n1 U[out][(N2 n1 + N1 n2) mod M] W_N1^{n1 k1}
  {
    constexpr int ITEMS = 16 / (N1 * FOLD);           // N_OUT * N2 items on 6 * N1 * N2 (or NEST_RUNS_V * FOLD * N1 * N2) threads
    static_assert(ITEMS >= 1 && ITEMS * N1 * FOLD == 16, "fold and N1 must divide 16");
    float vr[ITEMS][N1], vi[ITEMS][N1];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int it = tid + j * nthr;
      const int out = it / N2, rho = it - out * N2, n2 = n2r[rho];
      float ur[N1], ui[N1];
#pragma unroll
      for (int n1 = 0; n1 < N1; n1++) {
        const c32 *up = U + (size_t)out * UP + (N2 * n1 + N1 * n2) % M;
        c32 u = up[0];
        if constexpr (FV) {
#pragma unroll
          for (int f = 1; f < FOLD; f++) { const c32 t = up[f * M]; u.re += t.re; u.im += t.im; }    // virtual branches r + M f
        }
        ur[n1] = u.re; ui[n1] = u.im;
      }
      if constexpr (N1 == 1) { vr[j][0] = ur[0]; vi[j][0] = ui[0]; }
      else if constexpr (N1 == 2) {
        vr[j][0] = ur[0] + ur[1]; vi[j][0] = ui[0] + ui[1];
        vr[j][1] = ur[0] - ur[1]; vi[j][1] = ui[0] - ui[1];
      } else {
        const float er = ur[0] + ur[2], ei = ui[0] + ui[2], fr = ur[0] - ur[2], fi = ui[0] - ui[2];
        const float gr = ur[1] + ur[3], gi = ui[1] + ui[3], hr = ur[1] - ur[3], hi = ui[1] - ui[3];
        vr[j][0] = er + gr; vi[j][0] = ei + gi;
        vr[j][1] = fr + hi; vi[j][1] = fi - hr;       // f - j h
        vr[j][2] = er - gr; vi[j][2] = ei - gi;
        vr[j][3] = fr - hi; vi[j][3] = fi + hr;       // f + j h
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int it = tid + j * nthr;
      const int out = it / N2, rho = it - out * N2, n2 = n2r[rho];
#pragma unroll
      for (int k1 = 0; k1 < N1; k1++) U[(size_t)out * UP + k1 * N2 + n2] = c32{vr[j][k1], vi[j][k1]};
    }
    __syncthreads();
  }

  // ---- 3. N2-point DFTs at the channel bins + |Z|^2 over the valid outputs.  item = (og, cg): outputs og + 32 i
  const int n_cg = ncol / NEST_NCOL;
  const int n_par[2] = {(P.n_noise + 1) / 2, P.n_noise / 2};            // outputs of each parity in a slot
  constexpr int OG = FV ? 16 : 32;                    // output groups: a thread takes outputs og + OG i
  for (int item = tid; item < OG * n_cg; item += nthr) {
    const int og = item & (OG - 1), cg = item / OG;
    const int col0 = cg * NEST_NCOL, k1 = col0 / P.CPC;
    constexpr int NI = N_OUT / OG;                    // outputs per thread
    float zr[NI][NEST_NCOL], zi[NI][NEST_NCOL];
#pragma unroll
    for (int i = 0; i < NI; i++)
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) { zr[i][j] = 0.0f; zi[i][j] = 0.0f; }
    const c32 *vrow = U + (size_t)og * UP + k1 * N2;
    const c32 *wrow = WBs + col0;
#pragma unroll 5
    for (int n2 = 0; n2 < N2; n2++) {
      c32 v[NI], w[NEST_NCOL];
#pragma unroll
      for (int i = 0; i < NI; i++) v[i] = vrow[(size_t)OG * i * UP + n2];
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) w[j] = wrow[n2 * ncol + j];
#pragma unroll
      for (int i = 0; i < NI; i++)
#pragma unroll
        for (int j = 0; j < NEST_NCOL; j++) {
          zr[i][j] = fmaf(v[i].re, w[j].re, zr[i][j]); zr[i][j] = fmaf(-v[i].im, w[j].im, zr[i][j]);
          zi[i][j] = fmaf(v[i].re, w[j].im, zi[i][j]); zi[i][j] = fmaf(v[i].im, w[j].re, zi[i][j]);
        }
    }
    float e[NEST_NCOL];
#pragma unroll
    for (int j = 0; j < NEST_NCOL; j++) e[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < NI; i++) {
      const int out = og + OG * i;
      float wgt;
      if (sub) {
        const int u = i0 + out;
        wgt = u < P.n_used ? P.weights[u] : 0.0f;
      } else {
        const int pp = out / NEST_TO, idx = i0 + (out - pp * NEST_TO);
        wgt = idx < n_par[pp] ? 1.0f : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) e[j] = fmaf(wgt, zr[i][j] * zr[i][j] + zi[i][j] * zi[i][j], e[j]);
    }
#pragma unroll
    for (int j = 0; j < NEST_NCOL; j++) epart[og * ncol + col0 + j] = e[j];
  }
  __syncthreads();
  for (int col = tid; col < ncol; col += nthr) {
    float sum = 0.0f;
    for (int og = 0; og < OG; og++) sum += epart[og * ncol + col];
    P.E2[(size_t)tile_index * ncol + col] = sum;
  }
}

__global__ void k_nest_reduce(NestPlan P, int B)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * P.nch) return;
  const int b = idx / P.nch, c = idx - b * P.nch;
  const int col = P.chan_col[c];
  double sum = 0.0;
  for (int t = 0; t < P.tiles_per_slot; t++) sum += (double)P.E2[((size_t)b * P.tiles_per_slot + t) * P.ncol + col];
  P.esum[idx] = sum;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// k_nest2: the folded estimator (fold 2) in small blocks.  k_nest's tile is 48 outputs on 600 threads with 180 KB of
// shared memory: one block per SM, and by ncu only ~27 % of a tile's time is the tap loop at full FMA rate -- the rest
// is the pipeline fill (nothing to do until the first 100 KB have landed), the chunk barriers and the DFT phases, which
// nothing overlaps.  Here a block is ONE run of 16 outputs on 2 M threads (thread = virtual branch), chunks of 8 steps,
// a ring of 4 chunks (51 KB: three chunks of look-ahead), taps as plain floats (paired in a register), the DFT matrix
// fetched by a bulk copy into the dead ring while the branch sums are folded: ~70 KB, three blocks per SM, each in a
// different phase.  Same arithmetic per output as k_nest<.., 2> (same tap array, same fold, same DFT, same weights).
constexpr int N2_CH = 8;             // steps per chunk
constexpr int N2_RC = 4;             // ring slots (chunks)
constexpr int N2_OUT = NEST_R;       // outputs per tile
constexpr int N2_OG = 8;             // output groups of the N2-point DFT stage

struct Nest2Smem { size_t ring, wb_off, taps, n2r, epart, bar, total; };

__host__ __device__ inline Nest2Smem nest2_layout(const NestPlan &P)
{
  Nest2Smem L{};
  size_t o = 0;
  auto take = [&o](size_t bytes, size_t align) { o = (o + align - 1) / align * align; const size_t r = o; o += bytes; return r; };
  const int MV = 2 * P.M;
  const size_t ring = (size_t)N2_RC * N2_CH * MV * sizeof(c32);
  const size_t u = (size_t)N2_OUT * (MV + 1) * sizeof(c32);               // U / V tile, row pitch MV + 1
  L.wb_off = (u + 127) / 128 * 128;                                       // the DFT matrix follows the U tile in the dead ring
  const size_t after = L.wb_off + (size_t)P.N2 * P.ncol * sizeof(c32);
  L.ring = take(ring > after ? ring : after, 128);
  L.taps = take((size_t)2 * N2_CH * MV * sizeof(float), 128);
  L.n2r = take((size_t)P.N2 * sizeof(int), 16);
  L.epart = take((size_t)N2_OG * P.ncol * sizeof(float), 16);
  L.bar = take((N2_RC + 3) * 8, 8);
  L.total = o;
  return L;
}

template <int V> struct IntC { static constexpr int value = V; };

template <int N1, int MT>
__global__ void __launch_bounds__(2 * MT, 3) k_nest2(NestPlan P, NestResume R)
{
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < R.n_blocks) {
    // resume of clock-recovery chains (see k_nest); with windows to resume the block claims its SM for the duration
    __shared__ int s_live;
    if (tid == 0) {
      s_live = (int)blockIdx.x * R.blk < *R.W.tail.n_list;
      if (s_live && R.sm_flag) atomicAdd(&R.sm_flag[smid & 255], R.blk / 32);
    }
    __syncthreads();
    const bool live = s_live != 0;
    if (R.demC && R.blk == 32)
      mm_stateless_block<32, true>(R.G, R.W, R.mmse, R.demT, 2, reinterpret_cast<MmSave *>(R.save), R.W.tail.list, -1,
                                   smem, (int)blockIdx.x, R.demC, R.pitchC);
    else if (R.demC)
      mm_stateless_block<NEST_RESUME_BLK, true>(R.G, R.W, R.mmse, R.demT, 2, reinterpret_cast<MmSave *>(R.save), R.W.tail.list, -1,
                                                smem, (int)blockIdx.x, R.demC, R.pitchC);
    else
      mm_stateless_block<NEST_RESUME_BLK>(R.G, R.W, R.mmse, R.demT, 2, reinterpret_cast<MmSave *>(R.save), R.W.tail.list, -1,
                                          smem, (int)blockIdx.x);
    if (live && R.sm_flag && tid < R.blk && (tid & 31) == 0) atomicSub(&R.sm_flag[smid & 255], 1);   // one per chain warp
    return;
  }
  if (R.sm_flag) {
    // an SM whose resume block is still running its chains: wait (bounded: ~4 ms) instead of taking issue slots from them
    if (tid == 0) {
      const volatile int *f = R.sm_flag + (smid & 255);
      for (int spin = 0; spin < 4000 && *f > 0; spin++) __nanosleep(1000);
    }
    __syncthreads();
  }
  const int tile_index = (int)blockIdx.x - R.n_blocks;
  const Nest2Smem L = nest2_layout(P);
  c32 *ring = reinterpret_cast<c32 *>(smem + L.ring);
  float *taps = reinterpret_cast<float *>(smem + L.taps);
  c32 *WBs = reinterpret_cast<c32 *>(smem + L.ring + L.wb_off);
  int *n2r = reinterpret_cast<int *>(smem + L.n2r);
  float *epart = reinterpret_cast<float *>(smem + L.epart);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);          // [0..3] ring slots, [4..5] tap buffers, [6] DFT matrix

  constexpr int M = MT, N2 = MT / N1, MV = 2 * MT;
  const int ncol = P.ncol, nthr = MV;
  const int b = tile_index / P.tiles_per_slot, tile = tile_index - b * P.tiles_per_slot;
  const int i0 = tile * N2_OUT;                                        // first output (of the sub-sampled sequence) of the tile
  const long n_base = (long)b * P.S + P.fns + (long)MV * i0;           // sample of step 0, virtual branch 0
  const int n_chunks = P.q_rows_v / N2_CH;
  constexpr unsigned ring_bytes = N2_CH * MV * sizeof(c32), tap_bytes = N2_CH * MV * sizeof(float);

  if (tid == 0) {
    for (int i = 0; i < N2_RC + 3; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto load_ring = [&](int m) {
    uint64_t *bb = &bar[m % N2_RC];
    mbar_expect_tx(bb, ring_bytes);
    tma_bulk_g2s(ring + (size_t)(m % N2_RC) * N2_CH * MV, P.xr + n_base + (long)m * N2_CH * MV, ring_bytes, bb);
  };
  auto load_taps = [&](int c) {
    uint64_t *bb = &bar[N2_RC + (c & 1)];
    mbar_expect_tx(bb, tap_bytes);
    tma_bulk_g2s(taps + (size_t)(c & 1) * N2_CH * MV, P.hq1 + (size_t)c * N2_CH * MV, tap_bytes, bb);
  };
  if (tid == 0) {
    for (int m = 0; m < N2_RC && m < n_chunks; m++) load_ring(m);
    load_taps(0);
    if (n_chunks > 1) load_taps(1);
  }
  for (int i = tid; i < N2; i += nthr) n2r[i] = P.n2_of_rho[i];

  // ---- 1. branch sums: outputs i0 + o, o < 16, virtual branch tid.  Step g = 8 c + s brings sample row g and tap row g;
  // acc[o] += x[g] h[g - o]: the last 16 taps sit in a register window whose indices are static because chunks are
  // processed in pairs (H index = g mod 16).
  u64 acc[NEST_R], H[NEST_R];
#pragma unroll
  for (int o = 0; o < NEST_R; o++) { acc[o] = 0ull; H[o] = 0ull; }
  const u64 *ring64 = reinterpret_cast<const u64 *>(ring);
  auto chunk = [&](auto ofs_t, int c) {
    constexpr int OFS = decltype(ofs_t)::value;
    mbar_wait(&bar[c % N2_RC], (unsigned)((c / N2_RC) & 1));
    mbar_wait(&bar[N2_RC + (c & 1)], (unsigned)((c >> 1) & 1));
    const u64 *xp = ring64 + (size_t)(c % N2_RC) * N2_CH * MV + tid;
    const float *tp = taps + (size_t)(c & 1) * N2_CH * MV + tid;
#pragma unroll
    for (int s = 0; s < N2_CH; s++) {
      const u64 X = xp[s * MV];
      const float h = tp[s * MV];
      H[OFS + s] = pk_pack(h, h);
#pragma unroll
      for (int o = 0; o < NEST_R; o++) acc[o] = pk_fma(X, H[(OFS + s - o) & (NEST_R - 1)], acc[o]);
    }
    __syncthreads();                                  // ring chunk c and tap chunk c are free
    if (tid == 0) {
      fence_proxy_async();
      if (c + N2_RC < n_chunks) load_ring(c + N2_RC);
      if (c + 2 < n_chunks) load_taps(c + 2);
    }
  };
  int c = 0;
  for (; c + 1 < n_chunks; c += 2) { chunk(IntC<0>{}, c); chunk(IntC<N2_CH>{}, c + 1); }
  if (c < n_chunks) chunk(IntC<0>{}, c);
  // the ring is dead: the DFT matrix arrives behind the U tile while the branch sums are folded
  const unsigned wb_bytes = (unsigned)(N2 * ncol * sizeof(c32));
  if (tid == 0) { mbar_expect_tx(&bar[N2_RC + 2], wb_bytes); tma_bulk_g2s(WBs, P.WB, wb_bytes, &bar[N2_RC + 2]); }
  constexpr int UP = MV + 1;
  c32 *U = ring;
#pragma unroll
  for (int o = 0; o < NEST_R; o++) reinterpret_cast<u64 *>(U)[(size_t)o * UP + tid] = acc[o];
  __syncthreads();

  // ---- 2. fold the two virtual branches of a branch and N1-point DFTs, in place: V[out][k1 * N2 + n2]
  {
    constexpr int ITEMS = 16 / (N1 * 2);              // 16 N2 items on 2 N1 N2 threads
    float vr[ITEMS][N1], vi[ITEMS][N1];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int it = tid + j * nthr;
      const int out = it / N2, rho = it - out * N2, n2 = n2r[rho];
      float ur[N1], ui[N1];
#pragma unroll
      for (int n1 = 0; n1 < N1; n1++) {
        const c32 *up = U + (size_t)out * UP + (N2 * n1 + N1 * n2) % M;
        const c32 a = up[0], t = up[M];
        ur[n1] = a.re + t.re; ui[n1] = a.im + t.im;
      }
      if constexpr (N1 == 1) { vr[j][0] = ur[0]; vi[j][0] = ui[0]; }
      else if constexpr (N1 == 2) {
        vr[j][0] = ur[0] + ur[1]; vi[j][0] = ui[0] + ui[1];
        vr[j][1] = ur[0] - ur[1]; vi[j][1] = ui[0] - ui[1];
      } else {
        const float er = ur[0] + ur[2], ei = ui[0] + ui[2], fr = ur[0] - ur[2], fi = ui[0] - ui[2];
        const float gr = ur[1] + ur[3], gi = ui[1] + ui[3], hr = ur[1] - ur[3], hi = ui[1] - ui[3];
        vr[j][0] = er + gr; vi[j][0] = ei + gi;
        vr[j][1] = fr + hi; vi[j][1] = fi - hr;       // f - j h
        vr[j][2] = er - gr; vi[j][2] = ei - gi;
        vr[j][3] = fr - hi; vi[j][3] = fi + hr;       // f + j h
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int it = tid + j * nthr;
      const int out = it / N2, rho = it - out * N2, n2 = n2r[rho];
#pragma unroll
      for (int k1 = 0; k1 < N1; k1++) U[(size_t)out * UP + k1 * N2 + n2] = c32{vr[j][k1], vi[j][k1]};
    }
    __syncthreads();
  }

  // ---- 3. N2-point DFTs at the channel bins + weighted |Z|^2.  item = (og, cg): outputs og and og + 8, 5 columns
  mbar_wait(&bar[N2_RC + 2], 0);
  const int n_cg = ncol / NEST_NCOL;
  for (int item = tid; item < N2_OG * n_cg; item += nthr) {
    const int og = item & (N2_OG - 1), cg = item / N2_OG;
    const int col0 = cg * NEST_NCOL, k1 = col0 / P.CPC;
    constexpr int NI = N2_OUT / N2_OG;
    float zr[NI][NEST_NCOL], zi[NI][NEST_NCOL];
#pragma unroll
    for (int i = 0; i < NI; i++)
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) { zr[i][j] = 0.0f; zi[i][j] = 0.0f; }
    const c32 *vrow = U + (size_t)og * UP + k1 * N2;
    const c32 *wrow = WBs + col0;
#pragma unroll 5
    for (int n2 = 0; n2 < N2; n2++) {
      c32 v[NI], w[NEST_NCOL];
#pragma unroll
      for (int i = 0; i < NI; i++) v[i] = vrow[(size_t)N2_OG * i * UP + n2];
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) w[j] = wrow[n2 * ncol + j];
#pragma unroll
      for (int i = 0; i < NI; i++)
#pragma unroll
        for (int j = 0; j < NEST_NCOL; j++) {
          zr[i][j] = fmaf(v[i].re, w[j].re, zr[i][j]); zr[i][j] = fmaf(-v[i].im, w[j].im, zr[i][j]);
          zi[i][j] = fmaf(v[i].re, w[j].im, zi[i][j]); zi[i][j] = fmaf(v[i].im, w[j].re, zi[i][j]);
        }
    }
    float e[NEST_NCOL];
#pragma unroll
    for (int j = 0; j < NEST_NCOL; j++) e[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < NI; i++) {
      const int u = i0 + og + N2_OG * i;
      const float wgt = u < P.n_used ? P.weights[u] : 0.0f;
#pragma unroll
      for (int j = 0; j < NEST_NCOL; j++) e[j] = fmaf(wgt, zr[i][j] * zr[i][j] + zi[i][j] * zi[i][j], e[j]);
    }
#pragma unroll
    for (int j = 0; j < NEST_NCOL; j++) epart[og * ncol + col0 + j] = e[j];
  }
  __syncthreads();
  for (int col = tid; col < ncol; col += nthr) {
    float sum = 0.0f;
    for (int og = 0; og < N2_OG; og++) sum += epart[og * ncol + col];
    P.E2[(size_t)tile_index * ncol + col] = sum;
  }
}

size_t nest_smem_bytes(const NestPlan &P) { return P.v2 ? nest2_layout(P).total : nest_layout(P).total; }

#define NEST_DISPATCH(CALL) do { \
    if (P.fold == 2) { if (P.N1 == 4 && P.M == 100) { CALL(4, 100, 2); } else if (P.N1 == 4) { CALL(4, 0, 2); } else if (P.N1 == 2) { CALL(2, 0, 2); } else { CALL(1, 0, 2); } } \
    else if (P.N1 == 4 && P.M == 100) { CALL(4, 100, 1); } else if (P.N1 == 4) { CALL(4, 0, 1); } else if (P.N1 == 2) { CALL(2, 0, 1); } else { CALL(1, 0, 1); } } while (0)

int nest_setup(const NestPlan &P)
{
  if (2 * P.D != P.M || P.N1 * P.N2 != P.M || (P.N1 != 1 && P.N1 != 2 && P.N1 != 4)) return -1;
  if (P.M > 100 || P.q_rows % CH != 0 || P.q_rows < P.Q + CH || P.ncol % NEST_NCOL != 0) return -1;
  if (((CH * P.M * sizeof(float2)) & 15) != 0) return -1;                    // TMA bulk copies move multiples of 16 bytes
  if (P.fold != 1 && P.fold != 2) return -1;
  if (P.fold > 1) {
    const int MV = P.fold * P.M, Qv = (P.Q * P.M + MV - 1) / MV;
    if (P.stride != 2 * P.fold || !P.weights || P.n_used <= 8 || P.q_rows_v < Qv + NEST_R - 1) return -1;
    if (!P.v2 && (P.q_rows_v % CH != 0 || NEST_RUNS_V * MV > 2 * NEST_K * 100)) return -1;   // chunking and launch bound of k_nest
  } else if (P.stride != 1 && !(P.stride == 2 && P.weights && P.n_used > 8)) return -1;
  if ((((long)P.S * sizeof(c32)) & 15) != 0 || (((long)P.fns * sizeof(c32)) & 15) != 0 || ((NEST_TO * P.M * sizeof(c32)) & 15) != 0)
    return -1;                                                             // ... from 16-byte aligned addresses
  if (nest_smem_bytes(P) > 227 * 1024) return -1;
  if (P.v2) {
    if (P.fold != 2 || P.N1 != 4 || P.M != 100 || !P.hq1 || P.q_rows_v % N2_CH != 0) return -1;
    if (((N2_CH * 2 * P.M * sizeof(float)) & 15) != 0 || ((N2_OUT * 2 * P.M * sizeof(c32)) & 15) != 0) return -1;
    if (3 * (nest_smem_bytes(P) + 1024) > 228 * 1024) return -1;            // three blocks per SM
    return cudaFuncSetAttribute((const void *)k_nest2<4, 100>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nest_smem_bytes(P)) == cudaSuccess ? 0 : -1;
  }
  cudaError_t e = cudaSuccess;
#define NEST_OPT(N1_, MT_, F_) e = cudaFuncSetAttribute((const void *)k_nest<N1_, MT_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nest_smem_bytes(P))
  NEST_DISPATCH(NEST_OPT);
#undef NEST_OPT
  return e == cudaSuccess ? 0 : -1;
}

void launch_nest_prerot(const NestPlan &P, const c32 *x, long n_samples, cudaStream_t s)
{
  if (n_samples <= 0) return;
  k_nest_prerot<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>(x, P.xr, n_samples, P.phasor, P.period);
}

bool nest_can_resume(const NestPlan &P)
{
  const int threads = P.v2 ? 2 * P.M : P.fold > 1 ? NEST_RUNS_V * P.fold * P.M : 2 * NEST_K * P.M;
  return threads >= NEST_RESUME_BLK && nest_smem_bytes(P) >= mm_smem_bytes(NEST_RESUME_BLK);
}

void launch_nest(const NestPlan &P, int B, cudaStream_t s, const NestResume *resume)
{
  NestResume R{};
  if (resume && nest_can_resume(P)) R = *resume;
  const dim3 grid((unsigned)(B * P.tiles_per_slot + R.n_blocks));
  const size_t smem = nest_smem_bytes(P);
  if (P.v2) {
    if (R.n_blocks == 0) R.sm_flag = nullptr;
    if (R.sm_flag) cudaMemsetAsync(R.sm_flag, 0, 256 * sizeof(int), s);
    k_nest2<4, 100><<<grid, 2 * P.M, smem, s>>>(P, R);
    const int n = B * P.nch;
    k_nest_reduce<<<(n + 127) / 128, 128, 0, s>>>(P, B);
    return;
  }
  const int threads = P.fold > 1 ? NEST_RUNS_V * P.fold * P.M : 2 * NEST_K * P.M;
#define NEST_RUN(N1_, MT_, F_) k_nest<N1_, MT_, F_><<<grid, threads, smem, s>>>(P, R)
  NEST_DISPATCH(NEST_RUN);
#undef NEST_RUN
  const int n = B * P.nch;
  k_nest_reduce<<<(n + 127) / 128, 128, 0, s>>>(P, B);
}

}  // namespace btb200
