// rx_firdl.cu -- "delay line" exact-order noise FIR for LISTED windows (lazy squelch), D = 50.
// Compiled with --fmad=false like the rest of the bit-exact path.
//
// multi_block::check_snr (reference lib/multi_block.cc:253-275) runs a 20001-tap
// freq_xlating_fir_filter_ccf at decimation D over the slot; out[j] = sum_k x[s0 + j*D + k] * t_c[k],
// k ASCENDING, one rounding per fp32 operation (DESIGN.md "Exactness contract").
//
// The tiled kernel (rx_kernels.cu) reads every (output, tap) input from shared memory and restages
// its whole input span for every tap chunk.  Here
//  * a thread owns two ADJACENT outputs j, j+1 (D samples apart) for all <= 4 channels of the group.
//    The input of output j at tap k + D is the input of output j+1 at tap k, so it is kept in a D-deep
//    register delay line: per tap two LDS.32 (the new sample, re and im planes) and one broadcast LDS.128
//    per channel feed 2 x NC complex MACs.  The two outputs are the two halves of the packed fp32 operations
//    (FMUL2 / FFMA2(-1) / FADD2, each half rounding like the scalar instruction);
//  * inputs live in a shared-memory RING of rows of 100 samples (+1 pad slot: lanes are 100 samples
//    apart, i.e. 101 slots = conflict-free).  A tap chunk needs only KT new samples, fetched with
//    cp.async (LDGSTS: it splits re/im into the two planes) while the current chunk is being computed;
//  * the next chunk's tap bank -- one contiguous run of (c, c, d, d) entries per channel -- is staged by
//    the TMA engine: one `cp.async.bulk` per channel issued by a single thread, completion through an
//    mbarrier (expect_tx / try_wait.parity), double-buffered.
// Summation order per accumulator is untouched, so results are bit-identical to noise_fir_point().
#include "rx_kernels.cuh"
#include "rx_packed.cuh"
#include "rx_tma.cuh"
#include <cstdlib>

namespace btb200 {

namespace {

constexpr int DL_D = 50;              // decimation = delay-line depth (100 Msps configuration)
constexpr int DL_W = 7;               // warps per block
constexpr int DL_NT = DL_W * 32;
constexpr int DL_TJ = 64 * DL_W;      // outputs per block: lane l of warp w owns outputs 64w + 2l, 64w + 2l + 1
constexpr int DL_CG = LAZY_CG;        // channels per group (upper bound; buffer stride)
constexpr int DL_KT = 200;            // taps per chunk (multiple of 2 * DL_D)
constexpr int DL_ROW = 2 * DL_D;      // samples per ring row
constexpr int DL_PITCH = DL_ROW + 1;  // slots per ring row
// rows: the span of a chunk (DL_TJ * D + KT samples) + the chunk in flight
constexpr int DL_M = (DL_TJ * DL_D + 2 * DL_KT) / DL_ROW + 4;

struct DlJob {
  const c32 *x; long n_x;
  const float4 *taps4;   // [nch][N] (c, c, d, d)
  c32 *out;              // [group][n_noise][DL_CG]
  const int *groups;     // [group][1 + DL_CG]: slot, channels (live ones first, -1 = unused)
  int N, S, fns, n_noise, tiles;
};

__device__ __forceinline__ void cp_async4z(void *smem_dst, const void *gsrc, bool valid)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;                       // src-size 0: the slot is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// ring slot of local sample n (relative to the block's first input sample)
__device__ __forceinline__ int dl_slot(int n) { return ((n / DL_ROW) % DL_M) * DL_PITCH + (n % DL_ROW); }

template <int NC>
__device__ __forceinline__ void dl_body(const DlJob &J, float *ra, float *rb, float4 *tbuf, uint64_t *tbar,
                                        const int *ch, long s0, int nj, c32 *outp)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pairidx = 32 * w + lane;                     // outputs 2*pairidx, 2*pairidx + 1
  float da[DL_D], db[DL_D];                              // delay line: inputs of output A for the next D taps
  u64 are[NC], aim[NC];                                  // lo half = output A, hi half = output B
#pragma unroll
  for (int ci = 0; ci < NC; ci++) { are[ci] = 0ull; aim[ci] = 0ull; }

  // samples [n0, n1) of the block's span and the taps of chunk c -> shared memory (asynchronously)
  auto fetch = [&](int n0, int n1, int c) {
    for (int n = n0 + (int)threadIdx.x; n < n1; n += DL_NT) {
      const long g = s0 + n;
      const bool ok = g < J.n_x;
      const float *src = reinterpret_cast<const float *>(J.x + (ok ? g : 0));
      const int e = dl_slot(n);
      cp_async4z(ra + e, src, ok);
      cp_async4z(rb + e, src + 1, ok);
    }
    cp_commit();
    if (threadIdx.x == 0) {
      // tap bank of chunk c, [channel][KT]: one TMA bulk copy per channel
      float4 *tb = tbuf + (c & 1) * (DL_KT * DL_CG);
      const int k0 = c * DL_KT;
      const unsigned kt = (unsigned)((J.N - k0) < DL_KT ? (J.N - k0) : DL_KT);
      fence_proxy_async();
      mbar_expect_tx(&tbar[c & 1], (unsigned)NC * kt * (unsigned)sizeof(float4));
#pragma unroll
      for (int ci = 0; ci < NC; ci++)
        tma_bulk_g2s(tb + ci * DL_KT, J.taps4 + (size_t)ch[ci] * J.N + k0, kt * (unsigned)sizeof(float4), &tbar[c & 1]);
    }
  };

  const int nchunks = (J.N + DL_KT - 1) / DL_KT;
  fetch(0, DL_TJ * DL_D + DL_KT, 0);
  cp_wait_all();
  __syncthreads();
  {
    // fill the delay line: inputs of output A at taps 0..D-1 = samples 100*pairidx + kp
    const float *pa = ra + pairidx * DL_PITCH, *pb = rb + pairidx * DL_PITCH;
#pragma unroll
    for (int kp = 0; kp < DL_D; kp++) { da[kp] = pa[kp]; db[kp] = pb[kp]; }
  }
  for (int c = 0; c < nchunks; c++) {
    const int k0 = c * DL_KT;
    const int kt = (J.N - k0) < DL_KT ? (J.N - k0) : DL_KT;
    if (c + 1 < nchunks) {
      const int f = DL_TJ * DL_D + (c + 1) * DL_KT;      // frontier: samples below f are (being) loaded
      fetch(f, f + DL_KT, c + 1);
    }
    const float4 *tb = tbuf + (c & 1) * (DL_KT * DL_CG);
    mbar_wait(&tbar[c & 1], (unsigned)(c >> 1) & 1u);      // the taps of this chunk have landed
    const int nseg = kt / DL_D, rem = kt - nseg * DL_D;
    for (int u = 0; u < nseg; u++) {
      // new samples: output B at taps k0 + 50u + kp -> local sample 100*pairidx + 50 + k0 + 50u + kp
      const int q = DL_D + k0 + DL_D * u;
      int row = pairidx + q / DL_ROW;
      row = row >= DL_M ? row - DL_M : row;
      row = row >= DL_M ? row - DL_M : row;
      const int e = row * DL_PITCH + (q % DL_ROW);
      const float *pa = ra + e, *pb = rb + e;
      const float4 *tp = tb + DL_D * u;
#pragma unroll
      for (int kp = 0; kp < DL_D; kp++) {
        const float an = pa[kp], bn = pb[kp];
        const u64 Va = pk_pack(da[kp], an), Vb = pk_pack(db[kp], bn);
        const u64 Vn = pk_neg(Vb);                                     // (-b) * c = -(b * c) exactly
#pragma unroll
        for (int ci = 0; ci < NC; ci++) {
          const ulonglong2 T = *reinterpret_cast<const ulonglong2 *>(tp + ci * DL_KT + kp);   // (c, c), (d, d)
          const u64 pr = pk_xsubp(pk_mul(Va, T.x), pk_mul(Vb, T.y));   // a*c - b*d
          const u64 pi = pk_xsubp(pk_mul(Va, T.y), pk_mul(Vn, T.x));   // a*d - (-b)*c: the roundings of a*d + b*c
          are[ci] = pk_add(are[ci], pr);
          aim[ci] = pk_add(aim[ci], pi);
        }
        da[kp] = an; db[kp] = bn;
      }
    }
    // tail (N is not a multiple of D; only in the last chunk): both inputs straight from the ring
    for (int kp = 0; kp < rem; kp++) {
      const int k = k0 + DL_D * nseg + kp;
      const int eA = dl_slot(DL_ROW * pairidx + k), eB = dl_slot(DL_ROW * pairidx + DL_D + k);
      const u64 Va = pk_pack(ra[eA], ra[eB]), Vb = pk_pack(rb[eA], rb[eB]), Vn = pk_neg(Vb);
#pragma unroll
      for (int ci = 0; ci < NC; ci++) {
        const ulonglong2 T = *reinterpret_cast<const ulonglong2 *>(tb + ci * DL_KT + DL_D * nseg + kp);
        const u64 pr = pk_xsubp(pk_mul(Va, T.x), pk_mul(Vb, T.y));
        const u64 pi = pk_xsubp(pk_mul(Va, T.y), pk_mul(Vn, T.x));
        are[ci] = pk_add(are[ci], pr);
        aim[ci] = pk_add(aim[ci], pi);
      }
    }
    cp_wait_all();
    __syncthreads();
  }
  const int jA = 2 * pairidx;
#pragma unroll
  for (int ci = 0; ci < NC; ci++) {
    if (jA < nj) outp[(long)jA * DL_CG + ci] = c32{pk_lo(are[ci]), pk_lo(aim[ci])};
    if (jA + 1 < nj) outp[(long)(jA + 1) * DL_CG + ci] = c32{pk_hi(are[ci]), pk_hi(aim[ci])};
  }
}

__global__ void __launch_bounds__(DL_NT, 1) k_fir_dl(DlJob J)
{
  extern __shared__ __align__(16) unsigned char dl_smem[];
  __shared__ int s_ch[DL_CG];
  __shared__ __align__(8) uint64_t s_tbar[2];          // mbarriers of the two tap buffers
  if (threadIdx.x == 0) {
    mbar_init(&s_tbar[0], 1);
    mbar_init(&s_tbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  const int q = blockIdx.x / J.tiles, jt = blockIdx.x - q * J.tiles;
  const int j0 = jt * DL_TJ;
  const int nj = (J.n_noise - j0) < DL_TJ ? (J.n_noise - j0) : DL_TJ;
  const int *g = J.groups + (size_t)q * (1 + DL_CG);
  const int b = g[0];
  if (threadIdx.x < DL_CG) s_ch[threadIdx.x] = g[1 + threadIdx.x];
  int nc = 0;
#pragma unroll
  for (int i = 0; i < DL_CG; i++) nc += (g[1 + i] >= 0) ? 1 : 0;
  __syncthreads();
  float4 *tbuf = reinterpret_cast<float4 *>(dl_smem);                      // [2][<=CG][KT]
  float *ra = reinterpret_cast<float *>(tbuf + 2 * DL_KT * DL_CG);         // [M][PITCH] re plane
  float *rb = ra + DL_M * DL_PITCH;                                        // [M][PITCH] im plane
  const long s0 = (long)b * J.S + J.fns + (long)j0 * DL_D;
  c32 *outp = J.out + ((long)q * J.n_noise + j0) * DL_CG;
  switch (nc) {
    case 1: dl_body<1>(J, ra, rb, tbuf, s_tbar, s_ch, s0, nj, outp); break;
    case 2: dl_body<2>(J, ra, rb, tbuf, s_tbar, s_ch, s0, nj, outp); break;
    case 3: dl_body<3>(J, ra, rb, tbuf, s_tbar, s_ch, s0, nj, outp); break;
    case 4: dl_body<4>(J, ra, rb, tbuf, s_tbar, s_ch, s0, nj, outp); break;
    default: break;
  }
}

constexpr size_t DL_SMEM = (size_t)2 * DL_KT * DL_CG * sizeof(float4) + (size_t)DL_M * DL_PITCH * sizeof(c32);

}  // namespace

bool noise_fir_dl_supported(const Geom &G) { return G.D == DL_D; }

// groups: [n_groups][1 + LAZY_CG], live channels first.  NzL: [group][n_noise][LAZY_CG].
// taps4: [nch][Nn] float4 (c, c, d, d) = the rotated noise taps with both parts duplicated.
int launch_noise_fir_dl(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                        c32 *NzL, cudaStream_t s)
{
  if (n_groups <= 0) return 0;
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    if (cudaFuncSetAttribute((const void *)k_fir_dl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DL_SMEM) !=
        cudaSuccess)
      return -1;
    done[dev & 63] = true;
  }
  DlJob J{};
  J.x = W.x; J.n_x = (long)(W.B - 1) * G.S + G.H; J.taps4 = reinterpret_cast<const float4 *>(T.noise_taps4);
  J.out = NzL; J.groups = groups;
  J.N = G.Nn; J.S = G.S; J.fns = G.fns; J.n_noise = G.n_noise; J.tiles = (G.n_noise + DL_TJ - 1) / DL_TJ;
  k_fir_dl<<<(unsigned)n_groups * (unsigned)J.tiles, DL_NT, DL_SMEM, s>>>(J);
  return 0;
}

}  // namespace btb200
