// rx_kernels.cu -- sm_100a kernels of the receive path.  Compiled with
// --fmad=false (see rx_math.cuh): every fp32 operation rounds once.
#include "rx_kernels.cuh"
#include <cstdio>
#include <cstdlib>

#include "rx_packed.cuh"
#include "rx_tma.cuh"
#include "rx_mm.cuh"

namespace btb200 {

// ===========================================================================
// v1 baseline kernels: one thread per output, straight from the bodies.
// ===========================================================================

__global__ void k_chan_fir_v1(Geom G, const c32 *__restrict__ x, const c32 *__restrict__ rt,
                              c32 *__restrict__ Y, long Gtot)
{
  // c fastest so that a warp writes contiguous Y[g][c..]
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Gtot * G.nch) return;
  const long g = idx / G.nch;
  const int c = (int)(idx - g * G.nch);
  Y[idx] = chan_fir_point(G, x, rt + (long)c * G.Nc, g);
}

__global__ void k_noise_fir_v1(Geom G, const c32 *__restrict__ x, const c32 *__restrict__ rt,
                               c32 *__restrict__ Nz, int B)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long per_b = (long)G.n_noise * G.nch;
  if (idx >= per_b * B) return;
  const int b = (int)(idx / per_b);
  const long r = idx - (long)b * per_b;
  const int j = (int)(r / G.nch);
  const int c = (int)(r - (long)j * G.nch);
  Nz[idx] = noise_fir_point(G, x, rt + (long)c * G.Nn, b, j);
}

// One thread per channel-window: sequential fp64 accumulation in index order.
__global__ void k_energy(Geom G, DevBatch W, int device_gate)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W.B * G.nch) return;
  const int b = idx / G.nch, c = idx - b * G.nch;
  double on, off;
  window_energy(G, W.Y, W.Nz, W.phc, W.phn, b, c, b * W.bp_stride, &on, &off);
  W.energy[idx] = on;
  W.noise[idx] = off;
  if (device_gate) {
    // Device-side squelch with a guard band: CUDA's log10 is within 2 ulp of
    // libm's, the guard is 1e-6 dB.  Windows inside the band are processed and
    // the host settles them with libm (exact reference arithmetic).
    const double snr = 10.0 * log10(on / off);
    W.pass[idx] = (snr >= G.squelch_db - 1e-6) ? 1 : 0;
  }
}

__global__ void k_demod(Geom G, DevBatch W, const float *__restrict__ atan_tab)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int bc = blockIdx.y;
  if (i >= G.n_dem) return;
  if (!W.pass[bc]) return;
  const int b = bc / G.nch, c = bc - b * G.nch;
  W.dem[(long)bc * G.n_dem_pad + i] = window_demod_point(G, W.Y, W.phc, atan_tab, b, c, b * W.bp_stride, i);
}

// stateless: one thread per channel-window, constructor state every time
__global__ void k_mm_stateless(Geom G, DevBatch W, const float *__restrict__ mmse)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W.B * G.nch) return;
  if (!W.pass[idx]) { W.nsym[idx] = 0; return; }
  MmState st{G.mu0, G.mm.omega_mid, 0.0f};
  W.nsym[idx] = window_mm(G, mmse, W.dem + (long)idx * G.n_dem_pad, st,
                          W.bits + (long)idx * G.bw, W.soft ? W.soft + (long)idx * G.n_dem_pad : nullptr);
}

// chained: the reference's single serial chain over (slot, channel)
__global__ void k_mm_chained(Geom G, DevBatch W, const float *__restrict__ mmse)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  MmState st = *W.mm_state;
  for (int idx = 0; idx < W.B * G.nch; idx++) {
    if (!W.pass[idx]) { W.nsym[idx] = 0; continue; }
    W.nsym[idx] = window_mm(G, mmse, W.dem + (long)idx * G.n_dem_pad, st,
                            W.bits + (long)idx * G.bw, W.soft ? W.soft + (long)idx * G.n_dem_pad : nullptr);
  }
  *W.mm_state = st;
}

// multi_hopper channel loop (lib/multi_hopper_impl.cc:93-137, 152-209) on the shared chained state:
// channels first..first+n-1 of slot 0 in order; per channel ONE sniff_ac over min(nsym-68, 625) lags;
// the loop ends after the first channel whose packet has LAP == stop_lap and a header.
// res[c] = {processed, nsym, ac_index, lap}
__global__ void k_mm_chained_list(Geom G, DevBatch W, const float *__restrict__ mmse, const uint64_t *__restrict__ ac_lut,
                                  BchDev bch, int first, int n, uint32_t stop_lap, int4 *__restrict__ res)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  MmState st = *W.mm_state;
  bool stopped = false;
  for (int c = 0; c < G.nch; c++) {
    int4 r = make_int4(0, 0, -1, 0);
    if (c >= first && c < first + n && !stopped) {
      r.x = 1;
      if (W.pass[c]) {
        uint32_t *row = W.bits + (long)c * G.bw;
        const int nsym = window_mm(G, mmse, W.dem + (long)c * G.n_dem_pad, st, row,
                                   W.soft ? W.soft + (long)c * G.n_dem_pad : nullptr);
        r.y = nsym;
        if (nsym >= 68) {
          const int latest = (nsym - 68 < 625) ? nsym - 68 : 625;
          for (int lag = 0; lag < latest; lag++) {
            uint64_t lo; uint32_t hi, lap;
            bits_window(row, lag, &lo, &hi);
            int ne;
            if ((G.search & 4) ? br_lag_test_bch(bch, lo, hi, &lap, &ne) : br_lag_test(ac_lut, lo, hi, &lap)) { r.z = lag; r.w = (int)lap; break; }
          }
          if (r.z >= 0 && (uint32_t)r.w == stop_lap) {
            const int len = (nsym - r.z) < 3125 ? (nsym - r.z) : 3125;
            if (header_present_bits(row, r.z, len)) stopped = true;
          }
        }
      }
    }
    W.nsym[c] = r.y;
    res[c] = r;
  }
  *W.mm_state = st;
}

struct HitEmitter {
  const Geom &G; const DevBatch &W; int b, c, nsym;
  bool relative = false;      // lazy tail: n_symbols is relative to the (not yet known) symbol count; no symbols yet
  int *staged = nullptr;      // device-driven tail: running count of this window's hits (one emitting thread per window)
  __device__ void operator()(int kind, int offset, int n_symbols, uint32_t lap) const
  {
    if (staged) {
      const int i = (*staged)++;
      if (i < TAIL_MAXW) {
        DevHit h;
        h.b = b; h.chi = (int16_t)c; h.kind = (int16_t)kind; h.offset = offset; h.n_symbols = n_symbols;
        h.lap = lap; h.sym_offset = 0; h.sym_count = 0;
        W.tail.stage[((long)b * G.nch + c) * TAIL_MAXW + i] = h;
      }
      return;
    }
    const unsigned slot = atomicAdd(W.hit_count, 1u);
    if (slot >= W.hit_cap) return;
    int cnt = n_symbols < 3125 ? n_symbols : 3125;
    if (cnt < 0 || relative) cnt = 0;
    const unsigned long long so = atomicAdd(W.arena_used, (unsigned long long)cnt);
    DevHit h;
    h.b = b; h.chi = (int16_t)c; h.kind = (int16_t)kind; h.offset = offset; h.n_symbols = n_symbols;
    h.lap = lap; h.sym_offset = so; h.sym_count = (so + cnt <= W.arena_cap) ? (uint32_t)cnt : 0u;
    W.hits[slot] = h;
  }
};

__global__ void k_search_v1(Geom G, DevTables T, DevBatch W)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W.B * G.nch) return;
  const int nsym = W.nsym[idx];
  if (nsym <= 0) return;
  const int b = idx / G.nch, c = idx - b * G.nch;
  HitEmitter em{G, W, b, c, nsym};
  window_search(G, T.ac_lut, T.le_hdr_lut, W.bits + (long)idx * G.bw, nsym, T.le_index[c], T.le_white[c], em, &T.bch);
}

// copy the symbols of every hit into the arena, one byte per symbol
__global__ void k_gather(Geom G, DevBatch W)
{
  unsigned n = *W.hit_count;
  if (n > W.hit_cap) n = W.hit_cap;
  for (unsigned h = blockIdx.x; h < n; h += gridDim.x) {
    const DevHit hit = W.hits[h];
    const uint32_t *row = W.bits + ((long)hit.b * G.nch + hit.chi) * G.bw;
    uint8_t *dst = W.arena + hit.sym_offset;
    for (unsigned i = threadIdx.x; i < hit.sym_count; i += blockDim.x) {
      const int s = hit.offset + (int)i;
      dst[i] = (row[s >> 5] >> (s & 31)) & 1;
    }
  }
}

// ===========================================================================
// Tiled exact-order FIR (channel DDC and noise DDC share it).
//
// out[j][c] = sum_{k<N} x[s0 + j*D + k] * t_c[k], k ASCENDING per output -- the
// oracle's summation order, so results are bit-identical to the v1 kernels.
//
// Mapping (DESIGN.md "FIR kernel"): a block owns 16 channels x TJ = 2*R*W
// outputs.  Lane = (channel in group: 16) x (output half: 2); each thread keeps R
// outputs in registers.  Per tap the warp issues ONE conflict-free LDS.64 for
// the 16 taps (k-major, channel-minor layout) and R two-address multicast LDS.64
// for the inputs -> R+1 shared-memory wavefronts per 8*R FP32 instructions, so
// the kernel is FP32-issue bound, not shared-memory bound.  The input span and the
// tap bank of the group are staged once per tap chunk.
// ===========================================================================
struct FirJob {
  const c32 *x; long n_x;
  const c32 *taps;      // mode 0/1: [ngroups][N][16] group-interleaved; mode 2: [nch][N] per channel
  const float4 *taps4;  // packed kernel, modes 0/1, optional: [ngroups][N][16] (c, c, d, d) -- staged by TMA bulk copy
  c32 *out;             // mode 0/1: rows of nch, channel fastest; mode 2: [group][n_noise][CG]
  int N, D, nch, KT;
  int mode;             // 0: channel grid (tile t = outputs t*TJ..), 1: noise, every slot x channel group,
                        // 2: noise, listed (slot, <=CG channels) groups only (lazy squelch)
  long Gtot; int fcs;
  int S, fns, n_noise, tiles_per_slot;
  const int *groups;    // mode 2: [ngroups][1+CG] = slot, channel indices (-1 = unused)
  int group_fast;       // modes 0/1: blockIdx.x = channel group, blockIdx.y = tile, so the groups that share an
                        // input span run back to back and the span is read from DRAM once (L2 serves the rest)
};

template <int CG, int R, int W, int MINB = ((CG == 16) ? 1 : 2)>
__global__ void __launch_bounds__(W * 32, MINB) k_fir_tiled(FirJob J)
{
  constexpr int NH = 32 / CG;                 // output sub-groups per warp
  constexpr int TJ = NH * R * W;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  c32 *ts = reinterpret_cast<c32 *>(smem_raw);          // [KT][CG]
  c32 *xs = ts + (size_t)J.KT * CG;                     // [(TJ-1)*D + KT]
  __shared__ int s_ch[CG];
  long s0;
  int nj, ostride;
  c32 *outp;
  const unsigned bx = J.group_fast ? blockIdx.y : blockIdx.x;     // tile index
  const unsigned by = J.group_fast ? blockIdx.x : blockIdx.y;     // channel group
  if (J.mode == 0) {
    const long g0 = (long)bx * TJ;
    const long left = J.Gtot - g0;
    nj = left < TJ ? (int)left : TJ;
    s0 = J.fcs + g0 * J.D;
    outp = J.out + g0 * J.nch;
    ostride = J.nch;
  } else {
    const int q = bx / J.tiles_per_slot, jt = bx - q * J.tiles_per_slot;
    const int j0 = jt * TJ;
    nj = (J.n_noise - j0) < TJ ? (J.n_noise - j0) : TJ;
    int b = q;
    if (J.mode == 2) {
      b = J.groups[q * (1 + CG)];
      if (threadIdx.x < CG) s_ch[threadIdx.x] = J.groups[q * (1 + CG) + 1 + threadIdx.x];
      outp = J.out + ((long)q * J.n_noise + j0) * CG;
      ostride = CG;
    } else {
      outp = J.out + ((long)b * J.n_noise + j0) * J.nch;
      ostride = J.nch;
    }
    s0 = (long)b * J.S + J.fns + (long)j0 * J.D;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, cg = lane % CG, h = lane / CG;
  const int jj0 = NH * w + h;                // r-th output of this thread: jj0 + r*NH*W (interleaved, so the
  float ar[R], ai[R];                        // sub-groups of a warp read different banks)
#pragma unroll
  for (int r = 0; r < R; r++) { ar[r] = 0.0f; ai[r] = 0.0f; }
  const int rstride = NH * W * J.D;
  for (int k0 = 0; k0 < J.N; k0 += J.KT) {
    const int kt = (J.N - k0) < J.KT ? (J.N - k0) : J.KT;
    __syncthreads();
    if (J.mode == 2) {
      for (int i = threadIdx.x; i < kt * CG; i += W * 32) {
        const int ci = i / kt, k = i - ci * kt;          // consecutive threads -> consecutive k: coalesced
        const int ch = s_ch[ci];
        ts[k * CG + ci] = (ch >= 0) ? J.taps[(size_t)ch * J.N + k0 + k] : c32{0.0f, 0.0f};
      }
    } else {
      const c32 *tg = J.taps + ((size_t)by * J.N + k0) * CG;
      for (int i = threadIdx.x; i < kt * CG; i += W * 32) ts[i] = tg[i];
    }
    const int sp = (TJ - 1) * J.D + kt;
    const long base = s0 + k0;
    for (int i = threadIdx.x; i < sp; i += W * 32) {
      const long n = base + i;
      xs[i] = (n < J.n_x) ? J.x[n] : c32{0.0f, 0.0f};
    }
    __syncthreads();
    const c32 *tp = ts + cg;
    {
      const c32 *xp = xs + jj0 * J.D;
#pragma unroll 4
      for (int k = 0; k < kt; k++) {
        const c32 t = tp[k * CG];
#pragma unroll
        for (int r = 0; r < R; r++) {
          const c32 v = xp[r * rstride + k];
          cmac(ar[r], ai[r], v.re, v.im, t.re, t.im);
        }
      }
    }
  }
  const int c = (J.mode == 2) ? cg : (int)by * CG + cg;
  const bool live = (J.mode == 2) ? true : (c < J.nch);
  if (live) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int jj = jj0 + r * NH * W;
      if (jj < nj) outp[(long)jj * ostride + c] = c32{ar[r], ai[r]};
    }
  }
}

// ===========================================================================
// Packed-FP32 variant of the tiled FIR (Blackwell FMUL2/FADD2: two fp32 lanes per
// instruction).  tools/ubench_f32x2.cu measures the 2-register-operand packed forms
// (mul.rn.f32x2, add/sub.rn.f32x2) at the full issue rate on B200, i.e. twice the
// scalar FMUL/FADD throughput, while each half still rounds exactly like the
// scalar instruction -- so the oracle's one-rounding-per-operation order is kept.
//
// A thread pairs output r with output r+R/2: the staged input entry e holds
// (re[e], re[e+DELTA], im[e], im[e+DELTA]) with DELTA = the sample distance between
// the two outputs, so ONE LDS.128 feeds both; the tap entry is (c, c, d, d).
//   RE pair: (a*c) - (b*d)   IM pair: (a*d) + (b*c)   then acc += ...   = 8 packed
// instructions per 2 complex MACs (scalar: 16).
// ===========================================================================
// (pk_mul / pk_add / pk_xsubp / pk_neg: rx_packed.cuh, with the note on ptxas contraction)
// DT > 0: the decimation is the compile-time constant DT and the staged span is SKEWED -- one pad slot per DT
// samples (sample i sits at i + i/DT) -- so that the NH outputs a warp reads at once, DT samples apart, are
// DT+1 slots apart: with DT = 50 the 16 addresses of a 2-channel group then spread over all eight 16-byte bank
// groups (2 wavefronts per load instead of 4; the kernel is shared-memory bound otherwise: ncu in profiles/).
template <int CG, int R, int W, int MINB = 1, int DT = 0>
__global__ void __launch_bounds__(W * 32, MINB) k_fir_packed(FirJob J)
{
  constexpr int NH = 32 / CG, TJ = NH * R * W, RP = R / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *ts = reinterpret_cast<float4 *>(smem_raw);     // [KT][CG]  (c, c, d, d)
  float4 *xs = ts + (size_t)J.KT * CG;                   // [HS]      (re[e], re[e+DELTA], im[e], im[e+DELTA])
  __shared__ int s_ch[CG];
  __shared__ __align__(8) uint64_t s_tbar;               // mbarrier of the TMA tap-bank copies
  const bool tma_taps = (J.taps4 != nullptr) && (J.mode != 2);
  unsigned tphase = 0;
  if (tma_taps && threadIdx.x == 0) {
    mbar_init(&s_tbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  long s0;
  int nj, ostride;
  c32 *outp;
  const unsigned bx = J.group_fast ? blockIdx.y : blockIdx.x;     // tile index
  const unsigned by = J.group_fast ? blockIdx.x : blockIdx.y;     // channel group
  if (J.mode == 0) {
    const long g0 = (long)bx * TJ;
    const long left = J.Gtot - g0;
    nj = left < TJ ? (int)left : TJ;
    s0 = J.fcs + g0 * J.D;
    outp = J.out + g0 * J.nch;
    ostride = J.nch;
  } else {
    const int q = bx / J.tiles_per_slot, jt = bx - q * J.tiles_per_slot;
    const int j0 = jt * TJ;
    nj = (J.n_noise - j0) < TJ ? (J.n_noise - j0) : TJ;
    int b = q;
    if (J.mode == 2) {
      b = J.groups[q * (1 + CG)];
      if (threadIdx.x < CG) s_ch[threadIdx.x] = J.groups[q * (1 + CG) + 1 + threadIdx.x];
      outp = J.out + ((long)q * J.n_noise + j0) * CG;
      ostride = CG;
    } else {
      outp = J.out + ((long)b * J.n_noise + j0) * J.nch;
      ostride = J.nch;
    }
    s0 = (long)b * J.S + J.fns + (long)j0 * J.D;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, cg = lane % CG, h = lane / CG;
  const int jj0 = NH * w + h;
  const int rstride = NH * W * J.D;            // samples between consecutive outputs of a thread
  const int delta = RP * rstride;              // samples between the two outputs of a pair
  const int hs_base = (RP * NH * W - 1) * J.D; // entries = hs_base + kt
  u64 are[RP], aim[RP];
#pragma unroll
  for (int r = 0; r < RP; r++) { are[r] = 0ull; aim[r] = 0ull; }
  for (int k0 = 0; k0 < J.N; k0 += J.KT) {
    const int kt = (J.N - k0) < J.KT ? (J.N - k0) : J.KT;
    __syncthreads();
    if (J.mode == 2) {
      for (int i = threadIdx.x; i < kt * CG; i += W * 32) {
        const int ci = i / kt, k = i - ci * kt;
        const int ch = s_ch[ci];
        const c32 t = (ch >= 0) ? J.taps[(size_t)ch * J.N + k0 + k] : c32{0.0f, 0.0f};
        ts[k * CG + ci] = make_float4(t.re, t.re, t.im, t.im);
      }
    } else if (tma_taps) {
      // the chunk's tap bank is one contiguous run of (c, c, d, d) entries: a single TMA bulk copy, issued before
      // the input span is staged by the threads and waited for after it
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(&s_tbar, (unsigned)(kt * CG) * (unsigned)sizeof(float4));
        tma_bulk_g2s(ts, J.taps4 + ((size_t)by * J.N + k0) * CG, (unsigned)(kt * CG) * (unsigned)sizeof(float4), &s_tbar);
      }
    } else {
      const c32 *tg = J.taps + ((size_t)by * J.N + k0) * CG;
      for (int i = threadIdx.x; i < kt * CG; i += W * 32) { const c32 t = tg[i]; ts[i] = make_float4(t.re, t.re, t.im, t.im); }
    }
    const int hs = hs_base + kt;
    const long base = s0 + k0;
    for (int i = threadIdx.x; i < hs; i += W * 32) {
      const long n0 = base + i, n1 = n0 + delta;
      const c32 v0 = (n0 < J.n_x) ? J.x[n0] : c32{0.0f, 0.0f};
      const c32 v1 = (n1 < J.n_x) ? J.x[n1] : c32{0.0f, 0.0f};
      xs[DT > 0 ? i + i / (DT > 0 ? DT : 1) : i] = make_float4(v0.re, v1.re, v0.im, v1.im);
    }
    if (tma_taps) { mbar_wait(&s_tbar, tphase); tphase ^= 1u; }
    __syncthreads();
    const float4 *tp = ts + cg;
    if (DT > 0) {
      constexpr int D1 = DT + 1;
      const float4 *xq = xs + jj0 * D1;
      const int rs = NH * W * D1;
      int k = 0;
      for (; k + DT <= kt; k += DT, xq += D1) {            // whole segments: constant trip count
#pragma unroll 5
        for (int kp = 0; kp < DT; kp++) {
          const ulonglong2 T = *reinterpret_cast<const ulonglong2 *>(tp + (k + kp) * CG);
          const u64 ncc = pk_neg(T.x);
#pragma unroll
          for (int r = 0; r < RP; r++) {
            const ulonglong2 V = *reinterpret_cast<const ulonglong2 *>(xq + r * rs + kp);
            const u64 pr = pk_xsubp(pk_mul(V.x, T.x), pk_mul(V.y, T.y));
            const u64 pi = pk_xsubp(pk_mul(V.x, T.y), pk_mul(V.y, ncc));
            are[r] = pk_add(are[r], pr);
            aim[r] = pk_add(aim[r], pi);
          }
        }
      }
      for (int kp = 0; k < kt; k++, kp++) {                 // tail of the last chunk
        const ulonglong2 T = *reinterpret_cast<const ulonglong2 *>(tp + k * CG);
        const u64 ncc = pk_neg(T.x);
#pragma unroll
        for (int r = 0; r < RP; r++) {
          const ulonglong2 V = *reinterpret_cast<const ulonglong2 *>(xq + r * rs + kp);
          const u64 pr = pk_xsubp(pk_mul(V.x, T.x), pk_mul(V.y, T.y));
          const u64 pi = pk_xsubp(pk_mul(V.x, T.y), pk_mul(V.y, ncc));
          are[r] = pk_add(are[r], pr);
          aim[r] = pk_add(aim[r], pi);
        }
      }
    } else {
      const float4 *xp = xs + jj0 * J.D;
#pragma unroll 4
      for (int k = 0; k < kt; k++) {
        const ulonglong2 T = *reinterpret_cast<const ulonglong2 *>(tp + k * CG);     // .x = (c,c)  .y = (d,d)
        const u64 ncc = pk_neg(T.x);                                                   // (-c,-c), exact
#pragma unroll
        for (int r = 0; r < RP; r++) {
          const ulonglong2 V = *reinterpret_cast<const ulonglong2 *>(xp + r * rstride + k);   // .x = (a,a')  .y = (b,b')
          const u64 pr = pk_xsubp(pk_mul(V.x, T.x), pk_mul(V.y, T.y));  // a*c - b*d
          const u64 pi = pk_xsubp(pk_mul(V.x, T.y), pk_mul(V.y, ncc));  // a*d - (b*(-c)) = a*d + b*c, same roundings
          are[r] = pk_add(are[r], pr);
          aim[r] = pk_add(aim[r], pi);
        }
      }
    }
  }
  const int c = (J.mode == 2) ? cg : (int)by * CG + cg;
  if (J.mode == 2 || c < J.nch) {
#pragma unroll
    for (int r = 0; r < RP; r++) {
      const int ja = jj0 + r * NH * W, jb = ja + RP * NH * W;
      if (ja < nj) outp[(long)ja * ostride + c] = c32{pk_lo(are[r]), pk_lo(aim[r])};
      if (jb < nj) outp[(long)jb * ostride + c] = c32{pk_hi(are[r]), pk_hi(aim[r])};
    }
  }
}

static size_t fir_packed_smem(int CG, int R, int W, int D, int KT, bool skew = false)
{
  const size_t hs = (size_t)((R / 2) * (32 / CG) * W - 1) * D + KT;
  return ((size_t)KT * CG + hs + (skew ? hs / D + 2 : 0)) * sizeof(float4);
}

template <int BLK>
__global__ void k_dmm_stateless(Geom G, DevBatch W, const float *__restrict__ mmse_g, const float *__restrict__ atan_g);
struct MmSave;
template <int BLK>
__global__ void k_mm_stateless_v2(Geom G, DevBatch W, const float *__restrict__ mmse_g, const float *__restrict__ demT,
                                  int mode, MmSave *__restrict__ save, const int4 *__restrict__ list, int n_list);

static size_t fir_smem(int CG, int R, int W, int D, int KT)
{
  const size_t span = (size_t)((32 / CG) * R * W - 1) * D + KT;
  return ((size_t)KT * CG + span) * sizeof(c32);
}
static int g_max_smem = 48 * 1024;

// function attributes are per device: remember which devices a kernel has been opted in on
struct OptIn {
  bool done[64] = {};
  bool need()
  {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

int fir_setup(int device)
{
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return -1;
  g_max_smem = v;
  auto opt_in = [&](const void *fn) {
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) return false;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v - (int)fa.sharedSizeBytes) == cudaSuccess;
  };
  if (!opt_in((const void *)k_fir_tiled<16, 8, 16>)) return -1;
  if (!opt_in((const void *)k_fir_tiled<16, 8, 14>)) return -1;
  if (!opt_in((const void *)k_dmm_stateless<64>)) return -1;
  if (!opt_in((const void *)k_mm_stateless_v2<64>)) return -1;
  if (!opt_in((const void *)k_fir_packed<16, 8, 16>)) return -1;
  if (!opt_in((const void *)k_fir_packed<16, 8, 14>)) return -1;
  return 0;
}

// largest tap chunk (multiple of 32, <= N rounded up) whose tile fits in shared memory
static int pick_kt(int CG, int R, int W, int D, int N, int blocks_per_sm)
{
  int kt = (N + 31) & ~31;
  const size_t budget = (size_t)(g_max_smem + 1024) / blocks_per_sm - 2048;   // 1 KB/block reserved + static
  while (kt > 32 && fir_smem(CG, R, W, D, kt) > budget) kt -= 32;
  return kt;
}

// ===========================================================================
// Fused demod + Mueller&Mueller clock recovery + slicer, stateless mode.
// One thread per channel-window, consecutive lanes = consecutive channels of a
// slot, so the Y rows a warp touches are contiguous.  Demod values are produced
// on demand into a 16-deep per-thread ring in shared memory (the M&M loop only
// ever looks 8 samples ahead), so the demod floats never travel through HBM.
// Same arithmetic, in the same order, as window_demod_point + window_mm.
// ===========================================================================
// Y rows and rotator phases are prefetched into per-thread shared-memory rings with
// cp.async (LDGSTS), PFD rows ahead of the clock-recovery position: the loop is a serial
// dependency chain per window, so without the prefetch every step would expose a DRAM
// round trip (Y does not fit in L2).
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int BLK>
__global__ void __launch_bounds__(BLK) k_dmm_stateless(Geom G, DevBatch W, const float *__restrict__ mmse_g,
                                                       const float *__restrict__ atan_g)
{
  constexpr int RD = 64;          // ring depth (rows)
  constexpr int PFD = 48;         // prefetch distance (rows) beyond the 8-sample interpolator window
  constexpr int LAG = 6;          // cp.async groups allowed in flight (advance <= 4 rows/step -> PFD/4 >= LAG)
  extern __shared__ __align__(16) unsigned char dmm_smem[];
  c32 (*ry)[BLK] = reinterpret_cast<c32 (*)[BLK]>(dmm_smem);                         // [RD][BLK]
  c32 (*rp)[BLK] = reinterpret_cast<c32 (*)[BLK]>(dmm_smem + sizeof(c32) * RD * BLK);
  float (*ring)[BLK] = reinterpret_cast<float (*)[BLK]>(dmm_smem + 2 * sizeof(c32) * RD * BLK);   // [16][BLK]
  float *s_mmse = reinterpret_cast<float *>(dmm_smem + 2 * sizeof(c32) * RD * BLK + sizeof(float) * 16 * BLK);
  float *s_atan = s_mmse + 129 * 8;
  for (int i = threadIdx.x; i < 129 * 8; i += BLK) s_mmse[i] = mmse_g[i];
  for (int i = threadIdx.x; i < 257; i += BLK) s_atan[i] = atan_g[i];
  __syncthreads();
  const int idx = blockIdx.x * BLK + threadIdx.x;
  if (idx >= W.B * G.nch) return;
  if (!W.pass[idx]) { W.nsym[idx] = 0; return; }
  const int b = idx / G.nch, c = idx - b * G.nch;
  const c32 *__restrict__ y = W.Y + ((long)b * G.gps) * G.nch + c;
  const c32 *__restrict__ p = W.phc + c;
  uint32_t *__restrict__ bits_row = W.bits + (long)idx * G.bw;
  float *dem_row = W.dem ? W.dem + (long)idx * G.n_dem_pad : nullptr;
  float *soft_row = W.soft ? W.soft + (long)idx * G.n_dem_pad : nullptr;
  MmState st{G.mu0, G.mm.omega_mid, 0.0f};
  unsigned ii = 0;
  int oo = 0;
  const unsigned ni = (unsigned)(G.n_dem - 8);
  int pnext = 0;                 // demod indices [pnext-16, pnext) live in the demod ring
  int pf = 0;                    // rows [pf-RD, pf) are in (or on their way to) the Y/phase rings
  c32 zprev{0.0f, 0.0f};
  uint32_t word = 0;
  const int tid = threadIdx.x;
  bool first = true;
  while (oo < G.n_dem && ii < ni) {
    // 1. keep the prefetch PFD rows ahead (row d feeds demod d and d+1)
    int want = (int)ii + 8 + PFD;
    if (want > G.n_ddc) want = G.n_ddc;
    for (; pf < want; pf++) {
      cp_async8(&ry[pf & (RD - 1)][tid], y + (long)pf * G.nch);
      cp_async8(&rp[pf & (RD - 1)][tid], p + (long)pf * G.nch);
    }
    cp_async_commit();
    if (first) { cp_async_wait<0>(); first = false; } else cp_async_wait<LAG>();
    // 2. demod values up to ii+7 (a backward step of the loop never exceeds the 16-deep ring:
    //    mu + omega + gain_mu*mm_val >= -1 for |demod| <= gain*pi)
    while (pnext < (int)ii + 8) {
      const int d = pnext;
      float val = 0.0f;                                                 // demod_out[0] is never written
      const c32 zc = crot(ry[d & (RD - 1)][tid], rp[d & (RD - 1)][tid]);
      if (d > 0) val = demod_point(s_atan, G.demod_gain, zc, zprev);
      zprev = zc;
      ring[d & 15][tid] = val;
      if (dem_row) dem_row[d] = val;
      pnext++;
    }
    float in8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) in8[k] = ring[(ii + k) & 15][tid];
    const float out = mmse_interp(s_mmse, in8, st.mu);
    if (soft_row) soft_row[oo] = out;
    if (!(out < 0)) word |= 1u << (oo & 31);
    if ((oo & 31) == 31) { bits_row[oo >> 5] = word; word = 0; }
    ii += (unsigned)mm_update(G.mm, st, out);
    oo++;
  }
  cp_async_wait<0>();
  if (oo & 31) bits_row[oo >> 5] = word;
  for (int w = (oo + 31) >> 5; w < G.bw; w++) bits_row[w] = 0;
  W.nsym[idx] = oo;
}

// ===========================================================================
// Stateless pipeline, second generation: demod of every window in parallel (transposed
// [b][i][c] so that clock recovery reads channel-contiguous rows), then one thread per
// window for the inherently serial Mueller & Mueller chain with a cp.async ring on the
// demod floats, then a warp per window for the access-code search.
// ===========================================================================
// block (x: tile of DM_TI output indices, y: slot, z: 32-channel group); warp w handles the run
// i0 + w*DM_RUN .. of DM_RUN consecutive outputs for its 32 channels, carrying the previous
// rotated sample, so each DDC output is rotated once and all loads/stores are channel-contiguous.
constexpr int DM_WARPS = 8, DM_RUN = 16, DM_TI = DM_WARPS * DM_RUN;
__global__ void __launch_bounds__(DM_WARPS * 32) k_demod_all(Geom G, DevBatch W, const float *__restrict__ atan_g,
                                                             float *__restrict__ demT, int i_end)
{
  __shared__ float s_atan[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) s_atan[i] = atan_g[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int c = blockIdx.z * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  if (c >= G.nch) return;
  if (!W.pass[b * G.nch + c]) return;
  const int i0 = blockIdx.x * DM_TI + w * DM_RUN;
  if (i0 >= i_end) return;
  const int i1 = (i0 + DM_RUN < i_end) ? i0 + DM_RUN : i_end;
  const c32 *y = W.Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = W.phc + (long)(b * W.bp_stride) * G.n_ddc * G.nch + c;
  float *d = demT + ((long)b * G.n_dem_pad) * G.nch + c;
  c32 prev{0.0f, 0.0f};
  int i = i0;
  if (i == 0) { d[0] = 0.0f; prev = crot(y[0], p[0]); i = 1; }          // demod_out[0] is never written by the reference
  else prev = crot(y[(long)(i - 1) * G.nch], p[(long)(i - 1) * G.nch]);
#pragma unroll 4
  for (; i < i1; i++) {
    const c32 cur = crot(y[(long)i * G.nch], p[(long)i * G.nch]);
    d[(long)i * G.nch] = demod_point(s_atan, G.demod_gain, cur, prev);
    prev = cur;
  }
}

// demod tail [i_begin, n_dem) of LISTED windows (lazy tail): one thread per output
__global__ void k_demod_list(Geom G, DevBatch W, const float *__restrict__ atan_g, float *__restrict__ demT,
                             const int4 *__restrict__ list, int i_begin)
{
  __shared__ float s_atan[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) s_atan[i] = atan_g[i];
  __syncthreads();
  const int4 it = list[blockIdx.y];
  const int b = it.x, c = it.y;
  const int i = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G.n_dem) return;
  const c32 *y = W.Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = W.phc + c;                                   // stateless: one phase table
  float val = 0.0f;
  if (i > 0) {
    const c32 cur = crot(y[(long)i * G.nch], p[(long)i * G.nch]);
    const c32 prev = crot(y[(long)(i - 1) * G.nch], p[(long)(i - 1) * G.nch]);
    val = demod_point(s_atan, G.demod_gain, cur, prev);
  }
  demT[((long)b * G.n_dem_pad + i) * G.nch + c] = val;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}

// the stateless clock-recovery loop lives in rx_mm.cuh (shared with rx_nest.cu)
template <int BLK>
__global__ void __launch_bounds__(BLK) k_mm_stateless_v2(Geom G, DevBatch W, const float *__restrict__ mmse_g,
                                                         const float *__restrict__ demT, int mode,
                                                         MmSave *__restrict__ save, const int4 *__restrict__ list, int n_list)
{
  extern __shared__ __align__(16) unsigned char mm_smem[];
  mm_stateless_block<BLK>(G, W, mmse_g, demT, mode, save, list, n_list, mm_smem, (int)blockIdx.x);
}

// Access-code search, one warp per channel-window (lib/multi_sniffer_impl.cc:107-148 +
// lib/packet_impl.cc:247-268, 471-510, 1452-1527): the 32 lanes test 32 consecutive lags of
// the packed symbol row (64-bit sliding window, preamble/Barker gate, affine sync-word LUT,
// popcount threshold), warp ballots collect one flag per lag, then the first-hit / skip
// rule of the reference is replayed on the flag words.
__global__ void k_search_warp(Geom G, DevTables T, DevBatch W)
{
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= W.B * G.nch) return;
  const int nsym = W.nsym[idx];
  const bool staged = W.tail.stage != nullptr;
  if (nsym <= 0) { if (staged && lane == 0) W.tail.cnt[idx] = 0; return; }
  const int b = idx / G.nch, c = idx - b * G.nch;
  const uint32_t *__restrict__ row = W.bits + (long)idx * G.bw;
  const uint32_t wl = row[lane < G.bw ? lane : 0];           // words 0..31 cover lags 0..(625+72)
  const int le_idx = T.le_index[c];
  const uint32_t le_white = T.le_white[c];
  uint32_t br_mask = 0, le_mask = 0;                         // lane q keeps the flags of lags 32q..32q+31
  for (int q = 0; q < 20; q++) {
    const uint32_t w0 = __shfl_sync(0xffffffffu, wl, q), w1 = __shfl_sync(0xffffffffu, wl, q + 1);
    const uint32_t w2 = __shfl_sync(0xffffffffu, wl, q + 2), w3 = __shfl_sync(0xffffffffu, wl, q + 3);
    const uint64_t a = w0 | ((uint64_t)w1 << 32), bb = w2 | ((uint64_t)w3 << 32);
    const uint64_t lo = lane ? (a >> lane) | (bb << (64 - lane)) : a;
    const uint32_t hi = (uint32_t)(bb >> lane) & 0xff;
    const int lag = 32 * q + lane;
    uint32_t lap;
    int bch_err;
    const bool fb = (G.search & 1) && lag < 625 &&
                    ((G.search & 4) ? br_lag_test_bch(T.bch, lo, hi, &lap, &bch_err) : br_lag_test(T.ac_lut, lo, hi, &lap));
    const bool fl = (G.search & 2) && le_idx >= 0 && lag < 625 && le_lag_test(T.le_hdr_lut, lo, le_white, le_idx >= 37);
    const uint32_t mb = __ballot_sync(0xffffffffu, fb), ml = __ballot_sync(0xffffffffu, fl);
    if (lane == q) { br_mask = mb; le_mask = ml; }
  }
  // replay of the search loops (warp-uniform control flow; lane 0 emits)
  // lazy tail: the rows hold the first G.sym_target symbols; the true count is >= 1386 for this geometry,
  // so both limits are 625, and symbol counts are emitted RELATIVE to it (the host adds the true count)
  const int nsym_eff = G.early ? 0 : nsym;
  int n_staged = 0;
  HitEmitter em{G, W, b, c, nsym, G.early != 0, staged ? &n_staged : nullptr};
  int len = G.early ? (1 << 20) : nsym;
  if (G.search & 1) {
    const int limit0 = (len - 68 < 625) ? len - 68 : 625;
    int start = 0;
    while (limit0 - start >= 0) {
      int found = -1;
      for (int wi = start >> 5; wi < 20 && found < 0; wi++) {
        uint32_t m = __shfl_sync(0xffffffffu, br_mask, wi);
        if (wi == (start >> 5)) m &= ~0u << (start & 31);
        if (m) found = wi * 32 + __ffs(m) - 1;
      }
      if (found < 0 || found >= limit0) break;
      if (lane == 0) {
        uint64_t lo; uint32_t hi;
        bits_window(row, found, &lo, &hi);
        if (G.search & 4) {
          uint32_t lap2 = 0; int ne = 0;
          br_lag_test_bch(T.bch, lo, hi, &lap2, &ne);                 // the corrected LAP and the corrected-bit count
          em(0, found, nsym_eff - found, (lap2 & 0xffffff) | ((uint32_t)ne << 24));
        } else
          em(0, found, nsym_eff - found, ((uint32_t)(lo >> 38) & 0xffffff) | ((uint32_t)br_lag_errors(T.ac_lut, lo, hi) << 24));
      }
      start = found + 68;
    }
    len = (G.early ? (1 << 20) : nsym) - start;
  }
  if ((G.search & 2) && le_idx >= 0) {
    const int limit0 = (len - 68 < 625) ? len - 68 : 625;
    const int len_le = G.early ? len - (1 << 20) : len;          // relative to the true symbol count in lazy-tail mode
    int start = 0;
    while (limit0 - start >= 0) {
      int found = -1;
      for (int wi = start >> 5; wi < 20 && found < 0; wi++) {
        uint32_t m = __shfl_sync(0xffffffffu, le_mask, wi);
        if (wi == (start >> 5)) m &= ~0u << (start & 31);
        if (m) found = wi * 32 + __ffs(m) - 1;
      }
      if (found < 0 || found >= limit0) break;
      if (lane == 0) {
        uint64_t lo; uint32_t hi;
        bits_window(row, found, &lo, &hi);
        em(1, found, len_le - found, (uint32_t)(lo >> 8));
      }
      start = found + 40;
    }
  }
  if (staged && lane == 0) W.tail.cnt[idx] = n_staged < TAIL_MAXW ? n_staged : TAIL_MAXW;
}

// ---- device-driven tail ------------------------------------------------------------------------------------------
// positions of every window's hits in the final list (= the reference's visiting order: slot, channel, then the order
// the search loops produced them) and the list of windows that have hits; one block, contiguous chunks per thread
__global__ void __launch_bounds__(1024) k_tail_scan(Geom G, DevBatch W)
{
  // one block, a contiguous chunk of windows per WARP: coalesced reads, warp totals, a 32-entry prefix, then a second
  // coalesced pass with warp scans (shuffles / ballots) that writes positions and the compacted window list
  __shared__ int s_h[32], s_w[32];
  constexpr unsigned FULL = 0xffffffffu;
  const int n = W.B * G.nch, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per_warp = ((n + 31) / 32 + 31) / 32 * 32;
  const int i0 = wid * per_warp < n ? wid * per_warp : n, i1 = (i0 + per_warp < n) ? i0 + per_warp : n;
  int h = 0, w = 0;
  for (int i = i0 + lane; i < i1; i += 32) { const int c = W.tail.cnt[i]; h += c; w += c > 0; }
  h = __reduce_add_sync(FULL, h); w = __reduce_add_sync(FULL, w);
  if (lane == 0) { s_h[wid] = h; s_w[wid] = w; }
  __syncthreads();
  int hb = 0, wb = 0, ht = 0, wt = 0;                   // hits / windows with hits before this warp's chunk, and in total
  for (int k = 0; k < 32; k++) {
    const int a = s_h[k], b = s_w[k];
    if (k < wid) { hb += a; wb += b; }
    ht += a; wt += b;
  }
  for (int i = i0; i < i1; i += 32) {
    const int idx = i + lane;
    const int c = idx < i1 ? W.tail.cnt[idx] : 0;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
    const unsigned m = __ballot_sync(FULL, c > 0);
    if (idx < i1) {
      W.tail.base[idx] = hb + inc - c;
      if (c > 0) W.tail.list[wb + __popc(m & ((1u << lane) - 1u))] = make_int4(idx / G.nch, idx % G.nch, 0, 0);
    }
    hb += __shfl_sync(FULL, inc, 31);
    wb += __popc(m);
  }
  if (threadIdx.x == 0) { W.hit_count[0] = (unsigned)ht; W.hit_count[1] = (unsigned)wt; *W.tail.n_list = wt; }
}

// the hits of every listed window, in order, with their final symbol counts
__global__ void k_tail_compact(Geom G, DevBatch W)
{
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= *W.tail.n_list) return;
  const int4 it = W.tail.list[l];
  const int key = it.x * G.nch + it.y;
  const int cnt = W.tail.cnt[key], base = W.tail.base[key];
  const int nsym = W.nsym[key];
  for (int i = 0; i < cnt; i++) {
    DevHit h = W.tail.stage[(long)key * TAIL_MAXW + i];
    if (G.early) h.n_symbols += nsym;                   // the search ran on the prefix: counts were relative
    int c = h.n_symbols < 3125 ? h.n_symbols : 3125;
    if (c < 0) c = 0;
    h.sym_count = (uint32_t)c;
    if ((unsigned)(base + i) < W.hit_cap) W.tail.sorted[base + i] = h;
  }
}

// arena layout: exclusive scan of the symbol counts in list order (one block)
__global__ void __launch_bounds__(1024) k_tail_offsets(DevBatch W)
{
  __shared__ unsigned long long s_sum[1024];
  unsigned n = W.hit_count[0];
  if (n > W.hit_cap) n = W.hit_cap;
  const int t = threadIdx.x;
  const unsigned per = (n + 1023) / 1024, i0 = t * per, i1 = (i0 + per < n) ? i0 + per : n;
  unsigned long long sum = 0;
  for (unsigned i = i0; i < i1; i++) sum += W.tail.sorted[i].sym_count;
  s_sum[t] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned long long a = (t >= d) ? s_sum[t - d] : 0;
    __syncthreads();
    s_sum[t] += a;
    __syncthreads();
  }
  unsigned long long off = s_sum[t] - sum;
  for (unsigned i = i0; i < i1; i++) {
    DevHit &h = W.tail.sorted[i];
    h.sym_offset = off;
    if (off + h.sym_count > W.arena_cap) h.sym_count = 0;
    off += h.sym_count ? h.sym_count : 0;
  }
  if (t == 1023) *W.arena_used = s_sum[1023] < W.arena_cap ? s_sum[1023] : W.arena_cap;
}

__global__ void k_gather_sorted(Geom G, DevBatch W)
{
  unsigned n = W.hit_count[0];
  if (n > W.hit_cap) n = W.hit_cap;
  for (unsigned h = blockIdx.x; h < n; h += gridDim.x) {
    const DevHit hit = W.tail.sorted[h];
    const uint32_t *row = W.bits + ((long)hit.b * G.nch + hit.chi) * G.bw;
    uint8_t *dst = W.arena + hit.sym_offset;
    for (unsigned i = threadIdx.x; i < hit.sym_count; i += blockDim.x) {
      const int s = hit.offset + (int)i;
      dst[i] = (row[s >> 5] >> (s & 31)) & 1;
    }
  }
}

// Exact energies of LISTED channel-windows (lazy squelch), one warp per window: the 32 lanes
// rotate and square 32 consecutive outputs, the fp64 sum is then taken strictly in index
// order (every lane carries the same chain) -- bit-identical to window_energy().
__global__ void k_energy_list_warp(Geom G, DevBatch W, const int4 *__restrict__ list, int n_list,
                                   const c32 *__restrict__ NzL, int cgw, double *__restrict__ e_on, double *__restrict__ e_off)
{
  const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (l >= n_list) return;
  const int4 it = list[l];
  const int b = it.x, c = it.y;
  const c32 *y = W.Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = W.phc + c;
  // the loads of the next 32 samples are issued before the serial fp64 chain of the current 32 (their latency,
  // rows 632 bytes apart, would otherwise add to every step of a latency-bound kernel)
  double e = 0.0;
  {
    c32 yv = (lane < G.n_ddc) ? y[(long)lane * G.nch] : c32{0.0f, 0.0f};
    c32 pv = (lane < G.n_ddc) ? p[(long)lane * G.nch] : c32{0.0f, 0.0f};
    for (int i0 = 0; i0 < G.n_ddc; i0 += 32) {
      const float m = (i0 + lane < G.n_ddc) ? mag2(crot(yv, pv)) : 0.0f;
      const int in = i0 + 32 + lane;
      if (in < G.n_ddc) { yv = y[(long)in * G.nch]; pv = p[(long)in * G.nch]; }
      const int cnt = (G.n_ddc - i0) < 32 ? (G.n_ddc - i0) : 32;
      for (int j = 0; j < cnt; j++) e += (double)__shfl_sync(0xffffffffu, m, j);
    }
  }
  if (!NzL) { if (lane == 0) { e_on[l] = e / G.n_ddc; e_off[l] = 0.0; } return; }    // on-channel only
  const c32 *z = NzL + ((long)it.z * G.n_noise) * cgw + it.w;
  const c32 *q = W.phn + c;
  double n = 0.0;
  {
    c32 zv = (lane < G.n_noise) ? z[(long)lane * cgw] : c32{0.0f, 0.0f};
    c32 qv = (lane < G.n_noise) ? q[(long)lane * G.nch] : c32{0.0f, 0.0f};
    for (int j0 = 0; j0 < G.n_noise; j0 += 32) {
      const float m = (j0 + lane < G.n_noise) ? mag2(crot(zv, qv)) : 0.0f;
      const int jn = j0 + 32 + lane;
      if (jn < G.n_noise) { zv = z[(long)jn * cgw]; qv = q[(long)jn * G.nch]; }
      const int cnt = (G.n_noise - j0) < 32 ? (G.n_noise - j0) : 32;
      for (int k = 0; k < cnt; k++) n += (double)__shfl_sync(0xffffffffu, m, k);
    }
  }
  if (lane == 0) { e_on[l] = e / G.n_ddc; e_off[l] = n / G.n_noise; }
}

// every window passes (lazy squelch: the squelch is settled afterwards, exactly, for hit windows only)
__global__ void k_fill_pass(int *pass, int n, int v)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pass[i] = v;
}

// Exact energies of LISTED channel-windows (lazy squelch).  list[l] = {b, chi, group, slot_in_group}.
__global__ void k_energy_list(Geom G, DevBatch W, const int4 *__restrict__ list, int n_list,
                              const c32 *__restrict__ NzL, int cgw, double *__restrict__ e_on, double *__restrict__ e_off)
{
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_list) return;
  const int4 it = list[l];
  const int b = it.x, c = it.y;
  double e = 0.0;
  const c32 *y = W.Y + ((long)b * G.gps) * G.nch + c;
  const c32 *p = W.phc + c;                            // stateless: one table
  for (int i = 0; i < G.n_ddc; i++) e += mag2(crot(y[(long)i * G.nch], p[(long)i * G.nch]));
  e_on[l] = e / G.n_ddc;
  double n = 0.0;
  const c32 *z = NzL + ((long)it.z * G.n_noise) * cgw + it.w;
  const c32 *q = W.phn + c;
  for (int j = 0; j < G.n_noise; j++) n += mag2(crot(z[(long)j * cgw], q[(long)j * G.nch]));
  e_off[l] = n / G.n_noise;
}

static int pick_kt_packed(int CG, int R, int W, int D, int N, int blocks_per_sm = 1, bool skew = false)
{
  int kt = (N + 31) & ~31;
  if (skew) {                     // whole multiples of D per chunk
    const size_t budget = (size_t)(g_max_smem + 1024) / blocks_per_sm - 2048;
    kt = (N / D) * D; if (kt < D) kt = D;
    while (kt > D && fir_packed_smem(CG, R, W, D, kt, true) > budget) kt -= D;
    return kt;
  }
  const size_t budget = (size_t)(g_max_smem + 1024) / blocks_per_sm - 2048;
  while (kt > 32 && fir_packed_smem(CG, R, W, D, kt) > budget) kt -= 32;
  return kt;
}

// ===========================================================================
// launchers
// ===========================================================================
static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

constexpr int CHAN_R = 8, CHAN_W = 16, CHAN_TJ = 2 * CHAN_R * CHAN_W;   // 16 channels x 256 outputs per block

long chan_fir_tiles(const Geom &G, const DevBatch &W)
{
  const long Gtot = (long)(W.B - 1) * G.gps + G.n_ddc;
  return (Gtot + CHAN_TJ - 1) / CHAN_TJ;
}
// input samples (from W.x[0]) the tiles below `tile_end` read
long chan_fir_samples(const Geom &G, long tile_end) { return (long)G.fcs + (tile_end * CHAN_TJ - 1) * G.D + G.Nc; }

// tiles [tile0, tile1) of the channel FIR (tuned kernels): the job is the whole-batch job with its origin moved
void launch_chan_fir_range(const Geom &G, const DevTables &T, const DevBatch &W, int impl, long tile0, long tile1,
                           cudaStream_t s)
{
  const long Gtot = (long)(W.B - 1) * G.gps + G.n_ddc;
  const long g0 = tile0 * CHAN_TJ, g1 = (tile1 * CHAN_TJ < Gtot) ? tile1 * CHAN_TJ : Gtot;
  if (g1 <= g0) return;
  constexpr int R = CHAN_R, Wp = CHAN_W, TJ = CHAN_TJ;
  FirJob J{};
  J.x = W.x + g0 * G.D; J.n_x = (long)(W.B - 1) * G.S + G.H - g0 * G.D; J.taps = T.chan_tg; J.out = W.Y + g0 * G.nch;
  J.taps4 = reinterpret_cast<const float4 *>(T.chan_tg4);
  J.N = G.Nc; J.D = G.D; J.nch = G.nch;
  J.mode = 0; J.Gtot = g1 - g0; J.fcs = G.fcs;
  const unsigned ntile = cdiv(g1 - g0, TJ), ngrp = (unsigned)((G.nch + 15) / 16);
  J.group_fast = ntile <= 65535u;
  dim3 grid(J.group_fast ? ngrp : ntile, J.group_fast ? ntile : ngrp);
  if (impl == IMPL_TILED_SCALAR) {
    J.KT = pick_kt(16, R, Wp, G.D, G.Nc, 1);
    k_fir_tiled<16, R, Wp><<<grid, Wp * 32, fir_smem(16, R, Wp, G.D, J.KT), s>>>(J);
  } else {
    J.KT = pick_kt_packed(16, R, Wp, G.D, G.Nc);
    k_fir_packed<16, R, Wp><<<grid, Wp * 32, fir_packed_smem(16, R, Wp, G.D, J.KT), s>>>(J);
  }
}

void launch_chan_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s)
{
  const long Gtot = (long)(W.B - 1) * G.gps + G.n_ddc;
  if (impl == IMPL_BASELINE) {
    k_chan_fir_v1<<<cdiv(Gtot * G.nch, 128), 128, 0, s>>>(G, W.x, T.chan_rtaps, W.Y, Gtot);
    return;
  }
  launch_chan_fir_range(G, T, W, impl, 0, chan_fir_tiles(G, W), s);
}

void launch_noise_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s)
{
  if (impl == IMPL_BASELINE) {
    const long n = (long)W.B * G.n_noise * G.nch;
    k_noise_fir_v1<<<cdiv(n, 128), 128, 0, s>>>(G, W.x, T.noise_rtaps, W.Nz, W.B);
    return;
  }
  constexpr int R = 8, Wp = 14, TJ = 2 * R * Wp;
  FirJob J{};
  J.x = W.x; J.n_x = (long)(W.B - 1) * G.S + G.H; J.taps = T.noise_tg; J.out = W.Nz;
  J.N = G.Nn; J.D = G.D; J.nch = G.nch;
  J.mode = 1; J.S = G.S; J.fns = G.fns; J.n_noise = G.n_noise; J.tiles_per_slot = (G.n_noise + TJ - 1) / TJ;
  const unsigned ntile = (unsigned)(W.B * J.tiles_per_slot), ngrp = (unsigned)((G.nch + 15) / 16);
  J.group_fast = ntile <= 65535u;
  dim3 grid(J.group_fast ? ngrp : ntile, J.group_fast ? ntile : ngrp);
  if (impl == IMPL_TILED_SCALAR) {
    J.KT = pick_kt(16, R, Wp, G.D, G.Nn < 512 ? G.Nn : 512, 1);
    k_fir_tiled<16, R, Wp><<<grid, Wp * 32, fir_smem(16, R, Wp, G.D, J.KT), s>>>(J);
  } else {
    J.KT = pick_kt_packed(16, R, Wp, G.D, G.Nn < 256 ? G.Nn : 256);
    k_fir_packed<16, R, Wp><<<grid, Wp * 32, fir_packed_smem(16, R, Wp, G.D, J.KT), s>>>(J);
  }
}

void launch_energy(const Geom &G, const DevTables &T, const DevBatch &W, int device_gate, cudaStream_t s)
{
  (void)T;
  k_energy<<<cdiv((long)W.B * G.nch, 64), 64, 0, s>>>(G, W, device_gate);
}

void launch_demod(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  dim3 grid(cdiv(G.n_dem, 256), (unsigned)(W.B * G.nch));
  k_demod<<<grid, 256, 0, s>>>(G, W, T.atan_tab);
}

void launch_mm(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  if (G.stateless) k_mm_stateless<<<cdiv((long)W.B * G.nch, 32), 32, 0, s>>>(G, W, T.mmse);
  else             k_mm_chained<<<1, 32, 0, s>>>(G, W, T.mmse);
}

void launch_mm_chained_list(const Geom &G, const DevTables &T, const DevBatch &W, int first, int n, unsigned stop_lap,
                            int *res4, cudaStream_t s)
{
  k_mm_chained_list<<<1, 32, 0, s>>>(G, W, T.mmse, T.ac_lut, T.bch, first, n, stop_lap, reinterpret_cast<int4 *>(res4));
}

void launch_search(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  k_search_v1<<<cdiv((long)W.B * G.nch, 64), 64, 0, s>>>(G, T, W);
}

void launch_gather(const Geom &G, const DevBatch &W, cudaStream_t s)
{
  k_gather<<<148, 128, 0, s>>>(G, W);
}

void launch_demod_mm_v2(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, cudaStream_t s)
{
  const int i_end = G.early ? G.ne_dem : G.n_dem;
  if (!G.dem_grid) {               // polyphase mode: the channelizer's epilogue already wrote the demod floats
    dim3 grid(cdiv(i_end, DM_TI), (unsigned)W.B, (unsigned)((G.nch + 31) / 32));
    k_demod_all<<<grid, DM_WARPS * 32, 0, s>>>(G, W, T.atan_tab, demT, i_end);
  }
  constexpr int BLK = 64;
  const size_t smem = mm_smem_bytes(BLK);
  k_mm_stateless_v2<BLK><<<cdiv((long)W.B * G.nch, BLK), BLK, smem, s>>>(G, W, T.mmse, demT, G.early ? 1 : 0,
                                                                            reinterpret_cast<MmSave *>(W.mm_save), nullptr, 0);
}

// lazy tail, second pass: demod tail + clock-recovery resume of the LISTED windows (list4[l] = {b, chi, ., .})
void launch_mm_resume_list(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, const int *list4,
                           int n_list, cudaStream_t s)
{
  if (n_list <= 0) return;
  const int tail = G.n_dem - G.ne_dem;
  if (tail > 0 && !G.dem_grid) {       // polyphase mode: the grid already holds every demod float
    dim3 grid(cdiv(tail, 128), (unsigned)n_list);
    k_demod_list<<<grid, 128, 0, s>>>(G, W, T.atan_tab, demT, reinterpret_cast<const int4 *>(list4), G.ne_dem);
  }
  constexpr int BLK = 64;
  const size_t smem = mm_smem_bytes(BLK);
  k_mm_stateless_v2<BLK><<<cdiv(n_list, BLK), BLK, smem, s>>>(G, W, T.mmse, demT, 2, reinterpret_cast<MmSave *>(W.mm_save),
                                                               reinterpret_cast<const int4 *>(list4), n_list);
}

void launch_tail_scan(const Geom &G, const DevBatch &W, cudaStream_t s)
{
  k_tail_scan<<<1, 1024, 0, s>>>(G, W);
}

void launch_tail_resume(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, cudaStream_t s)
{
  // Two warps per block: the resumed chains are latency-bound (one dependent chain of ~80 instructions per symbol)
  // and their demod fetches are one sector per lane, so more than two warps on an SM queue up behind the load/store
  // unit (measured per 512 slots / 2268 windows: 64-thread blocks 0.9 ms, 32-thread 0.9 ms, 128-thread 2.5 ms).
  constexpr int BLK = 64;
  const size_t smem = mm_smem_bytes(BLK);
  // the list length lives on the device: launch for the worst case, blocks past the list end return at once
  k_mm_stateless_v2<BLK><<<cdiv((long)W.B * G.nch, BLK), BLK, smem, s>>>(G, W, T.mmse, demT, 2, reinterpret_cast<MmSave *>(W.mm_save),
                                                                            W.tail.list, -1);
}

void launch_tail_finish(const Geom &G, const DevBatch &W, cudaStream_t s)
{
  k_tail_compact<<<cdiv((long)W.B * G.nch, 128), 128, 0, s>>>(G, W);
  k_tail_offsets<<<1, 1024, 0, s>>>(W);
  k_gather_sorted<<<296, 128, 0, s>>>(G, W);
}

void launch_search_warp(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  k_search_warp<<<cdiv((long)W.B * G.nch * 32, 256), 256, 0, s>>>(G, T, W);
}

void launch_dmm_stateless(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  constexpr int BLK = 64;
  const size_t smem = 2 * sizeof(c32) * 64 * BLK + sizeof(float) * 16 * BLK + sizeof(float) * (129 * 8 + 257);
  k_dmm_stateless<BLK><<<cdiv((long)W.B * G.nch, BLK), BLK, smem, s>>>(G, W, T.mmse, T.atan_tab);
}

void launch_fill_pass(const DevBatch &W, int n, int v, cudaStream_t s)
{
  k_fill_pass<<<cdiv(n, 256), 256, 0, s>>>(W.pass, n, v);
}

// Deferred exact noise FIR over listed (slot, <= CG channels) groups (lazy squelch), tiled packed kernel; the
// default for D = 50 is the delay-line kernel of rx_firdl.cu -- see launch_noise_fir_list().
template <int CG, int R, int Wp, int BPS, int KTMAX = 512, int DT = 0>
static void launch_list_packed(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                               c32 *NzL, cudaStream_t s)
{
  constexpr int TJ = (32 / CG) * R * Wp;
  static OptIn opt;
  if (opt.need()) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, (const void *)k_fir_packed<CG, R, Wp, BPS, DT>);
    cudaFuncSetAttribute((const void *)k_fir_packed<CG, R, Wp, BPS, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         g_max_smem - (int)fa.sharedSizeBytes);
  }
  FirJob J{};
  J.x = W.x; J.n_x = (long)(W.B - 1) * G.S + G.H; J.taps = T.noise_rtaps; J.out = NzL;
  J.N = G.Nn; J.D = G.D; J.nch = G.nch;
  J.KT = pick_kt_packed(CG, R, Wp, G.D, G.Nn < KTMAX ? G.Nn : KTMAX, BPS, DT > 0);
  J.mode = 2; J.S = G.S; J.fns = G.fns; J.n_noise = G.n_noise; J.tiles_per_slot = (G.n_noise + TJ - 1) / TJ;
  J.groups = groups;
  dim3 grid((unsigned)(n_groups * J.tiles_per_slot), 1);
  k_fir_packed<CG, R, Wp, BPS, DT><<<grid, Wp * 32, fir_packed_smem(CG, R, Wp, G.D, J.KT, DT > 0), s>>>(J);
}

static int lazy_cfg()
{
  static int cfg = -1;
  if (cfg < 0) { const char *e = getenv("BTB200_LAZY_CFG"); cfg = e ? atoi(e) : 6; }
  return cfg;
}

int lazy_group_channels(const Geom &G)
{
  return (lazy_cfg() >= 6 && noise_fir_dl_supported(G)) ? 4 : 2;
}

// Measured on B200 (512 slots, 2 260 hit windows x 17 M complex MAC), ms per launch (BTB200_LAZY_CFG selects):
//   6  delay line + cp.async ring (rx_firdl.cu), <= 4 ch x 448 outputs, 7 warps ............... 13.4  (default, D = 50)
//   0  packed, skewed, 2 ch x 448 outputs (850 = 2 tiles, 95 %), 14 warps, one block/SM ....... 17.0  (kept as cross-check)
//   3  packed 2 ch x 256 outputs, 2 blocks/SM (default for D != 50) ........................... 21.7
// Tried and removed: same shape as 0 without the skew 18.9; skewed 2 ch x 288 outputs (3 tiles, 9 warps) 19.0;
// scalar 4 ch x 256 outputs 23.1; scalar 2 ch x 256 outputs 21.8; 144-output tiles 22.3; 320-output tiles with
// R = 4 24.8; run-time-D skew 25-35; scalar delay line (instruction-fetch bound) 19.1.
void launch_noise_fir_list(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                           c32 *NzL, cudaStream_t s)
{
  const int cfg = lazy_cfg();
  if (cfg >= 6 && noise_fir_dl_supported(G) && T.noise_taps4) {
    launch_noise_fir_dl(G, T, W, groups, n_groups, NzL, s);
    return;
  }
  if (cfg == 0 && G.D == 50) launch_list_packed<2, 2, 14, 1, 1000, 50>(G, T, W, groups, n_groups, NzL, s);
  else launch_list_packed<2, 4, 4, 2>(G, T, W, groups, n_groups, NzL, s);
}

void launch_energy_list(const Geom &G, const DevBatch &W, const int *list4, int n_list, const c32 *NzL,
                        double *e_on, double *e_off, cudaStream_t s)
{
  k_energy_list_warp<<<cdiv((long)n_list * 32, 128), 128, 0, s>>>(G, W, reinterpret_cast<const int4 *>(list4), n_list, NzL, lazy_group_channels(G), e_on, e_off);
}

}  // namespace btb200
