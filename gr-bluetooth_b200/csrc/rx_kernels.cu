// rx_kernels.cu -- sm_100a kernels of the receive path.  Compiled with
// --fmad=false (see rx_math.cuh): every fp32 operation rounds once.
#include "rx_kernels.cuh"
#include <cstdio>

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

struct HitEmitter {
  const Geom &G; const DevBatch &W; int b, c, nsym;
  __device__ void operator()(int kind, int offset, int n_symbols, uint32_t lap) const
  {
    const unsigned slot = atomicAdd(W.hit_count, 1u);
    if (slot >= W.hit_cap) return;
    int cnt = n_symbols < 3125 ? n_symbols : 3125;
    if (cnt < 0) cnt = 0;
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
  window_search(G, T.ac_lut, T.le_hdr_lut, W.bits + (long)idx * G.bw, nsym, T.le_index[c], T.le_white[c], em);
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
  c32 *out;             // mode 0/1: rows of nch, channel fastest; mode 2: [group][n_noise][CG]
  int N, D, nch, KT;
  int mode;             // 0: channel grid (tile t = outputs t*TJ..), 1: noise, every slot x channel group,
                        // 2: noise, listed (slot, <=CG channels) groups only (lazy squelch)
  long Gtot; int fcs;
  int S, fns, n_noise, tiles_per_slot;
  const int *groups;    // mode 2: [ngroups][1+CG] = slot, channel indices (-1 = unused)
};

template <int CG, int R, int W>
__global__ void __launch_bounds__(W * 32, (CG == 16) ? 1 : 2) k_fir_tiled(FirJob J)
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
  if (J.mode == 0) {
    const long g0 = (long)blockIdx.x * TJ;
    const long left = J.Gtot - g0;
    nj = left < TJ ? (int)left : TJ;
    s0 = J.fcs + g0 * J.D;
    outp = J.out + g0 * J.nch;
    ostride = J.nch;
  } else {
    const int q = blockIdx.x / J.tiles_per_slot, jt = blockIdx.x - q * J.tiles_per_slot;
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
      const c32 *tg = J.taps + ((size_t)blockIdx.y * J.N + k0) * CG;
      for (int i = threadIdx.x; i < kt * CG; i += W * 32) ts[i] = tg[i];
    }
    const int sp = (TJ - 1) * J.D + kt;
    const long base = s0 + k0;
    for (int i = threadIdx.x; i < sp; i += W * 32) {
      const long n = base + i;
      xs[i] = (n < J.n_x) ? J.x[n] : c32{0.0f, 0.0f};
    }
    __syncthreads();
    const c32 *xp = xs + jj0 * J.D;
    const c32 *tp = ts + cg;
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
  const int c = (J.mode == 2) ? cg : (int)blockIdx.y * CG + cg;
  const bool live = (J.mode == 2) ? true : (c < J.nch);
  if (live) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int jj = jj0 + r * NH * W;
      if (jj < nj) outp[(long)jj * ostride + c] = c32{ar[r], ai[r]};
    }
  }
}

static size_t fir_smem(int CG, int R, int W, int D, int KT) { return ((size_t)KT * CG + (size_t)((32 / CG) * R * W - 1) * D + KT) * sizeof(c32); }
static int g_max_smem = 48 * 1024;

int fir_setup(int device)
{
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return -1;
  g_max_smem = v;
  if (cudaFuncSetAttribute(k_fir_tiled<16, 8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, v) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_fir_tiled<16, 8, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, v) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_fir_tiled<4, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, v) != cudaSuccess) return -1;
  return 0;
}

// largest tap chunk (multiple of 32, <= N rounded up) whose tile fits in shared memory
static int pick_kt(int CG, int R, int W, int D, int N, int blocks_per_sm)
{
  int kt = (N + 31) & ~31;
  const size_t budget = (size_t)(g_max_smem + 1024) / blocks_per_sm - 2048;
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
template <int BLK>
__global__ void __launch_bounds__(BLK) k_dmm_stateless(Geom G, DevBatch W, const float *__restrict__ mmse_g,
                                                       const float *__restrict__ atan_g)
{
  __shared__ float s_mmse[129 * 8];
  __shared__ float s_atan[257];
  __shared__ float ring[16][BLK];
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
  int pnext = 0;                 // demod indices [pnext-16, pnext) live in the ring
  int prev_idx = -1;
  c32 zprev{0.0f, 0.0f};
  uint32_t word = 0;
  const int tid = threadIdx.x;
  while (oo < G.n_dem && ii < ni) {
    if ((int)ii < pnext - 16) { pnext = (int)ii; prev_idx = -1; }       // stepped far backwards: refill
    while (pnext < (int)ii + 8) {
      const int d = pnext;
      float val = 0.0f;                                                 // demod_out[0] is never written
      if (d > 0) {
        const c32 zc = crot(y[(long)d * G.nch], p[(long)d * G.nch]);
        if (prev_idx != d - 1) zprev = crot(y[(long)(d - 1) * G.nch], p[(long)(d - 1) * G.nch]);
        val = demod_point(s_atan, G.demod_gain, zc, zprev);
        zprev = zc;
        prev_idx = d;
      }
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
  if (oo & 31) bits_row[oo >> 5] = word;
  for (int w = (oo + 31) >> 5; w < G.bw; w++) bits_row[w] = 0;
  W.nsym[idx] = oo;
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

// ===========================================================================
// launchers
// ===========================================================================
static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

void launch_chan_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s)
{
  const long Gtot = (long)(W.B - 1) * G.gps + G.n_ddc;
  if (impl == IMPL_BASELINE) {
    k_chan_fir_v1<<<cdiv(Gtot * G.nch, 128), 128, 0, s>>>(G, W.x, T.chan_rtaps, W.Y, Gtot);
    return;
  }
  constexpr int R = 8, Wp = 16, TJ = 2 * R * Wp;   // 16 channels x 256 outputs per block
  FirJob J{};
  J.x = W.x; J.n_x = (long)(W.B - 1) * G.S + G.H; J.taps = T.chan_tg; J.out = W.Y;
  J.N = G.Nc; J.D = G.D; J.nch = G.nch; J.KT = pick_kt(16, R, Wp, G.D, G.Nc, 1);
  J.mode = 0; J.Gtot = Gtot; J.fcs = G.fcs;
  dim3 grid(cdiv(Gtot, TJ), (unsigned)((G.nch + 15) / 16));
  k_fir_tiled<16, R, Wp><<<grid, Wp * 32, fir_smem(16, R, Wp, G.D, J.KT), s>>>(J);
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
  J.KT = pick_kt(16, R, Wp, G.D, G.Nn < 512 ? G.Nn : 512, 1);
  J.mode = 1; J.S = G.S; J.fns = G.fns; J.n_noise = G.n_noise; J.tiles_per_slot = (G.n_noise + TJ - 1) / TJ;
  dim3 grid((unsigned)(W.B * J.tiles_per_slot), (unsigned)((G.nch + 15) / 16));
  k_fir_tiled<16, R, Wp><<<grid, Wp * 32, fir_smem(16, R, Wp, G.D, J.KT), s>>>(J);
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

void launch_search(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  k_search_v1<<<cdiv((long)W.B * G.nch, 64), 64, 0, s>>>(G, T, W);
}

void launch_gather(const Geom &G, const DevBatch &W, cudaStream_t s)
{
  k_gather<<<148, 128, 0, s>>>(G, W);
}

void launch_dmm_stateless(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s)
{
  constexpr int BLK = 64;
  k_dmm_stateless<BLK><<<cdiv((long)W.B * G.nch, BLK), BLK, 0, s>>>(G, W, T.mmse, T.atan_tab);
}

void launch_fill_pass(const DevBatch &W, int n, int v, cudaStream_t s)
{
  k_fill_pass<<<cdiv(n, 256), 256, 0, s>>>(W.pass, n, v);
}

void launch_noise_fir_list(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                           c32 *NzL, cudaStream_t s)
{
  constexpr int CG = LAZY_CG, R = 8, Wp = 4, TJ = (32 / CG) * R * Wp;
  FirJob J{};
  J.x = W.x; J.n_x = (long)(W.B - 1) * G.S + G.H; J.taps = T.noise_rtaps; J.out = NzL;
  J.N = G.Nn; J.D = G.D; J.nch = G.nch;
  J.KT = pick_kt(CG, R, Wp, G.D, G.Nn < 512 ? G.Nn : 512, 2);
  J.mode = 2; J.S = G.S; J.fns = G.fns; J.n_noise = G.n_noise; J.tiles_per_slot = (G.n_noise + TJ - 1) / TJ;
  J.groups = groups;
  dim3 grid((unsigned)(n_groups * J.tiles_per_slot), 1);
  k_fir_tiled<CG, R, Wp><<<grid, Wp * 32, fir_smem(CG, R, Wp, G.D, J.KT), s>>>(J);
}

void launch_energy_list(const Geom &G, const DevBatch &W, const int *list4, int n_list, const c32 *NzL,
                        double *e_on, double *e_off, cudaStream_t s)
{
  k_energy_list<<<cdiv(n_list, 32), 32, 0, s>>>(G, W, reinterpret_cast<const int4 *>(list4), n_list, NzL, LAZY_CG, e_on, e_off);
}

}  // namespace btb200
