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
// launchers
// ===========================================================================
static inline unsigned cdiv(long a, long b) { return (unsigned)((a + b - 1) / b); }

void launch_chan_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s)
{
  const long Gtot = (long)(W.B - 1) * G.gps + G.n_ddc;
  (void)impl;
  k_chan_fir_v1<<<cdiv(Gtot * G.nch, 128), 128, 0, s>>>(G, W.x, T.chan_rtaps, W.Y, Gtot);
}

void launch_noise_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s)
{
  (void)impl;
  const long n = (long)W.B * G.n_noise * G.nch;
  k_noise_fir_v1<<<cdiv(n, 128), 128, 0, s>>>(G, W.x, T.noise_rtaps, W.Nz, W.B);
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

int kernel_launches_per_batch() { return 7; }

}  // namespace btb200
