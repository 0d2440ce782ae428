// rx_pfb.cu -- throughput-mode channelizer (BTB200_DDC_POLYPHASE): the 79 per-channel
// freq_xlating_fir_filter_ccf objects of the reference (lib/multi_block.cc:329-341, run at :194-204)
// restated as ONE real-tap polyphase bank + DFT at the channel bins, with the FM demod
// (lib/multi_block.cc:158-168) and the on-channel energy (:206-218) fused into its epilogue.
//
// NOT on the bit-exact path (that is rx_kernels.cu, BTB200_DDC_EXACT): fp32 with FMA contraction and its own
// summation order; demod floats agree with the reference's to ~1e-5 (median) at 100 Msps -- the reference's own
// taps carry fp32 phase errors of that size (tests/test_pfb_model.py, tests/test_gpu_polyphase.py).
//
// Math.  Channel c of the reference computes, on the decimation grid n0 = fcs + g D,
//     y_c[g] = sum_k x[n0 + k] h[Nc-1-k] e^{j th_c (Nc-1-k)},   th_c = 2 pi (a_c + phi) / M,
// a_c integer MHz offset, phi the fractional offset common to all channels, M = samples per MHz.  With
// x'[n] = x[n] e^{-j 2 pi phi n / M} and h'[k] = h[Nc-1-k]:
//     u_g[r] = sum_{q<Q} x'[n0 + r + M q] h'[r + M q]             M branches of Q = ceil(Nc / M) real taps
//     Z_c[g] = sum_{r<M} e^{-j 2 pi a_c r / M} u_g[r]              M-point DFT, evaluated at the channel bins
//     y_c[g] = (unit-modulus factor that advances by a constant per grid step) * Z_c[g]
// The demod only sees y[i] conj(y[i-1]), where the rotator and all those factors collapse into the constant
// kappa_c = e^{-j 2 pi a_c D / M} (= +-1 for D = M / 2); the energy only sees |y| = |Z|.
//
// The DFT is a Good-Thomas prime-factor DFT, M = N1 N2 with gcd(N1, N2) = 1 (100 = 4 x 25): input index
// r = (N2 n1 + N1 n2) mod M, bin k <-> (k mod N1, k mod N2), no twiddles between the stages.  The N1-point
// stage (N1 in {1, 2, 4}: additions only) is fused with the branch sums; the N2-point stage is a small complex
// matrix product per residue class k1 that only produces the bins of actual channels.
//
// One block = one tile of 63 grid points (+ the one before, for the differential product) x all channels:
//   0. the tile's input span (3850 samples at 100 Msps) arrives by ONE TMA bulk copy (cp.async.bulk + mbarrier)
//      while the threads fetch the tables; optional pre-rotation by phi in shared memory;
//   1. branch sums with the taps in registers: consecutive same-parity grid points are one tap apart (2 D = M), so a
//      thread slides R = 8 outputs over R + Q - 1 loaded samples; N1-point DFTs in registers; V[k1][n2][t] to smem;
//   2. N2-point DFTs: thread = 4 grid points x 5 channels of one residue class, 25 iterations of 80 FMA on
//      4 + 5 conflict-free shared-memory loads;
//   3. Z tile to shared memory; demod (fast_atan2f table) written channel-contiguous = one contiguous run per
//      tile; per-tile sums of |Z|^2 for the window energies.
#include "rx_pfb.cuh"
#include "rx_tma.cuh"
#include <cstdio>
#include <cstring>

namespace btb200 {

namespace {

constexpr int PFB_THREADS = 256;
constexpr int PFB_R = 8;             // outputs a thread slides in stage 1
constexpr int VP = PFB_TT + 1;       // row pitch of V (odd: conflict-free stores from the branch stage)

struct PfbSmem { size_t xs, v, tab, wb, hq, atan, kap, cch, n2r, tab_bytes, bar, total; };

__host__ __device__ inline PfbSmem pfb_layout(const PfbPlan &P)
{
  PfbSmem L{};
  size_t o = 0;
  auto take = [&o](size_t bytes, size_t align) { o = (o + align - 1) / align * align; const size_t r = o; o += bytes; return r; };
  const int ZP = P.ncol | 1;
  const size_t vbytes = (size_t)P.N1 * P.N2 * VP * sizeof(c32), zbytes = (size_t)PFB_TT * ZP * sizeof(c32);
  // the staged input span; the epilogue reuses it for the demod tile of the channel-major copy ([ncol][TT + 1] floats)
  const size_t xbytes = (size_t)P.span * sizeof(c32), dbytes = (size_t)P.ncol * (PFB_TT + 1) * sizeof(float);
  L.xs = take(xbytes > dbytes ? xbytes : dbytes, 128);
  L.v = take(vbytes > zbytes ? vbytes : zbytes, 16);
  // the tables sit in shared memory exactly as in the global blob (pfb_pack_tables): one bulk copy fetches them
  L.tab = take(0, 128);
  L.wb = take((size_t)P.N2 * P.ncol * sizeof(c32), 16);
  L.hq = take((size_t)P.Q * P.M * sizeof(float), 16);
  L.atan = take(260 * sizeof(float), 16);
  L.kap = take((size_t)P.ncol * sizeof(c32), 16);
  L.cch = take((size_t)P.ncol * sizeof(int), 16);
  L.n2r = take((size_t)((P.N2 + 3) & ~3) * sizeof(int), 16);
  L.tab_bytes = ((o + 15) & ~(size_t)15) - L.tab;
  o = L.tab + L.tab_bytes;
  L.bar = take(8, 8);
  L.total = o;
  return L;
}

// gr::fast_atan2f, tolerance-mode demod: branch-free (the reference's early returns and quadrant
// ladder become selects -- divergent branches cost this epilogue more than the arithmetic), approximate reciprocal,
// floor and fraction of the table position from one round-towards-zero add of 2^23 instead of two conversions.
__device__ __forceinline__ float atan2_tab(const float *__restrict__ T, float y, float x)
{
  const float ya = fabsf(y), xa = fabsf(x);
  const float mx = fmaxf(ya, xa), mn = fminf(ya, xa);
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
  const float z = mn * rc;                               // NaN for x = y = 0: selected away at the end
  const float alpha = z * 255.0f;
  const float tz = __fadd_rz(alpha, 8388608.0f);         // 2^23 + floor(alpha), alpha in [0, 255]
  const int idx = __float_as_int(tz) & 0xff;
  const float frac = alpha - (tz - 8388608.0f);
  const float t0 = T[idx], t1 = T[idx + 1];
  float base = fmaf(t1 - t0, frac, t0);
  base = (z < 0.003921569f) ? z : base;
  // quadrant fix-up of the reference, folded: (xa > ya ? base : pi/2 - base), mirrored for x < 0, signed like y
  const float pi = 3.14159265358979323846f, hp = 1.57079632679489661923f;
  float angle = (xa > ya) ? base : hp - base;
  angle = (x < 0.0f) ? pi - angle : angle;
  angle = copysignf(angle, y);
  return (mx > 0.0f) ? angle : 0.0f;
}

template <int N1, int Q>
__global__ void __launch_bounds__(PFB_THREADS, 2) k_pfb(PfbPlan P, const c32 *__restrict__ x, long n_samples, long Gtot,
                                                        long tile0)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const PfbSmem L = pfb_layout(P);
  c32 *xs = reinterpret_cast<c32 *>(smem + L.xs);
  c32 *V = reinterpret_cast<c32 *>(smem + L.v);
  c32 *WBs = reinterpret_cast<c32 *>(smem + L.wb);
  float *hqs = reinterpret_cast<float *>(smem + L.hq);
  float *atans = reinterpret_cast<float *>(smem + L.atan);
  c32 *kaps = reinterpret_cast<c32 *>(smem + L.kap);
  int *cch = reinterpret_cast<int *>(smem + L.cch);
  int *n2r = reinterpret_cast<int *>(smem + L.n2r);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);

  const int tid = threadIdx.x;
  const long tile = tile0 + blockIdx.x;
  const int seg = (int)(tile / P.tps), jt = (int)(tile - (long)seg * P.tps);
  const int in_seg0 = jt * PFB_T;                                  // first own point, relative to the segment
  const long g_first = (long)seg * P.gps + in_seg0;
  if (in_seg0 >= P.gps || g_first >= Gtot) return;
  int n_own = PFB_T;
  if (n_own > P.gps - in_seg0) n_own = P.gps - in_seg0;
  if ((long)n_own > Gtot - g_first) n_own = (int)(Gtot - g_first);
  const long gs = g_first - 1;                                     // grid point of local index t = 0
  const long s0 = (long)P.fcs + gs * P.D;                          // first sample of the staged span
  const int M = P.M, D = P.D, N2 = P.N2, ncol = P.ncol;

  // ---- 0. stage the input span (TMA bulk copy when it is aligned and inside the batch) and the tables
  const bool bulk = s0 >= 0 && s0 + P.span <= n_samples && ((s0 | P.span) & 1) == 0;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  if (tid == 0) {
    const unsigned tbytes = (unsigned)L.tab_bytes, bytes = bulk ? (unsigned)P.span * (unsigned)sizeof(c32) : 0u;
    mbar_expect_tx(bar, tbytes + bytes);
    tma_bulk_g2s(smem + L.tab, P.tables, tbytes, bar);
    if (bulk) tma_bulk_g2s(xs, x + s0, bytes, bar);
  }
  if (!bulk) {
    for (int i = tid; i < P.span; i += PFB_THREADS) {
      const long n = s0 + i;
      xs[i] = (n >= 0 && n < n_samples) ? x[n] : c32{0.0f, 0.0f};
    }
  }
  mbar_wait(bar, 0);
  __syncthreads();
  if (P.xr) {
    // the estimator's rotated copy of the tile's own samples: local indices [D, D + n_own D) of the staged span
    // a thread's samples are PFB_THREADS apart: one phasor from the table, then a fixed rotation per step (a dozen
    // steps: the recurrence stays within 1e-6 of the table), so the loop has no dependent global loads
    const long n0 = s0 + D;
    const int cnt = n_own * D;
    if (tid < cnt) {
      c32 w = P.phasor[(int)((n0 + tid) % P.period)];
      const c32 st = P.phasor[PFB_THREADS % P.period];
      for (int i = tid; i < cnt; i += PFB_THREADS) {
        const long n = n0 + i;
        const c32 v = xs[D + i];
        if (n < n_samples) P.xr[n] = c32{v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re};
        w = c32{w.re * st.re - w.im * st.im, w.re * st.im + w.im * st.re};
      }
    }
    if (tile == 0) {
      // the samples in front of the first grid point (nobody's own)
      for (long n = tid; n < n0 && n < n_samples; n += PFB_THREADS) {
        const c32 v = x[n], w = P.phasor[(int)(n % P.period)];
        P.xr[n] = c32{v.re * w.re - v.im * w.im, v.re * w.im + v.im * w.re};
      }
    }
    if (P.phi_step != 0.0f) __syncthreads();
  }
  if (P.phi_step != 0.0f) {
    // x'[i] = x[i] e^{-j 2 pi phi i / M}: the phase origin is the tile's (any origin common to Z[g] and Z[g-1] does)
    for (int i = tid; i < P.span; i += PFB_THREADS) {
      float sn, cs;
      sincospif(P.phi_step * (float)i, &sn, &cs);
      const c32 v = xs[i];
      xs[i] = c32{v.re * cs - v.im * sn, v.re * sn + v.im * cs};
    }
    __syncthreads();
  }

  // ---- 1. branch sums + N1-point DFTs.  item = (rho, parity, run): local grid points t = p + 2 (i0 + o), o < R
  constexpr int R = PFB_R;
  const int runs = PFB_TT / 2 / R;
  for (int item = tid; item < N2 * 2 * runs; item += PFB_THREADS) {
    const int rho = item % N2, pr = item / N2;
    const int p = pr & 1, i0 = (pr >> 1) * R;
    const int n2 = n2r[rho];
    float ar[N1][R], ai[N1][R], h[N1][Q];
    int base[N1];
#pragma unroll
    for (int n1 = 0; n1 < N1; n1++) {
      const int r = (N2 * n1 + N1 * n2) % M;
      base[n1] = p * D + r + M * i0;
#pragma unroll
      for (int q = 0; q < Q; q++) h[n1][q] = hqs[q * M + r];
#pragma unroll
      for (int o = 0; o < R; o++) { ar[n1][o] = 0.0f; ai[n1][o] = 0.0f; }
    }
#pragma unroll
    for (int s = 0; s < R + Q - 1; s++) {
#pragma unroll
      for (int n1 = 0; n1 < N1; n1++) {
        const c32 X = xs[base[n1] + M * s];
#pragma unroll
        for (int o = 0; o < R; o++) {
          const int q = s - o;
          if (q >= 0 && q < Q) { ar[n1][o] = fmaf(X.re, h[n1][q], ar[n1][o]); ai[n1][o] = fmaf(X.im, h[n1][q], ai[n1][o]); }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < R; o++) {
      const int t = p + 2 * (i0 + o);
      if constexpr (N1 == 1) {
        V[n2 * VP + t] = c32{ar[0][o], ai[0][o]};
      } else if constexpr (N1 == 2) {
        V[n2 * VP + t] = c32{ar[0][o] + ar[1][o], ai[0][o] + ai[1][o]};
        V[(N2 + n2) * VP + t] = c32{ar[0][o] - ar[1][o], ai[0][o] - ai[1][o]};
      } else {
        // W_4 = -j:  k1 = 1: u0 - j u1 - u2 + j u3;  k1 = 3: u0 + j u1 - u2 - j u3
        const float er = ar[0][o] + ar[2][o], ei = ai[0][o] + ai[2][o];     // u0 + u2
        const float fr = ar[0][o] - ar[2][o], fi = ai[0][o] - ai[2][o];     // u0 - u2
        const float gr = ar[1][o] + ar[3][o], gi = ai[1][o] + ai[3][o];     // u1 + u3
        const float hr = ar[1][o] - ar[3][o], hi = ai[1][o] - ai[3][o];     // u1 - u3
        V[n2 * VP + t] = c32{er + gr, ei + gi};
        V[(N2 + n2) * VP + t] = c32{fr + hi, fi - hr};               // f - j h
        V[(2 * N2 + n2) * VP + t] = c32{er - gr, ei - gi};
        V[(3 * N2 + n2) * VP + t] = c32{fr - hi, fi + hr};           // f + j h
      }
    }
  }
  __syncthreads();

  // ---- 2. N2-point DFTs at the channel bins: thread = local points {tg + 16 i} x columns col0 .. col0 + 4
  const int n_cg = ncol / PFB_NCOL;
  const bool actB = tid < 16 * n_cg;
  const int tg = tid & 15, cg = tid >> 4;
  float zr[4][PFB_NCOL], zi[4][PFB_NCOL];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < PFB_NCOL; j++) { zr[i][j] = 0.0f; zi[i][j] = 0.0f; }
  if (actB) {
    const int col0 = cg * PFB_NCOL;
    const int k1 = col0 / P.CPC;
    const c32 *vrow = V + (size_t)k1 * N2 * VP + tg;
    const c32 *wrow = WBs + col0;
#pragma unroll 5
    for (int n2 = 0; n2 < N2; n2++) {
      c32 v[4], w[PFB_NCOL];
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = vrow[n2 * VP + 16 * i];
#pragma unroll
      for (int j = 0; j < PFB_NCOL; j++) w[j] = wrow[n2 * ncol + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < PFB_NCOL; j++) {
          zr[i][j] = fmaf(v[i].re, w[j].re, zr[i][j]); zr[i][j] = fmaf(-v[i].im, w[j].im, zr[i][j]);
          zi[i][j] = fmaf(v[i].re, w[j].im, zi[i][j]); zi[i][j] = fmaf(v[i].im, w[j].re, zi[i][j]);
        }
    }
  }
  __syncthreads();                                   // every read of V is done: the Z tile takes its place
  const int ZP = ncol | 1;
  c32 *zs = V;
  if (actB) {
    const int col0 = cg * PFB_NCOL;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < PFB_NCOL; j++) zs[(tg + 16 * i) * ZP + col0 + j] = c32{zr[i][j], zi[i][j]};
  }
  __syncthreads();

  // ---- 3. demod of the own points (local t = 1 .. n_own) and the tile's energy sums.  thread = (column, row group):
  // rows t = 1 + rg, 1 + rg + RG, ...; the |Z|^2 of its rows ride along and are reduced across the row groups.
  {
    const int RG = PFB_THREADS / ncol;                   // row groups (3 for 80 columns)
    constexpr int DTP = PFB_TT + 1;                      // pitch of the demod tile (channel-major copy)
    const int col = tid % ncol, rg = tid / ncol;
    float sa = 0.0f, sb = 0.0f;
    if (rg < RG) {
      const int ch = cch[col];
      const c32 k = kaps[col];
      float *drow = P.dem + (gs + 1 + rg) * (long)P.nch + ch;
      float *dt = reinterpret_cast<float *>(xs) + col * DTP;   // the staged input is dead: demod tile [col][t], odd pitch
      // three rows per pass: the six loads go out together, then the arithmetic, then the stores (the shared-memory
      // store of the demod tile would otherwise fence the next row's loads behind it)
      constexpr int UB = 3;
      for (int t = 1 + rg; t <= n_own; t += UB * RG, drow += (long)UB * RG * P.nch) {
        c32 z1[UB], z0[UB];
#pragma unroll
        for (int u = 0; u < UB; u++) {
          const int tt = t + u * RG <= n_own ? t + u * RG : t;
          z1[u] = zs[tt * ZP + col]; z0[u] = zs[(tt - 1) * ZP + col];
        }
        float d[UB];
#pragma unroll
        for (int u = 0; u < UB; u++) {
          const float m = z1[u].re * z1[u].re + z1[u].im * z1[u].im;
          if (t + u * RG <= n_own) {
            sa += m;
            if (in_seg0 + t + u * RG - 1 < P.rem) sb += m;
          }
          const float pr = z1[u].re * z0[u].re + z1[u].im * z0[u].im, pi = z1[u].im * z0[u].re - z1[u].re * z0[u].im;    // z1 conj(z0)
          const float qr = pr * k.re - pi * k.im, qi = pr * k.im + pi * k.re;
          d[u] = P.gain * atan2_tab(atans, qi, qr);
        }
        if (ch >= 0) {
#pragma unroll
          for (int u = 0; u < UB; u++)
            if (t + u * RG <= n_own) {
              drow[(long)u * RG * P.nch] = d[u];
              if (P.demC) dt[t + u * RG] = d[u];
            }
        }
      }
    }
    __syncthreads();                                     // the Z tile is dead: its first rows carry the partial sums
    if (P.demC) {
      // channel-major copy: element (column, t) of the demod tile, a warp = 32 consecutive grid points of one channel;
      // independent load/store pairs, four in flight per thread
      const float *dtile = reinterpret_cast<const float *>(xs);
#pragma unroll 4
      for (int e = tid; e < ncol * PFB_TT; e += PFB_THREADS) {
        const int cl = e / PFB_TT, t = e - cl * PFB_TT;
        const int ch = cch[cl];
        if (ch >= 0 && t >= 1 && t <= n_own) P.demC[(long)ch * P.pitchC + gs + t] = dtile[cl * DTP + t];
      }
    }
    float *part = reinterpret_cast<float *>(zs);
    if (rg < RG) { part[(rg * ncol + col) * 2] = sa; part[(rg * ncol + col) * 2 + 1] = sb; }
    __syncthreads();
    if (tid < ncol) {
      float ta = 0.0f, tb = 0.0f;
      for (int g2 = 0; g2 < RG; g2++) { ta += part[(g2 * ncol + tid) * 2]; tb += part[(g2 * ncol + tid) * 2 + 1]; }
      P.E[(tile * ncol + tid) * 2] = ta;
      P.E[(tile * ncol + tid) * 2 + 1] = tb;
    }
  }
}

// window energies from the per-tile sums: window b = segments b .. b + nfull - 1 and the first `rem` points of
// segment b + nfull (lib/multi_block.cc:206-218: mean |y|^2 over the whole window)
__global__ void k_pfb_energy(PfbPlan P, int B, const int *__restrict__ chan_col, double *__restrict__ e_on)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * P.nch) return;
  const int b = idx / P.nch, c = idx - b * P.nch;
  const int col = chan_col[c];
  // the tiles of segments b .. b + nfull - 1 are consecutive: four independent partial sums keep four loads in flight
  const float *e = P.E + ((long)b * P.tps * P.ncol + col) * 2;
  const long st = (long)P.ncol * 2;
  const int nt = P.nfull * P.tps;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = 0;
  for (; k + 3 < nt; k += 4) {
    s0 += (double)e[(long)k * st]; s1 += (double)e[(long)(k + 1) * st];
    s2 += (double)e[(long)(k + 2) * st]; s3 += (double)e[(long)(k + 3) * st];
  }
  for (; k < nt; k++) s0 += (double)e[(long)k * st];
  double sum = (s0 + s1) + (s2 + s3);
  if (P.rem > 0)
    for (int j = 0; j * PFB_T < P.rem; j++) sum += (double)P.E[(((long)(b + P.nfull) * P.tps + j) * P.ncol + col) * 2 + 1];
  e_on[idx] = sum / (double)P.n_ddc;
}

// 4 samples per thread; 16-byte loads when the source allows it
__global__ void k_i16_to_c32(const int16_t *__restrict__ src, c32 *__restrict__ dst, long n_samples, int vec)
{
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n_samples) return;
  if (vec && i + 4 <= n_samples) {
    const int4 v = *reinterpret_cast<const int4 *>(src + 2 * i);      // 4 samples: (re, im) int16 pairs
    const short2 a = *reinterpret_cast<const short2 *>(&v.x), b = *reinterpret_cast<const short2 *>(&v.y);
    const short2 c = *reinterpret_cast<const short2 *>(&v.z), d = *reinterpret_cast<const short2 *>(&v.w);
    float4 *o = reinterpret_cast<float4 *>(dst + i);
    o[0] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
    o[1] = make_float4((float)c.x, (float)c.y, (float)d.x, (float)d.y);
  } else {
    for (long n = i; n < i + 4 && n < n_samples; n++) dst[n] = c32{(float)src[2 * n], (float)src[2 * n + 1]};
  }
}

}  // namespace

size_t pfb_smem_bytes(const PfbPlan &P) { return pfb_layout(P).total; }

size_t pfb_table_bytes(const PfbPlan &P) { return pfb_layout(P).tab_bytes; }

void pfb_pack_tables(const PfbPlan &P, const c32 *WB, const float *hq, const float *atan_tab, const c32 *kappa,
                     const int *col_chan, const int *n2_of_rho, unsigned char *blob)
{
  const PfbSmem L = pfb_layout(P);
  memset(blob, 0, L.tab_bytes);
  memcpy(blob + (L.wb - L.tab), WB, (size_t)P.N2 * P.ncol * sizeof(c32));
  memcpy(blob + (L.hq - L.tab), hq, (size_t)P.Q * P.M * sizeof(float));
  memcpy(blob + (L.atan - L.tab), atan_tab, 257 * sizeof(float));
  memcpy(blob + (L.kap - L.tab), kappa, (size_t)P.ncol * sizeof(c32));
  memcpy(blob + (L.cch - L.tab), col_chan, (size_t)P.ncol * sizeof(int));
  memcpy(blob + (L.n2r - L.tab), n2_of_rho, (size_t)P.N2 * sizeof(int));
}

template <int N1, int Q>
static int pfb_optin(const PfbPlan &P)
{
  return cudaFuncSetAttribute((const void *)k_pfb<N1, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)pfb_smem_bytes(P)) == cudaSuccess ? 0 : -1;
}

#define PFB_DISPATCH(CALL)                                            \
  do {                                                                \
    if (P.Q == 7) {                                                   \
      if (P.N1 == 4) { CALL(4, 7); } else if (P.N1 == 2) { CALL(2, 7); } else { CALL(1, 7); } \
    } else {                                                          \
      if (P.N1 == 4) { CALL(4, 8); } else if (P.N1 == 2) { CALL(2, 8); } else { CALL(1, 8); } \
    }                                                                 \
  } while (0)

int pfb_setup(const PfbPlan &P)
{
  if ((P.Q != 7 && P.Q != 8) || 2 * P.D != P.M || (P.N1 != 1 && P.N1 != 2 && P.N1 != 4) || P.N1 * P.N2 != P.M) return -1;
  if (P.ncol % PFB_NCOL != 0 || 16 * (P.ncol / PFB_NCOL) > PFB_THREADS || P.ncol > PFB_THREADS) return -1;
  if (pfb_smem_bytes(P) > 227 * 1024) return -1;
  int rc = 0;
#define PFB_OPT(N1_, Q_) rc = pfb_optin<N1_, Q_>(P)
  PFB_DISPATCH(PFB_OPT);
#undef PFB_OPT
  return rc;
}

long pfb_tiles(const PfbPlan &P, int B)
{
  const long nseg = (long)(B - 1) + P.nfull + (P.rem > 0 ? 1 : 0);
  return nseg * P.tps;
}

long pfb_samples(const PfbPlan &P, long tile_end)
{
  if (tile_end <= 0) return 0;
  const long tile = tile_end - 1;
  const long seg = tile / P.tps, jt = tile - seg * P.tps;
  long in_seg_last = (jt + 1) * PFB_T;
  if (in_seg_last > P.gps) in_seg_last = P.gps;
  const long g_last = seg * P.gps + in_seg_last - 1;
  return (long)P.fcs + g_last * P.D + (long)P.Q * P.M;
}

void launch_pfb(const PfbPlan &P, const c32 *x, long n_samples, int B, long tile0, long tile1, cudaStream_t s)
{
  if (tile1 <= tile0) return;
  const long Gtot = (long)(B - 1) * P.gps + P.n_ddc;
  const size_t smem = pfb_smem_bytes(P);
  const dim3 grid((unsigned)(tile1 - tile0));
#define PFB_RUN(N1_, Q_) k_pfb<N1_, Q_><<<grid, PFB_THREADS, smem, s>>>(P, x, n_samples, Gtot, tile0)
  PFB_DISPATCH(PFB_RUN);
#undef PFB_RUN
}

void launch_pfb_energy(const PfbPlan &P, int B, double *e_on, cudaStream_t s)
{
  const int n = B * P.nch;
  k_pfb_energy<<<(n + 127) / 128, 128, 0, s>>>(P, B, P.chan_col, e_on);
}

void launch_i16_to_c32(const int16_t *src, c32 *dst, long n_samples, cudaStream_t s)
{
  if (n_samples <= 0) return;
  const int vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const long threads = (n_samples + 3) / 4;
  k_i16_to_c32<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(src, dst, n_samples, vec);
}

}  // namespace btb200
