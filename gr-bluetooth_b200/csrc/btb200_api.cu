// btb200_api.cu -- the extern "C" ABI declared in include/btb200.h.
// Host orchestration only: all arithmetic on the sample path runs in the
// kernels of rx_kernels.cu; there is no CPU fallback.
#include "../../include/btb200.h"
#include "plan.hpp"
#include "rx_kernels.cuh"
#include "rx_fast.cuh"
#include "rx_pfb.cuh"
#include "rx_nest.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace btb200;

namespace {
constexpr int kNumEvents = 8;
constexpr uint32_t kDefaultMaxSlots = 64;
constexpr unsigned kHitCap = 1u << 18;
constexpr int kCopyParts = 4;
constexpr unsigned long long kArenaCap = 256ull << 20;
}

namespace {
// BTB200_TRACE=1: host-side timeline of collect() on stderr
struct Trace {
  bool on = std::getenv("BTB200_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char *what)
  {
    if (!on) return;
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "[btb200 trace] %8.1f us  %s\n", us, what);
  }
};
}  // namespace

struct btb200_ctx {
  btb200_config cfg{};
  Plan plan;
  Geom G{};
  DevTables T{};
  DevBatch W{};
  int device = 0, sm_count = 0;
  uint32_t max_slots = 0;
  cudaStream_t stream = nullptr;       // compute stream: shared by all contexts of a device (see device_stream())
  bool owns_stream = false;
  cudaStream_t copy_stream = nullptr;  // host->device input copies, overlapping whatever the compute stream runs
  cudaEvent_t ev_sync = nullptr, ev_h2d = nullptr;
  cudaEvent_t ev_part[4] = {};
  cudaEvent_t ev[kNumEvents + 1]{};
  cudaEvent_t evl[3]{};      // lazy squelch: noise FIR / energies
  cudaEvent_t ev_tail = nullptr;
  cudaEvent_t ev_user[2] = {};
  cudaEvent_t ev_res = nullptr;       // throughput mode: end of the resume
  bool lazy_timed = false;
  // device allocations
  std::vector<void *> allocs;
  c32 *d_x = nullptr;
  size_t x_cap = 0;          // samples
  c32 *d_phc = nullptr, *d_phn = nullptr;
  float *d_soft = nullptr;
  float *d_dem = nullptr;
  // lazy squelch (stateless): exact energies only for windows with hits
  bool lazy = false;
  c32 *d_NzL = nullptr;
  int *d_groups = nullptr, *d_list = nullptr;
  double *d_eon = nullptr, *d_eoff = nullptr;
  int *h_groups = nullptr, *h_list = nullptr;
  double *h_eon = nullptr, *h_eoff = nullptr;
  size_t list_cap = 0, group_cap = 0;
  DevBatch pendW{};
  struct CollectState {
    bool begun = false;
    unsigned nh = 0, dropped = 0;
    unsigned long long used = 0;
    std::vector<uint32_t> keys, need_exact;
    std::vector<uint8_t> est_flag;      // per channel-window: 1 = snr from the fast estimate
    int nl = 0;                         // listed windows whose exact energies are on their way
  } cb;
  cudaEvent_t ev_up = nullptr;
  Trace trace;
  // lazy tail: clock recovery stops after the searchable prefix; hit windows are resumed in collect()
  bool early = false, pend_early = false;
  Geom pendG{};
  cudaStream_t stream2 = nullptr;
  int *d_list2 = nullptr, *h_list2 = nullptr, *h_nsym = nullptr;
  int *d_res4 = nullptr;
  // fast guarded snr
  bool fast_snr = false;
  FastNoisePlan F{};
  double *h_esum = nullptr;
  std::vector<int> win_mask;         // pass flags for the next submit (btb200_set_window_mask), empty: none
  int *h_mask = nullptr;             // pinned staging of the mask
  bool use_nest = false;             // rx_nest.cu (fused polyphase + DFT) instead of the two kernels of rx_fast.cu
  long nest_reach = 0;               // samples from a window's first one that the estimator's tiles read
  int *d_smflag = nullptr;           // k_nest2: per-SM "resume chains running" flags
  PfbDesign nfd;
  NestPlan NP{};
  double phi = 0;                    // common fractional MHz offset of the noise DDCs
  std::vector<double> fast_off;      // [B][nch] fast off-channel energy of the last batch
  // polyphase (throughput) mode: rx_pfb.cu
  bool poly = false;
  bool dense_tail = false;           // throughput mode: recent batches had hits in so many windows that the lazy tail does not pay
  PfbDesign pfd;
  PfbPlan PF{};
  double *d_eon_all = nullptr, *h_eon_all = nullptr;   // [B][nch] on-channel window energies from the channelizer
  int16_t *d_x16 = nullptr;          // staging of int16 input (btb200_submit_i16)
  // pinned host
  double *h_energy = nullptr, *h_noise = nullptr;
  int *h_pass = nullptr;
  unsigned *h_counts = nullptr;      // [0] hits, [2..3] arena_used (64-bit)
  DevHit *h_hits = nullptr;
  uint8_t *h_arena = nullptr;
  c32 *h_ph = nullptr;               // chained-mode phase staging
  size_t h_ph_cap = 0;
  // stream state
  std::vector<Rotator> rot_c, rot_n;
  MmState mm{};
  bool pending = false;
  uint32_t pend_slots = 0;
  uint64_t pend_first_slot = 0;
  uint32_t last_slots = 0;
  float timing[8]{};
  uint64_t launches = 0;
  int impl = IMPL_TUNED;
  std::string last_error;
};

namespace {

thread_local std::string g_create_error;


#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      char buf_[512];                                                                     \
      std::snprintf(buf_, sizeof buf_, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->last_error = buf_;                                                             \
      return BTB200_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

template <class T>
int dev_alloc(btb200_ctx *ctx, T **p, size_t count)
{
  void *v = nullptr;
  cudaError_t e = cudaMalloc(&v, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    ctx->last_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? BTB200_ERR_NOMEM : BTB200_ERR_CUDA;
  }
  ctx->allocs.push_back(v);
  *p = (T *)v;
  return 0;
}

template <class T>
int upload_raw(btb200_ctx *ctx, const T **dst, const void *src, size_t count)
{
  T *p = nullptr;
  int rc = dev_alloc(ctx, &p, count);
  if (rc) return rc;
  CK(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *dst = p;
  return 0;
}
template <class T>
int upload(btb200_ctx *ctx, const T **dst, const std::vector<T> &src)
{
  return upload_raw<T>(ctx, dst, src.data(), src.size());
}

void reset_stream_state(btb200_ctx *ctx)
{
  const Plan &P = ctx->plan;
  ctx->rot_c.assign(P.nch, Rotator{});
  ctx->rot_n.assign(P.nch, Rotator{});
  for (int c = 0; c < P.nch; c++) {
    ctx->rot_c[c].incr = P.chan_incr[c];
    ctx->rot_n[c].incr = P.noise_incr[c];
  }
  ctx->mm = MmState{P.mu0, P.omega_mid, 0.0f};
}


// Tables of the fast noise estimator (rx_fast.cu).  Returns 0; leaves ctx->fast_snr false when the
// configuration does not fit the polyphase model (non-integer MHz rate, D not dividing M).
int setup_fast(btb200_ctx *ctx)
{
  const Plan &P = ctx->plan;
  const double Md = P.fs / 1e6;
  const int M = (int)std::llround(Md);
  if (std::fabs(Md - M) > 1e-9 || M < 1 || M % P.D != 0) return 0;
  // noise DDC offsets in MHz: a_c + phi, a_c integer, phi common to all channels
  const double f0 = (2402e6 + P.ch_lo * 1e6 + 790000.0 - P.fc) / 1e6;
  const double a0 = std::floor(f0);
  const double phi = f0 - a0;
  // phasor period: phi/M as a fraction with denominator 1000*M
  const long num = std::llround(phi * 1000.0);
  if (std::fabs(phi * 1000.0 - num) > 1e-6) return 0;
  long den = 1000L * M, g = num ? num : den;
  for (long a = den, b = g; b;) { long t = a % b; a = b; b = t; g = a; }
  const long period = den / (num ? g : den);
  if (period > (1 << 20)) return 0;
  FastNoisePlan &F = ctx->F;
  F.M = M; F.C = M / P.D; F.Q = (P.Nn + M - 1) / M; F.period = (int)period;
  F.nchp = (P.nch + 15) & ~15;
  std::vector<float> hpad((size_t)(F.Q + 16) * M, 0.0f);
  for (int k = 0; k < P.Nn; k++) hpad[(size_t)(k / M + 8) * M + (k % M)] = P.noise_proto[P.Nn - 1 - k];   // h'[k] = h[N-1-k]
  std::vector<c32> ph((size_t)period);
  for (long n = 0; n < period; n++) {
    const double a = -2.0 * M_PI * phi * (double)n / M;
    ph[(size_t)n] = c32{(float)std::cos(a), (float)std::sin(a)};
  }
  std::vector<c32> tw((size_t)M * F.nchp, c32{0.0f, 0.0f});
  for (int c = 0; c < P.nch; c++) {
    const long ac = (long)a0 + c;
    for (int r = 0; r < M; r++) {
      const long m = ((ac * r) % M + M) % M;
      const double a = -2.0 * M_PI * (double)m / M;
      tw[(size_t)r * F.nchp + c] = c32{(float)std::cos(a), (float)std::sin(a)};
    }
  }
  int rc;
  if ((rc = upload(ctx, &F.hpad, hpad))) return rc;
  if ((rc = upload(ctx, &F.phasor, ph))) return rc;
  if ((rc = upload(ctx, &F.twid, tw))) return rc;
  const size_t B = ctx->max_slots;
  if ((rc = dev_alloc(ctx, &F.esum, B * P.nch))) return rc;
  {
    // fused estimator (rx_nest.cu) when the configuration fits it; else the two-kernel version of rx_fast.cu
    PfbDesign &N = ctx->nfd;
    NestPlan &K = ctx->NP;
    if (!std::getenv("BTB200_NO_NEST") && N.design_noise(P, NEST_NCOL, 16) == 0 && std::fabs(N.phi - phi) < 1e-9) {
      K.M = N.M; K.D = N.D; K.Q = N.Q; K.q_rows = N.q_rows; K.N1 = N.N1; K.N2 = N.N2; K.CPC = N.CPC; K.ncol = N.ncol;
      K.nch = P.nch; K.S = P.S; K.fns = P.fns; K.n_noise = P.n_noise;
      // throughput mode: the even outputs with end-corrected weights (half the work, rx_nest.cuh); the guarded mode of
      // the exact path keeps every output (its guard band assumes the tighter estimate)
      std::vector<float> wts;
      // BTB200_NEST_FOLD: 0 = every output, 1 = the even outputs, 2 = every 4th output (default where it fits)
      int fold_req = 2;
      if (const char *e = std::getenv("BTB200_NEST_FOLD")) fold_req = std::atoi(e);
      if (std::getenv("BTB200_NEST_DENSE")) fold_req = 0;
      const bool sub_ok = ctx->poly && (P.n_noise % 2) == 0 && P.n_noise >= 32;
      if (sub_ok && fold_req >= 2 && P.n_noise >= 128 && NEST_RUNS_V * 2 * N.M <= 2 * NEST_K * 100) {
        // every 4th output, weights by least squares over the band of |y|^2 (+-90 kHz of the 2 Msps output rate)
        K.fold = 2;
        K.stride = 4;
        const double omega = 2.0 * M_PI * 90e3 * P.D / P.fs;
        if (nest_quadrature(P.n_noise, K.stride, 2, 12, omega, wts) < 0) return BTB200_ERR_ARG;
        K.n_used = (int)wts.size();
        const int MV = K.fold * N.M, Qv = (P.Nn + MV - 1) / MV;
        K.q_rows_v = (Qv + 15 + 15) / 16 * 16;
        K.tiles_per_slot = (K.n_used + NEST_R * NEST_RUNS_V - 1) / (NEST_R * NEST_RUNS_V);
        const char *v2e = std::getenv("BTB200_NEST_V2");
        if (N.N1 == 4 && N.M == 100 && !(v2e && v2e[0] == '0')) {
          // small blocks, three per SM (rx_nest.cu: k_nest2): single runs of 16 outputs, chunks of 8 tap rows
          K.v2 = 1;
          K.q_rows_v = (Qv + 15 + 7) / 8 * 8;
          K.tiles_per_slot = (K.n_used + NEST_R - 1) / NEST_R;
        }
      } else if (sub_ok && fold_req >= 1) {
        K.stride = 2;
        K.n_used = P.n_noise / 2 + 1;
        wts.assign((size_t)K.n_used, 2.0f);
        const size_t n = wts.size();
        wts[0] = 1.3125f; wts[1] = 2.25f; wts[2] = 1.9375f;
        wts[n - 3] = 1.9375f; wts[n - 2] = 2.25f; wts[n - 1] = 0.3125f;
        K.tiles_per_slot = (K.n_used + 2 * NEST_TO - 1) / (2 * NEST_TO);
      } else {
        K.stride = 1;
        K.tiles_per_slot = ((P.n_noise + 1) / 2 + NEST_TO - 1) / NEST_TO;
      }
      K.period = F.period; K.phasor = F.phasor; K.esum = F.esum;
      // the last tile of a slot reads (outputs of the tiles + tap rows + the ring's look-ahead) rows of K.fold * M samples
      // past the slot's first noise sample
      const long reach = K.v2
          ? (long)P.fns + ((long)K.tiles_per_slot * NEST_R + K.q_rows_v + 8 * 4) * K.fold * K.M
          : K.fold > 1
          ? (long)P.fns + ((long)K.tiles_per_slot * NEST_R * NEST_RUNS_V + K.q_rows_v + 16 * (NEST_RUNS_V + 1)) * K.fold * K.M
          : (long)P.fns + ((long)K.tiles_per_slot * (K.stride == 2 ? 2 : 1) * NEST_TO + K.q_rows + 16 * (2 * NEST_K + 2)) * K.M;
      if (K.stride >= 2) { if ((rc = upload(ctx, &K.weights, wts))) return rc; }
      ctx->nest_reach = reach;
      if (K.v2) {
        std::vector<float> h1((size_t)K.q_rows_v * K.fold * K.M, 0.0f);
        for (size_t i = 0; i < N.hq.size() && i < h1.size(); i++) h1[i] = N.hq[i];
        for (size_t i = h1.size(); i < N.hq.size(); i++) if (N.hq[i] != 0.0f) return BTB200_ERR_ARG;   // every tap inside the rows read
        if ((rc = upload(ctx, &K.hq1, h1))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_smflag, 256))) return rc;
        CK(cudaMemset(ctx->d_smflag, 0, 256 * sizeof(int)));
      }
      bool nest_ok = reach <= P.H && nest_setup(K) == 0;
      if (!nest_ok && K.v2) {
        // the small-block variant does not fit: the 48-output tiles of k_nest
        const int MV = K.fold * N.M, Qv = (P.Nn + MV - 1) / MV;
        K.v2 = 0;
        K.q_rows_v = (Qv + 15 + 15) / 16 * 16;
        K.tiles_per_slot = (K.n_used + NEST_R * NEST_RUNS_V - 1) / (NEST_R * NEST_RUNS_V);
        ctx->nest_reach = (long)P.fns + ((long)K.tiles_per_slot * NEST_R * NEST_RUNS_V + K.q_rows_v + 16 * (NEST_RUNS_V + 1)) * K.fold * K.M;
        nest_ok = ctx->nest_reach <= P.H && nest_setup(K) == 0;
      }
      if (nest_ok) {
        {
          // flat in the tap index k = row * (row length) + branch: the folded mode reads the same array with rows of fold * M
          size_t n_h2 = N.hq.size();
          if (K.fold > 1 && (size_t)K.q_rows_v * K.fold * K.M > n_h2) n_h2 = (size_t)K.q_rows_v * K.fold * K.M;
          std::vector<float2> h2(n_h2, make_float2(0.0f, 0.0f));
          for (size_t i = 0; i < N.hq.size(); i++) h2[i] = make_float2(N.hq[i], N.hq[i]);
          if ((rc = upload(ctx, &K.hq2, h2))) return rc;
        }
        if ((rc = upload(ctx, &K.n2_of_rho, N.n2_of_rho))) return rc;
        if ((rc = upload_raw<c32>(ctx, &K.WB, N.WB.data(), N.WB.size()))) return rc;
        if ((rc = upload(ctx, &K.col_chan, N.col_chan))) return rc;
        if ((rc = upload(ctx, &K.chan_col, N.chan_col))) return rc;
        if ((rc = dev_alloc(ctx, &K.xr, (B - 1) * (size_t)P.S + P.H))) return rc;
        if ((rc = dev_alloc(ctx, &K.E2, B * K.tiles_per_slot * (size_t)K.ncol))) return rc;
        ctx->use_nest = true;
      }
    }
  }
  if (!ctx->use_nest) { if ((rc = dev_alloc(ctx, &F.U, B * P.n_noise * (size_t)M))) return rc; }
  CK(cudaMallocHost(&ctx->h_esum, B * P.nch * sizeof(double)));
  ctx->phi = phi;
  ctx->fast_snr = true;
  return 0;
}

// off-channel energy estimate of every window of the batch -> F.esum
void enqueue_noise_estimate(btb200_ctx *ctx, const Geom &G, const DevBatch &W, long n_samples, cudaStream_t s,
                            const NestResume *resume = nullptr)
{
  if (ctx->use_nest) {
    const bool rotated = ctx->poly && ctx->PF.xr == ctx->NP.xr;   // the channelizer already wrote the rotated copy
    if (!rotated) launch_nest_prerot(ctx->NP, W.x, n_samples, s);
    launch_nest(ctx->NP, W.B, s, resume);
    ctx->launches += rotated ? 2 : 3;
  } else {
    launch_noise_fast(ctx->F, W.x, W.B, G.S, G.fns, G.D, G.n_noise, G.nch, s);
    ctx->launches += 2;
  }
}

// the pass flags of a lazy-squelch batch: all ones, or the caller's window mask (consumed)
int enqueue_pass_flags(btb200_ctx *ctx, const DevBatch &W, size_t nbc, cudaStream_t s)
{
  if (ctx->win_mask.size() == nbc) {
    if (!ctx->h_mask) CK(cudaMallocHost(&ctx->h_mask, (size_t)ctx->max_slots * ctx->plan.nch * sizeof(int)));
    std::memcpy(ctx->h_mask, ctx->win_mask.data(), nbc * sizeof(int));
    CK(cudaMemcpyAsync(W.pass, ctx->h_mask, nbc * sizeof(int), cudaMemcpyHostToDevice, s));
  } else {
    launch_fill_pass(W, (int)nbc, 1, s);
  }
  ctx->win_mask.clear();
  ctx->launches++;
  return 0;
}

// Tables and buffers of the polyphase channelizer (rx_pfb.cu)
int setup_pfb(btb200_ctx *ctx)
{
  const Plan &P = ctx->plan;
  PfbDesign &F = ctx->pfd;
  if (F.design(P, PFB_T, PFB_TT, PFB_NCOL) != 0) {
    ctx->last_error = "polyphase mode needs an even integer number of samples per MHz and <= 8 taps per branch";
    return BTB200_ERR_ARG;
  }
  PfbPlan &K = ctx->PF;
  K.M = F.M; K.D = F.D; K.Q = F.Q; K.N1 = F.N1; K.N2 = F.N2;
  K.gps = P.grid_per_slot; K.n_ddc = P.n_ddc; K.nfull = F.nfull; K.rem = F.rem;
  K.fcs = P.fcs; K.nch = P.nch; K.tps = F.tps; K.CPC = F.CPC; K.ncol = F.ncol; K.span = F.span;
  K.gain = P.demod_gain;
  K.phi_step = (float)(-2.0 * F.phi / F.M);
  int rc;
  {
    std::vector<unsigned char> blob(pfb_table_bytes(K));
    pfb_pack_tables(K, reinterpret_cast<const c32 *>(F.WB.data()), F.hq.data(), P.atan_tab.data(),
                    reinterpret_cast<const c32 *>(F.kappa.data()), F.col_chan.data(), F.n2_of_rho.data(), blob.data());
    if ((rc = upload(ctx, &K.tables, blob))) return rc;
  }
  if ((rc = upload(ctx, &K.chan_col, F.chan_col))) return rc;
  const size_t B = ctx->max_slots;
  const size_t Gtot = (B - 1) * (size_t)P.grid_per_slot + P.n_ddc;
  if ((rc = dev_alloc(ctx, &K.dem, Gtot * P.nch))) return rc;
  if (!std::getenv("BTB200_NO_DEMC")) {
    // channel-major copy for the resume of the clock-recovery chains (rx_mm.cuh, CM); rows padded so that the
    // 16-byte groups around a window's ends stay inside the allocation
    K.pitchC = (long)((Gtot + 3) / 4 * 4 + 8);
    if ((rc = dev_alloc(ctx, &K.demC, (size_t)K.pitchC * P.nch))) return rc;
  }
  if ((rc = dev_alloc(ctx, &K.E, (size_t)pfb_tiles(K, (int)B) * K.ncol * 2))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_eon_all, B * P.nch))) return rc;
  CK(cudaMallocHost(&ctx->h_eon_all, B * P.nch * sizeof(double)));
  if (pfb_setup(K) != 0) { ctx->last_error = "polyphase channelizer: configuration outside the kernel's limits"; return BTB200_ERR_ARG; }
  // device-driven tail (rx_kernels.cuh, TailBufs)
  TailBufs &Tb = ctx->W.tail;
  const size_t nw = B * P.nch;
  if ((rc = dev_alloc(ctx, &Tb.stage, nw * TAIL_MAXW))) return rc;
  if ((rc = dev_alloc(ctx, &Tb.cnt, nw))) return rc;
  if ((rc = dev_alloc(ctx, &Tb.base, nw))) return rc;
  if ((rc = dev_alloc(ctx, &Tb.list, nw))) return rc;
  if ((rc = dev_alloc(ctx, &Tb.n_list, 1))) return rc;
  if ((rc = dev_alloc(ctx, &Tb.sorted, (size_t)kHitCap))) return rc;
  return 0;
}

// One compute stream per device, shared by every context: the kernels of this path are sized to fill the
// GPU (one block per SM), so batches of different contexts gain nothing from running concurrently -- measured:
// two contexts on private streams 29.6 ms per batch pair step against 24.5 ms serialised -- while issue order on
// one stream gives the pipeline of the double-buffered submit()/collect() loop for free:
// K(k), K(k+1), deferred-noise(k), K(k+2), deferred-noise(k+1), ...   BTB200_PRIVATE_STREAM=1 restores private streams.
cudaStream_t device_stream(int device)
{
  static std::mutex mu;
  static cudaStream_t streams[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  cudaStream_t &s = streams[device & 63];
  if (!s && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
  return s;
}

// wait for everything this context has enqueued so far (not for what other contexts enqueue later)
int sync_here(btb200_ctx *ctx)
{
  cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev_sync);
  if (e != cudaSuccess) {
    ctx->last_error = std::string("sync: ") + cudaGetErrorString(e);
    return BTB200_ERR_CUDA;
  }
  return 0;
}
#define SYNC_HERE() do { if (int rc_ = sync_here(ctx)) return rc_; } while (0)

int setup(btb200_ctx *ctx)
{
  const Plan &P = ctx->plan;
  Geom &G = ctx->G;
  G.S = P.S; G.H = P.H; G.D = P.D; G.Nc = P.Nc; G.Nn = P.Nn; G.fcs = P.fcs; G.fns = P.fns;
  G.nch = P.nch; G.n_ddc = P.n_ddc; G.n_noise = P.n_noise; G.n_dem = P.n_dem; G.gps = P.grid_per_slot;
  G.n_dem_pad = (P.n_dem + 3) & ~3;
  G.bw = (P.n_dem + 31) / 32 + 4;
  G.ch_lo = P.ch_lo;
  G.demod_gain = P.demod_gain;
  G.mm = MmConst{P.gain_mu, P.gain_omega, P.omega_mid, P.omega_lim};
  G.mu0 = P.mu0;
  G.squelch_db = P.squelch_db;
  G.search = ctx->cfg.search;
  G.stateless = ctx->cfg.mm_mode == BTB200_MM_STATELESS;
  G.early = 0; G.ne_dem = P.n_dem; G.sym_target = P.n_dem;
  G.dem_grid = 0; G.dem_rows = G.n_dem_pad;
  ctx->poly = ctx->cfg.ddc_mode == BTB200_DDC_POLYPHASE;
  if (ctx->poly && (!G.stateless || ctx->cfg.squelch_mode == BTB200_SQUELCH_EAGER)) {
    ctx->last_error = "BTB200_DDC_POLYPHASE needs BTB200_MM_STATELESS and the lazy squelch";
    return BTB200_ERR_ARG;
  }
  if (ctx->poly) { G.dem_grid = 1; G.dem_rows = P.grid_per_slot; }

  if (std::getenv("BTB200_PRIVATE_STREAM")) {
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  } else {
    ctx->stream = device_stream(ctx->device);
    if (!ctx->stream) { ctx->last_error = "cannot create the device compute stream"; return BTB200_ERR_CUDA; }
  }
  CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ctx->ev_up, cudaEventDisableTiming));
  for (auto &e : ctx->ev_part) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : ctx->ev) CK(cudaEventCreate(&e));
  for (auto &e : ctx->evl) CK(cudaEventCreate(&e));
  CK(cudaEventCreateWithFlags(&ctx->ev_tail, cudaEventDisableTiming));
  for (auto &e : ctx->ev_user) CK(cudaEventCreate(&e));
  CK(cudaEventCreate(&ctx->ev_res));

  int rc;
  if ((rc = upload_raw<c32>(ctx, &ctx->T.chan_rtaps, P.chan_rtaps.data(), P.chan_rtaps.size()))) return rc;
  if ((rc = upload_raw<c32>(ctx, &ctx->T.noise_rtaps, P.noise_rtaps.data(), P.noise_rtaps.size()))) return rc;
  if ((rc = upload(ctx, &ctx->T.mmse, P.mmse))) return rc;
  if ((rc = upload(ctx, &ctx->T.atan_tab, P.atan_tab))) return rc;
  if ((rc = upload(ctx, &ctx->T.ac_lut, P.ac_lut))) return rc;
  if (ctx->cfg.search & BTB200_SEARCH_BR_BCH) {
    // libbtbb-style access-code test (rx_math.cuh: br_lag_test_bch): syndrome tables of the sync word's (64,30) code
    const uint32_t bch = ctx->cfg.bch;
    BchTables bt;
    if (bt.build((int)(bch >> 28) & 7) != 0) { ctx->last_error = "BTB200_SEARCH_BR_BCH: max_ac_errors must be 0..2"; return BTB200_ERR_ARG; }
    BchDev &B = ctx->T.bch;
    if ((rc = upload(ctx, &B.par, bt.par))) return rc;
    if (bt.syn.empty()) { bt.syn.push_back(0); bt.err.push_back(0); }
    if ((rc = upload(ctx, &B.syn, bt.syn))) return rc;
    if ((rc = upload(ctx, &B.err, bt.err))) return rc;
    B.n = bt.max_err ? (int)bt.syn.size() : 0;
    B.max_err = bt.max_err;
    B.lap = (bch & (1u << 24)) ? (bch & 0xffffffu) : 0xffffffffu;
    B.target = (bch & (1u << 24)) ? sync_word(bch & 0xffffffu) : 0;
  }
  {
    // channel-group-interleaved tap banks for the tiled FIR: [group][k][16], zero taps for padding channels
    const int ng = (P.nch + 15) / 16;
    std::vector<c32> tg((size_t)ng * P.Nc * 16, c32{0.0f, 0.0f});
    for (int c = 0; c < P.nch; c++)
      for (int k = 0; k < P.Nc; k++) {
        const cf32 t = P.chan_rtaps[(size_t)c * P.Nc + k];
        tg[((size_t)(c / 16) * P.Nc + k) * 16 + (c % 16)] = c32{t.re, t.im};
      }
    if ((rc = upload(ctx, &ctx->T.chan_tg, tg))) return rc;
    {
      std::vector<float> t4(tg.size() * 4);
      for (size_t i = 0; i < tg.size(); i++) { t4[4 * i] = t4[4 * i + 1] = tg[i].re; t4[4 * i + 2] = t4[4 * i + 3] = tg[i].im; }
      const float *d4 = nullptr;
      if ((rc = upload(ctx, &d4, t4))) return rc;
      ctx->T.chan_tg4 = d4;
    }
    tg.assign((size_t)ng * P.Nn * 16, c32{0.0f, 0.0f});
    for (int c = 0; c < P.nch; c++)
      for (int k = 0; k < P.Nn; k++) {
        const cf32 t = P.noise_rtaps[(size_t)c * P.Nn + k];
        tg[((size_t)(c / 16) * P.Nn + k) * 16 + (c % 16)] = c32{t.re, t.im};
      }
    if ((rc = upload(ctx, &ctx->T.noise_tg, tg))) return rc;
    if (ctx->cfg.mm_mode == BTB200_MM_STATELESS && ctx->cfg.squelch_mode != BTB200_SQUELCH_EAGER && P.D == 50) {
      // delay-line noise FIR (rx_firdl.cu): rotated taps with both parts duplicated, (c, c, d, d)
      std::vector<float> t4((size_t)P.nch * P.Nn * 4);
      for (size_t i = 0; i < (size_t)P.nch * P.Nn; i++) {
        const cf32 t = P.noise_rtaps[i];
        t4[4 * i] = t4[4 * i + 1] = t.re; t4[4 * i + 2] = t4[4 * i + 3] = t.im;
      }
      const float *d4 = nullptr;
      if ((rc = upload(ctx, &d4, t4))) return rc;
      ctx->T.noise_taps4 = d4;
    }
    if (fir_setup(ctx->device) != 0) { ctx->last_error = "cannot opt in to large dynamic shared memory"; return BTB200_ERR_CUDA; }
  }
  std::vector<uint8_t> hdr(4 * 256);
  for (int w = 0; w < 4; w++)
    for (int v = 0; v < 256; v++) hdr[w * 256 + v] = (uint8_t)le_hdr_dist((uint32_t)v, w);
  if ((rc = upload(ctx, &ctx->T.le_hdr_lut, hdr))) return rc;
  if ((rc = upload(ctx, &ctx->T.le_index, P.le_index))) return rc;
  std::vector<uint32_t> white(P.nch, 0);
  for (int c = 0; c < P.nch; c++)
    for (int i = 0; i < 16; i++) white[c] |= (uint32_t)P.le_white16[(size_t)c * 16 + i] << i;
  if ((rc = upload(ctx, &ctx->T.le_white, white))) return rc;

  const size_t B = ctx->max_slots;
  const size_t nch = P.nch;
  const uint32_t sq = ctx->cfg.squelch_mode;    // 0 default (lazy), 1 eager, 2 lazy
  ctx->lazy = G.stateless && sq != 1;
  ctx->x_cap = (B - 1) * (size_t)P.S + P.H;
  if ((rc = dev_alloc(ctx, &ctx->d_x, ctx->x_cap))) return rc;
  DevBatch &W = ctx->W;
  if (!ctx->poly) { if ((rc = dev_alloc(ctx, &W.Y, ((B - 1) * P.grid_per_slot + P.n_ddc) * nch))) return rc; }
  if (!ctx->lazy) { if ((rc = dev_alloc(ctx, &W.Nz, B * P.n_noise * nch))) return rc; }
  else {
    ctx->group_cap = B * ((nch + 1) / 2);
    ctx->list_cap = B * nch;
    if ((rc = dev_alloc(ctx, &ctx->d_NzL, ctx->group_cap * P.n_noise * LAZY_CG))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_groups, ctx->group_cap * (1 + LAZY_CG)))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_list, ctx->list_cap * 4))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_eon, ctx->list_cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d_eoff, ctx->list_cap))) return rc;
  }
  ctx->early = false;
  if (ctx->lazy && ctx->cfg.tail_mode == BTB200_TAIL_LAZY && !ctx->cfg.keep_stages) {
    // Lazy tail.  The access-code search only looks at lags < 625 (72-symbol windows), i.e. at the first 697
    // symbols of a window; everything after that is payload for ac()/aa() and only matters for windows with a
    // hit.  Bound the loop's input advance per symbol rigorously: |out| <= max_imu sum|taps| * gain*pi,
    // |mm_val| <= 2|out|, advance <= floor(1 + omega_max + gain_mu*|mm_val|).
    double tsum = 0;
    for (int i = 0; i < 129; i++) {
      double a = 0;
      for (int k = 0; k < 8; k++) a += std::fabs((double)P.mmse[(size_t)i * 8 + k]);
      tsum = std::max(tsum, a);
    }
    const double outmax = 1.001 * tsum * std::fabs((double)P.demod_gain) * 3.1415927;
    const double adv = 1.0 + (double)P.omega_mid + (double)P.omega_lim + std::fabs((double)P.gain_mu) * 2.0 * outmax;
    const int max_adv = (int)std::floor(adv * 1.0001);      // mu < 1 is the "1.0 +" above
    const int sym_target = 704;                       // 625 + 72 rounded up to whole words
    const long ne = (long)sym_target * max_adv + 16;
    const long nsym_lb = std::min<long>(P.n_dem, (P.n_dem - 8) / max_adv);
    // both search limits must be 625 whatever the (unknown) symbol count: nsym >= 625 + 68 + 692
    if (ne + 64 < P.n_dem && nsym_lb >= 1400 && G.bw >= 32) {
      ctx->early = true;
      G.ne_dem = (int)ne; G.sym_target = sym_target;
    }
  }
  if (ctx->early) {
    void *sv = nullptr;
    if ((rc = dev_alloc(ctx, reinterpret_cast<char **>(&sv), B * nch * 24))) return rc;
    W.mm_save = sv;
    if ((rc = dev_alloc(ctx, &ctx->d_list2, ctx->list_cap * 4))) return rc;
    CK(cudaMallocHost(&ctx->h_list2, ctx->list_cap * 4 * sizeof(int)));
    CK(cudaMallocHost(&ctx->h_nsym, B * nch * sizeof(int)));
    // high priority: its few blocks take the first SMs the deferred noise FIR frees, and finish under it
    int pr_lo = 0, pr_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
    CK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, pr_hi));
  }
  if (ctx->poly && !ctx->stream2) {
    int pr_lo = 0, pr_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
    CK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, pr_hi));
  }
  const size_t bp = G.stateless ? 1 : B;
  if ((rc = dev_alloc(ctx, &ctx->d_phc, bp * P.n_ddc * nch))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->d_phn, bp * P.n_noise * nch))) return rc;
  if ((rc = dev_alloc(ctx, &W.energy, B * nch))) return rc;
  if ((rc = dev_alloc(ctx, &W.noise, B * nch))) return rc;
  if ((rc = dev_alloc(ctx, &W.pass, B * nch))) return rc;
  // demod floats only travel through HBM in chained mode (separate demod kernel) or for debug taps
  // demod floats: [b][c][i] rows in chained mode (separate kernels), transposed [b][i][c] in stateless mode
  if (ctx->poly) {
    if ((rc = setup_pfb(ctx))) return rc;
    ctx->d_dem = ctx->PF.dem;
  } else if ((rc = dev_alloc(ctx, &ctx->d_dem, B * nch * G.n_dem_pad))) return rc;
  W.dem = G.stateless ? nullptr : ctx->d_dem;
  if (ctx->cfg.keep_stages) { if ((rc = dev_alloc(ctx, &ctx->d_soft, B * nch * G.n_dem_pad))) return rc; }
  if ((rc = dev_alloc(ctx, &W.bits, B * nch * G.bw))) return rc;
  if ((rc = dev_alloc(ctx, &W.nsym, B * nch))) return rc;
  if ((rc = dev_alloc(ctx, &W.mm_state, 1))) return rc;
  if ((rc = dev_alloc(ctx, &W.hits, kHitCap))) return rc;
  unsigned *cnt = nullptr;
  if ((rc = dev_alloc(ctx, &cnt, 4))) return rc;
  W.hit_count = cnt;
  W.arena_used = reinterpret_cast<unsigned long long *>(cnt + 2);
  if ((rc = dev_alloc(ctx, &W.arena, kArenaCap))) return rc;
  W.hit_cap = kHitCap;
  W.arena_cap = kArenaCap;
  W.soft = ctx->d_soft;
  W.phc = ctx->d_phc;
  W.phn = ctx->d_phn;
  W.bp_stride = G.stateless ? 0 : 1;

  CK(cudaMallocHost(&ctx->h_energy, B * nch * sizeof(double)));
  CK(cudaMallocHost(&ctx->h_noise, B * nch * sizeof(double)));
  CK(cudaMallocHost(&ctx->h_pass, B * nch * sizeof(int)));
  CK(cudaMallocHost(&ctx->h_counts, 4 * sizeof(unsigned)));
  CK(cudaMallocHost(&ctx->h_hits, (size_t)kHitCap * sizeof(DevHit)));
  CK(cudaMallocHost(&ctx->h_arena, kArenaCap));
  if (ctx->lazy) {
    CK(cudaMallocHost(&ctx->h_groups, ctx->group_cap * (1 + LAZY_CG) * sizeof(int)));
    CK(cudaMallocHost(&ctx->h_list, ctx->list_cap * 4 * sizeof(int)));
    CK(cudaMallocHost(&ctx->h_eon, ctx->list_cap * sizeof(double)));
    CK(cudaMallocHost(&ctx->h_eoff, ctx->list_cap * sizeof(double)));
  }
  ctx->h_ph_cap = bp * (size_t)P.n_ddc * nch;
  CK(cudaMallocHost(&ctx->h_ph, ctx->h_ph_cap * sizeof(c32)));

  if (ctx->lazy && (ctx->cfg.snr_mode == BTB200_SNR_FAST_GUARDED || ctx->poly)) {
    int rcf = setup_fast(ctx);
    if (rcf) return rcf;
    if (ctx->poly && !ctx->fast_snr) {
      ctx->last_error = "polyphase mode: the noise estimator does not fit this configuration";
      return BTB200_ERR_ARG;
    }
    // throughput mode: the channelizer writes the estimator's rotated copy of the input along the way (rx_pfb.cuh)
    // when its tiles cover every sample the estimator reads of the batch's last window
    if (ctx->poly && ctx->use_nest && ctx->nest_reach > 0 && !std::getenv("BTB200_NO_FUSED_PREROT") &&
        (long)P.fcs + (long)P.n_ddc * P.D >= ctx->nest_reach) {
      ctx->PF.xr = ctx->NP.xr; ctx->PF.phasor = ctx->NP.phasor; ctx->PF.period = ctx->NP.period;
    }
  }
  reset_stream_state(ctx);
  if (G.stateless) {
    // one rotator table per DDC object, restarted at phase 1 for every window
    std::vector<Rotator> rc_ = ctx->rot_c, rn_ = ctx->rot_n;
    for (int c = 0; c < P.nch; c++) rc_[c].generate(reinterpret_cast<cf32 *>(ctx->h_ph) + c, P.n_ddc, P.nch);
    CK(cudaMemcpy(ctx->d_phc, ctx->h_ph, (size_t)P.n_ddc * nch * sizeof(c32), cudaMemcpyHostToDevice));
    for (int c = 0; c < P.nch; c++) rn_[c].generate(reinterpret_cast<cf32 *>(ctx->h_ph) + c, P.n_noise, P.nch);
    CK(cudaMemcpy(ctx->d_phn, ctx->h_ph, (size_t)P.n_noise * nch * sizeof(c32), cudaMemcpyHostToDevice));
  }
  return 0;
}

void teardown(btb200_ctx *ctx)
{
  if (ctx->stream && ctx->ev_sync) sync_here(ctx);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
  if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
  if (ctx->ev_up) cudaEventDestroy(ctx->ev_up);
  for (auto &e : ctx->ev_part) if (e) cudaEventDestroy(e);
  for (void *p : ctx->allocs) cudaFree(p);
  for (void *p : {(void *)ctx->h_energy, (void *)ctx->h_noise, (void *)ctx->h_pass, (void *)ctx->h_counts,
                  (void *)ctx->h_hits, (void *)ctx->h_arena, (void *)ctx->h_ph, (void *)ctx->h_groups,
                  (void *)ctx->h_list, (void *)ctx->h_eon, (void *)ctx->h_eoff, (void *)ctx->h_esum,
                  (void *)ctx->h_list2, (void *)ctx->h_nsym, (void *)ctx->h_eon_all, (void *)ctx->h_mask})
    if (p) cudaFreeHost(p);
  for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto &e : ctx->evl) if (e) cudaEventDestroy(e);
  if (ctx->ev_tail) cudaEventDestroy(ctx->ev_tail);
  for (auto &e : ctx->ev_user) if (e) cudaEventDestroy(e);
  if (ctx->ev_res) cudaEventDestroy(ctx->ev_res);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
}

}  // namespace

extern "C" {

const char *btb200_version(void) { return "btb200 0.1 (sm_100a)"; }

const char *btb200_strerror(int err)
{
  switch (err) {
    case BTB200_OK: return "ok";
    case BTB200_ERR_ARG: return "bad argument or configuration";
    case BTB200_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case BTB200_ERR_CUDA: return "CUDA error";
    case BTB200_ERR_NOMEM: return "out of device memory";
    case BTB200_ERR_SHORT_INPUT: return "input shorter than (n_slots-1)*S + H samples";
    case BTB200_ERR_TOO_MANY: return "n_slots exceeds max_slots_per_call";
    case BTB200_ERR_OVERFLOW: return "hit or symbol buffer too small";
    default: return "unknown error";
  }
}

const char *btb200_last_error(const btb200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int btb200_create(const btb200_config *cfg, btb200_ctx **out)
{
  if (!cfg || !out || cfg->abi_version != BTB200_ABI_VERSION) return BTB200_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = e != cudaSuccess ? cudaGetErrorString(e) : "no such device";
    return BTB200_ERR_NO_DEVICE;
  }
  btb200_ctx *ctx = new (std::nothrow) btb200_ctx();
  if (!ctx) return BTB200_ERR_NOMEM;
  ctx->cfg = *cfg;
  if (ctx->cfg.search == 0) ctx->cfg.search = BTB200_SEARCH_BR | BTB200_SEARCH_LE;
  if (ctx->cfg.search & BTB200_SEARCH_BR_BCH) ctx->cfg.search |= BTB200_SEARCH_BR;
  if (ctx->cfg.extra_history_symbols == 0) ctx->cfg.extra_history_symbols = 3125;
  ctx->max_slots = cfg->max_slots_per_call ? cfg->max_slots_per_call : kDefaultMaxSlots;
  ctx->device = cfg->device;
  const int drc = ctx->plan.design(cfg->sample_rate, cfg->center_freq, cfg->squelch_threshold,
                                   (int)ctx->cfg.extra_history_symbols);
  if (drc != 0 || (cfg->mm_mode != BTB200_MM_CHAINED && cfg->mm_mode != BTB200_MM_STATELESS) || cfg->device >= 64) {
    g_create_error = drc == -2 ? "sample rate not supported: 625 * fs / 1 MHz must be a multiple of the decimation (int)(fs / 1 MHz) / 2"
                               : "bad configuration";
    delete ctx;
    return BTB200_ERR_ARG;
  }
  int rc = BTB200_OK;
  if (cudaSetDevice(ctx->device) != cudaSuccess) rc = BTB200_ERR_NO_DEVICE;
  if (!rc) {
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, ctx->device);
    ctx->sm_count = prop.multiProcessorCount;
    rc = setup(ctx);
  }
  if (rc) {
    g_create_error = ctx->last_error;
    teardown(ctx);
    delete ctx;
    return rc;
  }
  *out = ctx;
  return BTB200_OK;
}

void btb200_destroy(btb200_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  teardown(ctx);
  delete ctx;
}

int btb200_get_info(const btb200_ctx *ctx, btb200_info *o)
{
  if (!ctx || !o) return BTB200_ERR_ARG;
  const Plan &P = ctx->plan;
  o->samples_per_slot = P.S; o->history = P.H; o->decimation = P.D;
  o->chan_taps = P.Nc; o->noise_taps = P.Nn;
  o->first_channel_sample = P.fcs; o->first_noise_sample = P.fns;
  o->channel_low = P.ch_lo; o->channel_high = P.ch_hi; o->n_channels = P.nch;
  o->ddc_out_per_window = P.n_ddc; o->noise_out_per_window = P.n_noise;
  o->demod_gain = P.demod_gain; o->omega_mid = P.omega_mid;
  o->max_slots_per_call = ctx->max_slots; o->sm_count = ctx->sm_count;
  return BTB200_OK;
}

int btb200_host_alloc(void **ptr, size_t bytes)
{
  if (!ptr) return BTB200_ERR_ARG;
  return cudaMallocHost(ptr, bytes) == cudaSuccess ? BTB200_OK : BTB200_ERR_NOMEM;
}
void btb200_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int btb200_get_mm_state(const btb200_ctx *ctx, float mm[3])
{
  if (!ctx || !mm) return BTB200_ERR_ARG;
  mm[0] = ctx->mm.mu; mm[1] = ctx->mm.omega; mm[2] = ctx->mm.last;
  return BTB200_OK;
}
int btb200_set_mm_state(btb200_ctx *ctx, const float mm[3])
{
  if (!ctx || !mm) return BTB200_ERR_ARG;
  ctx->mm = MmState{mm[0], mm[1], mm[2]};
  return BTB200_OK;
}
int btb200_reset(btb200_ctx *ctx)
{
  if (!ctx) return BTB200_ERR_ARG;
  // a batch still in flight (submitted, perhaps begun) is abandoned: let its work drain first
  if (ctx->pending) { sync_here(ctx); if (ctx->stream2) cudaStreamSynchronize(ctx->stream2); }
  reset_stream_state(ctx);
  ctx->pending = false;
  ctx->cb.begun = false;
  return BTB200_OK;
}

namespace {
enum InKind { IN_HOST_C32 = 0, IN_DEV_C32 = 1, IN_HOST_I16 = 2, IN_DEV_I16 = 3 };
}

static int submit_impl(btb200_ctx *ctx, const void *iq, int kind, size_t n_samples, uint64_t first_slot, uint32_t n_slots)
{
  if (!ctx || !iq || n_slots == 0) return BTB200_ERR_ARG;
  if (ctx->pending) return BTB200_ERR_ARG;
  if (n_slots > ctx->max_slots) return BTB200_ERR_TOO_MANY;
  const Plan &P = ctx->plan;
  Geom G = ctx->G;
  // lazy tail unless the stream is so busy that most windows would be resumed anyway (same results either way)
  G.early = (ctx->early && ctx->impl == IMPL_TUNED && !(ctx->poly && ctx->dense_tail)) ? 1 : 0;
  if (!G.early) { G.ne_dem = P.n_dem; G.sym_target = P.n_dem; }
  const size_t need = (size_t)(n_slots - 1) * P.S + P.H;
  if (n_samples < need) return BTB200_ERR_SHORT_INPUT;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBatch W = ctx->W;
  W.B = (int)n_slots;
  const size_t nbc = (size_t)n_slots * P.nch;
  const bool host = kind == IN_HOST_C32 || kind == IN_HOST_I16, i16 = kind == IN_HOST_I16 || kind == IN_DEV_I16;
  if (kind == IN_HOST_I16 && !ctx->d_x16) {
    int rc = dev_alloc(ctx, &ctx->d_x16, 2 * ctx->x_cap);
    if (rc) return rc;
  }

  // The front end (channel FIR of the exact mode / polyphase channelizer) runs in tile ranges so that it can follow
  // the input copy piece by piece.
  bool front_done = false;
  const long ntile = ctx->poly ? pfb_tiles(ctx->PF, (int)n_slots) : chan_fir_tiles(G, W);
  auto front_range = [&](long t0, long t1) {
    if (ctx->poly) launch_pfb(ctx->PF, W.x, (long)need, (int)n_slots, t0, t1, s);
    else launch_chan_fir_range(G, ctx->T, W, ctx->impl, t0, t1, s);
    ctx->launches++;
  };
  auto front_samples = [&](long t1) { return ctx->poly ? pfb_samples(ctx->PF, t1) : chan_fir_samples(G, t1); };
  if (kind == IN_DEV_C32) {
    CK(cudaEventRecord(ctx->ev[0], s));
    W.x = reinterpret_cast<const c32 *>(iq);
  } else {
    // host input: the copy runs on its own stream (it overlaps the compute stream's current work); compute waits for
    // it.  With the tuned kernels it is cut in kCopyParts pieces and the front end follows piece by piece, so only
    // the first piece of a batch is exposed when nothing else is queued (blocking btb200_process callers).
    // int16 input is converted to complex64 on the device (exact), piece by piece as well.
    CK(cudaEventRecord(ctx->ev[0], host ? ctx->copy_stream : s));
    W.x = ctx->d_x;
    const int parts = (host && ctx->impl != IMPL_BASELINE && G.stateless && ntile >= 64) ? kCopyParts : 1;
    size_t copied = 0;
    for (int c = 0; c < parts; c++) {
      const long t0 = ntile * c / parts, t1 = ntile * (c + 1) / parts;
      size_t upto = need;
      if (c != parts - 1) { upto = std::min<size_t>(need, ((size_t)front_samples(t1) + 3) & ~(size_t)3); }
      if (upto > copied) {
        const size_t n = upto - copied;
        if (kind == IN_HOST_C32)
          CK(cudaMemcpyAsync(ctx->d_x + copied, reinterpret_cast<const c32 *>(iq) + copied, n * sizeof(c32),
                             cudaMemcpyHostToDevice, ctx->copy_stream));
        else if (kind == IN_HOST_I16)
          CK(cudaMemcpyAsync(ctx->d_x16 + 2 * copied, reinterpret_cast<const int16_t *>(iq) + 2 * copied, n * 4,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
      }
      if (host) {
        CK(cudaEventRecord(ctx->ev_part[c], ctx->copy_stream));
        CK(cudaStreamWaitEvent(s, ctx->ev_part[c], 0));
      }
      if (upto > copied && i16) {
        const int16_t *src = (kind == IN_HOST_I16 ? ctx->d_x16 : reinterpret_cast<const int16_t *>(iq)) + 2 * copied;
        launch_i16_to_c32(src, ctx->d_x + copied, (long)(upto - copied), s);
        ctx->launches++;
      }
      if (upto > copied) copied = upto;
      if (parts > 1) front_range(t0, t1);
    }
    front_done = parts > 1;
    if (host) CK(cudaEventRecord(ctx->ev_h2d, ctx->copy_stream));
  }
  if (ctx->poly) {
    // throughput mode: polyphase channelizer with fused demod + energy, clock recovery of the searchable prefix,
    // access-code search, resume of the clock-recovery chains of the windows with hits, the noise estimate of every
    // window (it only feeds the squelch, which is settled on the host at collect time), then the hit list in the
    // reference's order, the arena layout and the symbol gather -- all laid out on the device, no host round trip
    // inside a batch.  Everything runs in order on the compute stream: the resume is a few dozen latency-bound warps,
    // and next to a compute-bound kernel (tried: second stream, high priority, beside the estimator or beside the
    // next batch) they are starved of issue slots and take 4-5x longer, which costs more than the overlap saves
    // (measured per step: 3.7 ms in order, 4.0 overlapped with the next batch, 6.5 beside the estimator).
    // BTB200_TAIL_STREAM2=1 keeps the overlapped variant for experiments.
    CK(cudaMemsetAsync(W.hit_count, 0, 4 * sizeof(unsigned), s));
    CK(cudaEventRecord(ctx->ev[1], s));
    if (!front_done) front_range(0, ntile);
    launch_pfb_energy(ctx->PF, (int)n_slots, ctx->d_eon_all, s); ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_eon_all, ctx->d_eon_all, nbc * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[2], s));
    if (int rcm = enqueue_pass_flags(ctx, W, nbc, s)) return rcm;
    launch_demod_mm_v2(G, ctx->T, W, ctx->d_dem, s); ctx->launches++;
    CK(cudaEventRecord(ctx->ev[3], s));
    launch_search_warp(G, ctx->T, W, s); ctx->launches++;
    launch_tail_scan(G, W, s); ctx->launches++;
    CK(cudaEventRecord(ctx->ev[4], s));
    // The resume of the windows with hits rides in the first blocks of the noise estimator's launch (rx_nest.cuh,
    // NestResume): those blocks own their SMs, so the latency-bound chains run at full speed while the other SMs work
    // through the estimator's tiles.  (A separate launch next to a compute-bound kernel starves the chains: 4-5x slower.)
    static const bool tail_inline = std::getenv("BTB200_TAIL_STREAM2") == nullptr;
    static const bool fuse_resume = std::getenv("BTB200_NO_FUSED_RESUME") == nullptr;
    cudaStream_t s2 = tail_inline ? s : ctx->stream2;
    NestResume nr{};
    const bool fused = G.early && fuse_resume && tail_inline && ctx->use_nest && nest_can_resume(ctx->NP);
    if (fused) {
      nr.G = G; nr.W = W; nr.mmse = ctx->T.mmse; nr.demT = ctx->d_dem; nr.save = W.mm_save;
      nr.demC = ctx->PF.demC; nr.pitchC = ctx->PF.pitchC;
      static const bool share_sm = std::getenv("BTB200_RESUME_SHARED") != nullptr;
      nr.sm_flag = (ctx->NP.v2 && !share_sm) ? ctx->d_smflag : nullptr;
      // chains per resume block: 64 (two warps on their SM), or 32 with BTB200_RESUME_BLK=32 (k_nest2 + channel-major copy)
      static const int resume_blk = (std::getenv("BTB200_RESUME_BLK") && std::atoi(std::getenv("BTB200_RESUME_BLK")) == 32) ? 32 : 64;
      nr.blk = (ctx->NP.v2 && ctx->PF.demC && resume_blk == 32) ? 32 : NEST_RESUME_BLK;
      nr.n_blocks = (int)((nbc + nr.blk - 1) / nr.blk);
    } else if (G.early) {
      if (!tail_inline) CK(cudaStreamWaitEvent(s2, ctx->ev[4], 0));
      launch_tail_resume(G, ctx->T, W, ctx->d_dem, s2); ctx->launches++;
      if (!tail_inline) CK(cudaEventRecord(ctx->ev_tail, s2));
    }
    CK(cudaEventRecord(ctx->ev_res, s));
    enqueue_noise_estimate(ctx, G, W, (long)need, s, fused ? &nr : nullptr);
    CK(cudaMemcpyAsync(ctx->h_esum, ctx->F.esum, nbc * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[5], s));
    if (G.early && !tail_inline) CK(cudaStreamWaitEvent(s, ctx->ev_tail, 0));
    launch_tail_finish(G, W, s); ctx->launches += 3;
    CK(cudaEventRecord(ctx->ev[6], s));
    CK(cudaMemcpyAsync(ctx->h_counts, W.hit_count, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[7], s));
    CK(cudaGetLastError());
    ctx->pending = true;
    ctx->pend_slots = n_slots;
    ctx->pend_first_slot = first_slot;
    ctx->pendW = W;
    ctx->pendG = G;
    ctx->pend_early = G.early != 0;
    return BTB200_OK;
  }
  const bool chan_fir_done = front_done;
  // chained mode advances host-side stream state (rotators) while it builds the batch: keep a copy and put it back
  // on every failure path, so that a transient CUDA error cannot desynchronise the stream
  struct Restore {
    btb200_ctx *c; std::vector<Rotator> rc, rn; MmState mm; bool armed;
    ~Restore() { if (armed) { c->rot_c = rc; c->rot_n = rn; c->mm = mm; } }
  } restore{ctx, {}, {}, ctx->mm, false};
  if (!G.stateless) { restore.rc = ctx->rot_c; restore.rn = ctx->rot_n; restore.armed = true; }
  if (!G.stateless) {
    // free-running rotators: the phases of this batch's windows, generated with
    // the same libm calls as the reference's rotator (hypotf renormalisation)
    cf32 *hp = reinterpret_cast<cf32 *>(ctx->h_ph);
    for (uint32_t b = 0; b < n_slots; b++)
      for (int c = 0; c < P.nch; c++)
        ctx->rot_c[c].generate(hp + ((size_t)b * P.n_ddc) * P.nch + c, P.n_ddc, P.nch);
    CK(cudaMemcpyAsync(ctx->d_phc, hp, (size_t)n_slots * P.n_ddc * P.nch * sizeof(c32), cudaMemcpyHostToDevice, s));
    SYNC_HERE();
    for (uint32_t b = 0; b < n_slots; b++)
      for (int c = 0; c < P.nch; c++)
        ctx->rot_n[c].generate(hp + ((size_t)b * P.n_noise) * P.nch + c, P.n_noise, P.nch);
    CK(cudaMemcpyAsync(ctx->d_phn, hp, (size_t)n_slots * P.n_noise * P.nch * sizeof(c32), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(W.mm_state, &ctx->mm, sizeof(MmState), cudaMemcpyHostToDevice, s));
  }
  CK(cudaMemsetAsync(W.hit_count, 0, 4 * sizeof(unsigned), s));
  CK(cudaEventRecord(ctx->ev[1], s));
  if (!chan_fir_done) { launch_chan_fir(G, ctx->T, W, ctx->impl, s); ctx->launches++; }
  CK(cudaEventRecord(ctx->ev[2], s));
  if (ctx->lazy) {
    // lazy squelch: every window is demodulated and searched; the squelch (and the snr
    // ac() prints) is settled exactly, afterwards, for the windows that produced hits
    if (ctx->fast_snr) {
      enqueue_noise_estimate(ctx, G, W, (long)need, s);
      CK(cudaMemcpyAsync(ctx->h_esum, ctx->F.esum, nbc * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(ctx->ev[3], s));
    if (int rcm = enqueue_pass_flags(ctx, W, nbc, s)) return rcm;
  } else {
    launch_noise_fir(G, ctx->T, W, ctx->impl, s); ctx->launches++;
    CK(cudaEventRecord(ctx->ev[3], s));
    launch_energy(G, ctx->T, W, G.stateless ? 1 : 0, s); ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_energy, W.energy, nbc * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_noise, W.noise, nbc * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (!G.stateless) {
      // chained mode: the squelch decides which windows advance the shared M&M
      // state, so it is settled with the reference's own libm arithmetic first
      SYNC_HERE();
      for (size_t i = 0; i < nbc; i++) {
        const double snr = 10.0 * std::log10(ctx->h_energy[i] / ctx->h_noise[i]);
        ctx->h_pass[i] = (snr >= P.squelch_db) ? 1 : 0;
      }
      CK(cudaMemcpyAsync(W.pass, ctx->h_pass, nbc * sizeof(int), cudaMemcpyHostToDevice, s));
    }
  }
  CK(cudaEventRecord(ctx->ev[4], s));
  if (G.stateless && ctx->impl == IMPL_TILED_SCALAR) {
    launch_dmm_stateless(G, ctx->T, W, s); ctx->launches++;       // fused demod+M&M variant
  } else if (G.stateless) {
    launch_demod_mm_v2(G, ctx->T, W, ctx->d_dem, s); ctx->launches += 2;
  } else {
    launch_demod(G, ctx->T, W, s);
    launch_mm(G, ctx->T, W, s);
    ctx->launches += 2;
  }
  CK(cudaEventRecord(ctx->ev[5], s));
  if (ctx->impl == IMPL_BASELINE) launch_search(G, ctx->T, W, s);
  else launch_search_warp(G, ctx->T, W, s);
  if (!G.early) launch_gather(G, W, s);         // lazy tail: symbols are gathered in collect(), after the resume
  ctx->launches += G.early ? 1 : 2;
  CK(cudaEventRecord(ctx->ev[6], s));
  CK(cudaMemcpyAsync(ctx->h_counts, W.hit_count, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  if (!G.stateless) CK(cudaMemcpyAsync(&ctx->mm, W.mm_state, sizeof(MmState), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(ctx->ev[7], s));
  CK(cudaGetLastError());
  restore.armed = false;
  ctx->pending = true;
  ctx->pend_slots = n_slots;
  ctx->pend_first_slot = first_slot;
  ctx->pendW = W;
  ctx->pendG = G;
  ctx->pend_early = G.early != 0;
  return BTB200_OK;
}

int btb200_set_window_mask(btb200_ctx *ctx, const uint8_t *mask, uint32_t n_slots)
{
  if (!ctx || !mask || n_slots == 0 || n_slots > ctx->max_slots) return BTB200_ERR_ARG;
  if (!ctx->lazy) return BTB200_ERR_ARG;             // stateless mode with the lazy squelch only
  const size_t n = (size_t)n_slots * ctx->plan.nch;
  ctx->win_mask.resize(n);
  for (size_t i = 0; i < n; i++) ctx->win_mask[i] = mask[i] ? 1 : 0;
  return BTB200_OK;
}

int btb200_submit(btb200_ctx *ctx, const float *iq, int iq_on_device, size_t n_samples,
                  uint64_t first_slot, uint32_t n_slots)
{
  return submit_impl(ctx, iq, iq_on_device ? IN_DEV_C32 : IN_HOST_C32, n_samples, first_slot, n_slots);
}

int btb200_submit_i16(btb200_ctx *ctx, const int16_t *iq, int iq_on_device, size_t n_samples,
                      uint64_t first_slot, uint32_t n_slots)
{
  return submit_impl(ctx, iq, iq_on_device ? IN_DEV_I16 : IN_HOST_I16, n_samples, first_slot, n_slots);
}

// First half of collect(): waits for the batch's search, fetches the hit list and ENQUEUES the deferred work
// (lazy tail resume on the second stream; deferred noise FIR + exact energies on the compute stream) without
// waiting for it.  A caller that keeps several batches in flight calls this, then submits the next batch
// (so that its kernels queue up behind the deferred work and its input copy overlaps it), then btb200_collect().
int btb200_collect_begin(btb200_ctx *ctx)
{
  if (!ctx || !ctx->pending || ctx->cb.begun) return BTB200_ERR_ARG;
  if (ctx->poly) { ctx->cb = btb200_ctx::CollectState{}; ctx->cb.begun = true; return BTB200_OK; }   // nothing is deferred to the host
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
  const Plan &P = ctx->plan;
  auto &cb = ctx->cb;
  cb = btb200_ctx::CollectState{};
  cb.begun = true;
  CK(cudaEventSynchronize(ctx->ev[7]));        // end of this batch's submit()
  Trace &tr = ctx->trace; tr = Trace{}; tr.mark("batch done");
  cb.nh = ctx->h_counts[0];
  std::memcpy(&cb.used, ctx->h_counts + 2, sizeof cb.used);
  if (cb.nh > kHitCap) { cb.dropped = cb.nh - kHitCap; cb.nh = kHitCap; }
  if (cb.used > kArenaCap) cb.used = kArenaCap;
  if (ctx->pend_early) cb.used = 0;
  const unsigned nh = cb.nh;
  // small transfers go through the copy stream: on the compute stream they would queue behind other batches
  if (nh) CK(cudaMemcpyAsync(ctx->h_hits, ctx->W.hits, (size_t)nh * sizeof(DevHit), cudaMemcpyDeviceToHost, cs));
  ctx->lazy_timed = false;
  if (!ctx->lazy) { if (nh) CK(cudaStreamSynchronize(cs)); return BTB200_OK; }

  const size_t nbc = (size_t)ctx->pend_slots * P.nch;
  if (ctx->poly) {
    // throughput mode: both energies of EVERY window are estimates that arrived with the batch
    for (size_t i = 0; i < nbc; i++) { ctx->h_energy[i] = ctx->h_eon_all[i]; ctx->h_noise[i] = ctx->h_esum[i] / P.n_noise; }
    cb.est_flag.assign(nbc, 1);
  } else
    for (size_t i = 0; i < nbc; i++) ctx->h_energy[i] = ctx->h_noise[i] = std::nan("");
  if (ctx->fast_snr) {
    ctx->fast_off.resize(nbc);
    for (size_t i = 0; i < nbc; i++) ctx->fast_off[i] = ctx->h_esum[i] / P.n_noise;
  }
  if (!nh) return BTB200_OK;
  CK(cudaStreamSynchronize(cs));
  // unique hit windows, in (slot, channel) order
  std::vector<uint32_t> &keys = cb.keys;
  keys.resize(nh);
  for (unsigned i = 0; i < nh; i++) keys[i] = (uint32_t)ctx->h_hits[i].b * (uint32_t)P.nch + (uint32_t)ctx->h_hits[i].chi;
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const int nl_all = (int)keys.size();
  if (ctx->pend_early) {
    // lazy tail: finish demod + clock recovery of the hit windows on the second stream, under the noise FIR
    for (int l = 0; l < nl_all; l++) {
      int *q = ctx->h_list2 + (size_t)l * 4;
      q[0] = (int)(keys[l] / P.nch); q[1] = (int)(keys[l] % P.nch); q[2] = 0; q[3] = 0;
    }
    cudaStream_t s2 = ctx->stream2;
    CK(cudaMemcpyAsync(ctx->d_list2, ctx->h_list2, (size_t)nl_all * 4 * sizeof(int), cudaMemcpyHostToDevice, s2));
    launch_mm_resume_list(ctx->pendG, ctx->T, ctx->pendW, ctx->d_dem, ctx->d_list2, nl_all, s2);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(ctx->h_nsym, ctx->pendW.nsym, nbc * sizeof(int), cudaMemcpyDeviceToHost, s2));
    tr.mark("resume launched");
  }
  if (ctx->poly) { CK(cudaGetLastError()); return BTB200_OK; }      // no exact fallback: the mode is hit-rate proof
  // pass 1 (fast mode): exact on-channel energy of every hit window, off-channel from the estimate
  std::vector<uint32_t> &need_exact = cb.need_exact;
  if (ctx->fast_snr) {
    for (int l = 0; l < nl_all; l++) {
      int *q = ctx->h_list + (size_t)l * 4;
      q[0] = (int)(keys[l] / P.nch); q[1] = (int)(keys[l] % P.nch); q[2] = 0; q[3] = 0;
    }
    CK(cudaMemcpyAsync(ctx->d_list, ctx->h_list, (size_t)nl_all * 4 * sizeof(int), cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(ctx->ev_up, cs));
    CK(cudaStreamWaitEvent(s, ctx->ev_up, 0));
    CK(cudaEventRecord(ctx->evl[0], s));
    CK(cudaEventRecord(ctx->evl[1], s));
    launch_energy_list(ctx->G, ctx->pendW, ctx->d_list, nl_all, nullptr, ctx->d_eon, ctx->d_eoff, s);
    ctx->launches++;
    CK(cudaEventRecord(ctx->evl[2], s));
    ctx->lazy_timed = true;
    CK(cudaMemcpyAsync(ctx->h_eon, ctx->d_eon, (size_t)nl_all * sizeof(double), cudaMemcpyDeviceToHost, s));
    SYNC_HERE();
    tr.mark("pass-1 energies back");
    const double guard = 5e-3;
    for (int l = 0; l < nl_all; l++) {
      const size_t bc = keys[l];
      const double on = ctx->h_eon[l], off = ctx->fast_off[bc];
      const double snr = 10.0 * std::log10(on / off);
      bool sure = std::isfinite(snr) && std::fabs(snr - P.squelch_db) > guard;
      if (sure) {
        const double d = snr * 10.0, t = d - std::floor(d);        // %.1f rounds at x.x5
        if (std::fabs(t - 0.5) <= 10.0 * guard) sure = false;
      }
      if (sure) { ctx->h_energy[bc] = on; ctx->h_noise[bc] = off; }
      else need_exact.push_back(keys[l]);
    }
    cb.est_flag.assign(nbc, 1);
    for (uint32_t k : need_exact) cb.est_flag[k] = 0;
  } else {
    need_exact = keys;
  }
  if (!need_exact.empty()) {
    // groups of up to CGR hit channels of one slot (they share the input span); groups with more
    // channels first: their blocks run longest
    const int CGR = lazy_group_channels(ctx->G);
    struct Grp { int b, n, c[LAZY_CG]; };
    std::vector<Grp> grp;
    for (uint32_t k : need_exact) {
      const int b = (int)(k / P.nch), c = (int)(k % P.nch);
      if (grp.empty() || grp.back().b != b || grp.back().n == CGR) grp.push_back(Grp{b, 0, {-1, -1, -1, -1}});
      grp.back().c[grp.back().n++] = c;
    }
    std::stable_sort(grp.begin(), grp.end(), [](const Grp &x, const Grp &y) { return x.n > y.n; });
    const int ng = (int)grp.size();
    int nl = 0;
    for (int gi = 0; gi < ng; gi++) {
      int *g = ctx->h_groups + (size_t)gi * (1 + CGR);
      g[0] = grp[gi].b;
      for (int i = 0; i < CGR; i++) g[1 + i] = grp[gi].c[i];
      for (int i = 0; i < grp[gi].n; i++) {
        int *l = ctx->h_list + (size_t)nl * 4;
        l[0] = grp[gi].b; l[1] = grp[gi].c[i]; l[2] = gi; l[3] = i;
        nl++;
      }
    }
    cb.nl = nl;
    CK(cudaMemcpyAsync(ctx->d_groups, ctx->h_groups, (size_t)ng * (1 + CGR) * sizeof(int), cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(ctx->d_list, ctx->h_list, (size_t)nl * 4 * sizeof(int), cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(ctx->ev_up, cs));
    CK(cudaStreamWaitEvent(s, ctx->ev_up, 0));
    if (!ctx->fast_snr) CK(cudaEventRecord(ctx->evl[0], s));
    launch_noise_fir_list(ctx->G, ctx->T, ctx->pendW, ctx->d_groups, ng, ctx->d_NzL, s);
    if (!ctx->fast_snr) CK(cudaEventRecord(ctx->evl[1], s));
    launch_energy_list(ctx->G, ctx->pendW, ctx->d_list, nl, ctx->d_NzL, ctx->d_eon, ctx->d_eoff, s);
    if (!ctx->fast_snr) { CK(cudaEventRecord(ctx->evl[2], s)); ctx->lazy_timed = true; }
    ctx->launches += 2;
    CK(cudaMemcpyAsync(ctx->h_eon, ctx->d_eon, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_eoff, ctx->d_eoff, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, s));
    tr.mark("deferred noise FIR enqueued");
  }
  CK(cudaGetLastError());
  return BTB200_OK;
}

namespace {
// borrowed symbols: the caller passes symbols == NULL and symbols_cap == UINT64_MAX and gets a pointer into the
// context's pinned arena (valid until the next submit on this context) instead of a copy
inline bool borrow_symbols(const btb200_hits *out) { return out && !out->symbols && out->symbols_cap == UINT64_MAX; }

int collect_poly(btb200_ctx *ctx, btb200_hits *out)
{
  const Plan &P = ctx->plan;
  cudaStream_t cs = ctx->copy_stream;
  ctx->pending = false;
  ctx->cb.begun = false;
  CK(cudaEventSynchronize(ctx->ev[7]));              // the batch's tail has finished; counts are on the host
  unsigned nh = ctx->h_counts[0], dropped = 0;
  if (nh > kHitCap) { dropped = nh - kHitCap; nh = kHitCap; }
  unsigned long long used = 0;
  std::memcpy(&used, ctx->h_counts + 2, sizeof used);
  if (used > kArenaCap) used = kArenaCap;
  {
    // adapt the tail policy to the traffic: resuming a window's clock recovery costs ~3x the share it has in a pass
    // over ALL windows (compact list, one sector per lane), so past ~a quarter of the windows the full pass is cheaper
    // (a full pass is one latency-bound wave of ~1.5 ms whatever the batch size, so it only pays for long batches)
    const double frac = (double)ctx->h_counts[1] / (double)((size_t)ctx->pend_slots * P.nch);
    if (frac > 0.25 && ctx->pend_slots >= 384) ctx->dense_tail = true;
    else if (frac < 0.15 || ctx->pend_slots < 384) ctx->dense_tail = false;
  }
  const bool want_sym = out && (out->symbols || borrow_symbols(out));
  if (nh) CK(cudaMemcpyAsync(ctx->h_hits, ctx->pendW.tail.sorted, (size_t)nh * sizeof(DevHit), cudaMemcpyDeviceToHost, cs));
  if (want_sym && used) CK(cudaMemcpyAsync(ctx->h_arena, ctx->W.arena, used, cudaMemcpyDeviceToHost, cs));
  CK(cudaEventRecord(ctx->ev[8], cs));
  const size_t nbc = (size_t)ctx->pend_slots * P.nch;
  for (size_t i = 0; i < nbc; i++) { ctx->h_energy[i] = ctx->h_eon_all[i]; ctx->h_noise[i] = ctx->h_esum[i] / P.n_noise; }
  ctx->fast_off.assign(ctx->h_noise, ctx->h_noise + nbc);
  CK(cudaEventSynchronize(ctx->ev[8]));
  ctx->last_slots = ctx->pend_slots;
  // order of the throughput mode's stream: channelizer | clock recovery (prefix) | search | resume | noise estimate |
  // tail (hit list, arena, gather, copies)
  cudaEventElapsedTime(&ctx->timing[0], ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->timing[1], ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->timing[4], ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ctx->timing[5], ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->timing[3], ctx->ev[4], ctx->ev_res);       // [3]: resume of the windows with hits
  cudaEventElapsedTime(&ctx->timing[2], ctx->ev_res, ctx->ev[5]);
  cudaEventElapsedTime(&ctx->timing[6], ctx->ev[5], ctx->ev[8]);
  cudaEventElapsedTime(&ctx->timing[7], ctx->ev[0], ctx->ev[8]);
  if (!out) return BTB200_OK;
  out->count = 0;
  out->overflow = dropped;
  out->symbols_used = 0;
  const bool borrow = borrow_symbols(out);
  if (borrow) { out->symbols = ctx->h_arena; out->symbols_used = used; }
  const DevHit *hh = ctx->h_hits;
  for (unsigned i = 0; i < nh; i++) {
    const DevHit &h = hh[i];                      // already in the reference's visiting order
    const size_t bc = (size_t)h.b * P.nch + h.chi;
    const double snr = 10.0 * std::log10(ctx->h_energy[bc] / ctx->h_noise[bc]);     // multi_block.cc:293
    if (!(snr >= P.squelch_db)) continue;
    if (out->count >= out->cap) { out->overflow++; continue; }
    btb200_hit &o = out->hits[out->count];
    o.slot = (uint32_t)(ctx->pend_first_slot + (uint64_t)h.b);
    o.channel = (uint16_t)(P.ch_lo + h.chi);
    o.kind = (uint16_t)h.kind;
    o.offset = h.offset;
    o.n_symbols = h.n_symbols;
    o.lap = h.kind == 0 ? (h.lap & 0xffffffu) : h.lap;       // BR: bits 24..31 carry the access code's symbol errors
    o.flags = ((std::fabs(snr - P.squelch_db) <= 1e-6) ? 1u : 0u) | 2u | 4u;
    o.snr = snr;
    o.sym_offset = 0;
    o.sym_count = 0;
    o.ac_errors = h.kind == 0 ? (h.lap >> 24) : 0u;
    if (borrow) { o.sym_offset = h.sym_offset; o.sym_count = h.sym_count; }
    else if (out->symbols && h.sym_count && out->symbols_used + h.sym_count <= out->symbols_cap) {
      std::memcpy(out->symbols + out->symbols_used, ctx->h_arena + h.sym_offset, h.sym_count);
      o.sym_offset = out->symbols_used;
      o.sym_count = h.sym_count;
      out->symbols_used += h.sym_count;
    }
    out->count++;
  }
  return BTB200_OK;
}
}  // namespace

int btb200_collect(btb200_ctx *ctx, btb200_hits *out)
{
  if (!ctx || !ctx->pending) return BTB200_ERR_ARG;
  if (ctx->poly) { CK(cudaSetDevice(ctx->device)); return collect_poly(ctx, out); }
  if (!ctx->cb.begun) { if (int rc = btb200_collect_begin(ctx)) return rc; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const Plan &P = ctx->plan;
  auto &cb = ctx->cb;
  Trace &tr = ctx->trace;
  ctx->pending = false;
  cb.begun = false;
  const unsigned nh = cb.nh, dropped = cb.dropped;
  unsigned long long used = cb.used;
  const std::vector<uint8_t> &est_flag = cb.est_flag;
  if (ctx->pend_early && nh) {
    // lazy tail, second half: the symbol counts are known once the resume has finished (it runs under the
    // deferred noise FIR): complete the hit records, lay the arena out, gather the symbols
    cudaStream_t s2 = ctx->stream2;
    CK(cudaStreamSynchronize(s2));
    tr.mark("resume done (host)");
    unsigned long long off = 0;
    for (unsigned i = 0; i < nh; i++) {
      DevHit &h = ctx->h_hits[i];
      h.n_symbols += ctx->h_nsym[(size_t)h.b * P.nch + h.chi];
      int cnt = h.n_symbols < 3125 ? h.n_symbols : 3125;
      if (cnt < 0) cnt = 0;
      h.sym_offset = off;
      h.sym_count = (off + (unsigned)cnt <= kArenaCap) ? (uint32_t)cnt : 0u;
      off += (unsigned)cnt;
    }
    used = off > kArenaCap ? kArenaCap : off;
    if (out && (out->symbols || borrow_symbols(out)) && used) {
      CK(cudaMemcpyAsync(ctx->W.hits, ctx->h_hits, (size_t)nh * sizeof(DevHit), cudaMemcpyHostToDevice, s2));
      launch_gather(ctx->pendG, ctx->pendW, s2); ctx->launches++;
      CK(cudaMemcpyAsync(ctx->h_arena, ctx->W.arena, used, cudaMemcpyDeviceToHost, s2));
    }
    CK(cudaEventRecord(ctx->ev_tail, s2));
    CK(cudaStreamWaitEvent(s, ctx->ev_tail, 0));     // the batch (and its timing) ends when the tail has
  } else if (used && out && (out->symbols || borrow_symbols(out))) {
    CK(cudaMemcpyAsync(ctx->h_arena, ctx->W.arena, used, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaEventRecord(ctx->ev[8], s));
  SYNC_HERE();
  tr.mark("all done");
  for (int l = 0; l < cb.nl; l++) {
    const size_t bc = (size_t)ctx->h_list[(size_t)l * 4] * P.nch + ctx->h_list[(size_t)l * 4 + 1];
    ctx->h_energy[bc] = ctx->h_eon[l];
    ctx->h_noise[bc] = ctx->h_eoff[l];
  }
  ctx->last_slots = ctx->pend_slots;
  for (int i = 0; i < 7; i++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]);
    ctx->timing[i] = ms;
  }
  {
    // [0] H2D [1] chan FIR [2] noise FIR [3] energy [4] demod+mm [5] search [6] D2H, [7] total
    float d2h = 0, tot = 0;
    cudaEventElapsedTime(&d2h, ctx->ev[6], ctx->ev[8]);
    cudaEventElapsedTime(&tot, ctx->ev[0], ctx->ev[8]);
    ctx->timing[6] = d2h;
    ctx->timing[7] = tot;
    if (ctx->lazy_timed) {
      // lazy squelch: [2]/[3] report the deferred noise FIR and energy kernels, [6] the rest of the tail
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, ctx->evl[0], ctx->evl[1]);
      cudaEventElapsedTime(&b, ctx->evl[1], ctx->evl[2]);
      ctx->timing[2] = a; ctx->timing[3] = b; ctx->timing[6] = d2h - a - b;
    }
  }
  if (!out) return BTB200_OK;

  // order = the reference's visiting order: slot, channel, BR before LE, ascending offset
  std::vector<unsigned> order(nh);
  for (unsigned i = 0; i < nh; i++) order[i] = i;
  const DevHit *hh = ctx->h_hits;
  std::sort(order.begin(), order.end(), [hh](unsigned a, unsigned b) {
    const DevHit &x = hh[a], &y = hh[b];
    if (x.b != y.b) return x.b < y.b;
    if (x.chi != y.chi) return x.chi < y.chi;
    if (x.kind != y.kind) return x.kind < y.kind;
    return x.offset < y.offset;
  });
  out->count = 0;
  out->overflow = dropped;
  out->symbols_used = 0;
  const bool borrow = borrow_symbols(out);
  if (borrow) { out->symbols = ctx->h_arena; out->symbols_used = used; }
  for (unsigned oi = 0; oi < nh; oi++) {
    const DevHit &h = hh[order[oi]];
    const size_t bc = (size_t)h.b * P.nch + h.chi;
    // the value ac() prints, with the reference's own arithmetic (multi_block.cc:293)
    const double snr = 10.0 * std::log10(ctx->h_energy[bc] / ctx->h_noise[bc]);
    if (!(snr >= P.squelch_db)) continue;          // guard-band window the exact compare rejects
    if (out->count >= out->cap) { out->overflow++; continue; }
    btb200_hit &o = out->hits[out->count];
    o.slot = (uint32_t)(ctx->pend_first_slot + (uint64_t)h.b);
    o.channel = (uint16_t)(P.ch_lo + h.chi);
    o.kind = (uint16_t)h.kind;
    o.offset = h.offset;
    o.n_symbols = h.n_symbols;
    o.lap = h.kind == 0 ? (h.lap & 0xffffffu) : h.lap;       // BR: bits 24..31 carry the access code's symbol errors
    o.flags = (std::fabs(snr - P.squelch_db) <= 1e-6) ? 1u : 0u;
    if (!est_flag.empty() && est_flag[bc]) o.flags |= 2u;
    if (ctx->poly) o.flags |= 4u;
    o.snr = snr;
    o.sym_offset = 0;
    o.sym_count = 0;
    o.ac_errors = h.kind == 0 ? (h.lap >> 24) : 0u;
    if (borrow) { o.sym_offset = h.sym_offset; o.sym_count = h.sym_count; }
    else if (out->symbols && h.sym_count && out->symbols_used + h.sym_count <= out->symbols_cap) {
      std::memcpy(out->symbols + out->symbols_used, ctx->h_arena + h.sym_offset, h.sym_count);
      o.sym_offset = out->symbols_used;
      o.sym_count = h.sym_count;
      out->symbols_used += h.sym_count;
    }
    out->count++;
  }
  return BTB200_OK;
}

int btb200_process(btb200_ctx *ctx, const float *iq, size_t n_samples, uint64_t first_slot,
                   uint32_t n_slots, btb200_hits *out)
{
  int rc = btb200_submit(ctx, iq, 0, n_samples, first_slot, n_slots);
  if (rc) return rc;
  return btb200_collect(ctx, out);
}

int btb200_process_i16(btb200_ctx *ctx, const int16_t *iq, size_t n_samples, uint64_t first_slot,
                       uint32_t n_slots, btb200_hits *out)
{
  int rc = btb200_submit_i16(ctx, iq, 0, n_samples, first_slot, n_slots);
  if (rc) return rc;
  return btb200_collect(ctx, out);
}

int btb200_process_device(btb200_ctx *ctx, const float *d_iq, size_t n_samples, uint64_t first_slot,
                          uint32_t n_slots, btb200_hits *out)
{
  int rc = btb200_submit(ctx, d_iq, 1, n_samples, first_slot, n_slots);
  if (rc) return rc;
  return btb200_collect(ctx, out);
}

int btb200_process_channels(btb200_ctx *ctx, const float *iq, size_t n_samples, uint64_t slot, int32_t first_channel,
                            int32_t n_channels, uint32_t stop_lap, btb200_chan_result *res, uint8_t *symbols,
                            size_t symbols_cap)
{
  (void)slot;
  if (!ctx || !iq || !res || n_channels < 0) return BTB200_ERR_ARG;
  const Plan &P = ctx->plan;
  const Geom &G = ctx->G;
  if (G.stateless || ctx->pending) return BTB200_ERR_ARG;
  const int first = first_channel - P.ch_lo;
  if (n_channels > 0 && (first < 0 || first + n_channels > P.nch)) return BTB200_ERR_ARG;
  if (n_samples < (size_t)P.H) return BTB200_ERR_SHORT_INPUT;
  if (n_channels == 0) return BTB200_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBatch W = ctx->W;
  W.B = 1;
  const size_t nch = (size_t)P.nch;
  // rotators advance only for channels whose DDC objects the reference actually calls: keep a copy
  std::vector<Rotator> save_c = ctx->rot_c, save_n = ctx->rot_n;
  CK(cudaMemcpyAsync(ctx->d_x, iq, (size_t)P.H * sizeof(c32), cudaMemcpyHostToDevice, s));
  W.x = ctx->d_x;
  cf32 *hp = reinterpret_cast<cf32 *>(ctx->h_ph);
  for (int c = 0; c < P.nch; c++) ctx->rot_c[c].generate(hp + c, P.n_ddc, P.nch);
  CK(cudaMemcpyAsync(ctx->d_phc, hp, (size_t)P.n_ddc * nch * sizeof(c32), cudaMemcpyHostToDevice, s));
  SYNC_HERE();
  for (int c = 0; c < P.nch; c++) ctx->rot_n[c].generate(hp + c, P.n_noise, P.nch);
  CK(cudaMemcpyAsync(ctx->d_phn, hp, (size_t)P.n_noise * nch * sizeof(c32), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(W.mm_state, &ctx->mm, sizeof(MmState), cudaMemcpyHostToDevice, s));
  launch_chan_fir(G, ctx->T, W, ctx->impl, s);
  launch_noise_fir(G, ctx->T, W, ctx->impl, s);
  launch_energy(G, ctx->T, W, 0, s);
  ctx->launches += 3;
  CK(cudaMemcpyAsync(ctx->h_energy, W.energy, nch * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ctx->h_noise, W.noise, nch * sizeof(double), cudaMemcpyDeviceToHost, s));
  SYNC_HERE();
  std::vector<double> snr(nch);
  for (size_t i = 0; i < nch; i++) {
    snr[i] = 10.0 * std::log10(ctx->h_energy[i] / ctx->h_noise[i]);
    const bool listed = (int)i >= first && (int)i < first + n_channels;
    ctx->h_pass[i] = (listed && snr[i] >= P.squelch_db) ? 1 : 0;
  }
  CK(cudaMemcpyAsync(W.pass, ctx->h_pass, nch * sizeof(int), cudaMemcpyHostToDevice, s));
  launch_demod(G, ctx->T, W, s);
  int *d_res = ctx->d_list ? ctx->d_list : nullptr;
  if (!ctx->d_res4) { int rc = dev_alloc(ctx, &ctx->d_res4, 4 * 128); if (rc) return rc; }
  d_res = ctx->d_res4;
  launch_mm_chained_list(G, ctx->T, W, first, n_channels, stop_lap, d_res, s);
  ctx->launches += 2;
  std::vector<int> r4(4 * nch);
  std::vector<uint32_t> rows(nch * (size_t)G.bw);
  CK(cudaMemcpyAsync(r4.data(), d_res, 4 * nch * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(rows.data(), W.bits, rows.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&ctx->mm, W.mm_state, sizeof(MmState), cudaMemcpyDeviceToHost, s));
  SYNC_HERE();
  ctx->last_slots = 1;
  size_t used = 0;
  for (int c = 0; c < P.nch; c++) {
    const bool listed = c >= first && c < first + n_channels;
    const bool processed = listed && r4[4 * c] != 0;
    if (!processed) { ctx->rot_c[c] = save_c[c]; ctx->rot_n[c] = save_n[c]; }    // DDC objects never called
    if (!listed) continue;
    btb200_chan_result &o = res[c - first];
    o.channel = P.ch_lo + c;
    o.processed = processed ? 1 : 0;
    o.pass = processed ? ctx->h_pass[c] : 0;
    o.n_symbols = r4[4 * c + 1];
    o.ac_index = processed ? r4[4 * c + 2] : -1;
    o.lap = (uint32_t)r4[4 * c + 3];
    o.snr = snr[c];
    o.sym_offset = 0; o.sym_count = 0; o.reserved = 0;
    if (o.ac_index >= 0 && symbols) {
      int cnt = o.n_symbols - o.ac_index;
      if (cnt > 3125) cnt = 3125;
      if (used + (size_t)cnt <= symbols_cap) {
        const uint32_t *row = &rows[(size_t)c * G.bw];
        for (int i = 0; i < cnt; i++) { const int b = o.ac_index + i; symbols[used + i] = (row[b >> 5] >> (b & 31)) & 1; }
        o.sym_offset = used; o.sym_count = (uint32_t)cnt;
        used += (size_t)cnt;
      }
    }
  }
  return BTB200_OK;
}

int btb200_timer_start(btb200_ctx *ctx)
{
  if (!ctx) return BTB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->ev_user[0], ctx->stream));
  return BTB200_OK;
}

int btb200_timer_stop(btb200_ctx *ctx, float *ms)
{
  if (!ctx || !ms) return BTB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->ev_user[1], ctx->stream));
  CK(cudaEventSynchronize(ctx->ev_user[1]));
  CK(cudaEventElapsedTime(ms, ctx->ev_user[0], ctx->ev_user[1]));
  return BTB200_OK;
}

int btb200_search_bits(btb200_ctx *ctx, const uint8_t *symbols, size_t n_symbols, uint32_t stride, btb200_hits *out)
{
  if (!ctx || !symbols || !out || stride == 0 || stride > 625 || ctx->pending) return BTB200_ERR_ARG;
  const Plan &P = ctx->plan;
  Geom G = ctx->G;
  G.early = 0;
  G.search = BTB200_SEARCH_BR | (ctx->cfg.search & BTB200_SEARCH_BR_BCH);
  const int wlen = (int)stride + 72;
  if (G.bw < (wlen + 31) / 32 + 4) return BTB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t nwin = (n_symbols + stride - 1) / stride;
  const size_t per_call = (size_t)ctx->max_slots * P.nch;
  std::vector<uint32_t> rows(per_call * G.bw);
  std::vector<int> nsym(per_call);
  out->count = 0; out->overflow = 0; out->symbols_used = 0;
  for (size_t w0 = 0; w0 < nwin; w0 += per_call) {
    const size_t nw = std::min(per_call, nwin - w0);
    const size_t nb = (nw + P.nch - 1) / P.nch;
    std::fill(rows.begin(), rows.end(), 0u);
    std::fill(nsym.begin(), nsym.end(), 0);
    for (size_t i = 0; i < nw; i++) {
      const size_t first = (w0 + i) * stride;
      const int len = (int)std::min<size_t>((size_t)wlen, n_symbols - first);
      nsym[i] = len;
      uint32_t *row = &rows[i * G.bw];
      for (int k = 0; k < len; k++) row[k >> 5] |= (uint32_t)(symbols[first + k] & 1) << (k & 31);
    }
    DevBatch W = ctx->W;
    W.B = (int)nb;
    CK(cudaMemcpyAsync(W.bits, rows.data(), nb * P.nch * G.bw * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(W.nsym, nsym.data(), nb * P.nch * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(W.hit_count, 0, 4 * sizeof(unsigned), s));
    launch_search_warp(G, ctx->T, W, s); ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_counts, W.hit_count, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    SYNC_HERE();
    unsigned nh = ctx->h_counts[0];
    if (nh > kHitCap) { out->overflow += nh - kHitCap; nh = kHitCap; }
    if (nh) CK(cudaMemcpy(ctx->h_hits, W.hits, (size_t)nh * sizeof(DevHit), cudaMemcpyDeviceToHost));
    std::vector<unsigned> order(nh);
    for (unsigned i = 0; i < nh; i++) order[i] = i;
    const DevHit *hh = ctx->h_hits;
    std::sort(order.begin(), order.end(), [hh](unsigned a, unsigned b) {
      const DevHit &x = hh[a], &y = hh[b];
      if (x.b != y.b) return x.b < y.b;
      if (x.chi != y.chi) return x.chi < y.chi;
      return x.offset < y.offset;
    });
    for (unsigned oi = 0; oi < nh; oi++) {
      const DevHit &h = hh[order[oi]];
      if (h.kind != 0) continue;
      if (out->count >= out->cap) { out->overflow++; continue; }
      btb200_hit &o = out->hits[out->count++];
      o = btb200_hit{};
      o.slot = (uint32_t)(w0 + (size_t)h.b * P.nch + h.chi);
      o.offset = h.offset; o.n_symbols = h.n_symbols; o.lap = h.lap & 0xffffffu; o.ac_errors = h.lap >> 24;
    }
  }
  ctx->last_slots = 0;
  return BTB200_OK;
}

int btb200_last_timing(const btb200_ctx *ctx, float ms[8])
{
  if (!ctx || !ms) return BTB200_ERR_ARG;
  std::memcpy(ms, ctx->timing, sizeof ctx->timing);
  return BTB200_OK;
}

uint64_t btb200_launch_count(const btb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int64_t btb200_get_stage(btb200_ctx *ctx, int stage, uint32_t b, uint32_t chi, void *dst, size_t cap)
{
  if (!ctx || !dst) return BTB200_ERR_ARG;
  const Plan &P = ctx->plan;
  const Geom &G = ctx->G;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return BTB200_ERR_NO_DEVICE;
  auto host_copy = [&](const void *src, size_t bytes) -> int64_t {
    if (bytes > cap) return BTB200_ERR_ARG;
    std::memcpy(dst, src, bytes);
    return (int64_t)bytes;
  };
  if (stage >= BTB200_STAGE_CHAN_TAPS) {
    if ((int)chi >= P.nch) return BTB200_ERR_ARG;
    switch (stage) {
      case BTB200_STAGE_CHAN_TAPS: return host_copy(&P.chan_rtaps[(size_t)chi * P.Nc], (size_t)P.Nc * 8);
      case BTB200_STAGE_NOISE_TAPS: return host_copy(&P.noise_rtaps[(size_t)chi * P.Nn], (size_t)P.Nn * 8);
      case BTB200_STAGE_MMSE_TABLE: return host_copy(P.mmse.data(), P.mmse.size() * 4);
      case BTB200_STAGE_ATAN_TABLE: return host_copy(P.atan_tab.data(), P.atan_tab.size() * 4);
      case BTB200_STAGE_AC_LUT: return host_copy(P.ac_lut.data(), P.ac_lut.size() * 8);
      default: return BTB200_ERR_ARG;
    }
  }
  if (b >= ctx->last_slots || (int)chi >= P.nch) return BTB200_ERR_ARG;
  const size_t bc = (size_t)b * P.nch + chi;
  auto dev_copy = [&](const void *src, size_t bytes) -> int64_t {
    if (bytes > cap) return BTB200_ERR_ARG;
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
    return (int64_t)bytes;
  };
  switch (stage) {
    case BTB200_STAGE_ENERGY: return host_copy(&ctx->h_energy[bc], 8);
    case BTB200_STAGE_NOISE: return host_copy(&ctx->h_noise[bc], 8);
    case BTB200_STAGE_NOISE_FAST:
      if (!ctx->fast_snr || ctx->fast_off.size() <= bc) return BTB200_ERR_ARG;
      return host_copy(&ctx->fast_off[bc], 8);
    case BTB200_STAGE_SNR: {
      const double snr = 10.0 * std::log10(ctx->h_energy[bc] / ctx->h_noise[bc]);
      return host_copy(&snr, 8);
    }
    case BTB200_STAGE_PASS: return dev_copy(ctx->W.pass + bc, 4);
    case BTB200_STAGE_NSYM: return dev_copy(ctx->W.nsym + bc, 4);
    case BTB200_STAGE_BITS: {
      int nsym = 0;
      if (cudaMemcpy(&nsym, ctx->W.nsym + bc, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
      if ((size_t)nsym > cap) return BTB200_ERR_ARG;
      std::vector<uint32_t> row(G.bw);
      if (cudaMemcpy(row.data(), ctx->W.bits + bc * G.bw, (size_t)G.bw * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return BTB200_ERR_CUDA;
      uint8_t *o = (uint8_t *)dst;
      for (int i = 0; i < nsym; i++) o[i] = (row[i >> 5] >> (i & 31)) & 1;
      return nsym;
    }
    case BTB200_STAGE_DDC: {
      if (ctx->poly) return BTB200_ERR_ARG;      // the polyphase front end never materialises the DDC outputs
      // rotated DDC output of the window = Y[b*gps + i][chi] * phase[i] (same fp32 ops as the kernels)
      if ((size_t)P.n_ddc * 8 > cap) return BTB200_ERR_ARG;
      std::vector<c32> y(P.n_ddc), ph(P.n_ddc);
      const size_t pitch = (size_t)P.nch * sizeof(c32);
      if (cudaMemcpy2D(y.data(), sizeof(c32), ctx->W.Y + ((size_t)b * P.grid_per_slot) * P.nch + chi, pitch,
                       sizeof(c32), P.n_ddc, cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
      const size_t bp = G.stateless ? 0 : b;
      if (cudaMemcpy2D(ph.data(), sizeof(c32), ctx->d_phc + (bp * P.n_ddc) * P.nch + chi, pitch,
                       sizeof(c32), P.n_ddc, cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
      c32 *o = (c32 *)dst;
      for (int i = 0; i < P.n_ddc; i++) o[i] = crot(y[i], ph[i]);
      return (int64_t)P.n_ddc * 8;
    }
    case BTB200_STAGE_DEMOD:
      if (!ctx->d_dem || ctx->impl == IMPL_TILED_SCALAR) return BTB200_ERR_ARG;
      if (G.stateless) {          // transposed [b][i][c]; polyphase mode: window b is a view of the global grid
        if ((size_t)P.n_dem * 4 > cap) return BTB200_ERR_ARG;
        if (cudaMemcpy2D(dst, 4, ctx->d_dem + ((size_t)b * G.dem_rows) * P.nch + chi, (size_t)P.nch * 4, 4, P.n_dem,
                         cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
        if (G.dem_grid) *(float *)dst = 0.0f;
        return (int64_t)P.n_dem * 4;
      }
      return dev_copy(ctx->d_dem + bc * G.n_dem_pad, (size_t)P.n_dem * 4);
    case BTB200_STAGE_SOFT: {
      if (!ctx->d_soft) return BTB200_ERR_ARG;
      int nsym = 0;
      if (cudaMemcpy(&nsym, ctx->W.nsym + bc, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return BTB200_ERR_CUDA;
      return dev_copy(ctx->d_soft + bc * G.n_dem_pad, (size_t)nsym * 4);
    }
    default: return BTB200_ERR_ARG;
  }
}

// debug hook (tools/trace_pipeline.py): when the last batch of `ctx` started its kernels, finished its search, finished
// its tail and finished its copies, in ms after `base`'s btb200_timer_start()
BTB200_API int btb200_debug_timeline(btb200_ctx *ctx, btb200_ctx *base, float out[5])
{
  if (!ctx || !base || !out) return BTB200_ERR_ARG;
  const int idx[5] = {0, 1, 6, 7, 8};
  for (int i = 0; i < 5; i++)
    if (cudaEventElapsedTime(&out[i], base->ev_user[0], ctx->ev[idx[i]]) != cudaSuccess) out[i] = -1.0f;
  return BTB200_OK;
}

// test hook: select baseline (0) or tuned (1) kernels
BTB200_API int btb200_set_impl(btb200_ctx *ctx, int impl)
{
  if (!ctx) return BTB200_ERR_ARG;
  ctx->impl = impl;
  return BTB200_OK;
}

}  // extern "C"
