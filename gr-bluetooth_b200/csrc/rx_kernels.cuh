// rx_kernels.cuh -- launch interface of the sm_100a kernels (rx_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "rx_bodies.cuh"

namespace btb200 {

struct DevTables {
  const c32 *chan_rtaps;     // [nch][Nc]
  const c32 *noise_rtaps;    // [nch][Nn]
  const c32 *chan_tg;        // [ngroups][Nc][16]  channel-group-interleaved taps for the tiled FIR
  const c32 *noise_tg;       // [ngroups][Nn][16]
  const void *chan_tg4;      // [ngroups][Nc][16] float4 (c, c, d, d): the channel tap banks as the TMA copies them
  const void *noise_taps4;   // [nch][Nn] float4 (c, c, d, d) for the delay-line noise FIR (D = 50), else null
  const float *mmse;         // [129*8]
  const float *atan_tab;     // [257]
  const uint64_t *ac_lut;    // [769]
  const uint8_t *le_hdr_lut; // [4*256]
  const int8_t *le_index;    // [nch]
  const uint32_t *le_white;  // [nch] 16 whitening bits
  BchDev bch;                // BTB200_SEARCH_BR_BCH: libbtbb-style access-code test (rx_math.cuh)
};

// Device-driven tail of the throughput mode: the search kernel stages each window's hits in the window's own slots (the
// warp replays the reference's search loops in order, so they arrive in visiting order: BR before LE, ascending lag),
// a scan over the windows turns the counts into positions, and everything after that -- the list of windows to resume,
// the hit list in the reference's order, the layout of the symbol arena -- is laid out by small kernels without a host
// round trip (btb200_api.cu, polyphase mode).  stage == nullptr: hits are appended with atomics and sorted on the host.
constexpr int TAIL_MAXW = 32;        // hits a window can hold: <= 10 access codes (68 apart) + <= 16 LE (40 apart) in 625 lags
struct TailBufs {
  DevHit *stage = nullptr;           // [B*nch][TAIL_MAXW]
  int *cnt = nullptr;                // [B*nch] hits of a window
  int *base = nullptr;               // [B*nch] position of the window's first hit in `sorted`
  int4 *list = nullptr;              // windows with hits, in (slot, channel) order: {b, chi, 0, 0}
  int *n_list = nullptr;             // [1]
  DevHit *sorted = nullptr;          // the hit list in the reference's visiting order, symbol counts and arena offsets final
};

struct DevBatch {
  const c32 *x;              // input
  c32 *Y;                    // [G][nch]
  c32 *Nz;                   // [B][n_noise][nch]
  const c32 *phc, *phn;      // rotator tables
  int bp_stride;             // 0: one table for every window (stateless), 1: per-window tables
  double *energy, *noise;    // [B][nch]
  int *pass;                 // [B][nch]  (device decision, or host-provided in chained mode)
  float *dem;                // [B][nch][n_dem_pad]
  float *soft;               // optional [B][nch][n_dem_pad]
  uint32_t *bits;            // [B][nch][bw]
  int *nsym;                 // [B][nch]
  MmState *mm_state;         // [1] chained-mode state in/out
  void *mm_save;             // [B][nch] saved clock-recovery loop state (lazy tail)
  DevHit *hits;              // [hit_cap]
  unsigned *hit_count;       // [1]
  unsigned long long *arena_used;   // [1]
  uint8_t *arena;            // [arena_cap]
  unsigned hit_cap;
  unsigned long long arena_cap;
  int B;
  TailBufs tail;
};

// implementation selectors (tests compare tuned kernels against the v1 baseline)
enum { IMPL_BASELINE = 0, IMPL_TUNED = 1, IMPL_TILED_SCALAR = 2 };

void launch_chan_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s);
// the channel FIR in tile ranges (input copy and filtering overlap, btb200_submit): tiles of the tuned kernels,
// and the number of input samples the tiles below `tile_end` read
long chan_fir_tiles(const Geom &G, const DevBatch &W);
long chan_fir_samples(const Geom &G, long tile_end);
void launch_chan_fir_range(const Geom &G, const DevTables &T, const DevBatch &W, int impl, long tile0, long tile1,
                           cudaStream_t s);
void launch_noise_fir(const Geom &G, const DevTables &T, const DevBatch &W, int impl, cudaStream_t s);
void launch_energy(const Geom &G, const DevTables &T, const DevBatch &W, int device_gate, cudaStream_t s);
void launch_demod(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s);
void launch_mm(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s);
void launch_mm_chained_list(const Geom &G, const DevTables &T, const DevBatch &W, int first, int n, unsigned stop_lap,
                            int *res4, cudaStream_t s);
void launch_search(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s);
void launch_gather(const Geom &G, const DevBatch &W, cudaStream_t s);
// device-driven tail (W.tail.stage != nullptr): positions + resume list; resume of the listed windows (lazy tail);
// final hit list (W.tail.sorted, count in W.hit_count[0]), arena layout (W.arena_used) and symbol gather
void launch_tail_scan(const Geom &G, const DevBatch &W, cudaStream_t s);
void launch_tail_resume(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, cudaStream_t s);
void launch_tail_finish(const Geom &G, const DevBatch &W, cudaStream_t s);
void launch_demod_mm_v2(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, cudaStream_t s);
void launch_mm_resume_list(const Geom &G, const DevTables &T, const DevBatch &W, float *demT, const int *list4,
                           int n_list, cudaStream_t s);
void launch_search_warp(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s);
void launch_dmm_stateless(const Geom &G, const DevTables &T, const DevBatch &W, cudaStream_t s);
void launch_fill_pass(const DevBatch &W, int n, int v, cudaStream_t s);
// lazy squelch: noise FIR for listed (slot, <=LAZY_CG channels) groups, then exact energies of listed windows
constexpr int LAZY_CG = 4;            // upper bound of channels per group (buffer sizing)
int lazy_group_channels(const Geom &G);   // channels per group of the selected configuration
void launch_noise_fir_list(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                           c32 *NzL, cudaStream_t s);
void launch_energy_list(const Geom &G, const DevBatch &W, const int *list4, int n_list, const c32 *NzL,
                        double *e_on, double *e_off, cudaStream_t s);
// delay-line variant (rx_firdl.cu), D = 50 only; groups of up to LAZY_CG channels, live channels first
bool noise_fir_dl_supported(const Geom &G);
int launch_noise_fir_dl(const Geom &G, const DevTables &T, const DevBatch &W, const int *groups, int n_groups,
                        c32 *NzL, cudaStream_t s);
int  fir_setup(int device);   // opt in to large dynamic shared memory

}  // namespace btb200
