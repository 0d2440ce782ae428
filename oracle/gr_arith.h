/*
 * oracle/gr_arith.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatement, in plain C, of the GNU Radio 3.7 arithmetic that gr-bluetooth's
 * receive path calls but that is NOT part of /root/reference (third-party
 * dependency: GNU Radio >= 3.7, CMakeLists.txt:86-87 of the reference, no exact
 * pin, VOLK dot products underneath).  Call sites in the reference:
 *   firdes::low_pass ................. lib/multi_block.cc:65-69, 75-79
 *   freq_xlating_fir_filter_ccf ...... lib/multi_block.cc:194-204, 269-275, 331-340
 *   complex_to_mag_squared ........... lib/multi_block.cc:206-213, 278-282
 *   mmse_fir_interpolator_ff ......... lib/multi_block.cc:97,103,133,139
 *   gr::fast_atan2f .................. lib/multi_block.cc:166
 *   gr::branchless_clip .............. lib/multi_block.cc:144
 *
 * Parity status of THIS boundary: "parity unpinned" -- the reference ships no
 * test that pins GNU Radio's floats; this file follows
 * SURVEY.md Appendix A (A.2-A.7) and is pinned indirectly: compiled under the
 * verbatim reference sources (oracle/shim + oracle/Makefile -> oracle/_ref) it
 * reproduces the LAPs documented in the reference's doc/README.first:45-67 and
 * the survey-time stdout digests (tests/test_oracle_ref.py).
 *
 * All arithmetic is IEEE binary32 unless stated, one rounding per operation
 * (build with -ffp-contract=off), sums in ascending index order.
 *
 * Used by: oracle/shim/ (the GNU Radio stand-in under the verbatim reference
 * build) and oracle/btb_oracle.c (the C restatement of the hot path).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use it.
 */
#ifndef BTB_ORACLE_GR_ARITH_H
#define BTB_ORACLE_GR_ARITH_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } gra_c32;

/* ---- A.2 firdes::low_pass(gain, fs, fc, tw, WIN_HANN) ------------------- */
static inline int gra_lowpass_ntaps(double fs, double tw)
{
  int ntaps = (int)(44.0 * fs / (22.0 * tw));   /* Hann: 44 dB max attenuation */
  if ((ntaps & 1) == 0) ntaps++;
  return ntaps;
}

/* taps must hold gra_lowpass_ntaps(fs, tw) floats */
static inline void gra_lowpass(double gain, double fs, double fc, double tw, float *taps)
{
  int ntaps = gra_lowpass_ntaps(fs, tw);
  float *w = (float *)malloc(sizeof(float) * (size_t)ntaps);
  int M = (ntaps - 1) / 2;
  double fwT0 = 2 * M_PI * fc / fs;
  for (int n = 0; n < ntaps; n++)
    w[n] = (float)(0.5 - 0.5 * cos((2 * M_PI * n) / (float)(ntaps - 1)));
  for (int n = -M; n <= M; n++) {
    if (n == 0)
      taps[n + M] = (float)(fwT0 / M_PI * w[n + M]);
    else
      taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
  }
  double fmax = taps[0 + M];
  for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M];
  gain /= fmax;
  for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain);
  free(w);
}

/* ---- A.3 freq_xlating_fir_filter_ccf ------------------------------------ */
typedef struct {
  int      ntaps;
  int      decim;
  gra_c32 *rtaps;       /* band-pass taps, REVERSED (rtaps[k] multiplies in[i*D+k]) */
  gra_c32  phase;       /* rotator state: persists across work() calls */
  gra_c32  incr;
  unsigned counter;
} gra_fxlat;

/* std::abs(gr_complex) == cabsf == hypotf; std::exp(gr_complex(0,x)) == cexpf == (cosf x, sinf x) */
static inline float gra_cabsf(gra_c32 z) { return hypotf(z.re, z.im); }
static inline gra_c32 gra_expj(float x) { gra_c32 r; r.re = cosf(x); r.im = sinf(x); return r; }

static inline void gra_fxlat_init(gra_fxlat *f, int decim, const float *taps, int ntaps,
                                  double center_freq, double fs)
{
  f->ntaps = ntaps;
  f->decim = decim;
  f->rtaps = (gra_c32 *)malloc(sizeof(gra_c32) * (size_t)ntaps);
  float fwT0 = (float)(2 * M_PI * center_freq / fs);
  for (int i = 0; i < ntaps; i++) {
    gra_c32 e = gra_expj(i * fwT0);                    /* std::exp(gr_complex(0, i*fwT0)) */
    /* float tap times complex: (t*re, t*im) */
    gra_c32 c;
    c.re = taps[i] * e.re;
    c.im = taps[i] * e.im;
    f->rtaps[ntaps - 1 - i] = c;
  }
  gra_c32 inc = gra_expj(-fwT0 * decim);
  float a = gra_cabsf(inc);
  f->incr.re = inc.re / a;
  f->incr.im = inc.im / a;
  f->phase.re = 1.0f;
  f->phase.im = 0.0f;
  f->counter = 0;
}

static inline void gra_fxlat_free(gra_fxlat *f) { free(f->rtaps); f->rtaps = NULL; }

static inline int gra_fxlat_ninput_to_noutput(const gra_fxlat *f, int ninput)
{
  int n = ninput - f->ntaps + 1;
  if (n < 0) n = 0;
  return n / f->decim;
}

/* acc = sum_k in[k] * rtaps[k], k ascending, complex multiply-add in fp32 */
static inline gra_c32 gra_dot_cc(const gra_c32 *in, const gra_c32 *rtaps, int ntaps)
{
  float ar = 0.0f, ai = 0.0f;
  for (int k = 0; k < ntaps; k++) {
    float a = in[k].re, b = in[k].im, c = rtaps[k].re, d = rtaps[k].im;
    float pr = a * c - b * d;
    float pi = a * d + b * c;
    ar = ar + pr;
    ai = ai + pi;
  }
  gra_c32 r = { ar, ai };
  return r;
}

/* one rotator step: z = in*phase; phase *= incr; renormalise every 512 */
static inline gra_c32 gra_rotate(gra_c32 in, gra_c32 *phase, gra_c32 incr, unsigned *counter)
{
  (*counter)++;
  gra_c32 z;
  z.re = in.re * phase->re - in.im * phase->im;
  z.im = in.re * phase->im + in.im * phase->re;
  gra_c32 p;
  p.re = phase->re * incr.re - phase->im * incr.im;
  p.im = phase->re * incr.im + phase->im * incr.re;
  if ((*counter % 512) == 0) {
    float a = gra_cabsf(p);
    p.re = p.re / a;
    p.im = p.im / a;
  }
  *phase = p;
  return z;
}

static inline int gra_fxlat_work(gra_fxlat *f, int nout, const gra_c32 *in, gra_c32 *out)
{
  for (int i = 0; i < nout; i++) {
    gra_c32 acc = gra_dot_cc(in + (size_t)i * f->decim, f->rtaps, f->ntaps);
    out[i] = gra_rotate(acc, &f->phase, f->incr, &f->counter);
  }
  return nout;
}

/* ---- A.4 complex_to_mag_squared ----------------------------------------- */
static inline float gra_mag2(gra_c32 z) { return z.re * z.re + z.im * z.im; }

/* ---- A.5 mmse_fir_interpolator_ff --------------------------------------- */
#define GRA_MMSE_NTAPS 8
#define GRA_MMSE_NSTEPS 128

/* Table regenerated in closed form (least squares against an ideal band-limited
 * (|f| <= 0.25) fractional delay), each element rounded through "%.5e" exactly
 * like GNU Radio's generated interpolator_taps.h. */
static inline void gra_mmse_table(float t[GRA_MMSE_NSTEPS + 1][GRA_MMSE_NTAPS])
{
  for (int s = 0; s <= GRA_MMSE_NSTEPS; s++) {
    double R[8][9];
    double tau = -(double)s / GRA_MMSE_NSTEPS;
    for (int i = 0; i < 8; i++) {
      double pi_ = i - 4;
      for (int j = 0; j < 8; j++) {
        double d = pi_ - (double)(j - 4);
        R[i][j] = (d == 0.0) ? 0.5 : sin(2 * M_PI * 0.25 * d) / (M_PI * d);
      }
      double d = pi_ - tau;
      R[i][8] = (d == 0.0) ? 0.5 : sin(2 * M_PI * 0.25 * d) / (M_PI * d);
    }
    /* Gaussian elimination with partial pivoting */
    for (int c = 0; c < 8; c++) {
      int piv = c;
      for (int r = c + 1; r < 8; r++)
        if (fabs(R[r][c]) > fabs(R[piv][c])) piv = r;
      if (piv != c)
        for (int k = 0; k < 9; k++) { double tmp = R[c][k]; R[c][k] = R[piv][k]; R[piv][k] = tmp; }
      for (int r = c + 1; r < 8; r++) {
        double m = R[r][c] / R[c][c];
        for (int k = c; k < 9; k++) R[r][k] -= m * R[c][k];
      }
    }
    double h[8];
    for (int r = 7; r >= 0; r--) {
      double acc = R[r][8];
      for (int k = r + 1; k < 8; k++) acc -= R[r][k] * h[k];
      h[r] = acc / R[r][r];
    }
    for (int i = 0; i < 8; i++) {
      char buf[40];
      double v = h[i];
      if (fabs(v) < 5e-10) v = 0.0;           /* exact-zero entries of rows 0 and 128 */
      snprintf(buf, sizeof buf, "%.5e", v);
      t[s][i] = strtof(buf, NULL);
    }
  }
}

/* interpolate(in, mu): sum_k in[k]*t[imu][7-k], k ascending. Returns 0 and
 * sets *bad when imu is outside 0..128 (GNU Radio throws). */
static inline float gra_mmse_interpolate(const float t[GRA_MMSE_NSTEPS + 1][GRA_MMSE_NTAPS],
                                         const float *in, float mu, int *bad)
{
  int imu = (int)rint(mu * GRA_MMSE_NSTEPS);
  if (imu < 0 || imu > GRA_MMSE_NSTEPS) { if (bad) *bad = 1; return 0.0f; }
  float acc = 0.0f;
  for (int k = 0; k < GRA_MMSE_NTAPS; k++)
    acc = acc + in[k] * t[imu][GRA_MMSE_NTAPS - 1 - k];
  return acc;
}

/* ---- A.6 gr::fast_atan2f ------------------------------------------------- */
static inline void gra_atan_table(float T[257])
{
  for (int i = 0; i < 256; i++) T[i] = (float)atan(i / 255.0);
  T[256] = T[255];
}

static inline float gra_fast_atan2f(const float T[257], float y, float x)
{
  float x_abs, y_abs, z, alpha, angle, base_angle;
  int index;
  y_abs = fabsf(y);
  x_abs = fabsf(x);
  if (!((y_abs > 0.0f) || (x_abs > 0.0f))) return 0.0f;
  if (y_abs < x_abs) z = y_abs / x_abs;
  else               z = x_abs / y_abs;
  if (z < 0.003921569f) {
    base_angle = z;
  } else {
    alpha = z * 255.0f;
    index = ((int)alpha) & 0xff;
    alpha = alpha - (float)index;
    base_angle = T[index];
    base_angle = base_angle + (T[index + 1] - T[index]) * alpha;
  }
  if (x_abs > y_abs) {
    if (x >= 0.0f) {
      angle = (y >= 0.0f) ? base_angle : -base_angle;
    } else {
      angle = (float)3.14159265358979323846;
      if (y >= 0.0f) angle = angle - base_angle;
      else           angle = base_angle - angle;
    }
  } else {
    if (y >= 0.0f) {
      angle = (float)1.57079632679489661923;
      if (x >= 0.0f) angle = angle - base_angle;
      else           angle = angle + base_angle;
    } else {
      angle = (float)-1.57079632679489661923;
      if (x >= 0.0f) angle = angle + base_angle;
      else           angle = angle - base_angle;
    }
  }
  return angle;
}

/* ---- A.7 gr::branchless_clip -------------------------------------------- */
static inline float gra_branchless_clip(float x, float clip)
{
  float x1 = fabsf(x + clip);
  float x2 = fabsf(x - clip);
  x1 = x1 - x2;
  return (float)(0.5 * x1);
}

#ifdef __cplusplus
}
#endif
#endif
