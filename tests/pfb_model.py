"""TEST INFRASTRUCTURE: numpy model of the polyphase (WOLA) channelizer of the throughput mode.

Not the oracle (that is oracle/, the reference's own arithmetic) and not the product (that is
csrc/rx_pfb.cu): an independent statement of the SAME mathematics the kernel implements, used by
the CPU tier to check the algebra (branch folding, DFT bins, the constant of the differential
product) against the oracle's direct-form DDC, and to measure how far tolerance-level floats are
from the reference's.

Reference structure being restated (lib/multi_block.cc:180-228, 329-341): channel c's DDC is
    z_c[i] = rot_c[i] * sum_k x[fcs + i*D + k] * rt_c[k],   rt_c[k] = h[Nc-1-k] e^{j th_c (Nc-1-k)}
with th_c = 2 pi f_c / fs, f_c = a_c + phi MHz (a_c integer, phi common to all channels).  With
x'[n] = x[n] e^{-j 2 pi phi n / M} (M = fs / 1 MHz) and h'[k] = h[Nc-1-k]:
    u_g[r]  = sum_q x'[n0 + r + M q] h'[r + M q],   n0 = fcs + g D            (M branches, real taps)
    Yt_c[g] = sum_r e^{-j 2 pi a_c r / M} u_g[r]                              (DFT at bin a_c)
    z_c[i] z_c[i-1]^* = Yt_c[g] Yt_c[g-1]^* e^{-j 2 pi a_c D / M}             (all other factors cancel)
    |z_c[i]| = |Yt_c[g]|
"""
import numpy as np


class PfbModel:
    def __init__(self, P, dtype=np.complex64):
        """P: oracle.Plan (geometry + the reference's prototype).  dtype: complex64 mimics the kernel's
        precision, complex128 is the exact-arithmetic limit of the model."""
        self.P = P
        self.ct = np.dtype(dtype)
        self.rt = np.float32 if self.ct == np.complex64 else np.float64
        M = P.fs / 1e6
        self.M = int(round(M))
        assert abs(M - self.M) < 1e-9, "polyphase mode needs an integer number of samples per MHz"
        f0 = (2402e6 + P.ch_lo * 1e6 - P.fc) / 1e6
        self.a0 = int(np.floor(f0 + 1e-9))
        self.phi = f0 - self.a0
        if abs(self.phi) < 1e-9:
            self.phi = 0.0
        self.a = self.a0 + np.arange(P.nch)
        self.Q = (P.Nc + self.M - 1) // self.M
        h = P.chan_proto().astype(np.float64)[::-1]                     # h'[k] = h[Nc-1-k]
        hp = np.zeros(self.Q * self.M)
        hp[:P.Nc] = h
        self.hq = hp.reshape(self.Q, self.M).astype(self.rt)            # [q][r]
        r = np.arange(self.M)
        self.W = np.exp(-2j * np.pi * np.outer(r, self.a % self.M) / self.M).astype(self.ct)   # [r][c]
        self.kappa = np.exp(-2j * np.pi * self.a * P.D / self.M).astype(self.ct)               # [c]

    def prerotate(self, x, n_origin=0):
        if self.phi == 0.0:
            return x.astype(self.ct)
        n = np.arange(len(x), dtype=np.float64) + n_origin
        return (x.astype(np.complex128) * np.exp(-2j * np.pi * self.phi * n / self.M)).astype(self.ct)

    def channelize(self, x, n_grid):
        """x: samples, x[0] = first sample of the first window; -> Yt [n_grid][nch]."""
        P, M, Q = self.P, self.M, self.Q
        xp = self.prerotate(x)
        need = P.fcs + (n_grid - 1) * P.D + Q * M
        if len(xp) < need:
            xp = np.concatenate([xp, np.zeros(need - len(xp), self.ct)])
        out = np.empty((n_grid, P.nch), self.ct)
        step = 2048
        for g0 in range(0, n_grid, step):
            g1 = min(n_grid, g0 + step)
            idx = (P.fcs + np.arange(g0, g1) * P.D)[:, None, None] + (np.arange(Q) * M)[None, :, None] + np.arange(M)[None, None, :]
            u = (xp[idx] * self.hq[None, :, :]).sum(axis=1, dtype=self.ct)        # [g][r]
            out[g0:g1] = u @ self.W
        return out

    def demod(self, Yt, atan_table):
        """Demod floats on the grid: d[g] for g >= 1 (d[0] = 0), fast_atan2f like the reference."""
        prod = (Yt[1:] * np.conj(Yt[:-1]) * self.kappa[None, :]).astype(self.ct)
        d = np.zeros(Yt.shape, np.float32)
        d[1:] = np.float32(self.P.demod_gain) * fast_atan2f(atan_table, prod.imag.astype(np.float32), prod.real.astype(np.float32))
        return d


def fast_atan2f(T, y, x):
    """Vectorised gr::fast_atan2f, float32."""
    T = np.asarray(T, np.float32)
    y = np.asarray(y, np.float32)
    x = np.asarray(x, np.float32)
    ya, xa = np.abs(y), np.abs(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(ya < xa, ya / xa, xa / ya).astype(np.float32)
    z = np.nan_to_num(z, nan=0.0)
    alpha = (z * np.float32(255.0)).astype(np.float32)
    idx = alpha.astype(np.int32) & 0xFF
    frac = (alpha - idx.astype(np.float32)).astype(np.float32)
    base = (T[idx] + (T[idx + 1] - T[idx]) * frac).astype(np.float32)
    base = np.where(z < np.float32(0.003921569), z, base)
    pi = np.float32(3.14159265358979323846)
    h = np.float32(1.57079632679489661923)
    ang = np.where(xa > ya,
                   np.where(x >= 0, np.where(y >= 0, base, -base), np.where(y >= 0, pi - base, base - pi)),
                   np.where(y >= 0, np.where(x >= 0, h - base, h + base), np.where(x >= 0, -h + base, -h - base)))
    ang = np.where((ya > 0) | (xa > 0), ang, np.float32(0.0))
    return ang.astype(np.float32)
