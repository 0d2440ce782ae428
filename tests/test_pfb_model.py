"""CPU tier: the mathematics of the throughput-mode (polyphase) channelizer against the oracle.

* tests/pfb_model.py (numpy, brute-force DFT) restates the algebra of rx_pfb.cu independently;
* emul_pfb (tests/emul) drives the PRODUCT's host tables (PfbDesign in plan.cpp: branch taps, Good-Thomas index
  maps, column order, DFT matrix, kappa) in the kernel's stage order;
both must reproduce the oracle's direct-form DDC (lib/multi_block.cc:180-228) up to float rounding, on every
BASELINE geometry.  What is compared: the demod floats (the only thing downstream stages see) and the window energy.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import FILES, ROOT, load_excerpt
from oracle import oracle as O
from pfb_model import PfbModel, fast_atan2f

EMUL = os.path.join(ROOT, "tests", "emul", "libbtb_emul.so")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(EMUL):
        import __graft_entry__ as ge
        ge.build()
    L = C.CDLL(EMUL)
    L.emul_pfb.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                           C.POINTER(C.c_double), C.c_void_p]
    return L


def synth_window(fs, fc, seed=3):
    import gr_bluetooth_b200  # noqa: F401
    from gr_bluetooth_b200 import synth
    P = O.Plan(fs, fc)
    nsl = (P.H + P.S - 1) // P.S + 1
    iq, _ = synth.generate(fs, fc, nsl, seed=seed, occupancy=0.3)
    return P, iq[:P.H]


def demod_from_Z(P, Z, kappa):
    prod = Z[1:] * np.conj(Z[:-1]) * kappa[None, :]
    d = np.zeros(Z.shape, np.float32)
    d[1:] = np.float32(P.demod_gain) * fast_atan2f(P.atan_table(), prod.imag.astype(np.float32), prod.real.astype(np.float32))
    return d


CASES = [("headset3", None), ("headset2", None), ("headset1", None), ("keyboard1", None),
         ("synth30", (30e6, 2414e6)), ("synth100", (100e6, 2441e6))]


@pytest.mark.parametrize("name,geom", CASES)
def test_polyphase_equals_direct_form(emul, name, geom):
    if geom is None:
        ex = load_excerpt(name, "stateless")
        fs, fc = FILES[name]
        P = O.Plan(fs, fc)
        k = min(ex["nslots"] - 1, 12)
        x = np.concatenate([np.zeros(P.H - 1, np.complex64), ex["iq"]])
        win = x[k * P.S:k * P.S + P.H]
    else:
        P, win = synth_window(*geom)
    n_grid = P.n_ddc if P.fs < 50e6 else 1500
    hits, d = P.window(win, slot=0, stateless=True)
    ok = np.where(d["pass_"] > 0)[0]
    assert len(ok) >= 1
    ref = d["demod"][ok][:, 1:n_grid - 1]
    wrap = 2 * np.pi * P.demod_gain

    def cmp(dem):
        dd = np.abs(dem.T[ok][:, 1:n_grid - 1] - ref)
        return np.minimum(dd, np.abs(dd - wrap))

    # 1. independent numpy model, fp32
    m = PfbModel(P, np.complex64)
    Zm = m.channelize(win, n_grid)
    dm = cmp(m.demod(Zm, P.atan_table()))
    # 2. the product's tables in the kernel's stage order, fp64
    Z = np.zeros((n_grid, P.nch, 2))
    kap = np.zeros((P.nch, 2), np.float32)
    dims = np.zeros(12, np.int32)
    phi = C.c_double()
    w = np.ascontiguousarray(win, np.complex64)
    rc = emul.emul_pfb(P.fs, P.fc, 3125, w.ctypes.data, len(w), n_grid, Z.ctypes.data, kap.ctypes.data, C.byref(phi),
                       dims.ctypes.data)
    assert rc == 0
    M, D, Q, N1, N2 = [int(v) for v in dims[:5]]
    assert M == int(round(P.fs / 1e6)) and 2 * D == M and N1 * N2 == M and np.gcd(N1, N2) == 1
    assert abs(phi.value - m.phi) < 1e-9 and int(dims[11]) == m.a0
    Zp = Z[..., 0] + 1j * Z[..., 1]
    kappa = kap[:, 0] + 1j * kap[:, 1]
    assert np.allclose(kappa, m.kappa, atol=1e-6)
    dp = cmp(demod_from_Z(P, Zp, kappa))
    # energies (whole window) from the model
    if n_grid == P.n_ddc:
        e = (np.abs(Zm.astype(np.complex128)) ** 2).mean(axis=0)
        assert np.max(np.abs(e[ok] / d["energy"][ok] - 1)) < 2e-5
    # |Z| of the two formulations agree to fp32 rounding (relative to the rms level: bins between channels are small)
    scale = np.sqrt(np.mean(np.abs(Zp) ** 2))
    assert np.max(np.abs(np.abs(Zp) - np.abs(Zm))) < 2e-4 * scale
    # stated tolerance of the throughput mode's demod floats (range +-4 at 100 Msps): the deviation from the reference's
    # direct form is dominated by the reference's own fp32 tap phases (i * theta in float, up to 1.6e3 rad at 100 Msps)
    tol_med, tol_999 = (2e-5, 5e-3) if P.fs >= 30e6 else (3e-6, 5e-4)
    for dd in (dm, dp):
        assert np.median(dd) < tol_med, np.median(dd)
        assert np.quantile(dd, 0.999) < tol_999, np.quantile(dd, 0.999)


def exact_noise_energies(P, iq, b, n_out):
    """|y_j|^2, j < n_out, of window b's noise DDCs (lib/multi_block.cc:253-287) for every channel, float64 (FFT convolution
    with the oracle's prototype; the DDC's unit-modulus rotator does not change |y|)."""
    h = np.asarray(P.noise_proto(), np.float64)[::-1]
    Nn, D = len(h), P.D
    L = (n_out - 1) * D + Nn
    n0 = b * P.S + P.fns
    seg = iq[n0:n0 + L].astype(np.complex128)
    F = 1 << int(np.ceil(np.log2(L + Nn)))
    Hf = np.fft.fft(h, F)
    n = np.arange(L)
    g = np.empty((P.nch, n_out))
    for c in range(P.nch):
        f = 2402e6 + (P.ch_lo + c) * 1e6 + 790e3 - P.fc
        y = np.fft.ifft(np.fft.fft(seg * np.exp(-2j * np.pi * f / P.fs * n), F) * Hf)[Nn - 1:Nn - 1 + n_out * D:D]
        g[c] = np.abs(y) ** 2
    return g


@pytest.mark.parametrize("s,n_extra,n_free,khz", [(4, 2, 12, 90.0), (2, 2, 8, 90.0)])
def test_subsampled_noise_energy_quadrature(emul, s, n_extra, n_free, khz):
    """The estimator of the throughput mode sums |y|^2 over every s-th noise-DDC output with the weights of
    nest_quadrature() (plan.cpp).  With exact |y_j|^2 (float64, the oracle's prototype) of the benchmark's synthetic
    traffic -- light and with 25 dB bursts in half of all channel-slots -- the weighted sub-sampled sum reproduces the
    reference's sum over all 850 outputs to < 2e-5 relative (the mode's tolerance on the snr is 5e-3 dB = 1.2e-3);
    what the kernel adds on top is fp32 rounding."""
    from gr_bluetooth_b200 import synth
    fs, fc = 100e6, 2441e6
    P = O.Plan(fs, fc)
    N = P.n_noise
    emul.emul_nest_quadrature.restype = C.c_double
    emul.emul_nest_quadrature.argtypes = [C.c_int] * 4 + [C.c_double, C.c_void_p, C.c_void_p]
    w = np.zeros(N, np.float32)
    n_used = C.c_int32(0)
    omega = 2 * np.pi * khz * 1e3 * P.D / fs
    res = emul.emul_nest_quadrature(N, s, n_extra, n_free, omega, w.ctypes.data, C.byref(n_used))
    n = n_used.value
    assert n == (N - 1) // s + 1 + n_extra and 0 <= res < 1e-3
    w = w[:n].astype(np.float64)
    assert np.all(w[n_free:n - n_free] == s) and abs(w.sum() - N) < 1e-3       # exact for a constant
    worst = 0.0
    for occ, snr, seed in ((0.05, 17.0, 5), (0.5, 25.0, 6)):
        iq, _ = synth.generate(fs, fc, 4, seed=seed, occupancy=occ, snr_db=snr)
        for b in (0, 1):
            g = exact_noise_energies(P, iq, b, s * n)
            ref = g[:, :N].sum(axis=1)
            est = (g[:, ::s][:, :n] * w[None, :]).sum(axis=1)
            worst = max(worst, float(np.max(np.abs(est / ref - 1))))
    print("stride %d: worst relative deviation of the weighted sub-sampled sum %.2e (design residual %.1e)" % (s, worst, res))
    assert worst < 2e-5


@pytest.mark.parametrize("N,s", [(850, 4), (850, 2), (1250, 4), (128, 4), (333, 4), (64, 2)])
def test_quadrature_weights_other_lengths(emul, N, s):
    """nest_quadrature() for the window lengths of the other rates (n_noise = 850 at 100 / 30 / 8 Msps geometries with
    D-dependent omega, odd and short lengths as edge cases): exact for a constant, in-band tones to the design residual,
    interior weights pinned, nothing wild at the ends."""
    emul.emul_nest_quadrature.restype = C.c_double
    emul.emul_nest_quadrature.argtypes = [C.c_int] * 4 + [C.c_double, C.c_void_p, C.c_void_p]
    w = np.zeros(N + 8, np.float32)
    n_used = C.c_int32(0)
    omega = 2 * np.pi * 90e3 / 2e6
    res = emul.emul_nest_quadrature(N, s, 2, 12, omega, w.ctypes.data, C.byref(n_used))
    n = n_used.value
    assert n == (N - 1) // s + 1 + 2 and 0 <= res < 2e-3
    w = w[:n].astype(np.float64)
    assert abs(w.sum() - N) < 2e-3 and np.abs(w).max() < 12 * s
    if n > 24:
        assert np.all(w[12:n - 12] == s)
    rng = np.random.default_rng(N + s)
    j, m = np.arange(N), np.arange(n) * s
    for om in rng.uniform(-omega, omega, 50):
        ph = rng.uniform(0, 2 * np.pi)
        full = np.cos(om * j + ph).sum()
        sub = (w * np.cos(om * m + ph)).sum()
        assert abs(full - sub) < max(2.5 * res, 1e-4) + 1e-9, (om, full, sub)
    # invalid arguments are refused
    assert emul.emul_nest_quadrature(8, 4, 2, 12, omega, w.ctypes.data, C.byref(n_used)) < 0
    assert emul.emul_nest_quadrature(N, 0, 2, 12, omega, w.ctypes.data, C.byref(n_used)) < 0


@pytest.mark.parametrize("fs,fc", [(100e6, 2441e6), (30e6, 2414e6)])
def test_folded_noise_estimator_tables_against_oracle(emul, fs, fc):
    """The noise estimator of the throughput mode driven by the product's host tables (design_noise: flat tap array,
    Good-Thomas maps, DFT matrix, column order; nest_quadrature weights) in the order rx_nest.cu uses them -- fold * M
    virtual branches, fold, N1- and N2-point DFTs, weighted |Z|^2 -- reproduces the oracle's off-channel energy
    (lib/multi_block.cc:253-287: 20001-tap DDC at f_ch + 790 kHz, mean |y|^2 over the first slot's outputs) of every
    channel to < 5e-4, with every output and with every 4th one; the two evaluations agree to 2e-5."""
    from gr_bluetooth_b200 import synth
    emul.emul_nest.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    P = O.Plan(fs, fc)
    first, B = 7, 2
    iq, _ = synth.generate(fs, fc, first + B + 1, seed=41, occupancy=0.25, snr_db=22.0)
    o = P.run(iq, first_call=first, num_calls=B, stateless=True, threads=8, want_energy=True)
    w0 = first * P.S - (P.H - 1)
    seg = np.ascontiguousarray(iq[w0:w0 + (B - 1) * P.S + P.H], np.complex64)
    for b in range(B):
        est = {}
        for fold in (0, 2):
            e = np.zeros(P.nch)
            rc = emul.emul_nest(fs, fc, 3125, seg.ctypes.data, len(seg), b, fold, e.ctypes.data)
            assert rc == 0, rc
            est[fold] = e / P.n_noise
            dev = np.abs(est[fold] / o["noise"][b] - 1)
            assert dev.max() < 5e-4, (fold, b, dev.max())
        assert np.abs(est[2] / est[0] - 1).max() < 2e-5
