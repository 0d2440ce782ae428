"""CPU tier: the per-thread bodies of the baseline CUDA kernels (rx_bodies.cuh / rx_math.cuh)
and the product's host design step (plan.cpp), compiled for the host by tests/emul, against
the oracle.  This checks the kernel LOGIC without a GPU; the -m gpu tier checks the kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, FILES, load_excerpt
from oracle import oracle as O

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


class EHit(C.Structure):
    _fields_ = [("slot", C.c_int32), ("channel", C.c_int16), ("kind", C.c_int16), ("offset", C.c_int32),
                ("n_symbols", C.c_int32), ("lap", C.c_uint32), ("snr", C.c_double)]


class EOut(C.Structure):
    _fields_ = [("energy", C.c_void_p), ("noise", C.c_void_p), ("nsym", C.c_void_p), ("bits", C.c_void_p),
                ("bits_stride", C.c_int32), ("hits", C.c_void_p), ("hits_cap", C.c_int32), ("nhits", C.c_int32),
                ("mm", C.c_float * 3)]


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libbtb_emul.so")
    srcs = [os.path.join(EMUL_DIR, "emul.cpp"), os.path.join(ROOT, "gr-bluetooth_b200", "csrc", "plan.cpp")]
    deps = srcs + [os.path.join(ROOT, "gr-bluetooth_b200", "csrc", f) for f in ("rx_bodies.cuh", "rx_math.cuh", "plan.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               *srcs, "-o", so])
    L = C.CDLL(so)
    L.emul_run.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                           C.c_int, C.POINTER(EOut)]
    L.emul_sync_word.restype = C.c_uint64
    L.emul_sync_word.argtypes = [C.c_uint32]
    L.emul_plan_info.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.emul_plan_tables.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 6
    return L


def test_sync_word_lut_source(emul):
    # SURVEY.md Appendix B check values
    assert emul.emul_sync_word(0x9E8B33) == 0x4E7A2CCE331A3AE2
    assert emul.emul_sync_word(0x24D952) == 0xB093654ABEDEF6FA
    assert emul.emul_sync_word(0) == 0xB0000002C7820E7E
    rng = np.random.default_rng(5)
    for lap in rng.integers(0, 1 << 24, 500):
        bits = O.acgen_bits(int(lap))
        assert emul.emul_sync_word(int(lap)) == sum(int(bits[4 + i]) << i for i in range(64))


@pytest.mark.parametrize("fs,fc,extra", [(2e6, 2476e6, 3125), (4e6, 2476e6, 3125), (8e6, 2476.5e6, 3125),
                                         (30e6, 2414e6, 3125), (100e6, 2441e6, 3125), (100e6, 2441e6, 68),
                                         (8e6, 2476.5e6, 68)])
def test_product_design_equals_oracle(emul, fs, fc, extra):
    P = O.Plan(fs, fc, extra_symbols=extra)
    info = (C.c_int32 * 16)()
    assert emul.emul_plan_info(fs, fc, 10.0, extra, info) == 0
    assert list(info[:14]) == [P.S, P.H, P.D, P.Nc, P.Nn, P.fcs, P.fns, P.ch_lo, P.ch_hi, P.nch, P.n_ddc,
                               P.n_noise, P.n_ddc - 1, P.S // P.D]
    for chi in sorted({0, P.nch // 2, P.nch - 1}):
        ct = np.zeros(P.Nc, np.complex64); nt = np.zeros(P.Nn, np.complex64)
        mm = np.zeros((129, 8), np.float32); at = np.zeros(257, np.float32)
        lut = np.zeros(769, np.uint64); inc = np.zeros(4, np.float32)
        assert emul.emul_plan_tables(fs, fc, extra, chi, ct.ctypes.data, nt.ctypes.data, mm.ctypes.data,
                                     at.ctypes.data, lut.ctypes.data, inc.ctypes.data) == 0
        assert np.array_equal(ct.view(np.uint32), P.chan_rtaps(chi).view(np.uint32))
        assert np.array_equal(nt.view(np.uint32), P.noise_rtaps(chi).view(np.uint32))
        assert np.array_equal(mm, P.mmse_table()) and np.array_equal(at, P.atan_table())
        ci, ni = P.rot_incr(chi), P.rot_incr(chi, noise=True)
        assert np.array_equal(inc, np.array([ci.real, ci.imag, ni.real, ni.imag], np.float32))


def run_emul(emul, P, x, B, first, stateless, mm0):
    nch, stride = P.nch, P.n_ddc
    en = np.zeros((B, nch)); nz = np.zeros((B, nch)); nsym = np.zeros((B, nch), np.int32)
    bits = np.zeros((B, nch, stride), np.uint8); hits = (EHit * 4096)()
    eo = EOut(en.ctypes.data, nz.ctypes.data, nsym.ctypes.data, bits.ctypes.data, stride, C.addressof(hits), 4096, 0,
              (C.c_float * 3)(*mm0))
    x = np.ascontiguousarray(x, np.complex64)
    assert emul.emul_run(P.fs, P.fc, P.snr_db, P.extra_symbols, int(stateless), 3, x.ctypes.data, B, first,
                         C.byref(eo)) == 0
    got = [(h.slot, h.channel, h.kind, h.offset, h.n_symbols, h.lap, h.snr) for h in hits[:eo.nhits]]
    return dict(energy=en, noise=nz, nsym=nsym, bits=bits, hits=got, mm=list(eo.mm))


def oracle_hits(o):
    return [(int(h["slot"]), int(h["channel"]), int(h["kind"]), int(h["offset"]), int(h["len"]), int(h["lap"]),
             float(h["snr"])) for h in o["hits"]]


@pytest.mark.parametrize("name,first,B", [("headset3", 0, 100), ("headset3", 180, 12), ("keyboard1", 0, 12),
                                          ("headset1", 0, 9), ("headset2", 28, 8)])
@pytest.mark.parametrize("stateless", [True, False])
def test_kernel_bodies_equal_oracle(emul, name, first, B, stateless):
    ex = load_excerpt(name, "chained")
    P = O.Plan(ex["fs"], ex["fc"])
    iq = ex["iq"]
    st = O.State(P)
    if not stateless and first:
        P.run(iq, first_call=0, num_calls=first, stateless=False, state=st)
    mm0 = list(st.mm)
    o = P.run(iq, first_call=first, num_calls=B, stateless=stateless, state=None if stateless else st,
              want_bits=True, want_energy=True)
    w0 = first * P.S - (P.H - 1)
    n = (B - 1) * P.S + P.H
    x = np.zeros(n, np.complex64)
    lo = max(w0, 0)
    x[lo - w0:] = iq[lo:w0 + n]
    e = run_emul(emul, P, x, B, first, stateless, mm0)
    assert np.array_equal(e["energy"], o["energy"], equal_nan=True)
    assert np.array_equal(e["noise"], o["noise"], equal_nan=True)
    assert np.array_equal(e["nsym"], o["nsym"]) and np.array_equal(e["bits"], o["bits"])
    assert e["hits"] == oracle_hits(o)
    if not stateless:
        assert np.array_equal(np.array(e["mm"], np.float32), st.mm)


def test_search_edge_cases(emul):
    """Hand-built symbol rows through the oracle's search: hit at lag 0, two packets 68 apart
    (the skip rule), a hit exactly at the last searched lag, too-short rows."""
    ac = O.acgen_bits(0x9E8B33)[:72]
    row = np.zeros(1000, np.uint8)
    row[0:72] = ac
    row[68 + 30:68 + 30 + 72] = O.acgen_bits(0x24D952)[:72]
    assert O.sniff_ac(row, 625) == 0
    assert O.sniff_ac(row[68:], 625 - 68) == 30
    row2 = np.zeros(1000, np.uint8)
    row2[624:624 + 72] = ac
    assert O.sniff_ac(row2, 625) == 624 and O.sniff_ac(row2, 624) == -1
    assert O.sniff_ac(np.zeros(80, np.uint8), 0) == -1
