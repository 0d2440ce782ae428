"""ctypes binding of the C restatement (oracle/btb_oracle.c).

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product path never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libbtb_oracle.so")


class Info(C.Structure):
    _fields_ = [("fs", C.c_double), ("fc", C.c_double), ("snr_db", C.c_double),
                ("extra_symbols", C.c_int), ("S", C.c_int), ("H", C.c_int), ("D", C.c_int),
                ("Nc", C.c_int), ("Nn", C.c_int), ("fcs", C.c_int), ("fns", C.c_int),
                ("ch_lo", C.c_int), ("ch_hi", C.c_int), ("nch", C.c_int),
                ("n_ddc", C.c_int), ("n_noise", C.c_int),
                ("demod_gain", C.c_float), ("omega_mid", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("slot", C.c_int32), ("channel", C.c_int16), ("kind", C.c_int16),
                ("offset", C.c_int32), ("len", C.c_int32), ("lap", C.c_uint32),
                ("snr", C.c_double)]


HIT_DTYPE = np.dtype([("slot", "<i4"), ("channel", "<i2"), ("kind", "<i2"), ("offset", "<i4"),
                      ("len", "<i4"), ("lap", "<u4"), ("snr", "<f8")], align=True)


class ChanResult(C.Structure):
    _fields_ = [("chi", C.c_int32), ("processed", C.c_int32), ("pass_", C.c_int32), ("nsym", C.c_int32),
                ("ac_index", C.c_int32), ("lap", C.c_uint32), ("snr", C.c_double)]


class Debug(C.Structure):
    _fields_ = [("energy", C.c_void_p), ("noise", C.c_void_p), ("snr", C.c_void_p),
                ("pass_", C.c_void_p), ("nsym", C.c_void_p), ("bits", C.c_void_p),
                ("ddc", C.c_void_p), ("demod", C.c_void_p), ("soft", C.c_void_p)]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f))
                                             for f in ("btb_oracle.c", "btb_oracle.h", "gr_arith.h")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.btbo_plan_create.restype = C.c_void_p
        L.btbo_plan_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.btbo_plan_free.argtypes = [C.c_void_p]
        L.btbo_plan_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        for name in ("btbo_chan_proto", "btbo_noise_proto", "btbo_mmse_table", "btbo_atan_table"):
            getattr(L, name).restype = C.POINTER(C.c_float)
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("btbo_chan_rtaps", "btbo_noise_rtaps"):
            getattr(L, name).restype = C.POINTER(C.c_float)
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.btbo_rot_incr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.btbo_state_create.restype = C.c_void_p
        L.btbo_state_create.argtypes = [C.c_void_p]
        L.btbo_state_free.argtypes = [C.c_void_p]
        L.btbo_state_get_mm.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btbo_state_set_mm.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btbo_acgen_bits.argtypes = [C.c_uint32, C.c_void_p]
        L.btbo_acgen_bytes.argtypes = [C.c_uint32, C.c_void_p]
        L.btbo_check_ac.argtypes = [C.c_void_p, C.c_uint32]
        L.btbo_sniff_ac.argtypes = [C.c_void_p, C.c_int]
        L.btbo_sniff_aa.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.btbo_lut.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.btbo_bch_lag.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.btbo_find_ac_bch.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.btbo_le_index.argtypes = [C.c_double]
        L.btbo_demod.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.btbo_mm_cr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.btbo_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(Debug)]
        L.btbo_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long,
                               C.c_long, C.c_long, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.POINTER(C.c_int),
                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.btbo_window_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                                       C.c_void_p, C.c_void_p]
        L.btbo_header_present.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Plan:
    """multi_block constructor state (lib/multi_block.cc:40-120, 299-342)."""

    def __init__(self, fs, fc, snr_db=10.0, extra_symbols=3125):
        self.L = lib()
        self.h = self.L.btbo_plan_create(fs, fc, snr_db, extra_symbols)
        self.info = Info()
        self.L.btbo_plan_info(self.h, C.byref(self.info))
        for f, _ in Info._fields_:
            setattr(self, f, getattr(self.info, f))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.btbo_plan_free(self.h)
            self.h = None

    def chan_proto(self):
        return np.ctypeslib.as_array(self.L.btbo_chan_proto(self.h), (self.Nc,)).copy()

    def noise_proto(self):
        return np.ctypeslib.as_array(self.L.btbo_noise_proto(self.h), (self.Nn,)).copy()

    def chan_rtaps(self, chi):
        return np.ctypeslib.as_array(self.L.btbo_chan_rtaps(self.h, chi), (self.Nc * 2,)).copy().view(np.complex64)

    def noise_rtaps(self, chi):
        return np.ctypeslib.as_array(self.L.btbo_noise_rtaps(self.h, chi), (self.Nn * 2,)).copy().view(np.complex64)

    def rot_incr(self, chi, noise=False):
        out = (C.c_float * 2)()
        self.L.btbo_rot_incr(self.h, chi, int(noise), out)
        return np.complex64(complex(out[0], out[1]))

    def mmse_table(self):
        return np.ctypeslib.as_array(self.L.btbo_mmse_table(self.h), (129, 8)).copy()

    def atan_table(self):
        return np.ctypeslib.as_array(self.L.btbo_atan_table(self.h), (257,)).copy()

    def demod(self, ddc_out):
        x = np.ascontiguousarray(ddc_out, dtype=np.complex64)
        out = np.zeros(len(x) - 1, np.float32)
        self.L.btbo_demod(self.h, _ptr(x), _ptr(out), len(out))
        return out

    def mm_cr(self, demod, mm=None):
        mm = np.array([0.32, self.omega_mid, 0.0] if mm is None else mm, np.float32)
        x = np.ascontiguousarray(demod, dtype=np.float32)
        out = np.zeros(len(x), np.float32)
        n = self.L.btbo_mm_cr(self.h, _ptr(mm), _ptr(x), len(x), _ptr(out), len(x))
        return out[:n], mm

    def window(self, win, slot=0, stateless=True, state=None, debug=True):
        """One work() call on H complex samples; returns (hits, dbg dict)."""
        I = self
        w = np.ascontiguousarray(win, dtype=np.complex64)
        assert len(w) == I.H
        hits = np.zeros(256, HIT_DTYPE)
        nh = C.c_int(0)
        d = {}
        dbg = Debug()
        if debug:
            d = dict(energy=np.zeros(I.nch), noise=np.zeros(I.nch), snr=np.zeros(I.nch),
                     pass_=np.zeros(I.nch, np.int32), nsym=np.zeros(I.nch, np.int32),
                     bits=np.zeros((I.nch, I.H), np.uint8),
                     ddc=np.zeros((I.nch, I.n_ddc), np.complex64),
                     demod=np.zeros((I.nch, I.n_ddc - 1), np.float32),
                     soft=np.zeros((I.nch, I.n_ddc - 1), np.float32))
            for k, v in d.items():
                setattr(dbg, k, v.ctypes.data)
        st = state.h if state is not None else None
        if not stateless and st is None:
            raise ValueError("chained mode needs a State")
        rc = self.L.btbo_window(self.h, st, _ptr(w), slot, 1 if stateless else 0,
                                _ptr(hits), len(hits), C.byref(nh), C.byref(dbg) if debug else None)
        assert rc == 0
        return hits[:nh.value].copy(), d

    def window_list(self, win, state, chis, stop_lap=0xFFFFFFFF):
        """One multi_hopper work() call (multi_hopper_impl.cc:76-209) on the chained state.
        -> (list of ChanResult, symbols [n][H])"""
        w = np.ascontiguousarray(win, dtype=np.complex64)
        assert len(w) == self.H
        ch = np.ascontiguousarray(chis, dtype=np.int32)
        res = (ChanResult * len(ch))()
        sym = np.zeros((len(ch), self.H), np.uint8)
        rc = self.L.btbo_window_list(self.h, state.h, _ptr(w), _ptr(ch), len(ch), stop_lap, C.byref(res), _ptr(sym))
        assert rc == 0
        return list(res), sym

    def run(self, iq, first_call=0, num_calls=None, stateless=True, threads=1, state=None,
            iq_first=0, n_total=None, want_bits=False, want_energy=False, hits_cap=1 << 16):
        """Scheduler emulation over a sample array."""
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        n_total = len(x) + iq_first if n_total is None else n_total
        ncalls_all = (n_total + self.S - 1) // self.S
        if num_calls is None:
            num_calls = ncalls_all - first_call
        hits = np.zeros(hits_cap, HIT_DTYPE)
        nh = C.c_int(0)
        stride = self.n_ddc
        bits = np.zeros((num_calls, self.nch, stride), np.uint8) if want_bits else None
        nsym = np.zeros((num_calls, self.nch), np.int32) if want_bits else None
        en = np.zeros((num_calls, self.nch)) if want_energy else None
        nz = np.zeros((num_calls, self.nch)) if want_energy else None
        own = None
        if not stateless and state is None:
            own = state = State(self)
        rc = self.L.btbo_run(self.h, state.h if state is not None else None, _ptr(x), iq_first, len(x),
                             first_call, num_calls, 1 if stateless else 0, threads,
                             _ptr(hits), hits_cap, C.byref(nh), _ptr(bits), stride, _ptr(nsym),
                             _ptr(en), _ptr(nz))
        assert rc == 0, rc
        out = dict(hits=hits[:nh.value].copy())
        if want_bits:
            out.update(bits=bits, nsym=nsym)
        if want_energy:
            out.update(energy=en, noise=nz)
        return out


class State:
    def __init__(self, plan):
        self.plan = plan
        self.h = plan.L.btbo_state_create(plan.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.plan.L.btbo_state_free(self.h)
            self.h = None

    @property
    def mm(self):
        out = (C.c_float * 3)()
        self.plan.L.btbo_state_get_mm(self.h, out)
        return np.array(out[:], np.float32)


def acgen_bits(lap):
    out = np.zeros(72, np.uint8)
    lib().btbo_acgen_bits(lap, _ptr(out))
    return out


def acgen_bytes(lap):
    out = np.zeros(9, np.uint8)
    lib().btbo_acgen_bytes(lap, _ptr(out))
    return bytes(out)


def check_ac(stream, lap):
    s = np.ascontiguousarray(stream, np.uint8)
    return bool(lib().btbo_check_ac(_ptr(s), lap))


def sniff_ac(stream, limit=None):
    s = np.ascontiguousarray(stream, np.uint8)
    limit = len(s) - 72 if limit is None else limit
    return lib().btbo_sniff_ac(_ptr(s), limit)


LAP_ANY = 0xFFFFFFFF


def bch_lag(stream, max_ac_errors=1, lap=LAP_ANY):
    """libbtbb-style test of the lag at stream[0] (>= 68 symbols): (accepted, lap, corrected bits).  PARITY UNPINNED."""
    s = np.ascontiguousarray(stream, np.uint8)
    assert len(s) >= 68
    lap_out, n_err = C.c_uint32(0), C.c_int(0)
    ok = lib().btbo_bch_lag(_ptr(s), max_ac_errors, C.c_uint32(lap), C.byref(lap_out), C.byref(n_err))
    return bool(ok), int(lap_out.value), int(n_err.value)


def find_ac_bch(stream, search_length, lap=LAP_ANY, max_ac_errors=1):
    """btbb_find_ac restated (lib/multi_LAP_impl.cc:93): first accepted lag < search_length or -1, LAP, corrected bits."""
    s = np.ascontiguousarray(stream, np.uint8)
    assert len(s) >= search_length + 67
    lap_out, n_err = C.c_uint32(0), C.c_int(0)
    off = lib().btbo_find_ac_bch(_ptr(s), search_length, C.c_uint32(lap), max_ac_errors, C.byref(lap_out), C.byref(n_err))
    return int(off), int(lap_out.value), int(n_err.value)


def sniff_aa(stream, freq, limit=None):
    s = np.ascontiguousarray(stream, np.uint8)
    limit = len(s) - 72 if limit is None else limit
    return lib().btbo_sniff_aa(_ptr(s), limit, freq)


def lut(which):
    out = np.zeros(512, np.uint8)
    n = lib().btbo_lut(which, _ptr(out), 512)
    return out[:n].copy()


def le_index(freq):
    return lib().btbo_le_index(freq)


def sniffdem(symbols):
    """Loop sniff_ac over a symbol array, skip 68 after every hit (btref sniffdem)."""
    s = np.ascontiguousarray(symbols, np.uint8)
    s = np.concatenate([s, np.zeros(80, np.uint8)])
    n = len(symbols)
    pos, out = 0, []
    L = lib()
    while pos + 72 < n:
        i = L.btbo_sniff_ac(s[pos:].ctypes.data_as(C.c_void_p), n - 72 - pos)
        if i < 0:
            break
        lap = int(sum(int(s[pos + i + 38 + b]) << b for b in range(24)))
        out.append((pos + i, lap))
        pos += i + 68
    return out
