"""gr_bluetooth_b200 -- B200-native multi-channel Bluetooth receive path.

Host-side mirror of the reference's Python surface (swig/gr_bluetooth.i:35-45
exposes gr_bluetooth.multi_sniffer / multi_LAP / multi_hopper built by the
make() factories of include/gr_bluetooth/multi_*.h) on top of the C ABI in
include/btb200.h.  All arithmetic on the sample path runs in the sm_100a kernels
of libbtb200.so; there is NO CPU fallback -- constructing a block without a GPU
raises Btb200Error, and a missing library raises ImportError-like OSError.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbtb200.so")

ABI_VERSION = 1
MM_CHAINED, MM_STATELESS = 0, 1
SEARCH_BR, SEARCH_LE, SEARCH_BR_BCH = 1, 2, 4


def bch_any(max_err=1):
    """btb200_config.bch for BTB200_SEARCH_BR_BCH with LAP_ANY (include/btb200.h: BTB200_BCH_ANY)."""
    return (max_err & 7) << 28


def bch_lap(lap, max_err=2):
    """btb200_config.bch for BTB200_SEARCH_BR_BCH with a given LAP (BTB200_BCH_LAP)."""
    return (lap & 0xFFFFFF) | (1 << 24) | ((max_err & 7) << 28)
SQUELCH_DEFAULT, SQUELCH_EAGER, SQUELCH_LAZY = 0, 1, 2
SNR_EXACT, SNR_FAST_GUARDED = 0, 1
TAIL_LAZY, TAIL_FULL = 0, 1
DDC_EXACT, DDC_POLYPHASE = 0, 1

STAGE = dict(noise_fast=10, energy=1, noise=2, snr=3, pass_=4, nsym=5, bits=6, ddc=7, demod=8, soft=9,
             chan_taps=20, noise_taps=21, mmse_table=22, atan_table=23, ac_lut=24)


class Btb200Error(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().btb200_strerror(code).decode()
        super().__init__("btb200 error %d: %s%s" % (code, msg, (" -- " + detail) if detail else ""))


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("sample_rate", C.c_double), ("center_freq", C.c_double),
                ("squelch_threshold", C.c_double), ("extra_history_symbols", C.c_uint32),
                ("mm_mode", C.c_int32), ("search", C.c_int32), ("device", C.c_int32),
                ("max_slots_per_call", C.c_uint32), ("keep_stages", C.c_uint32),
                ("squelch_mode", C.c_uint32), ("snr_mode", C.c_uint32), ("tail_mode", C.c_uint32),
                ("ddc_mode", C.c_uint32), ("bch", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("samples_per_slot", C.c_int32), ("history", C.c_int32), ("decimation", C.c_int32),
                ("chan_taps", C.c_int32), ("noise_taps", C.c_int32),
                ("first_channel_sample", C.c_int32), ("first_noise_sample", C.c_int32),
                ("channel_low", C.c_int32), ("channel_high", C.c_int32), ("n_channels", C.c_int32),
                ("ddc_out_per_window", C.c_int32), ("noise_out_per_window", C.c_int32),
                ("demod_gain", C.c_float), ("omega_mid", C.c_float),
                ("max_slots_per_call", C.c_uint32), ("sm_count", C.c_int32)]


HIT_DTYPE = np.dtype([("slot", "<u4"), ("channel", "<u2"), ("kind", "<u2"), ("offset", "<i4"),
                      ("n_symbols", "<i4"), ("lap", "<u4"), ("flags", "<u4"), ("snr", "<f8"),
                      ("sym_offset", "<u8"), ("sym_count", "<u4"), ("ac_errors", "<u4")], align=True)


class ChanResult(C.Structure):
    _fields_ = [("channel", C.c_int32), ("processed", C.c_int32), ("pass_", C.c_int32), ("n_symbols", C.c_int32),
                ("ac_index", C.c_int32), ("lap", C.c_uint32), ("snr", C.c_double), ("sym_offset", C.c_uint64),
                ("sym_count", C.c_uint32), ("reserved", C.c_uint32)]


class Hits(C.Structure):
    _fields_ = [("hits", C.c_void_p), ("cap", C.c_uint32), ("count", C.c_uint32), ("overflow", C.c_uint32),
                ("symbols", C.c_void_p), ("symbols_cap", C.c_uint64), ("symbols_used", C.c_uint64)]


EXPORTS = ["btb200_process_channels", "btb200_create", "btb200_destroy", "btb200_get_info", "btb200_process", "btb200_process_device",
           "btb200_submit", "btb200_submit_i16", "btb200_process_i16", "btb200_collect_begin", "btb200_collect", "btb200_host_alloc", "btb200_host_free",
           "btb200_get_mm_state", "btb200_set_mm_state", "btb200_reset", "btb200_get_stage",
           "btb200_search_bits", "btb200_timer_start", "btb200_timer_stop", "btb200_set_window_mask", "btb200_hop_candidates", "btb200_hop_select",
           "btb200_last_timing", "btb200_launch_count", "btb200_strerror", "btb200_last_error",
           "btb200_version"]

_lib = None


def lib():
    """Load libbtb200.so (built in-tree by __graft_entry__.build() / make)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("libbtb200.so not built (run `make -C gr-bluetooth_b200` or __graft_entry__.build()); "
                          "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.btb200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.btb200_destroy.argtypes = [C.c_void_p]
        L.btb200_destroy.restype = None
        L.btb200_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.btb200_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.POINTER(Hits)]
        L.btb200_process_device.argtypes = L.btb200_process.argtypes
        L.btb200_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_uint32]
        L.btb200_submit_i16.argtypes = L.btb200_submit.argtypes
        L.btb200_process_i16.argtypes = L.btb200_process.argtypes
        L.btb200_collect_begin.argtypes = [C.c_void_p]
        L.btb200_collect.argtypes = [C.c_void_p, C.POINTER(Hits)]
        L.btb200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.btb200_host_free.argtypes = [C.c_void_p]
        L.btb200_host_free.restype = None
        L.btb200_get_mm_state.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btb200_set_mm_state.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btb200_reset.argtypes = [C.c_void_p]
        L.btb200_get_stage.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        L.btb200_get_stage.restype = C.c_int64
        L.btb200_search_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(Hits)]
        L.btb200_set_window_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.btb200_hop_candidates.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_uint32)]
        L.btb200_hop_select.argtypes = [C.c_uint32, C.c_int, C.c_uint32]
        L.btb200_timer_start.argtypes = [C.c_void_p]
        L.btb200_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btb200_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.btb200_launch_count.argtypes = [C.c_void_p]
        L.btb200_launch_count.restype = C.c_uint64
        L.btb200_strerror.argtypes = [C.c_int]
        L.btb200_strerror.restype = C.c_char_p
        L.btb200_last_error.argtypes = [C.c_void_p]
        L.btb200_last_error.restype = C.c_char_p
        L.btb200_version.restype = C.c_char_p
        L.btb200_set_impl.argtypes = [C.c_void_p, C.c_int]
        L.btb200_process_channels.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int32, C.c_int32,
                                              C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


class PinnedBuffer:
    """Page-locked host memory (cudaMallocHost) viewed as a complex64 array, or (i16=True) as an int16 array of
    interleaved (re, im) pairs."""

    def __init__(self, n_samples, i16=False):
        self.ptr = C.c_void_p()
        rc = lib().btb200_host_alloc(C.byref(self.ptr), int(n_samples) * (4 if i16 else 8))
        if rc:
            raise Btb200Error(rc)
        if i16:
            buf = (C.c_int16 * (2 * int(n_samples))).from_address(self.ptr.value)
            self.array = np.frombuffer(buf, dtype=np.int16)
        else:
            buf = (C.c_float * (2 * int(n_samples))).from_address(self.ptr.value)
            self.array = np.frombuffer(buf, dtype=np.complex64)
        self._free = lib().btb200_host_free          # kept: module globals may be gone at interpreter shutdown

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class multi_block:
    """Common part of the blocks: the reference's gr::bluetooth::multi_block
    (include/gr_bluetooth/multi_block.h:40) seen through the C ABI."""

    EXTRA_SYMBOLS = 3125

    def __init__(self, sample_rate, center_freq, squelch_threshold, *, mm_mode=MM_CHAINED,
                 search=SEARCH_BR | SEARCH_LE, device=0, max_slots=64, keep_stages=False,
                 squelch=SQUELCH_DEFAULT, snr_mode=SNR_EXACT, tail=TAIL_LAZY, ddc=DDC_EXACT, bch=0):
        self._L = lib()
        cfg = Config(abi_version=ABI_VERSION, sample_rate=sample_rate, center_freq=center_freq,
                     squelch_threshold=squelch_threshold, extra_history_symbols=self.EXTRA_SYMBOLS,
                     mm_mode=mm_mode, search=search, device=device, max_slots_per_call=max_slots,
                     keep_stages=int(keep_stages), squelch_mode=squelch, snr_mode=snr_mode, tail_mode=tail,
                     ddc_mode=ddc, bch=bch)
        self._ctx = C.c_void_p()
        rc = self._L.btb200_create(C.byref(cfg), C.byref(self._ctx))
        if rc:
            raise Btb200Error(rc, self._L.btb200_last_error(None).decode())
        self.info = Info()
        self._L.btb200_get_info(self._ctx, C.byref(self.info))
        self.sample_rate, self.center_freq, self.squelch_threshold = sample_rate, center_freq, squelch_threshold
        self.mm_mode = mm_mode
        self.max_slots = self.info.max_slots_per_call
        self._cumulative_slots = 0
        self._hit_cap = 1 << 16
        self._hits = np.zeros(self._hit_cap, HIT_DTYPE)
        self._sym_cap = 64 << 20
        self._syms = np.zeros(self._sym_cap, np.uint8)
        # the constructor line the reference prints (lib/multi_block.cc:116)
        self.constructor_banner = None

    # -- gr::sync_block surface ------------------------------------------------
    def history(self):
        return self.info.history

    @property
    def samples_per_slot(self):
        return self.info.samples_per_slot

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.btb200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise Btb200Error(rc, self._L.btb200_last_error(self._ctx).decode())

    def _hits_struct(self, want_symbols):
        if want_symbols == "borrow":         # zero-copy: symbols stay in the context's pinned arena
            return Hits(hits=self._hits.ctypes.data, cap=self._hit_cap, count=0, overflow=0, symbols=None,
                        symbols_cap=0xFFFFFFFFFFFFFFFF, symbols_used=0)
        return Hits(hits=self._hits.ctypes.data, cap=self._hit_cap, count=0, overflow=0,
                    symbols=self._syms.ctypes.data if want_symbols else None,
                    symbols_cap=self._sym_cap, symbols_used=0)

    def _take(self, h, want_symbols):
        if want_symbols == "borrow":
            # views, valid until the next submit on this block: the hit records in this object's buffer, the symbols
            # in the library's pinned arena
            n = int(h.symbols_used)
            syms = np.frombuffer((C.c_uint8 * n).from_address(h.symbols), dtype=np.uint8) if n else np.zeros(0, np.uint8)
            return self._hits[:h.count], syms, int(h.overflow)
        hits = self._hits[:h.count].copy()
        syms = self._syms[:h.symbols_used].copy() if want_symbols else None
        return hits, syms, int(h.overflow)

    def process(self, iq, first_slot, n_slots, want_symbols=False):
        """n_slots work() calls; iq[0] is the first sample of the first window."""
        x = np.ascontiguousarray(iq, dtype=np.complex64)
        h = self._hits_struct(want_symbols)
        self._check(self._L.btb200_process(self._ctx, x.ctypes.data, len(x), first_slot, n_slots, C.byref(h)))
        return self._take(h, want_symbols)

    def process_i16(self, iq16, first_slot, n_slots, want_symbols=False):
        """Same with int16 input: interleaved (re, im) pairs, 2 * n_samples values."""
        x = np.ascontiguousarray(iq16, dtype=np.int16)
        h = self._hits_struct(want_symbols)
        self._check(self._L.btb200_process_i16(self._ctx, x.ctypes.data, len(x) // 2, first_slot, n_slots, C.byref(h)))
        return self._take(h, want_symbols)

    def submit_i16(self, ptr, on_device, n_samples, first_slot, n_slots):
        self._check(self._L.btb200_submit_i16(self._ctx, C.c_void_p(ptr), int(on_device), n_samples, first_slot, n_slots))

    def process_device(self, dptr, n_samples, first_slot, n_slots, want_symbols=False):
        h = self._hits_struct(want_symbols)
        self._check(self._L.btb200_process_device(self._ctx, C.c_void_p(dptr), n_samples, first_slot, n_slots,
                                                  C.byref(h)))
        return self._take(h, want_symbols)

    def process_channels(self, window, slot, first_channel, n_channels, stop_lap=0xFFFFFFFF):
        """One multi_hopper work() call (chained mode): channels first_channel.. in order, ends after the first
        channel whose packet carries stop_lap and a header.  -> (list of ChanResult, symbols uint8 array)"""
        x = np.ascontiguousarray(window, dtype=np.complex64)
        res = (ChanResult * max(n_channels, 1))()
        syms = np.zeros(max(n_channels, 1) * 3125, np.uint8)
        self._check(self._L.btb200_process_channels(self._ctx, x.ctypes.data, len(x), slot, first_channel, n_channels,
                                                    stop_lap, C.byref(res), syms.ctypes.data, len(syms)))
        return list(res)[:n_channels], syms

    def submit(self, ptr, on_device, n_samples, first_slot, n_slots):
        self._check(self._L.btb200_submit(self._ctx, C.c_void_p(ptr), int(on_device), n_samples, first_slot, n_slots))

    def collect_begin(self):
        """Optional first half of collect(): enqueue the deferred work of the pending batch without waiting."""
        self._check(self._L.btb200_collect_begin(self._ctx))

    def collect(self, want_symbols=False):
        h = self._hits_struct(want_symbols)
        self._check(self._L.btb200_collect(self._ctx, C.byref(h)))
        return self._take(h, want_symbols)

    def work(self, noutput_items, input_items, output_items=None):
        """gr::sync_block::work(): one window of history() samples in
        input_items[0], consumes one slot (lib/multi_sniffer_impl.cc:82-166).
        Returns (items consumed, hits)."""
        x = np.ascontiguousarray(input_items[0][:self.history()], dtype=np.complex64)
        hits, syms, _ = self.process(x, self._cumulative_slots, 1, want_symbols=True)
        self._cumulative_slots += 1
        self._last = (hits, syms)
        return self.samples_per_slot, hits

    def run_stream(self, samples, batch=None, want_symbols=False, first_call=0, num_calls=None):
        """Emulates the GNU Radio scheduler over a whole capture:
        zero history in front, call k sees samples [k*S-(H-1), k*S]."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        S, H = self.samples_per_slot, self.history()
        ncalls = (len(x) + S - 1) // S
        k1 = ncalls if num_calls is None else min(ncalls, first_call + num_calls)
        batch = batch or self.max_slots
        all_hits, all_syms, sym_base = [], [], 0
        k = first_call
        while k < k1:
            n = min(batch, k1 - k)
            w0 = k * S - (H - 1)
            need = (n - 1) * S + H
            if w0 >= 0:
                seg = x[w0:w0 + need]
            else:
                seg = np.concatenate([np.zeros(-w0, np.complex64), x[:w0 + need]])
            hits, syms, ovf = self.process(seg, k, n, want_symbols)
            if ovf:
                raise Btb200Error(-1, "hit overflow")
            if want_symbols and len(hits):
                hits["sym_offset"] += sym_base
                sym_base += len(syms)
                all_syms.append(syms)
            all_hits.append(hits)
            k += n
        self._cumulative_slots = k
        hits = np.concatenate(all_hits) if all_hits else np.zeros(0, HIT_DTYPE)
        if want_symbols:
            return hits, (np.concatenate(all_syms) if all_syms else np.zeros(0, np.uint8))
        return hits

    # -- state / debug -----------------------------------------------------------
    def get_mm_state(self):
        mm = (C.c_float * 3)()
        self._check(self._L.btb200_get_mm_state(self._ctx, mm))
        return np.array(mm[:], np.float32)

    def set_mm_state(self, mm):
        arr = (C.c_float * 3)(*[float(v) for v in mm])
        self._check(self._L.btb200_set_mm_state(self._ctx, arr))

    def reset(self):
        self._check(self._L.btb200_reset(self._ctx))
        self._cumulative_slots = 0

    def set_impl(self, impl):
        self._check(self._L.btb200_set_impl(self._ctx, impl))

    def stage(self, name, slot_in_batch=0, chan_index=0):
        I = self.info
        sizes = dict(noise_fast=(8, np.float64), energy=(8, np.float64), noise=(8, np.float64), snr=(8, np.float64), pass_=(4, np.int32),
                     nsym=(4, np.int32), bits=(I.ddc_out_per_window, np.uint8),
                     ddc=(I.ddc_out_per_window * 8, np.complex64), demod=(I.ddc_out_per_window * 4, np.float32),
                     soft=(I.ddc_out_per_window * 4, np.float32), chan_taps=(I.chan_taps * 8, np.complex64),
                     noise_taps=(I.noise_taps * 8, np.complex64), mmse_table=(129 * 8 * 4, np.float32),
                     atan_table=(257 * 4, np.float32), ac_lut=(769 * 8, np.uint64))
        cap, dt = sizes[name]
        buf = np.zeros(cap, np.uint8)
        n = self._L.btb200_get_stage(self._ctx, STAGE[name], slot_in_batch, chan_index, buf.ctypes.data, cap)
        if n < 0:
            raise Btb200Error(int(n), self._L.btb200_last_error(self._ctx).decode())
        return buf[:n].view(dt).copy()

    def search_bits(self, symbols, stride=625, with_errors=False):
        """Known-answer entry: the access-code search kernel on a caller-supplied symbol stream (one symbol per byte),
        cut into windows every `stride` symbols.  -> list of (absolute symbol offset, LAP[, ac_errors])."""
        sym = np.ascontiguousarray(symbols, dtype=np.uint8)
        h = self._hits_struct(False)
        self._check(self._L.btb200_search_bits(self._ctx, sym.ctypes.data, len(sym), stride, C.byref(h)))
        if h.overflow:
            raise Btb200Error(-7, "hit buffer too small")
        hits = self._hits[:h.count]
        if with_errors:
            return [(int(x["slot"]) * stride + int(x["offset"]), int(x["lap"]), int(x["ac_errors"])) for x in hits]
        return [(int(x["slot"]) * stride + int(x["offset"]), int(x["lap"])) for x in hits]

    def set_window_mask(self, mask):
        """mask[slot_in_batch, chan_index] != 0: the channel-windows the NEXT submit/process demodulates and searches."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        self._check(self._L.btb200_set_window_mask(self._ctx, m.ctypes.data, m.shape[0]))

    def timer_start(self):
        self._check(self._L.btb200_timer_start(self._ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self._L.btb200_timer_stop(self._ctx, C.byref(ms)))
        return float(ms.value)

    def last_timing(self):
        ms = (C.c_float * 8)()
        self._L.btb200_last_timing(self._ctx, ms)
        keys = ["h2d", "chan_fir", "noise_fir", "energy", "demod_mm", "search", "d2h", "total"]
        return dict(zip(keys, ms[:]))

    def launch_count(self):
        return int(self._L.btb200_launch_count(self._ctx))


class multi_sniffer(multi_block):
    """gr_bluetooth.multi_sniffer(sample_rate, center_freq, squelch_threshold, tun)
    -- include/gr_bluetooth/multi_sniffer.h:54, lib/multi_sniffer_impl.cc:42-74."""
    EXTRA_SYMBOLS = 3125

    def __init__(self, sample_rate, center_freq, squelch_threshold, tun=False, **kw):
        if tun:
            raise NotImplementedError("TUN/Wireshark output (lib/tun.cc) is outside the hot path")
        super().__init__(sample_rate, center_freq, squelch_threshold, **kw)

    @classmethod
    def make(cls, sample_rate, center_freq, squelch_threshold, tun=False, **kw):
        return cls(sample_rate, center_freq, squelch_threshold, tun, **kw)


class multi_LAP(multi_block):
    """gr_bluetooth.multi_LAP(sample_rate, center_freq, squelch_threshold)
    -- include/gr_bluetooth/multi_LAP.h:53.  Window geometry of
    lib/multi_LAP_impl.cc:54 (history + 68 symbols); the access-code search uses
    sniff_ac semantics (libbtbb's btbb_find_ac is external: parity unpinned)."""
    EXTRA_SYMBOLS = 68

    def __init__(self, sample_rate, center_freq, squelch_threshold, **kw):
        kw.setdefault("search", SEARCH_BR)
        super().__init__(sample_rate, center_freq, squelch_threshold, **kw)

    @classmethod
    def make(cls, sample_rate, center_freq, squelch_threshold, **kw):
        return cls(sample_rate, center_freq, squelch_threshold, **kw)


def format_hit_line(hit):
    """The prefix ac()/aa() print for a hit (lib/multi_sniffer_impl.cc:177-178, 213)."""
    if hit["kind"] == 0:
        return "time %6d, snr=%.1f, channel %2d, LAP %06x " % (hit["slot"] & 0x7ffffff, hit["snr"],
                                                               hit["channel"], hit["lap"])
    return "time %6d, snr=%.1f, " % (hit["slot"] & 0x7ffffff, hit["snr"])


def hop_candidates(address28, clock6, first_channel, afh=False, aliased=False, device=0):
    """GPU candidate search of the hop reversal -> ascending uint32 array of CLK1-27 candidates."""
    out = np.zeros(1 << 16, np.uint32)
    n = C.c_uint32()
    rc = lib().btb200_hop_candidates(device, address28, int(afh), int(aliased), clock6, first_channel, out.ctypes.data, len(out), C.byref(n))
    if rc:
        raise Btb200Error(rc)
    if n.value > len(out):
        out = np.zeros(n.value, np.uint32)
        lib().btb200_hop_candidates(device, address28, int(afh), int(aliased), clock6, first_channel, out.ctypes.data, len(out), C.byref(n))
    return out[:n.value].copy()


def version():
    return lib().btb200_version().decode()
