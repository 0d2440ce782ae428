"""The libbtbb-style access-code search (BTB200_SEARCH_BR_BCH): what multi_LAP / multi_UAP get from btbb_find_ac
(lib/multi_LAP_impl.cc:93, lib/multi_UAP_impl.cc:95).

libbtbb is an external library outside the reference tree, so there is nothing of the reference's to pin this to:
PARITY UNPINNED.  What IS checked: the oracle's brute-force restatement of the published algorithm (oracle/
btb_oracle.c: btbo_bch_lag -- flip up to max_ac_errors bits, regenerate the parity with the reference's own acgen
encoder) against the code's defining properties, and the product's table-driven test (csrc/rx_math.cuh:
br_lag_test_bch with the syndrome tables of plan.cpp) against that oracle lag for lag -- on the CPU through tests/emul
and on the GPU through the search kernel's known-answer entry."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, have_gpu
from oracle import oracle as O

EMUL = os.path.join(ROOT, "tests", "emul", "libbtb_emul.so")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(EMUL):
        import __graft_entry__ as ge
        ge.build()
    L = C.CDLL(EMUL)
    L.emul_bch_scan.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


@pytest.fixture(scope="module")
def channel37():
    z = np.load(os.path.join(GOLDEN, "channel37_bits.npz"))
    return np.unpackbits(z["packed"])[:int(z["n"])].astype(np.uint8)


def scan(emul, sym, max_err, lap=O.LAP_ANY, cap=4096):
    s = np.ascontiguousarray(sym, np.uint8)
    lags, laps, errs = np.zeros(cap, np.int32), np.zeros(cap, np.uint32), np.zeros(cap, np.int32)
    n = emul.emul_bch_scan(s.ctypes.data, len(s), max_err, lap, lags.ctypes.data, laps.ctypes.data, errs.ctypes.data, cap)
    assert 0 <= n <= cap
    return [(int(a), int(b), int(c)) for a, b, c in zip(lags[:n], laps[:n], errs[:n])]


def planted_stream(rng, n, plants):
    """noise with access codes planted: plants = [(lag, lap, flipped sync-word bits, flipped barker bits)]"""
    x = rng.integers(0, 2, n).astype(np.uint8)
    for lag, lap, flips, bflips in plants:
        x[lag:lag + 72] = O.acgen_bits(lap)
        for f in list(flips) + list(bflips):
            x[lag + 4 + f] ^= 1
    return x


def test_oracle_restatement_has_the_codes_properties():
    """Any sync word of acgen with up to t flipped bits among sync-word bits 0..56 (and up to 3 more in the 7 MSB/Barker
    symbols) is accepted with its LAP and the flip count; t + 1 flips are rejected (minimum distance 14); with a given
    LAP the test is the plain Hamming distance."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        lap = int(rng.integers(0, 1 << 24))
        ac = O.acgen_bits(lap)
        for t in (0, 1, 2):
            s = np.concatenate([ac, rng.integers(0, 2, 8).astype(np.uint8)])
            flips = rng.choice(57, size=t, replace=False)      # bits 0..56; bit 57 (LAP MSB) belongs to the Barker group
            for f in flips:
                s[4 + f] ^= 1
            for f in rng.choice(np.arange(57, 64), size=int(rng.integers(0, 4)), replace=False):
                s[4 + f] ^= 1                                  # up to 3 of the 7 MSB + Barker symbols: corrected, not counted
            ok, l, e = O.bch_lag(s, t)
            assert ok and l == lap and e == t, (hex(lap), t, ok, hex(l), e)
            if t:
                assert not O.bch_lag(s, t - 1)[0]
            ok2, _, d = O.bch_lag(np.concatenate([ac, s[72:]]), 2, lap)
            assert ok2 and d == 0
    # a given LAP: distance over all 64 symbols, Barker included
    ac = O.acgen_bits(0x24D952)
    s = ac.copy(); s[4 + 3] ^= 1; s[4 + 60] ^= 1
    assert O.bch_lag(s, 2, 0x24D952) == (True, 0x24D952, 2) and not O.bch_lag(s, 1, 0x24D952)[0]
    assert not O.bch_lag(ac, 2, 0x24D953)[0]


@pytest.mark.parametrize("max_err", [0, 1, 2])
def test_product_test_equals_oracle_lag_for_lag(emul, channel37, max_err):
    """csrc/rx_math.cuh's table-driven test on every lag of (a) the reference's demodulated capture channel37.dem and
    (b) noise with planted, damaged access codes == the oracle's brute force: same accepted lags, LAPs, error counts."""
    rng = np.random.default_rng(100 + max_err)
    plants = [(500, 0x9E8B33, (), ()), (1400, 0x24D952, (5,), (59,)), (2600, 0x123456, (7, 33), ()),
              (3900, 0xFFFFFF, (0, 56), (61, 63)), (5200, 0x000000, (12, 40, 41), ())]
    streams = {"channel37": channel37[:600000], "planted": planted_stream(rng, 7000, plants)}
    for name, x in streams.items():
        got = scan(emul, x, max_err)
        acc = {g[0] for g in got}
        for lag, lap, ne in got:
            assert O.bch_lag(x[lag:lag + 68], max_err) == (True, lap, ne), (name, lag)
        check = rng.integers(0, len(x) - 68, 4000) if name == "channel37" else np.arange(len(x) - 68)
        for lag in check:
            assert O.bch_lag(x[lag:lag + 68], max_err)[0] == (int(lag) in acc), (name, int(lag))
    planted = {g[0]: g for g in scan(emul, streams["planted"], max_err)}
    for lag, lap, flips, _ in plants:
        if len(flips) <= max_err:
            assert planted[lag] == (lag, lap, len(flips))
        else:
            assert lag not in planted
    if max_err == 1:
        # the capture's piconet: every packet sniff_ac finds with a clean sync word is found here too
        laps = [g[1] for g in scan(emul, channel37, 1)]
        assert laps.count(0x24D952) >= 25
    # a given LAP (multi_UAP's call): Hamming distance to that LAP's sync word
    got = scan(emul, streams["planted"], 2, lap=0x24D952)
    assert got == [(1400, 0x24D952, 2)] and scan(emul, streams["planted"], 1, lap=0x24D952) == []


def replay(accepted, n, stride):
    """the search kernel's window loop on a set of accepted lags: windows every `stride` symbols, first hit, skip 68"""
    out = set()
    acc = sorted(accepted)
    for w0 in range(0, n, stride):
        length = min(stride + 72, n - w0)
        limit = min(length - 68, 625)
        start = 0
        while limit - start >= 0:
            nxt = [a for a in acc if w0 + start <= a[0] < w0 + limit]
            if not nxt:
                break
            out.add(nxt[0])
            start = nxt[0][0] - w0 + 68
    return sorted(out)


@pytest.mark.gpu
@pytest.mark.parametrize("max_err", [1, 2])
def test_search_kernel_bch_equals_oracle(emul, channel37, max_err):
    """GPU: the search kernel with BTB200_SEARCH_BR_BCH on channel37.dem (3 997 342 symbols) and on planted noise ==
    the oracle-checked CPU scan replayed through the kernel's window loop: offsets, LAPs, corrected-bit counts."""
    import gr_bluetooth_b200 as g
    assert have_gpu()
    rng = np.random.default_rng(5)
    plants = [(700 + 900 * i, int(rng.integers(0, 1 << 24)), tuple(rng.choice(58, size=i % 3, replace=False)), ()) for i in range(40)]
    planted = planted_stream(rng, 40000, plants)
    for lap_mode in ("any", "lap"):
        kw = dict(search=g.SEARCH_BR | g.SEARCH_BR_BCH,
                  bch=g.bch_any(max_err) if lap_mode == "any" else g.bch_lap(0x24D952, max_err))
        blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_STATELESS, max_slots=64, **kw)
        for name, x in (("channel37", channel37), ("planted", planted)):
            acc = scan(emul, x, max_err, lap=O.LAP_ANY if lap_mode == "any" else 0x24D952, cap=1 << 16)
            for stride in (625, 557):
                got = sorted(set(blk.search_bits(x, stride=stride, with_errors=True)))
                assert got == replay(acc, len(x), stride), (name, lap_mode, stride, len(got))
        blk.close()
    # what the blocks ask for: multi_LAP = LAP_ANY with one correctable bit, multi_UAP = the piconet's LAP within 2
    blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_STATELESS, max_slots=64, search=g.SEARCH_BR | g.SEARCH_BR_BCH,
                          bch=g.bch_any(1))
    laps = [l for _, l in blk.search_bits(channel37)]
    assert laps.count(0x24D952) >= 25
    blk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,lap", [("headset1", 0x24D952), ("keyboard1", 0x4831DD)])
def test_cpp_multi_lap_block_uses_the_libbtbb_style_search(name, lap, tmp_path):
    """gr::bluetooth::multi_LAP::make() through btrx_b200 -L (lib/multi_LAP_impl.cc:65-114) asks for
    btbb_find_ac(LAP_ANY, max_ac_errs = 1) semantics: its "GOT PACKET" lines equal the oracle's prediction -- the
    reference's front end in the block's geometry (history + 68 symbols, chained state) followed by the oracle's
    brute-force btbb_find_ac over min(num_symbols - 68, 625) lags of every channel-window -- channel, LAP, corrected bits
    and slot; the capture's documented LAP (doc/README.first:45-67) is among them.  BTB200_AC_SEARCH=sniff_ac gives
    sniff_ac's (larger) set."""
    import re
    import subprocess
    from conftest import load_excerpt
    exe = os.path.join(ROOT, "gr-bluetooth_b200", "host", "btrx_b200")
    if not os.path.exists(exe):
        pytest.skip("btrx_b200 not built")
    ex = load_excerpt(name, "chained")
    path = tmp_path / "x.cfile"
    ex["iq"].tofile(path)

    def run(env):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([exe, "-f", repr(ex["fc"]), "-r", repr(ex["fs"]), "-i", str(path), "-L"], capture_output=True,
                             text=True, timeout=600, env=e)
        assert out.returncode == 0, out.stderr[-1500:]
        return [(int(c), int(l, 16), int(e_), int(s_)) for c, l, e_, s_ in
                re.findall(r"GOT PACKET: ch=(\d+), LAP=([0-9a-f]{6}), err=(\d+) at time slot (\d+)", out.stdout)]

    P = O.Plan(ex["fs"], ex["fc"], extra_symbols=68)
    o = P.run(ex["iq"], stateless=False, want_bits=True)
    ncall = o["nsym"].shape[0]
    want_bch, want_sniff = [], []
    for call in range(ncall):
        for chi in range(P.nch):
            ns = int(o["nsym"][call, chi])
            if ns < 68:
                continue
            bits = np.concatenate([o["bits"][call, chi, :ns], np.zeros(80, np.uint8)])
            lim = min(ns - 68, 625)
            off, l, ne = O.find_ac_bch(bits, lim, max_ac_errors=1)
            if off >= 0:
                want_bch.append((P.ch_lo + chi, l, ne, call))
            if O.sniff_ac(bits[:ns], lim) >= 0:
                want_sniff.append((P.ch_lo + chi, call))
    last = ncall - 2
    got = [h for h in run({}) if h[3] <= last]
    assert got == [h for h in want_bch if h[3] <= last] and len(got) >= 2, (got, want_bch)
    assert any(h[1] == lap for h in got) and all(h[2] <= 1 for h in got)
    sniff = [(c, s_) for c, _, _, s_ in run({"BTB200_AC_SEARCH": "sniff_ac"}) if s_ <= last]
    assert sniff == [h for h in want_sniff if h[1] <= last] and len(sniff) >= len(got)


@pytest.mark.parametrize("name,lap", [("headset1", 0x24D952), ("headset2", 0x24D952), ("headset3", 0x24D952), ("keyboard1", 0x4831DD)])
def test_multi_lap_semantics_on_the_bundled_captures(emul, name, lap):
    """multi_LAP as the reference defines it (lib/multi_LAP_impl.cc:65-114: block geometry history + 68 symbols, chained
    state, per channel-window ONE btbb_find_ac(LAP_ANY, 1) over min(num_symbols - 68, 625) lags) evaluated on the oracle's
    bit streams of the whole bundled capture with the product's table-driven test: the LAPs it prints are the documented ones
    (doc/README.first:45-67), the capture's own most often, each accepted lag is confirmed by the oracle's brute force, and the packets
    are a subset of what sniff_ac's rule reports on the same windows (clean sync words pass both)."""
    from conftest import FILES, full_capture
    iq = full_capture(name)
    if iq is None:
        pytest.skip("full capture not reachable")
    fs, fc = FILES[name]
    P = O.Plan(fs, fc, extra_symbols=68)
    o = P.run(iq, stateless=False, want_bits=True)
    found, sniffed = [], 0
    for call in range(o["nsym"].shape[0]):
        for chi in range(P.nch):
            ns = int(o["nsym"][call, chi])
            if ns < 68 + 4:
                continue
            lim = min(ns - 68, 625)
            bits = o["bits"][call, chi, :ns]
            acc = [a for a in scan(emul, bits[:lim + 67], 1) if a[0] < lim]
            s = O.sniff_ac(bits, lim)
            sniffed += s >= 0
            if acc:
                lag, l, ne = acc[0]
                assert O.bch_lag(bits[lag:lag + 68], 1) == (True, l, ne)
                found.append((call, chi, lag, l, ne))
                assert s >= 0 and s <= lag                     # sniff_ac reports this packet or an earlier code of the window
    from collections import Counter
    laps = Counter(f[3] for f in found)
    # (the keyboard capture also holds a few packets of the headset's piconet)
    assert len(found) >= 3 and laps.most_common(1)[0][0] == lap and set(laps) <= {0x24D952, 0x4831DD}, laps
    assert sniffed >= len(found)
    print("%s: %d packets with libbtbb semantics (%d with one corrected bit), %d with sniff_ac's" %
          (name, len(found), sum(f[4] for f in found), sniffed))
