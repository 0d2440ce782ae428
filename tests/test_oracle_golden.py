"""CPU tier: the C restatement (oracle/btb_oracle.c) against the committed outputs of the
reference's own code (tests/golden, produced by tests/golden/make_golden.py from
oracle/_ref/btref = lib/*.cc compiled verbatim)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, FILES, load_excerpt, golden_bits
from oracle import oracle as O
from oracle import ref as R


def test_acgen_known_answers(kats):
    # acgen(0x9E8B33) = 5475c58cc73345e72a contains the spec's GIAC sync word
    assert kats["acgen"]["9e8b33"] == "5475c58cc73345e72a"
    for lap, want in kats["acgen"].items():
        assert O.acgen_bytes(int(lap, 16)).hex() == want


def test_acgen_is_affine_in_lap():
    rng = np.random.default_rng(7)
    def sw(l):
        return int.from_bytes(np.packbits(O.acgen_bits(int(l))[4:68]).tobytes(), "big")
    for a, b, c in rng.integers(0, 1 << 24, (300, 3)):
        assert sw(a ^ b ^ c) == sw(a) ^ sw(b) ^ sw(c)


def test_check_ac_threshold():
    ac = O.acgen_bits(0x24D952)[:72].copy()
    assert O.check_ac(ac, 0x24D952)
    rng = np.random.default_rng(3)
    for nerr in range(0, 10):
        s = ac.copy()
        # errors inside the first 68 symbols but outside the LAP field, so the LAP still reads back
        pos = rng.choice(np.r_[0:38, 62:68], nerr, replace=False)
        s[pos] ^= 1
        assert O.check_ac(s, 0x24D952) == (nerr < 7)


def test_luts_match_reference_tables():
    names = {"classic.PREAMBLE_DISTANCE": 0, "classic.BARKER_DISTANCE": 1, "le.PREAMBLE_DISTANCE": 2,
             "le.ACCESS_ADDRESS_DISTANCE_0": 3, "le.ACCESS_ADDRESS_DISTANCE_1": 4,
             "le.ACCESS_ADDRESS_DISTANCE_2": 5, "le.ACCESS_ADDRESS_DISTANCE_3": 6,
             "le.ACCESS_HEADER_DISTANCE_LSB": 7, "le.ACCESS_HEADER_DISTANCE_MSB": 8,
             "le.DATA_HEADER_DISTANCE_LSB": 9, "le.DATA_HEADER_DISTANCE_MSB": 10}
    seen = 0
    for line in open(os.path.join(GOLDEN, "ref_tables.txt")):
        p = line.split()
        if p[0] in names:
            assert np.array_equal(np.array(p[2:], int), O.lut(names[p[0]])), p[0]
            seen += 1
        elif p[0] == "le.chan2index":
            ch, idx = int(p[1]), int(p[2])
            assert O.le_index(2402e6 + 2e6 * ch) == idx
    assert seen == len(names)
    assert O.le_index(2403e6) == -1          # odd-MHz channels carry no LE


def test_channel37_dem_hit_list(kats):
    z = np.load(os.path.join(GOLDEN, "channel37_bits.npz"))
    sym = np.unpackbits(z["packed"])[:int(z["n"])]
    hits = O.sniffdem(sym)
    assert [[o, l] for o, l in hits] == kats["channel37_hits"]
    assert len(hits) == 33 and hits[0] == (66136, 0xF2F57B) and hits[1] == (206587, 0x24D952)
    assert sum(1 for _, l in hits if l == 0x24D952) == 31


@pytest.mark.parametrize("name", list(FILES))
@pytest.mark.parametrize("mode", ["chained", "stateless"])
def test_excerpt_matches_reference(name, mode):
    ex = load_excerpt(name, mode)
    P = O.Plan(ex["fs"], ex["fc"])
    out = P.run(ex["iq"], num_calls=ex["nslots"], stateless=(mode == "stateless"), want_bits=True,
                want_energy=True, n_total=len(ex["iq"]) + P.S)
    assert np.array_equal(out["energy"], ex["energy"], equal_nan=True)
    assert np.array_equal(out["noise"], ex["noise"], equal_nan=True)
    assert np.array_equal(out["nsym"], ex["nsym"])
    for call in range(ex["nslots"]):
        for chi in range(P.nch):
            n = ex["nsym"][call, chi]
            assert np.array_equal(out["bits"][call, chi, :n], golden_bits(ex, call, chi))
    want = R.parse_stdout_hits(ex["stdout"])
    got = out["hits"]
    assert len(got) == len(want) and len(want) > 0
    for h, w in zip(got, want):
        assert (int(h["slot"]), int(h["kind"]), int(h["lap"]), "%.1f" % h["snr"]) == \
               (w["slot"], w["kind"], w["lap"], w["snr"])
        if w["kind"] == 0:
            assert int(h["channel"]) == w["channel"]


@pytest.mark.parametrize("name", ["headset3", "headset2"])
def test_stage_floats_match_reference(name):
    """DDC output, demod floats and soft symbols of the heavy-captured windows: bit-exact."""
    ex = load_excerpt(name, "chained")
    P = O.Plan(ex["fs"], ex["fc"])
    st = O.State(P)
    calls = sorted({int(k.split("_")[1]) for k in ex["heavy"]})
    S, H = P.S, P.H
    x = np.concatenate([np.zeros(H - 1, np.complex64), ex["iq"]])
    for k in range(calls[-1] + 1):
        hits, d = P.window(x[k * S:k * S + H], slot=k, stateless=False, state=st, debug=True)
        if k in calls:
            for chi in range(P.nch):
                assert np.array_equal(d["ddc"][chi], ex["heavy"]["ddc_%d_%d" % (k, chi)])
                if "demod_%d_%d" % (k, chi) in ex["heavy"]:
                    assert np.array_equal(d["demod"][chi], ex["heavy"]["demod_%d_%d" % (k, chi)])
                    n = d["nsym"][chi]
                    assert np.array_equal(d["soft"][chi][:n], ex["heavy"]["soft_%d_%d" % (k, chi)])


def test_geometry_table():
    """SURVEY.md section 8 derived constants."""
    rows = {(2e6, 2476e6): (74, 74, 1, 13, 401, 1250, 7901, 380, 7497, 850),
            (4e6, 2476e6): (73, 75, 2, 27, 801, 2500, 15801, 758, 7495, 850),
            (8e6, 2476.5e6): (71, 78, 4, 53, 1601, 5000, 31601, 1516, 7495, 850),
            (30e6, 2414e6): (0, 26, 15, 201, 6001, 18750, 118501, 5680, 7494, 850),
            (100e6, 2441e6): (0, 78, 50, 667, 20001, 62500, 395001, 18934, 7494, 850)}
    for (fs, fc), want in rows.items():
        P = O.Plan(fs, fc)
        assert (P.ch_lo, P.ch_hi, P.D, P.Nc, P.Nn, P.S, P.H, P.fcs, P.n_ddc, P.n_noise) == want
    P = O.Plan(100e6, 2441e6, extra_symbols=68)
    assert P.H == 89301
    P = O.Plan(2e6, 2476e6, extra_symbols=68)
    assert P.H == 1787


def test_mmse_table_check_values():
    """SURVEY.md A.5 check rows of GNU Radio's interpolator_taps.h."""
    t = O.Plan(2e6, 2476e6).mmse_table()
    row1 = [-1.54700e-04, 8.53777e-04, -2.76968e-03, 7.89295e-03, 9.98534e-01, -5.41054e-03, 1.24642e-03, -1.98993e-04]
    assert np.array_equal(t[1], np.array(row1, np.float32))
    assert np.array_equal(t[0], np.eye(8, dtype=np.float32)[4]) and np.array_equal(t[128], np.eye(8, dtype=np.float32)[3])
    assert np.array_equal(t[64][:4], np.array([-6.77751e-03, 3.94578e-02, -1.42658e-01, 6.09836e-01], np.float32))
    assert np.array_equal(t[64], t[64][::-1])


@pytest.mark.parametrize("name", list(FILES))
def test_lazy_tail_assumptions_hold_in_reference_streams(name):
    """The product's lazy tail rests on two facts about the clock-recovery loop, proven
    from the loop constants at create(): the input advance per symbol is at most 5, so (i) 704 symbols never need
    more than 704*5+16 demod outputs and (ii) a window yields at least (n_dem-8)/5 >= 1400 symbols, which keeps
    both search limits at 625.  Cross-check on the reference's own streams: the symbol counts of every window of
    the committed excerpts lie far above the bound and close to n_dem/2."""
    ex = load_excerpt(name, "stateless")
    P = O.Plan(ex["fs"], ex["fc"])
    n_dem = P.n_ddc - 1                      # demod outputs per window (multi_block.cc:158-168)
    nsym = ex["nsym"][ex["nsym"] > 0]
    assert len(nsym) > 0
    assert nsym.min() >= (n_dem - 8) // 5 >= 1400
    assert abs(float(nsym.mean()) - n_dem / 2) < 32
    assert 704 * 5 + 16 < n_dem
