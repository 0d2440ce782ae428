"""CPU tier, container only: the C restatement against the VERBATIM reference build
(oracle/_ref/btref = /root/reference/lib/*.cc compiled unmodified over oracle/shim) on the
full bundled captures, and the verbatim build against the survey-time digests (SURVEY.md
section 4).  Skipped where /root/reference is absent (the GPU box)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import FILES, REF_SAMPLES
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not (R.available() and os.path.isdir(REF_SAMPLES)),
                                reason="needs oracle/_ref/btref and /root/reference/samples")

SURVEY_MD5 = {"headset3": "944aca67f68f0929382c102a0597d63d", "headset1": "3dbbe1051f64ce8a573639e08cce5197",
              "keyboard1": "594ecd260a8be83f65a0af473f1a4220", "headset2": "6525be72fa218ed1d0cc7f06ec5adee6"}


@pytest.mark.parametrize("name", list(FILES))
def test_reference_build_reproduces_survey_digests(name, kats):
    fs, fc = FILES[name]
    r = R.sniff(os.path.join(REF_SAMPLES, name + ".cfile"), fs, fc)
    md5 = hashlib.md5(r["stdout"].encode()).hexdigest()
    assert md5 == SURVEY_MD5[name] == kats["stdout_md5"][name]


def test_reference_hopper_digest():
    r = R.sniff(os.path.join(REF_SAMPLES, "headset1.cfile"), 8e6, 2476.5e6, hop_lap=0x24D952)
    assert hashlib.md5(r["stdout"].encode()).hexdigest() == "a5dd1f5176e5c96ef4836ff30035fea3"


def test_documented_laps_found():
    """doc/README.first:45-67 of the reference: headset* -> 24d952, keyboard1 -> 4831dd."""
    for name, lap in (("headset3", 0x24D952), ("keyboard1", 0x4831DD)):
        fs, fc = FILES[name]
        r = R.sniff(os.path.join(REF_SAMPLES, name + ".cfile"), fs, fc)
        assert any(h["kind"] == 0 and h["lap"] == lap for h in R.parse_stdout_hits(r["stdout"]))


@pytest.mark.parametrize("name", list(FILES))
@pytest.mark.parametrize("stateless", [False, True])
def test_port_equals_reference_on_full_capture(name, stateless):
    fs, fc = FILES[name]
    path = os.path.join(REF_SAMPLES, name + ".cfile")
    r = R.sniff(path, fs, fc, stateless=stateless, dump=True)
    P = O.Plan(fs, fc)
    iq = np.fromfile(path, dtype=np.complex64)
    o = P.run(iq, stateless=stateless, want_bits=True, want_energy=True)
    want = [(h["slot"], h["kind"], h["lap"], h["snr"]) for h in R.parse_stdout_hits(r["stdout"])]
    got = [(int(h["slot"]), int(h["kind"]), int(h["lap"]), "%.1f" % h["snr"]) for h in o["hits"]]
    assert got == want and len(want) > 20
    ncalls = o["nsym"].shape[0]
    en = np.full((ncalls, P.nch), np.nan)
    nz = np.full((ncalls, P.nch), np.nan)
    nwin = 0
    for typ, call, rid, pay in r["records"]:
        chi = rid // 2
        if typ == R.REC_ENERGY:
            (nz if rid % 2 else en)[call, chi] = pay[0]
        elif typ == R.REC_BITS:
            n = o["nsym"][call, chi]
            assert n == len(pay) and np.array_equal(o["bits"][call, chi, :n], pay)
            nwin += 1
    assert nwin == int((o["nsym"] > 0).sum()) and nwin > 1000
    assert np.array_equal(en, o["energy"], equal_nan=True) and np.array_equal(nz, o["noise"], equal_nan=True)


def test_threaded_stateless_run_equals_serial():
    fs, fc = FILES["keyboard1"]
    iq = np.fromfile(os.path.join(REF_SAMPLES, "keyboard1.cfile"), dtype=np.complex64)[:40 * 5000]
    P = O.Plan(fs, fc)
    a = P.run(iq, stateless=True, threads=1, want_bits=True)
    b = P.run(iq, stateless=True, threads=4, want_bits=True)
    assert np.array_equal(a["hits"], b["hits"]) and np.array_equal(a["bits"], b["bits"])
