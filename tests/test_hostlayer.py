"""CPU tier: the product's native host packet layer (gr-bluetooth_b200/host/lib/bt_host.cc: header
and payload decode, UAP/CLK1-6 discovery, FHS, BLE printout = the reference's ac()/aa() call chains,
SURVEY.md 8f-1/2) driven by the ORACLE's hit list must print the reference's stdout, byte for byte."""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, FILES, REF_SAMPLES, load_excerpt
from oracle import oracle as O

HDIR = os.path.join(ROOT, "tests", "hostlayer")


@pytest.fixture(scope="module")
def harness():
    exe = os.path.join(HDIR, "_build", "hostlayer")
    srcs = [os.path.join(HDIR, "hostlayer_main.cc"), os.path.join(ROOT, "gr-bluetooth_b200", "host", "lib", "bt_host.cc")]
    deps = srcs + [os.path.join(ROOT, "gr-bluetooth_b200", "host", "lib", "bt_host.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *srcs, "-o", exe])
    return exe


def host_stdout(harness, P, iq, tmp_path, num_calls=None, tun_out=None):
    o = P.run(iq, num_calls=num_calls, stateless=False, want_bits=True, n_total=None if num_calls is None else len(iq) + P.S)
    path = tmp_path / "hits.bin"
    first_call = 0
    with open(path, "wb") as f:
        for h in o["hits"]:
            call, chi = int(h["slot"]) - first_call, int(h["channel"]) - P.ch_lo
            n = int(o["nsym"][call, chi])
            # len symbols from the hit's offset; the BR search may have consumed part of the window (h['len'])
            sym = o["bits"][call, chi, int(h["offset"]):n]
            ln = min(int(h["len"]), 3125, len(sym))
            f.write(struct.pack("<Iiddi", int(h["slot"]), int(h["kind"]), 2402e6 + 1e6 * int(h["channel"]), float(h["snr"]), ln))
            f.write(sym[:ln].astype(np.uint8).tobytes())
    out = subprocess.run([harness, str(path)] + ([str(tun_out)] if tun_out else []), capture_output=True, timeout=120)
    assert out.returncode == 0
    chist = P.Nc + P.D * 8
    banner = "history set to %d samples: channel=%d, noise=%d\n" % (P.S + max(chist, P.Nn), chist, P.Nn)
    return banner + out.stdout.decode()


@pytest.mark.parametrize("name", list(FILES))
def test_excerpt_stdout_identical(harness, name, tmp_path):
    ex = load_excerpt(name, "chained")
    P = O.Plan(ex["fs"], ex["fc"])
    got = host_stdout(harness, P, ex["iq"], tmp_path, num_calls=ex["nslots"])
    assert got == ex["stdout"]


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
@pytest.mark.parametrize("name", list(FILES))
def test_full_capture_stdout_digest(harness, name, kats, tmp_path):
    fs, fc = FILES[name]
    iq = np.fromfile(os.path.join(REF_SAMPLES, name + ".cfile"), dtype=np.complex64)
    got = host_stdout(harness, O.Plan(fs, fc), iq, tmp_path)
    assert hashlib.md5(got.encode()).hexdigest() == kats["stdout_md5"][name]


HOPPER_MD5 = "a5dd1f5176e5c96ef4836ff30035fea3"     # headset1, multi_hopper, LAP 24d952


@pytest.fixture(scope="module")
def hopper_harness():
    exe = os.path.join(HDIR, "_build", "hopper")
    obj = os.path.join(HDIR, "_build", "btb_oracle.o")
    srcs = [os.path.join(HDIR, "hopper_main.cc"), os.path.join(ROOT, "gr-bluetooth_b200", "host", "lib", "bt_host.cc")]
    deps = srcs + [os.path.join(ROOT, "gr-bluetooth_b200", "host", "lib", "bt_host.h"),
                   os.path.join(ROOT, "oracle", "btb_oracle.c"), os.path.join(ROOT, "oracle", "btb_oracle.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-c",
                               os.path.join(ROOT, "oracle", "btb_oracle.c"), "-o", obj])
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *srcs, obj, "-lpthread", "-lm", "-o", exe])
    return exe


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
def test_hopper_logic_reproduces_reference_digest(hopper_harness):
    """multi_hopper (BASELINE config 4 logic): native HopperHost + hop reversal over the oracle front end
    (oracle/btb_oracle.c: btbo_window_list) on headset1 -> UAP 0xaf after 7 packets, 26555 -> 408 -> 49 ->
    10 -> 2 -> 2 CLK1-27 candidates, offset 0x00a3c6f, then hop-along decodes: the reference's stdout."""
    out = subprocess.run([hopper_harness, "8e6", "2476.5e6", "24d952", os.path.join(REF_SAMPLES, "headset1.cfile")],
                         capture_output=True, timeout=300)
    assert out.returncode == 0
    text = out.stdout.decode()
    assert "Acquired CLK1-27 offset = 0x00a3c6f" in text and "26555 initial CLK1-27 candidates" in text
    assert hashlib.md5(out.stdout).hexdigest() == HOPPER_MD5


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
def test_hopper_logic_aliased_mode_equals_reference_build(hopper_harness):
    """`btrx --aliased` (apps/btrx:37-38, 154): the hop reversal works on the 25 aliased channels -- including the
    reference's quirk that the FIRST candidate set is still drawn with the previous flag (piconet_impl.cc:118-124) --
    same stdout as the verbatim reference build on headset1 (the capture is not aliased, so the candidates run out)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/btref not built")
    path = os.path.join(REF_SAMPLES, "headset1.cfile")
    want = subprocess.run([R.BTREF, "hop", "--fs", "8e6", "--fc", "2476.5e6", "--lap", "24d952", "--in", path, "--aliased"],
                          capture_output=True, timeout=300).stdout.decode()
    out = subprocess.run([hopper_harness, "8e6", "2476.5e6", "24d952", path, "aliased"], capture_output=True, timeout=300)
    assert out.returncode == 0 and out.stdout.decode() == want
    assert "26555 initial CLK1-27 candidates" in want and "no candidates remaining" in want


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
def test_hopper_logic_keyboard1_equals_reference_build(hopper_harness):
    """BASELINE config 4 input (keyboard1, LAP 4831dd): same stdout as the verbatim reference build."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/btref not built")
    path = os.path.join(REF_SAMPLES, "keyboard1.cfile")
    want = R.sniff(path, 8e6, 2476.5e6, hop_lap=0x4831DD)["stdout"]
    out = subprocess.run([hopper_harness, "8e6", "2476.5e6", "4831dd", path], capture_output=True, timeout=300)
    assert out.returncode == 0 and out.stdout.decode() == want
    assert "UAP = 0x61" in want


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
@pytest.mark.parametrize("name", list(FILES))
def test_wireshark_frames_equal_reference_sniffer(harness, name, tmp_path):
    """the TAP frames (lib/tun.cc:91-123 around classic_packet::tun_format,
    packet_impl.cc:1175-1202) the native host layer writes for tun = true are the reference's, byte for byte.
    The verbatim reference build writes them to a file through its own write_interface()."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/btref not built")
    fs, fc = FILES[name]
    path = os.path.join(REF_SAMPLES, name + ".cfile")
    want_file, got_file = tmp_path / "ref.frames", tmp_path / "got.frames"
    want_txt = R.sniff(path, fs, fc, tun_out=str(want_file))["stdout"]
    iq = np.fromfile(path, dtype=np.complex64)
    got_txt = host_stdout(harness, O.Plan(fs, fc), iq, tmp_path, tun_out=got_file)
    assert got_txt == want_txt
    want, got = want_file.read_bytes(), got_file.read_bytes()
    assert got == want
    assert len(want) >= 14 and want[12:14] == b"\xff\xf0"          # EtherType 0xFFF0, big endian


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
def test_wireshark_frames_equal_reference_hopper(hopper_harness, tmp_path):
    """multi_hopper hop-along frames on headset1 (UAP 0xaf: the reference keeps the address in an int, so the
    two NAP bytes of the destination MAC come out as ff:ff -- reproduced)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/btref not built")
    path = os.path.join(REF_SAMPLES, "headset1.cfile")
    want_file, got_file = tmp_path / "ref.frames", tmp_path / "got.frames"
    want_txt = R.sniff(path, 8e6, 2476.5e6, hop_lap=0x24D952, tun_out=str(want_file))["stdout"]
    out = subprocess.run([hopper_harness, "8e6", "2476.5e6", "24d952", path, str(got_file)], capture_output=True, timeout=300)
    assert out.returncode == 0 and out.stdout.decode() == want_txt
    want, got = want_file.read_bytes(), got_file.read_bytes()
    assert got == want and len(want) > 14
    assert want[:6] == bytes([0xFF, 0xFF, 0xAF, 0x24, 0xD9, 0x52])


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="needs the bundled captures")
@pytest.mark.parametrize("name,lap,uap", [("headset1", "24d952", 0xAF), ("keyboard1", "4831dd", 0x61)])
def test_multi_uap_logic_finds_documented_uap(hopper_harness, name, lap, uap):
    """multi_UAP (UapHost over the oracle front end): the UAP the reference documents for the bundled captures
    (doc/README.first:45-67) and that its hopper/sniffer derive; stops as soon as it is known."""
    out = subprocess.run([hopper_harness, "8e6", "2476.5e6", lap, os.path.join(REF_SAMPLES, name + ".cfile"), "uap"],
                         capture_output=True, timeout=300)
    assert out.returncode == 0
    text = out.stdout.decode()
    assert ("UAP = 0x%x found after" % uap) in text
    assert text.rstrip().splitlines()[-1] == "multi_UAP done: UAP 0x%02x" % uap
