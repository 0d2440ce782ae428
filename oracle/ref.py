"""Runner/parser for the verbatim reference build oracle/_ref/btref.

TEST INFRASTRUCTURE.  btref is the reference's own lib/*.cc compiled unmodified
(oracle/Makefile `ref`); it exists wherever it was built (this container) and
travels to the GPU box as a prebuilt binary.
"""
import os
import re
import subprocess
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BTREF = os.path.join(HERE, "_ref", "btref")
REFERENCE = "/root/reference"

REC_DDC, REC_ENERGY, REC_BITS, REC_SOFT, REC_DEMOD, REC_MU = 1, 2, 3, 4, 5, 6


def available():
    return os.path.exists(BTREF)


def build():
    if os.path.isdir(os.path.join(REFERENCE, "lib")):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return available()


def run(cmd, *args, **kw):
    p = subprocess.run([BTREF, cmd, *[str(a) for a in args]], capture_output=True, **kw)
    if p.returncode != 0:
        raise RuntimeError("btref failed: %s" % p.stderr.decode()[-2000:])
    return p.stdout.decode(), p.stderr.decode()


def parse_dump(path):
    """-> list of (type, call, id, payload ndarray)."""
    raw = np.fromfile(path, dtype=np.uint8)
    out, pos = [], 0
    while pos < len(raw):
        typ, call, rid, n = np.frombuffer(raw, "<u4", 4, pos)
        rid = int(np.int32(rid))
        pos += 16
        if typ == REC_DDC:
            pay = np.frombuffer(raw, np.complex64, n, pos); pos += 8 * n
        elif typ == REC_ENERGY:
            pay = np.frombuffer(raw, np.float64, 1, pos); pos += 8
        elif typ == REC_BITS:
            pay = np.frombuffer(raw, np.uint8, n, pos); pos += n
        elif typ in (REC_SOFT, REC_DEMOD, REC_MU):
            pay = np.frombuffer(raw, np.float32, n, pos); pos += 4 * n
        else:
            raise ValueError("bad record type %d at %d" % (typ, pos))
        out.append((int(typ), int(call), rid, pay))
    return out


HIT_RE = re.compile(r"^time\s+(\d+), snr=(-?[\d.]+|nan|-nan|inf), (?:channel\s+(\d+), LAP ([0-9a-f]{6})|BTLE index=(\d+), AA=([0-9a-f]{8}))")


def parse_stdout_hits(text):
    """Lines printed by multi_sniffer_impl::ac()/aa() (lib/multi_sniffer_impl.cc:177-178, 213-214).
    -> list of dicts(kind, slot, snr_str, channel|index, lap|aa)."""
    hits = []
    for line in text.splitlines():
        m = HIT_RE.match(line)
        if not m:
            continue
        if m.group(3) is not None:
            hits.append(dict(kind=0, slot=int(m.group(1)), snr=m.group(2), channel=int(m.group(3)),
                             lap=int(m.group(4), 16)))
        else:
            hits.append(dict(kind=1, slot=int(m.group(1)), snr=m.group(2), index=int(m.group(5)),
                             lap=int(m.group(6), 16)))
    return hits


def sniff(path, fs, fc, snr=10.0, i16=False, stateless=False, first_call=0, num_calls=None,
          dump=False, heavy=None, hop_lap=None, tun_out=None):
    """Run the reference multi_sniffer (or multi_hopper) over a file.
    -> dict(stdout, stderr, records (if dump))."""
    args = ["--fs", fs, "--fc", fc, "--snr", snr, "--in", path, "--first-call", first_call]
    if num_calls is not None:
        args += ["--num-calls", num_calls]
    if i16:
        args.append("--i16")
    if stateless:
        args.append("--stateless")
    tmp = None
    if dump:
        tmp = tempfile.NamedTemporaryFile(suffix=".btref", delete=False)
        tmp.close()
        args += ["--dump", tmp.name]
        if heavy:
            args += ["--heavy", "%d:%d" % heavy]
    if tun_out:
        args += ["--tun-out", tun_out]        # the Wireshark-interface frames (lib/tun.cc) go to this file
    cmd = "sniff"
    if hop_lap is not None:
        cmd = "hop"
        args += ["--lap", "%06x" % hop_lap]
    try:
        out, err = run(cmd, *args)
        res = dict(stdout=out, stderr=err)
        if dump:
            res["records"] = parse_dump(tmp.name)
        return res
    finally:
        if tmp:
            os.unlink(tmp.name)
