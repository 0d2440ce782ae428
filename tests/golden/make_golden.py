#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference (run in the build container,
where /root/reference exists and oracle/_ref/btref has been built by
`make -C oracle ref`).  The fixtures are OUTPUTS of the reference's own code
(lib/*.cc compiled verbatim) on excerpts of its bundled samples:

  ref_kats.json        acgen() for a list of LAPs, the channel37.dem hit list,
                       the survey-time stdout digests, per-file hit lists
  ref_tables.txt       the reference's detection LUTs (`btref tables`)
  channel37_bits.npz   samples/channel37.dem packed 8 symbols per byte
  <name>_<mode>.npz    IQ excerpt (int16: the captures are integer valued) of the
                       first N slots of samples/<name>.cfile + what the reference
                       computes for them: stdout, energies, bit streams, and for
                       a few windows the DDC output / demod / soft symbols
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref as R  # noqa: E402

SAMPLES = "/root/reference/samples"
FILES = {  # name: (fs, fc, slots in the committed excerpt, heavy-capture call range)
    "headset3": (2e6, 2476e6, 300, (86, 89)),
    "headset1": (8e6, 2476.5e6, 52, (6, 8)),
    "keyboard1": (8e6, 2476.5e6, 16, (5, 7)),
    "headset2": (4e6, 2476e6, 60, (32, 34)),
}
LAPS = [0x9E8B33, 0x9E8B00, 0x24D952, 0x4831DD, 0x000000, 0xFFFFFF, 0x800000, 0x000001, 0xF2F57B, 0x133BEC,
        0xFC6FEE, 0xAE3CCB, 0x123456, 0xABCDEF, 0x5A5A5A, 0xA5A5A5]


def main():
    assert R.build(), "oracle/_ref/btref not built"
    kats = {}
    out, _ = R.run("acgen", *["%06x" % l for l in LAPS])
    kats["acgen"] = {l.split()[0]: l.split()[1] for l in out.splitlines()}
    out, _ = R.run("sniffdem", os.path.join(SAMPLES, "channel37.dem"))
    kats["channel37_hits"] = [[int(l.split()[0]), int(l.split()[1], 16)] for l in out.splitlines()]
    out, _ = R.run("tables")
    open(os.path.join(HERE, "ref_tables.txt"), "w").write(out)
    dem = np.fromfile(os.path.join(SAMPLES, "channel37.dem"), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "channel37_bits.npz"), packed=np.packbits(dem), n=len(dem))

    kats["stdout_md5"] = {}
    kats["file_hits"] = {}
    for name, (fs, fc, nslots, heavy) in FILES.items():
        path = os.path.join(SAMPLES, name + ".cfile")
        full = R.sniff(path, fs, fc)
        kats["stdout_md5"][name] = hashlib.md5(full["stdout"].encode()).hexdigest()
        kats["file_hits"][name] = [[h["slot"], h["kind"], h.get("channel", -1), h["lap"], h["snr"]]
                                   for h in R.parse_stdout_hits(full["stdout"])]
        iq = np.fromfile(path, dtype=np.complex64)
        S = int(625 * fs / 1e6)
        n = nslots * S
        x = iq[:n]
        xi = np.empty(2 * n, np.int16)
        xi[0::2] = x.real.astype(np.int16)
        xi[1::2] = x.imag.astype(np.int16)
        assert np.array_equal(xi[0::2].astype(np.float32), x.real) and np.array_equal(xi[1::2].astype(np.float32), x.imag)
        for mode in ("chained", "stateless"):
            r = R.sniff(path, fs, fc, stateless=(mode == "stateless"), num_calls=nslots, dump=True, heavy=heavy)
            recs = r["records"]
            nddc = max(rid for _, _, rid, _ in recs) + 1
            nch = nddc // 2
            energy = np.full((nslots, nch), np.nan)
            noise = np.full((nslots, nch), np.nan)
            nsym = np.zeros((nslots, nch), np.int32)
            bits = {}
            heavy_d = {}
            for typ, call, rid, pay in recs:
                chi = rid // 2
                if typ == R.REC_ENERGY:
                    (noise if rid % 2 else energy)[call, chi] = pay[0]
                elif typ == R.REC_BITS:
                    nsym[call, chi] = len(pay)
                    bits[(call, chi)] = pay
                elif typ in (R.REC_DDC, R.REC_SOFT, R.REC_DEMOD, R.REC_MU):
                    if typ == R.REC_DDC and rid % 2:
                        continue
                    key = {R.REC_DDC: "ddc", R.REC_SOFT: "soft", R.REC_DEMOD: "demod", R.REC_MU: "mu"}[typ]
                    heavy_d["%s_%d_%d" % (key, call, chi)] = pay
            stride = int(nsym.max())
            packed = np.zeros((nslots, nch, (stride + 7) // 8), np.uint8)
            for (call, chi), pay in bits.items():
                pb = np.packbits(pay)
                packed[call, chi, :len(pb)] = pb
            np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (name, mode)),
                                iq_i16=xi if mode == "chained" else np.zeros(0, np.int16),
                                fs=fs, fc=fc, nslots=nslots, stdout=np.array(r["stdout"]),
                                energy=energy, noise=noise, nsym=nsym, bits_packed=packed, **heavy_d)
            print(name, mode, "slots", nslots, "hits", len(R.parse_stdout_hits(r["stdout"])))
    json.dump(kats, open(os.path.join(HERE, "ref_kats.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
