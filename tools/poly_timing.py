#!/usr/bin/env python
"""Stage times of one batch in a given mode (device-resident input): quick look, not the bench."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gr_bluetooth_b200 as g
from gr_bluetooth_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=512)
ap.add_argument("--ddc", default="poly")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--occupancy", type=float, default=0.05)
a = ap.parse_args()
FS, FC = 100e6, 2441e6
B = a.slots
S = 62500
iq, truth = synth.generate(FS, FC, B + 7, seed=1234, occupancy=a.occupancy)
blk = g.multi_sniffer.make(FS, FC, 10.0, False, mm_mode=g.MM_STATELESS, max_slots=B,
                           ddc=g.DDC_POLYPHASE if a.ddc == "poly" else g.DDC_EXACT)
H = blk.history()
w0 = 7 * S - (H - 1)
n_in = (B - 1) * S + H
d = torch.from_numpy(iq[w0:w0 + n_in].view(np.float32).copy()).cuda()
for i in range(a.iters):
    hits, syms, ovf = blk.process_device(d.data_ptr(), n_in, 7, B, want_symbols=True)
    tm = blk.last_timing()
    print(json.dumps({k: round(v, 3) for k, v in tm.items()}), "hits", len(hits), "Msps", round(B * S / tm["total"] / 1e3, 1))
found = {(int(h["channel"]), int(h["lap"])) for h in hits if h["kind"] == 0}
expect = {(t["channel"], t["lap"]) for t in truth if t["slot"] <= B - 2}
print("truth", len(expect), "found", len(expect & found), "BR hits", int((hits["kind"] == 0).sum()), "LE hits", int((hits["kind"] == 1).sum()))
