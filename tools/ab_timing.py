#!/usr/bin/env python
"""A/B stage times of one throughput-mode batch (device-resident input) under different environment settings of the
library: quick look for kernel work, not the bench.  Usage: ab_timing.py "BTB200_NEST_FOLD=1,BTB200_NO_DEMC=1" "" ..."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gr_bluetooth_b200 as g
from gr_bluetooth_b200 import synth

FS, FC, B, S = 100e6, 2441e6, int(os.environ.get("AB_SLOTS", "512")), 62500
occ = float(os.environ.get("AB_OCC", "0.05"))
iq, truth = synth.generate_range(FS, FC, 0, B + 7, seed=1234, occupancy=occ, as_int16=True)
iq = iq.astype(np.float32).view(np.complex64) if iq.dtype == np.int16 else iq
ref = None
for cfg in sys.argv[1:] or [""]:
    env = dict(kv.split("=") for kv in cfg.split(",") if kv)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    blk = g.multi_sniffer.make(FS, FC, 10.0, False, mm_mode=g.MM_STATELESS, max_slots=B, ddc=g.DDC_POLYPHASE)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    H = blk.history()
    w0 = 7 * S - (H - 1)
    n_in = (B - 1) * S + H
    d = torch.from_numpy(iq[w0:w0 + n_in].view(np.float32).copy()).cuda()
    best = None
    for i in range(5):
        hits, syms, ovf = blk.process_device(d.data_ptr(), n_in, 7, B, want_symbols=True)
        tm = blk.last_timing()
        if best is None or tm["total"] < best["total"]: best = tm
    key = [(int(h["slot"]), int(h["channel"]), int(h["kind"]), int(h["lap"]), int(h["offset"]), int(h["n_symbols"])) for h in hits]
    if ref is None: ref = (key, syms.copy())
    same = key == ref[0] and np.array_equal(syms, ref[1])
    print("%-40s %s hits %d same_as_first %s Msps %.0f" % (cfg or "(default)", json.dumps({k: round(v, 3) for k, v in best.items()}),
                                                       len(hits), same, B * S / best["total"] / 1e3), flush=True)
    blk.close(); del d
    torch.cuda.empty_cache()
