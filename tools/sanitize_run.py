"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel family once.
  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gr_bluetooth_b200 as g
from gr_bluetooth_b200 import synth

fs, fc, nslots = 100e6, 2441e6, 10
iq, _ = synth.generate(fs, fc, nslots, seed=3, laps=[0x9E8B33, 0x24D952], occupancy=0.1, snr_db=20.0, le_adv_occupancy=0.05)
first, B = 7, 3
for kw in (dict(squelch=g.SQUELCH_EAGER), dict(squelch=g.SQUELCH_LAZY), dict(squelch=g.SQUELCH_LAZY, snr_mode=g.SNR_FAST_GUARDED)):
    for impl in ([1] if os.environ.get("SAN_QUICK") else [0, 1, 2]):
        blk = g.multi_sniffer(fs, fc, 10.0, mm_mode=g.MM_STATELESS, max_slots=B, **kw)
        blk.set_impl(impl)
        S, H = blk.samples_per_slot, blk.history()
        w0 = first * S - (H - 1)
        hits, syms, _ = blk.process(iq[w0:w0 + (B - 1) * S + H], first, B, want_symbols=True)
        print(kw, impl, len(hits))
        blk.close()
# chained, small rate
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keyboard1_chained.npz"))
xi = z["iq_i16"].astype(np.float32)
x = (xi[0::2] + 1j * xi[1::2]).astype(np.complex64)
blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_CHAINED, max_slots=8)
print("chained", len(blk.run_stream(x[:12 * 5000])))
blk.close()
print("done")
