#!/usr/bin/env python
"""GPU timeline of the pipelined submit/collect loop (debug): per batch, when its kernels started, its search ended,
its tail ended and its copies ended, in ms after the first submit."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gr_bluetooth_b200 as g
from gr_bluetooth_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=512)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--nctx", type=int, default=3)
ap.add_argument("--host", action="store_true")
a = ap.parse_args()
FS, FC, B, S = 100e6, 2441e6, a.slots, 62500
x16, _ = synth.generate_range(FS, FC, 0, B + 7, as_int16=True)
blks = [g.multi_sniffer.make(FS, FC, 10.0, False, mm_mode=g.MM_STATELESS, max_slots=B, ddc=g.DDC_POLYPHASE) for _ in range(a.nctx)]
H = blks[0].history()
w0, n_in = 7 * S - (H - 1), (B - 1) * S + H
seg = x16[2 * w0:2 * (w0 + n_in)]
d = torch.from_numpy(seg.astype(np.float32)).cuda()
pins = [g.PinnedBuffer(n_in, i16=True) for _ in range(a.nctx)]
for p in pins:
    p.array[:] = seg
L = g.lib()
L.btb200_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
def submit(c):
    if a.host:
        blks[c].submit_i16(pins[c].ptr.value, False, n_in, 7, B)
    else:
        blks[c].submit(d.data_ptr(), True, n_in, 7, B)
for rep in range(2):
    rows = []
    stages = []
    torch.cuda.synchronize()
    blks[0].timer_start()
    import time
    t0 = time.perf_counter()
    n = a.steps
    for i in range(min(a.nctx - 1, n)):
        submit(i % a.nctx)
    for j in range(n):
        c = j % a.nctx
        blks[c].collect_begin()
        i = j + a.nctx - 1
        if i < n:
            submit(i % a.nctx)
        th = time.perf_counter()
        blks[c].collect(want_symbols="borrow")
        out = (C.c_float * 5)()
        L.btb200_debug_timeline(blks[c]._ctx, blks[0]._ctx, out)
        rows.append((j, [round(v, 2) for v in out], round((th - t0) * 1e3, 2), round((time.perf_counter() - t0) * 1e3, 2)))
        stages.append({k: round(v, 2) for k, v in blks[c].last_timing().items()})
    tot = blks[0].timer_stop()
    if rep == 1:
        for r in rows:
            print("batch %d: start_in %.2f start_k %.2f search_end %.2f tail_end %.2f copies_end %.2f | host collect from %.2f to %.2f" % (r[0], *r[1], r[2], r[3]))
        for st in stages[-3:]:
            print("  stages", st)
        print("total %.2f ms for %d steps -> %.2f ms/step" % (tot, n, tot / n))
