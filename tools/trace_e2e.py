"""Host-side timeline of the double-buffered end-to-end loop (submit/collect over two contexts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gr_bluetooth_b200 as g
import bench

NCTX = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 512
iq, truth, lead, S = bench.synth_batch(B, seed=1234)
blks = [g.multi_sniffer.make(bench.FS, bench.FC, bench.SNR_DB, False, mm_mode=g.MM_STATELESS, max_slots=B)
        for _ in range(NCTX)]
H = blks[0].history(); w0 = lead * S - (H - 1); n_in = (B - 1) * S + H
pinned = [g.PinnedBuffer(n_in) for _ in range(NCTX)]
for p in pinned:
    p.array[:] = iq[w0:w0 + n_in]
T0 = time.perf_counter()
log = []
def stamp(what):
    log.append((time.perf_counter() - T0, what))
def loop(n):
    for i in range(min(NCTX - 1, n)):
        blks[i % NCTX].submit(pinned[i % NCTX].ptr.value, False, n_in, lead, B); stamp("submit %d done" % i)
    for j in range(n):
        blks[j % NCTX].collect_begin(); stamp("begin %d done" % j)
        i = j + NCTX - 1
        if i < n:
            blks[i % NCTX].submit(pinned[i % NCTX].ptr.value, False, n_in, lead, B); stamp("submit %d done" % i)
        blks[j % NCTX].collect(want_symbols=True); stamp("collect %d done" % j)
loop(4)
torch.cuda.synchronize()
log.clear(); T0 = time.perf_counter()
N = 8
loop(N)
torch.cuda.synchronize()
tot = time.perf_counter() - T0
prev = 0.0
for t, w in log:
    print("%8.2f ms  (+%6.2f)  %s" % (t * 1e3, (t - prev) * 1e3, w)); prev = t
print("contexts %d: %.2f ms/step, %.1f Msps" % (NCTX, tot / N * 1e3, N * B * S / tot / 1e6))
