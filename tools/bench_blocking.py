"""Throughput of the BLOCKING call (btb200_process with host buffers: one batch in flight, nothing pipelined) --
what a gr::sync_block::work() caller gets.  The input copy is cut in parts that the channel FIR follows."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gr_bluetooth_b200 as g
import bench

B = 512
iq, truth, lead, S = bench.synth_batch(B, seed=1234)
blk = g.multi_sniffer.make(bench.FS, bench.FC, bench.SNR_DB, False, mm_mode=g.MM_STATELESS, max_slots=B)
H = blk.history(); w0 = lead * S - (H - 1); n_in = (B - 1) * S + H
pin = g.PinnedBuffer(n_in); pin.array[:] = iq[w0:w0 + n_in]
for i in range(3):
    blk.submit(pin.ptr.value, False, n_in, lead, B); blk.collect(want_symbols=True)
t = time.perf_counter()
N = 5
for i in range(N):
    blk.submit(pin.ptr.value, False, n_in, lead, B); h, _, _ = blk.collect(want_symbols=True)
dt = (time.perf_counter() - t) / N
print("blocking: %.2f ms/batch, %.1f Msps, hits %d" % (dt * 1e3, B * S / dt / 1e6, len(h)), blk.last_timing())
pin.close(); blk.close()
