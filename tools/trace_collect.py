"""Host-side timeline of btb200_collect (BTB200_TRACE=1): where the tail of a batch goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gr_bluetooth_b200 as g
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
B = 512
iq, truth, lead, S = bench.synth_batch(B, seed=1234)
blk = g.multi_sniffer.make(bench.FS, bench.FC, bench.SNR_DB, False, mm_mode=g.MM_STATELESS, max_slots=B,
                           snr_mode=g.SNR_FAST_GUARDED if mode == "fast" else g.SNR_EXACT)
H = blk.history(); w0 = lead * S - (H - 1); n_in = (B - 1) * S + H
d = torch.from_numpy(iq[w0:w0 + n_in].view(np.float32).copy()).cuda()
for i in range(4):
    sys.stderr.write("---- call %d ----\n" % i)
    hits, _, _ = blk.process_device(d.data_ptr(), n_in, lead, B, want_symbols=True)
    print(i, len(hits), blk.last_timing())
