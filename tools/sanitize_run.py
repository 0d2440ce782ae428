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
POLY_ONLY = bool(os.environ.get("SAN_POLY_ONLY"))      # only the throughput mode at 100 Msps and the libbtbb-style search
for kw in (() if POLY_ONLY else (dict(squelch=g.SQUELCH_EAGER), dict(squelch=g.SQUELCH_LAZY), dict(squelch=g.SQUELCH_LAZY, snr_mode=g.SNR_FAST_GUARDED))):
    for impl in ([1] if os.environ.get("SAN_QUICK") else [0, 1, 2]):
        blk = g.multi_sniffer(fs, fc, 10.0, mm_mode=g.MM_STATELESS, max_slots=B, **kw)
        blk.set_impl(impl)
        S, H = blk.samples_per_slot, blk.history()
        w0 = first * S - (H - 1)
        hits, syms, _ = blk.process(iq[w0:w0 + (B - 1) * S + H], first, B, want_symbols=True)
        print(kw, impl, len(hits))
        blk.close()
# throughput mode: polyphase channelizer, fused estimator + resume, device-driven tail, int16 input, window mask;
# both tail policies; also the 30 Msps geometry (N1 = 2) and an 8 Msps one (N1 = 1, pre-rotation by half a channel)
for tail in (g.TAIL_LAZY, g.TAIL_FULL):
    blk = g.multi_sniffer(fs, fc, 10.0, mm_mode=g.MM_STATELESS, max_slots=B, ddc=g.DDC_POLYPHASE, tail=tail)
    S, H = blk.samples_per_slot, blk.history()
    w0 = first * S - (H - 1)
    seg = iq[w0:w0 + (B - 1) * S + H]
    hits, syms, _ = blk.process(seg, first, B, want_symbols=True)
    h16, _, _ = blk.process_i16(np.round(seg.view(np.float32)).astype(np.int16), first, B, want_symbols="borrow")
    m = np.zeros((B, blk.info.n_channels), np.uint8); m[:, ::3] = 1
    blk.set_window_mask(m)
    hm, _, _ = blk.process(seg, first, B)
    print("poly", tail, len(hits), len(h16), len(hm))
    blk.close()
blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_STATELESS, max_slots=8, search=g.SEARCH_BR | g.SEARCH_BR_BCH, bch=g.bch_any(2))
print("bch", len(blk.search_bits(np.random.default_rng(2).integers(0, 2, 20000).astype(np.uint8))))
blk.close()
if POLY_ONLY:
    print("done")
    sys.exit(0)
for fs2, fc2 in ((30e6, 2414e6), (8e6, 2476.5e6)):
    iq2, _ = synth.generate(fs2, fc2, nslots, seed=4, occupancy=0.2, snr_db=20.0, le_adv_occupancy=0.05)
    blk = g.multi_sniffer(fs2, fc2, 10.0, mm_mode=g.MM_STATELESS, max_slots=B, ddc=g.DDC_POLYPHASE)
    S, H = blk.samples_per_slot, blk.history()
    w0 = first * S - (H - 1)
    print("poly", fs2, len(blk.process(iq2[w0:w0 + (B - 1) * S + H], first, B, want_symbols=True)[0]))
    blk.close()
print("hop", len(g.hop_candidates(0x0F24D952 & 0xFFFFFFF, 5, 10)))
blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_STATELESS, max_slots=8)
print("kat", len(blk.search_bits(np.random.default_rng(1).integers(0, 2, 20000).astype(np.uint8))))
blk.close()
# chained, small rate
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "keyboard1_chained.npz"))
xi = z["iq_i16"].astype(np.float32)
x = (xi[0::2] + 1j * xi[1::2]).astype(np.complex64)
blk = g.multi_sniffer(8e6, 2476.5e6, 10.0, mm_mode=g.MM_CHAINED, max_slots=8)
print("chained", len(blk.run_stream(x[:12 * 5000])))
blk.close()
print("done")
