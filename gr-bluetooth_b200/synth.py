"""Synthetic wideband Bluetooth IQ with ground truth.

complex64 samples at `fs` centred on `fc`: AWGN (sigma per component) plus GFSK bursts
(BT 0.5, modulation index 0.32, 1 Msym/s, bit 1 = positive deviation) on the classic
channels the band covers.  Each burst = 72-symbol access code of a LAP from `laps`
(spec structure, built by this module's own BCH encoder) + 54 header symbols (18 random
bits, each sent 3 times) + random payload.  Host-side tool used by bench.py and tests;
not on the measured path.
"""
import numpy as np

_PN = [0x03, 0xF2, 0xA3, 0x3D, 0xD6, 0x9B, 0x12, 0x1C, 0x10]
_G = [1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1]

DEFAULT_LAPS = [0x9E8B33] + [int(v) for v in np.random.default_rng(20240607).integers(0, 1 << 24, 63)]


def _pn(pos):
    return (_PN[pos >> 3] >> (7 - (pos & 7))) & 1


def access_code(lap):
    """72 air-order symbols: preamble, (64,30) BCH sync word, trailer (Bluetooth Core, vol 2 part B 6.3)."""
    info = [(lap >> i) & 1 for i in range(24)]
    msb = (lap >> 23) & 1
    info += [1, 1, 0, 0, 1, 0] if msb else [0, 0, 1, 1, 0, 1]
    reg = [0] * 34
    for i in range(29, -1, -1):
        fb = (info[i] ^ _pn(38 + i)) ^ reg[33]
        reg = [_G[0] & fb] + [reg[j - 1] ^ (_G[j] & fb) for j in range(1, 34)]
    sync = [reg[i] ^ _pn(4 + i) for i in range(34)] + info
    pre = [1, 0, 1, 0] if sync[0] else [0, 1, 0, 1]
    trl = [1, 0, 1, 0] if msb else [0, 1, 0, 1]
    return np.array(pre + sync + trl, np.uint8)


def band_channels(fs, fc):
    """Classic channels multi_block::set_channels covers (lib/multi_block.cc:306-342)."""
    center = (fc - 2402e6) / 1e6
    bw = fs / 1e6
    lo = max(int(center - bw / 2 + 0.45 + 1), 0)
    hi = min(int(center + bw / 2 - 0.45), 78)
    return lo, hi


def gfsk(bits, sps, bt=0.5, h=0.32):
    """Unit-amplitude complex baseband GFSK of a bit array at sps samples/symbol."""
    sps = int(sps)
    nrz = np.repeat(2.0 * bits.astype(np.float64) - 1.0, sps)
    alpha = np.sqrt(np.log(2.0) / 2.0) / bt
    t = np.arange(-2 * sps, 2 * sps + 1) / sps
    g = np.exp(-(np.pi * t / alpha) ** 2)
    g /= g.sum()
    f = np.convolve(nrz, g, mode="same")
    phase = np.cumsum(f) * (np.pi * h / sps)
    return np.exp(1j * phase)


def generate(fs, fc, nslots, seed=1234, laps=None, occupancy=0.05, snr_db=17.0, sigma=50.0,
             burst_seed=5678, return_f64=False):
    """-> (iq complex64 [nslots*S], truth list of dict(slot, channel, lap, start_sample, nsym))."""
    laps = DEFAULT_LAPS if laps is None else laps
    S = int(625 * fs / 1e6)
    sps = fs / 1e6
    n = nslots * S
    rng = np.random.default_rng(seed)
    iq = np.empty(n, np.complex64)
    v = iq.view(np.float32)
    chunk = 1 << 24
    for i in range(0, 2 * n, chunk):
        m = min(chunk, 2 * n - i)
        v[i:i + m] = rng.standard_normal(m, dtype=np.float32) * np.float32(sigma)
    amp = sigma * np.sqrt(2.0 * 10 ** (snr_db / 10.0) * 1e6 / fs)
    brng = np.random.default_rng(burst_seed)
    lo, hi = band_channels(fs, fc)
    truth = []
    for slot in range(nslots):
        for ch in range(lo, hi + 1):
            if brng.random() >= occupancy:
                continue
            lap = int(laps[int(brng.integers(0, len(laps)))])
            hdr = np.repeat(brng.integers(0, 2, 18).astype(np.uint8), 3)
            pay = brng.integers(0, 2, int(brng.integers(0, 367))).astype(np.uint8)
            # a few symbols of ramp-up before the access code, as a real transmitter does
            bits = np.concatenate([np.array([0, 1, 0, 1], np.uint8) ^ access_code(lap)[0] ^ 1, access_code(lap), hdr, pay])
            start = slot * S + int(brng.uniform(0, 200e-6) * fs)
            sig = gfsk(bits, sps)
            k = np.arange(len(sig))
            f_off = (2402e6 + ch * 1e6 - fc) / fs
            sig = amp * sig * np.exp(2j * np.pi * f_off * (start + k)) * np.exp(2j * np.pi * brng.random())
            end = min(start + len(sig), n)
            if end <= start:
                continue
            iq[start:end] += sig[:end - start].astype(np.complex64)
            truth.append(dict(slot=slot, channel=ch, lap=lap, start_sample=start, nsym=len(bits)))
    return iq, truth
