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


LE_ADV_CHANNELS = {0: 37, 24: 38, 78: 39}       # classic channel number -> LE advertising index
LE_ADV_AA = 0x8E89BED6


def ble_whiten(bits, chan_idx):
    """BLE data whitening (Core spec vol 6 part B 3.2): LFSR x^7 + x^4 + 1, position 0 = 1,
    positions 1..6 = channel index MSB..LSB, output from position 6."""
    reg = [1] + [(chan_idx >> (5 - i)) & 1 for i in range(6)]
    out = np.empty(len(bits), np.uint8)
    for i, b in enumerate(bits):
        o = reg[6]
        out[i] = b ^ o
        reg = [o, reg[0], reg[1], reg[2], reg[3] ^ o, reg[4], reg[5]]
    return out


def ble_adv_packet(chan_idx, rng):
    """Air-order bits of an ADV_IND: preamble, access address, whitened (header + payload + CRC bits)."""
    aa = [(LE_ADV_AA >> i) & 1 for i in range(32)]
    pre = [0, 1] * 4 if aa[0] == 0 else [1, 0] * 4
    n_data = int(rng.integers(0, 20))
    hdr0, hdr1 = 0x00, 6 + n_data                    # PDU type 0, TxAdd = RxAdd = 0; length
    pdu = [hdr0, hdr1] + [int(v) for v in rng.integers(0, 256, 6 + n_data)] + [int(v) for v in rng.integers(0, 256, 3)]
    bits = np.array([(byte >> i) & 1 for byte in pdu for i in range(8)], np.uint8)
    return np.concatenate([np.array(pre + aa, np.uint8), ble_whiten(bits, chan_idx)])


def generate(fs, fc, nslots, seed=1234, laps=None, occupancy=0.05, snr_db=17.0, sigma=50.0,
             burst_seed=5678, return_f64=False, le_adv_occupancy=0.0):
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
            if le_adv_occupancy > 0 and ch in LE_ADV_CHANNELS and brng.random() < le_adv_occupancy / occupancy:
                # a BLE advertising packet (h = 0.5) on an advertising channel
                bits = ble_adv_packet(LE_ADV_CHANNELS[ch], brng)
                bits = np.concatenate([bits[:2] ^ 1, bits])             # two ramp-up symbols
                start = slot * S + int(brng.uniform(0, 200e-6) * fs)
                sig = gfsk(bits, sps, h=0.5)
                k = np.arange(len(sig))
                f_off = (2402e6 + ch * 1e6 - fc) / fs
                sig = amp * sig * np.exp(2j * np.pi * f_off * (start + k)) * np.exp(2j * np.pi * brng.random())
                end = min(start + len(sig), n)
                if end > start:
                    iq[start:end] += sig[:end - start].astype(np.complex64)
                    truth.append(dict(slot=slot, channel=ch, lap=LE_ADV_AA, start_sample=start, nsym=len(bits), kind=1))
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
            truth.append(dict(slot=slot, channel=ch, lap=lap, start_sample=start, nsym=len(bits), kind=0))
    return iq, truth


def _draw_burst(brng, fs, S, slot, ch, laps, occupancy, le_adv_occupancy):
    """The random draws of one (slot, channel) cell, in generate()'s order.  -> None or a dict describing the burst."""
    if brng.random() >= occupancy:
        return None
    if le_adv_occupancy > 0 and ch in LE_ADV_CHANNELS and brng.random() < le_adv_occupancy / occupancy:
        bits = ble_adv_packet(LE_ADV_CHANNELS[ch], brng)
        bits = np.concatenate([bits[:2] ^ 1, bits])
        start = slot * S + int(brng.uniform(0, 200e-6) * fs)
        return dict(kind=1, bits=bits, start=start, h=0.5, phase=brng.random(), lap=LE_ADV_AA)
    lap = int(laps[int(brng.integers(0, len(laps)))])
    hdr = np.repeat(brng.integers(0, 2, 18).astype(np.uint8), 3)
    pay = brng.integers(0, 2, int(brng.integers(0, 367))).astype(np.uint8)
    ac = access_code(lap)
    bits = np.concatenate([np.array([0, 1, 0, 1], np.uint8) ^ ac[0] ^ 1, ac, hdr, pay])
    start = slot * S + int(brng.uniform(0, 200e-6) * fs)
    return dict(kind=0, bits=bits, start=start, h=0.32, phase=brng.random(), lap=lap)


def generate_range(fs, fc, slot0, nslots, seed=1234, laps=None, occupancy=0.05, snr_db=17.0, sigma=50.0,
                   burst_seed=5678, le_adv_occupancy=0.0, as_int16=False):
    """Slots [slot0, slot0 + nslots) of ONE unbounded synthetic stream defined by (seed, burst_seed): the noise of
    every slot comes from its own generator and the bursts from one sequential draw order, so any two calls agree
    sample for sample where their ranges overlap -- what time-sharding across GPUs needs.
    Same signal model as generate().  -> (iq, truth) with absolute slot numbers / sample positions in truth;
    iq is complex64 [nslots * S], or with as_int16 the ROUNDED samples as interleaved int16 [2 * nslots * S]."""
    laps = DEFAULT_LAPS if laps is None else laps
    S = int(625 * fs / 1e6)
    sps = fs / 1e6
    n = nslots * S
    iq = np.empty(n, np.complex64)
    v = iq.view(np.float32)
    for k in range(nslots):
        rng = np.random.default_rng([seed, slot0 + k])
        v[2 * k * S:2 * (k + 1) * S] = rng.standard_normal(2 * S, dtype=np.float32) * np.float32(sigma)
    amp = sigma * np.sqrt(2.0 * 10 ** (snr_db / 10.0) * 1e6 / fs)
    brng = np.random.default_rng(burst_seed)
    lo, hi = band_channels(fs, fc)
    truth = []
    base = slot0 * S
    for slot in range(slot0 + nslots):
        for ch in range(lo, hi + 1):
            bst = _draw_burst(brng, fs, S, slot, ch, laps, occupancy, le_adv_occupancy)
            if bst is None or slot < slot0 - 1:
                continue                       # a burst is shorter than a slot: older ones cannot reach the range
            start, bits = bst["start"], bst["bits"]
            end = start + len(bits) * int(sps)
            if end <= base or start >= base + n:
                continue
            sig = gfsk(bits, sps, h=bst["h"])
            k = np.arange(len(sig))
            f_off = (2402e6 + ch * 1e6 - fc) / fs
            sig = amp * sig * np.exp(2j * np.pi * f_off * (start + k)) * np.exp(2j * np.pi * bst["phase"])
            a, b = max(start, base), min(start + len(sig), base + n)
            iq[a - base:b - base] += sig[a - start:b - start].astype(np.complex64)
            if slot >= slot0:
                truth.append(dict(slot=slot, channel=ch, lap=bst["lap"], start_sample=start, nsym=len(bits), kind=bst["kind"]))
    if as_int16:
        return np.round(iq.view(np.float32)).astype(np.int16), truth
    return iq, truth
