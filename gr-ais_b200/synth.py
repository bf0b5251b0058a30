"""Synthetic AIS traffic for tests and benchmarks (not part of the demod path).

Builds ITU-R M.1371 style bursts from first principles in float64 -- payload ->
CRC-16/X.25 -> HDLC bit stuffing -> 0x7E flags -> 24-bit 0101.. training ->
NRZI -> GMSK (BT 0.4, h 0.5) -- so that "known payload in => same payload out"
pins the demodulator independently of any restated GNU Radio kernel
(KAT 1).  Also holds the CPU HDLC deframer / CRC checker
used for the packet-detect-rate sweep (the step after the path; the reference
uses digital.hdlc_deframer_bp(11, 64), python/radio.py:64).
"""
import numpy as np

FS = 48000
SPS = 5
SYMBOL_RATE = 9600
SEED = 20260925


# ---------------------------------------------------------------- framing

def crc16_x25(data: bytes) -> int:
    crc = 0xFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return crc ^ 0xFFFF


def bytes_to_bits_lsb(data: bytes) -> np.ndarray:
    return np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8), bitorder="little")


def bits_to_bytes_lsb(bits) -> bytes:
    return np.packbits(np.asarray(bits, dtype=np.uint8), bitorder="little").tobytes()


def hdlc_stuff(bits) -> np.ndarray:
    out, run = [], 0
    for b in bits:
        out.append(int(b))
        run = run + 1 if b else 0
        if run == 5:
            out.append(0)
            run = 0
    return np.array(out, dtype=np.uint8)


def frame_bits(payload: bytes, ramp_bits: int = 8, tail_bits: int = 24) -> np.ndarray:
    """Over-the-air data bits (before NRZI) of one AIS slot."""
    fcs = crc16_x25(payload)
    body = bytes_to_bits_lsb(payload + bytes([fcs & 0xFF, fcs >> 8]))
    flag = np.array([0, 1, 1, 1, 1, 1, 1, 0], dtype=np.uint8)
    training = np.tile(np.array([0, 1], dtype=np.uint8), 12)
    return np.concatenate([np.zeros(ramp_bits, np.uint8), training, flag, hdlc_stuff(body), flag,
                           np.zeros(tail_bits, np.uint8)])


def nrzi_encode(bits, level: int = 1) -> np.ndarray:
    """AIS NRZI: a 0 toggles the line level, a 1 holds it."""
    out = np.empty(len(bits), dtype=np.uint8)
    for i, b in enumerate(bits):
        if not b:
            level ^= 1
        out[i] = level
    return out


def nrzi_level_for_training(ramp_bits: int = 8) -> int:
    """Initial line level that makes the training sequence appear as 1,1,0,0,... levels."""
    for level in (0, 1):
        lv = nrzi_encode(np.concatenate([np.zeros(ramp_bits, np.uint8),
                                         np.tile(np.array([0, 1], np.uint8), 12)]), level)
        if list(lv[ramp_bits:ramp_bits + 4]) == [1, 1, 0, 0]:
            return level
    raise AssertionError


# ------------------------------------------------------------- modulation

def gaussian_pulse(bt: float, sps: int, span: int = 4) -> np.ndarray:
    """Frequency pulse g(t) = gaussian (*) rect(T), sampled at sps, unit area."""
    t = (np.arange(-span * sps, span * sps + 1)) / sps
    sigma = np.sqrt(np.log(2.0)) / (2 * np.pi * bt)
    fine = 32
    tf = (np.arange(-span * sps * fine, span * sps * fine + 1)) / (sps * fine)
    g = np.exp(-0.5 * (tf / sigma) ** 2)
    g /= g.sum()
    rect = np.ones(sps * fine) / (sps * fine)
    full = np.convolve(g, rect)
    centre = (len(full) - 1) / 2
    idx = centre + t * sps * fine
    q = np.interp(idx, np.arange(len(full)), full)
    return q / q.sum()


def gmsk_modulate(levels, sps: int = SPS, bt: float = 0.4, frac_delay: float = 0.0) -> np.ndarray:
    """float64 GMSK, modulation index 0.5: +-pi/2 phase advance per symbol."""
    nrz = 2.0 * np.asarray(levels, dtype=np.float64) - 1.0
    up = np.zeros(len(nrz) * sps)
    up[::sps] = nrz
    q = gaussian_pulse(bt, sps)
    freq = np.convolve(up, q)  # per-sample phase increment / (pi/2)
    phase = (np.pi / 2) * np.cumsum(freq)
    if frac_delay:
        n = np.arange(len(phase), dtype=np.float64)
        phase = np.interp(n - frac_delay, n, phase, left=phase[0], right=phase[-1])
    return np.exp(1j * phase)


def make_burst(payload: bytes, sps: int = SPS, frac_delay: float = 0.0) -> np.ndarray:
    bits = frame_bits(payload)
    levels = nrzi_encode(bits, nrzi_level_for_training())
    return gmsk_modulate(levels, sps, frac_delay=frac_delay)


# ---------------------------------------------------------------- records

def rng_for(channel: int, seed: int = SEED) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, channel]))


def random_payload(rng, nbytes: int = 21) -> bytes:
    return bytes(rng.integers(0, 256, nbytes, dtype=np.uint8).tolist())


def make_record(channel: int = 0, n: int = FS, nbursts: int = 4, snr_db: float = 30.0,
                cfo_hz: float = 0.0, amplitude: float = 1.0, seed: int = SEED,
                random_impairments: bool = False):
    """One channel record: AWGN + `nbursts` bursts at slot-aligned random offsets.

    snr_db is Es/N0 in the symbol bandwidth: per-sample complex noise variance is
    amplitude^2 * sps / 10^(snr_db/10).  Returns (iq complex64 [n], truth list of
    dict(start, payload, cfo_hz))."""
    rng = rng_for(channel, seed)
    sigma2 = amplitude ** 2 * SPS / (10.0 ** (snr_db / 10.0))
    x = np.sqrt(sigma2 / 2) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    truth = []
    slot = 1280
    first_slot = 2  # leave the AGC / correlator pipelines time to fill
    nslots = (n - 1600) // slot - first_slot
    if nbursts > 0 and nslots >= nbursts:
        slots = np.sort(rng.choice(nslots, size=nbursts, replace=False)) + first_slot
        for s in slots:
            payload = random_payload(rng)
            frac = float(rng.uniform(0, 1)) if random_impairments else 0.0
            cfo = float(rng.uniform(-500, 500)) if random_impairments else cfo_hz
            ph0 = float(rng.uniform(0, 2 * np.pi)) if random_impairments else 0.0
            b = make_burst(payload, frac_delay=frac)
            start = int(s) * slot + int(rng.integers(0, 64))
            m = min(len(b), n - start)
            k = np.arange(m)
            x[start:start + m] += amplitude * b[:m] * np.exp(1j * (2 * np.pi * cfo * k / FS + ph0))
            truth.append(dict(start=start, payload=payload, cfo_hz=cfo))
    return x.astype(np.complex64), truth


def replicate_record(base: np.ndarray, channels: int, rotate: int = 16) -> np.ndarray:
    """Config-2 style fan-out: channel c is the base record rotated by rotate*c samples."""
    out = np.empty((channels, len(base)), dtype=np.complex64)
    for c in range(channels):
        out[c] = np.roll(base, rotate * c)
    return out


# ------------------------------------------------------- deframe + verify

def hdlc_deframe(bits, min_bytes: int = 11, max_bytes: int = 64):
    """Bit-level HDLC deframer with CRC check; returns the payloads that pass."""
    out = []
    ones = 0
    frame = None
    for b in bits:
        b = int(b)
        if b:
            ones += 1
            if frame is not None:
                frame.append(1)
            continue
        # b == 0
        if ones == 6:  # flag 01111110 just ended
            if frame is not None and len(frame) >= 7:
                body = frame[:-7]  # drop the 0 + six 1s of the closing flag
                nbytes = len(body) // 8
                if len(body) % 8 == 0 and min_bytes <= nbytes - 2 <= max_bytes:
                    raw = bits_to_bytes_lsb(body)
                    if crc16_x25(raw[:-2]) == (raw[-2] | (raw[-1] << 8)):
                        out.append(raw[:-2])
            frame = []
        elif ones > 6:
            frame = None
        elif ones == 5:
            pass  # stuffed zero: drop it
        elif frame is not None:
            frame.append(0)
        ones = 0
    return out


def payloads_found(bits, truth):
    found = hdlc_deframe(bits)
    return [t["payload"] in found for t in truth]


# ------------------------------------------------------- wideband captures

def make_wideband(source: int = 0, rate: float = 250e3, n: int = 250000, nbursts: int = 3,
                  snr_db: float = 25.0, freqs=(-25e3, 25e3), seed: int = SEED):
    """One wideband capture as the reference's ais_rx sees it (python/radio.py:86-91): AIS
    channel k sits at freqs[k] Hz from the centre.  Bursts are GMSK at 9600 baud generated
    directly at `rate` (float64), with AWGN of per-sample variance rate/9600 / 10^(snr/10)
    (Es/N0 in the symbol bandwidth).  Returns (iq complex64 [n], truth list of
    dict(channel, start, payload))."""
    rng = rng_for(1000003 + source, seed)
    sps = rate / SYMBOL_RATE
    sigma2 = sps / (10.0 ** (snr_db / 10.0))
    x = np.sqrt(sigma2 / 2) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    truth = []
    slot = int(round(256 * sps))
    first_slot = 3
    nslots = (n - int(300 * sps)) // slot - first_slot
    t = np.arange(n, dtype=np.float64)
    for k, f in enumerate(freqs):
        if nslots < nbursts or nbursts <= 0:
            continue
        slots = np.sort(rng.choice(nslots, size=nbursts, replace=False)) + first_slot
        for s in slots:
            payload = random_payload(rng)
            bits = frame_bits(payload)
            levels = nrzi_encode(bits, nrzi_level_for_training())
            b = gmsk_modulate_rate(levels, sps)
            start = int(s) * slot + int(rng.integers(0, 64))
            m = min(len(b), n - start)
            x[start:start + m] += b[:m] * np.exp(2j * np.pi * f * t[start:start + m] / rate)
            truth.append(dict(channel=k, start=start, payload=payload))
    return x.astype(np.complex64), truth


def gmsk_modulate_rate(levels, sps: float, bt: float = 0.4) -> np.ndarray:
    """float64 GMSK (h = 0.5) at a non-integer number of samples per symbol: the phase is
    built at 8 samples per symbol and interpolated linearly onto the output grid."""
    fine = gmsk_phase(levels, 8, bt)
    nout = int(len(levels) * sps)
    pos = np.arange(nout) * (8.0 / sps)
    return np.exp(1j * np.interp(pos, np.arange(len(fine)), fine))


def gmsk_phase(levels, sps: int, bt: float = 0.4) -> np.ndarray:
    nrz = 2.0 * np.asarray(levels, dtype=np.float64) - 1.0
    up = np.zeros(len(nrz) * sps)
    up[::sps] = nrz
    return (np.pi / 2) * np.cumsum(np.convolve(up, gaussian_pulse(bt, sps)))
