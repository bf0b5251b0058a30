"""CPU tests of the oracle's rx rows (oracle/ais_oracle_rx.c): CRC, hdlc_deframer_bp,
pdu_to_nmea, firdes.low_pass, freq_xlating_fir_filter_ccf -- pinned by public AIVDM
sentences (tests/golden/aivdm_kat.json), standard check values and float64 truths."""
import json
import os

import numpy as np
import pytest

from gr_ais_b200 import synth
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load_kat():
    with open(os.path.join(HERE, "golden", "aivdm_kat.json")) as fh:
        return json.load(fh)


def nmea_checksum_ok(s):
    body = s[1:s.index("*")]
    x = 0
    for ch in body:
        x ^= ord(ch)
    return "%02X" % x == s[s.index("*") + 1:]


def dearmour(payload, npad):
    """6-bit ASCII armour -> bytes (ITU-R M.1371 / NMEA 0183 AIVDM), written independently."""
    bits = []
    for ch in payload:
        v = ord(ch) - 48
        if v > 40:
            v -= 8
        bits += [(v >> (5 - b)) & 1 for b in range(6)]
    if npad:
        bits = bits[:-npad]
    assert len(bits) % 8 == 0
    return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()


def sentence_fields(s):
    f = s[:s.index("*")].split(",")
    return dict(nfrag=int(f[1]), frag=int(f[2]), seq=f[3], chan=f[4], payload=f[5], npad=int(f[6]))


def test_kat_sentences_are_self_consistent():
    kat = load_kat()
    for s in kat["single"] + [x for m in kat["multi"] for x in m]:
        assert nmea_checksum_ok(s), s


def test_crc_check_value():
    # CRC-16/X.25 check value of the ASCII string "123456789"
    assert O.crc_ccitt(b"123456789") == 0x906E
    assert O.crc_ccitt(b"123456789") == synth.crc16_x25(b"123456789")


def test_pdu_to_nmea_reproduces_public_sentences():
    kat = load_kat()
    for s in kat["single"]:
        f = sentence_fields(s)
        pdu = dearmour(f["payload"], f["npad"])
        assert len(pdu) == 21
        assert O.pdu_to_nmea(pdu, f["chan"]) == s
    mmsi_s, mmsi = next(iter(kat["mmsi"].items()))
    pdu = dearmour(sentence_fields(mmsi_s)["payload"], 0)
    assert (int.from_bytes(pdu[:5], "big") >> 2) & 0x3FFFFFFF == mmsi


def test_pdu_to_nmea_fragments_like_the_public_type5():
    kat = load_kat()
    a, b = kat["multi"][0]
    fa, fb = sentence_fields(a), sentence_fields(b)
    pdu = dearmour(fa["payload"] + fb["payload"], fb["npad"])
    assert len(pdu) == 53  # 424 bits
    got = O.pdu_to_nmea(pdu, fa["chan"]).split("\n")
    assert len(got) == 2 and all(nmea_checksum_ok(g) for g in got)
    ga, gb = sentence_fields(got[0]), sentence_fields(got[1])
    assert (ga["nfrag"], ga["frag"], gb["nfrag"], gb["frag"]) == (2, 1, 2, 2)
    assert ga["seq"] == "" and gb["seq"] == ""  # the reference never numbers the sequence
    assert ga["payload"] == fa["payload"] and len(ga["payload"]) == 56
    assert gb["payload"] == fb["payload"] and gb["npad"] == 2 and ga["npad"] == 2


def test_pdu_to_nmea_padding_quirk():
    """lib/pdu_to_nmea_impl.cc:75-77 shifts the already left-aligned last group npad more
    times inside a uint8_t: its top bits fall off, and values >= 128 go negative in to_ascii's
    (signed) char arithmetic."""
    s = O.pdu_to_nmea(b"\xff", "A")           # 8 bits -> "?" + 2 bits, npad 4
    f = sentence_fields(s)
    assert f["npad"] == 4 and f["payload"][0] == "w"      # 63 -> 'w'
    assert f["payload"][1] == "0"                         # 0b110000 << 4 == 0 in uint8_t
    s = O.pdu_to_nmea(b"\xff\xff", "A")       # 16 bits: last group 4 bits, npad 2
    f = sentence_fields(s)
    assert f["npad"] == 2
    v = (0b111100 << 2) & 0xFF                # 240 -> char -16 -> not > 39 -> +48 = 32
    assert v == 240 and f["payload"][2] == chr((v - 256 + 48) & 0xFF)
    assert nmea_checksum_ok(s)


def air_bits(pdu):
    return synth.frame_bits(pdu)


def test_hdlc_deframes_kat_payloads_and_streams():
    kat = load_kat()
    pdus = [dearmour(sentence_fields(s)["payload"], 0) for s in kat["single"]]
    rng = np.random.default_rng(7)
    bits = np.concatenate([np.concatenate([rng.integers(0, 2, int(rng.integers(0, 90)), dtype=np.uint8),
                                           air_bits(p)]) for p in pdus])
    d = O.HdlcDeframer(11, 64)
    frames = d.work(bits)
    assert O.frames_payloads(frames)[-len(pdus):] == pdus or set(pdus) <= set(O.frames_payloads(frames))
    assert set(synth.hdlc_deframe(bits)) >= set(pdus)
    # ragged pieces give the same frames at the same absolute bit positions
    d2 = O.HdlcDeframer(11, 64)
    got, pos = [], 0
    while pos < len(bits):
        k = int(rng.integers(0, 200))
        got.append(d2.work(bits[pos:pos + k]))
        pos += k
    got = np.concatenate(got)
    assert np.array_equal(got["end_bit"], frames["end_bit"])
    assert O.frames_payloads(got) == O.frames_payloads(frames)


def test_hdlc_rejects_bad_crc_and_length_limits():
    pdu = bytes(range(21))
    bits = air_bits(pdu).copy()
    ok = O.HdlcDeframer(11, 64).work(bits)
    assert O.frames_payloads(ok) == [pdu]
    bad = bits.copy()
    bad[60] ^= 1
    assert len(O.HdlcDeframer(11, 64).work(bad)) == 0
    # length_min counts the CRC bytes: 9 payload bytes pass, 8 do not
    for n, want in ((9, 1), (8, 0)):
        assert len(O.HdlcDeframer(11, 64).work(air_bits(bytes(range(n))))) == want
    # the closing flag's first bits are shifted in before the delimiter is seen, so a frame may
    # hold at most length_max bytes including its CRC
    for n, want in ((62, 1), (63, 0)):
        assert len(O.HdlcDeframer(11, 64).work(air_bits(bytes(range(n))))) == want


def test_hdlc_bit_stuffing():
    pdu = b"\xff" * 12 + b"\x7e\x7e\x00\xff"
    got = O.HdlcDeframer(11, 64).work(air_bits(pdu))
    assert O.frames_payloads(got) == [pdu]


def test_firdes_low_pass_matches_scipy():
    from scipy.signal import firwin
    taps = O.firdes_low_pass(1.0, 250e3, 11e3, 1e3)
    assert len(taps) == 603 and taps.dtype == np.float32
    assert np.array_equal(taps, taps[::-1])
    assert abs(float(taps.astype(np.float64).sum()) - 1.0) < 1e-6
    ref = firwin(603, 11e3, window="hamming", fs=250e3)
    assert np.abs(taps - ref).max() < 2e-7
    assert len(O.firdes_low_pass(1.0, 1.2e6, 11e3, 1e3)) == 2891


@pytest.mark.parametrize("rate,freq", [(250e3, -25e3), (250e3, 25e3), (240e3, 0.0), (1.2e6, 25e3)])
def test_xlat_against_float64_truth(rate, freq):
    taps = O.firdes_low_pass(1.0, rate, 11e3, 1e3)
    D = int(rate / 48000)
    rng = np.random.default_rng(3)
    nout = 700
    n = len(taps) - 1 + nout * D
    t = np.arange(n)
    x = (np.exp(2j * np.pi * (freq + 1000.0) * t / rate)
         + 0.3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    f = O.FreqXlatingFir(D, taps, freq, rate)
    y, fir = f.work(x, fir=True)
    truth = f.f64(x)
    wf = float(np.float32(2 * np.pi * freq / rate))
    rot = np.exp(-1j * wf * D * np.arange(nout))
    # the filter proper: sequential float32 fma chains over up to 2891 taps vs float64
    assert np.abs(fir - truth / rot).max() < 1e-5
    # gr::blocks::rotator: |phase| wanders by ~3e-8 per item between its renormalisations
    # every 512 items, its angle by float rounding of the increment
    ph = y[np.abs(fir) > 0.1] / fir[np.abs(fir) > 0.1]
    assert np.abs(np.abs(ph) - 1).max() < 5e-5
    assert np.abs(np.angle(ph * np.conj(rot[np.abs(fir) > 0.1]))).max() < 5e-4
    # the +1 kHz tone comes out at +1 kHz of the decimated rate
    k = np.arange(nout)
    tone = np.exp(2j * np.pi * 1000.0 * k * D / rate)
    assert abs(np.vdot(tone[100:], y[100:])) / (nout - 100) > 0.9


def test_xlat_streams_like_one_call():
    rate, freq = 250e3, -25e3
    taps = O.firdes_low_pass(1.0, rate, 11e3, 1e3)
    D, nt = 5, len(taps)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(nt - 1 + 1500 * D) + 1j * rng.standard_normal(nt - 1 + 1500 * D)).astype(np.complex64)
    one = O.FreqXlatingFir(D, taps, freq, rate).work(x)
    f = O.FreqXlatingFir(D, taps, freq, rate)
    parts, pos = [], 0
    for k in (1, 511, 1, 600, 387):
        parts.append(f.work(x[pos * D: pos * D + nt - 1 + k * D]))
        pos += k
    assert np.array_equal(np.concatenate(parts), one)
    # conjugate-symmetric frequencies have exactly conjugate taps (one pass serves A and B)
    a = O.FreqXlatingFir(D, taps, -25e3, rate).ctaps
    b = O.FreqXlatingFir(D, taps, 25e3, rate).ctaps
    assert np.array_equal(a, np.conj(b))


def test_hdlc_oracle_agrees_with_the_independent_deframer_on_random_streams():
    """Two deframers written independently (the oracle's restatement of hdlc_deframer_bp's
    counters, synth.hdlc_deframe's flag/frame list) must publish the same payloads from the same
    bits: random noise with embedded frames, abort runs (seven or more ones), back-to-back and
    shared flags, and frames cut short."""
    rng = np.random.default_rng(11)
    flag = np.array([0, 1, 1, 1, 1, 1, 1, 0], np.uint8)
    for trial in range(40):
        parts = []
        for _ in range(int(rng.integers(3, 9))):
            kind = int(rng.integers(0, 6))
            pdu = bytes(rng.integers(0, 256, int(rng.integers(9, 62)), dtype=np.uint8).tolist())
            fcs = synth.crc16_x25(pdu)
            body = synth.hdlc_stuff(synth.bytes_to_bits_lsb(pdu + bytes([fcs & 0xFF, fcs >> 8])))
            if kind == 0:      # noise
                parts.append(rng.integers(0, 2, int(rng.integers(1, 400)), dtype=np.uint8))
            elif kind == 1:    # a plain frame
                parts += [flag, body, flag]
            elif kind == 2:    # two frames sharing one flag
                parts += [flag, body, flag, body, flag]
            elif kind == 3:    # an abort run inside a frame
                parts += [flag, body[:40], np.ones(int(rng.integers(7, 20)), np.uint8), flag]
            elif kind == 4:    # a frame cut short by noise
                parts += [flag, body[:int(rng.integers(1, len(body)))],
                          rng.integers(0, 2, 30, dtype=np.uint8)]
            else:              # idle flags
                parts += [flag] * int(rng.integers(1, 5))
        bits = np.concatenate(parts).astype(np.uint8)
        got = O.frames_payloads(O.HdlcDeframer(11, 64).work(bits, max_frames=512))
        want = synth.hdlc_deframe(bits, min_bytes=9, max_bytes=62)
        assert got == want, trial


def test_xlat_indexing_against_numpy_convolution():
    """History convention, decimation phase and rotator direction checked against an independent
    formulation: out[j] = conv(x, ctaps)[j*D + ntaps-1] * exp(-j fwT0 D j), with x = history + new."""
    rate, freq, D = 250e3, -25e3, 5
    taps = O.firdes_low_pass(1.0, rate, 11e3, 1e3)
    nt = len(taps)
    rng = np.random.default_rng(9)
    nout = 400
    x = (rng.standard_normal(nt - 1 + nout * D) + 1j * rng.standard_normal(nt - 1 + nout * D)).astype(np.complex64)
    f = O.FreqXlatingFir(D, taps, freq, rate)
    y, fir = f.work(x, fir=True)
    wf = float(np.float32(2 * np.pi * freq / rate))
    # GNU Radio forms the angle as the float product i * fwT0 (one rounding per tap)
    theta = (np.arange(nt, dtype=np.float32) * np.float32(wf)).astype(np.float64)
    ctaps = taps.astype(np.float64) * np.exp(1j * theta)
    assert np.abs(f.ctaps - ctaps).max() < 1e-7          # the band-pass taps GNU Radio composes
    full = np.convolve(x.astype(np.complex128), ctaps)
    want_fir = full[nt - 1:nt - 1 + nout * D:D]
    assert np.abs(fir - want_fir).max() < 1e-5
    want = want_fir * np.exp(-1j * wf * D * np.arange(nout))
    assert np.abs(y - want).max() < 1e-4
    # a tone at the channel centre lands at DC with the filter's unit gain
    t = np.arange(len(x))
    tone = np.exp(2j * np.pi * freq * t / rate).astype(np.complex64)
    z = O.FreqXlatingFir(D, taps, freq, rate).work(tone)
    assert np.abs(np.abs(z[200:]) - 1.0).max() < 1e-3
    assert np.abs(np.angle(z[201:] * np.conj(z[200:-1]))).max() < 1e-3
