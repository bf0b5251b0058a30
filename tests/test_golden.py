"""The committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from
the CPU oracle): the oracle must keep reproducing them bit for bit, and they must contain the
payloads that were transmitted (truth independent of the oracle)."""
import os

import numpy as np

from gr_ais_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.load(os.path.join(HERE, "golden", name))


def test_oracle_reproduces_chain_vectors(oracle):
    z = load("chain_kat.npz")
    for name in ("l120", "l140"):
        r = oracle.demod_chain(z[name + "_iq"], z[name + "_template"])
        assert np.array_equal(r["bits"], z[name + "_bits"])
        for f in ("offset", "key", "port", "value"):
            assert np.array_equal(r["tags"][f], z[name + "_tags"][f]), f
        sent = [bytes(z[name + "_payloads"][i:i + 21]) for i in (0, 21)]
        found = synth.hdlc_deframe(z[name + "_bits"])
        # (the 140-tap template loses the burst with the larger frequency offset)
        assert sum(p in found for p in sent) >= (2 if name == "l120" else 1)


def test_oracle_reproduces_rx_vectors(oracle):
    z = load("rx_kat.npz")
    rate, taps, x = float(z["rate"]), z["taps"], z["iq"]
    assert np.array_equal(taps, oracle.firdes_low_pass(1.0, rate, 11e3, 1e3))
    D = int(rate / 48000)
    sentences = []
    tmpl = oracle.gmsk_template_bits(np.array([1, 1, 0, 0] * 6, np.uint8))
    for k, (f, des) in enumerate(((-25e3, "A"), (25e3, "B"))):
        buf = np.concatenate([np.zeros(len(taps) - 1, np.complex64), x])
        nout = (len(buf) - (len(taps) - 1)) // D
        y = oracle.FreqXlatingFir(D, taps, f, rate).work(buf[:len(taps) - 1 + nout * D])
        assert np.array_equal(y, z["chan%d" % k])
        bits, _ = oracle.DemodStream(tmpl, oracle.chain_cfg()).work(y)
        assert np.array_equal(bits, z["bits%d" % k])
        frames = oracle.HdlcDeframer(11, 64).work(bits)
        assert np.array_equal(frames["end_bit"], z["end_bit%d" % k])
        sentences += [oracle.pdu_to_nmea(bytes(fr["data"][:fr["len"]]), des) for fr in frames]
    assert sentences == list(z["sentences"])
    # all four transmitted payloads came out, as checksummed sentences
    from test_rx_oracle import dearmour, nmea_checksum_ok, sentence_fields
    sent = {bytes(z["payloads"][i:i + 21]) for i in range(0, len(z["payloads"]), 21)}
    got = {dearmour(sentence_fields(s)["payload"], 0) for s in sentences}
    assert sent == got and all(nmea_checksum_ok(s) for s in sentences)
