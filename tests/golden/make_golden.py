#!/usr/bin/env python3
"""Regenerates tests/golden/chain_kat.npz and rx_kat.npz from the CPU oracle.

The reference ships no vectors and cannot run here, so these fixtures
freeze the ORACLE's answers on small seeded inputs: they catch any later drift of the oracle
itself (tests/test_golden.py, CPU) and give the GPU path committed vectors to match
(tests/test_gpu_golden.py).  Inputs are stored too, so the fixtures do not depend on numpy's
random streams.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gr_ais_b200 import synth  # noqa: E402
from gr_ais_b200.ais_demod import preamble_template  # noqa: E402
from oracle import oracle as O  # noqa: E402


def chain():
    n = 16384
    out = {}
    for name, kind in (("l120", "north_star"), ("l140", "intended")):
        x, truth = synth.make_record(7, n=n, nbursts=2, snr_db=20.0, random_impairments=True)
        tmpl = preamble_template(kind)
        r = O.demod_chain(x, tmpl)
        out[name + "_iq"] = x
        out[name + "_template"] = tmpl
        out[name + "_bits"] = r["bits"]
        out[name + "_tags"] = r["tags"]
        out[name + "_payloads"] = np.frombuffer(b"".join(t["payload"] for t in truth), np.uint8)
    np.savez_compressed(os.path.join(HERE, "chain_kat.npz"), **out)


def rx():
    rate, n = 240e3, 72000
    x, truth = synth.make_wideband(3, rate, n, nbursts=2, snr_db=22.0)
    taps = O.firdes_low_pass(1.0, rate, 11e3, 1e3)
    D = int(rate / 48000)
    tmpl = preamble_template("north_star", 5)
    out = {"iq": x, "rate": np.float64(rate), "taps": taps}
    sentences = []
    for k, (f, des) in enumerate(((-25e3, "A"), (25e3, "B"))):
        xl = O.FreqXlatingFir(D, taps, f, rate)
        buf = np.concatenate([np.zeros(len(taps) - 1, np.complex64), x])
        nout = (len(buf) - (len(taps) - 1)) // D
        y = xl.work(buf[:len(taps) - 1 + nout * D])
        bits, _ = O.DemodStream(tmpl, O.chain_cfg()).work(y)
        frames = O.HdlcDeframer(11, 64).work(bits)
        out["chan%d" % k] = y
        out["bits%d" % k] = bits
        out["end_bit%d" % k] = frames["end_bit"]
        for fr in frames:
            sentences.append(O.pdu_to_nmea(bytes(fr["data"][:fr["len"]]), des))
    out["sentences"] = np.array(sentences)
    out["payloads"] = np.frombuffer(b"".join(t["payload"] for t in truth), np.uint8)
    np.savez_compressed(os.path.join(HERE, "rx_kat.npz"), **out)


if __name__ == "__main__":
    O.build()
    chain()
    rx()
    for f in ("chain_kat.npz", "rx_kat.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
