"""The C++ GNU Radio block adapters (gr-ais_b200/gr_adapter) driven by the stub scheduler:
corr_est_cc -> msk_timing_recovery_cc, freqest, invert.  Outputs must equal the CPU oracle's
for the same sequence of work()/general_work() calls."""
import os
import struct
import subprocess

import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import synth

pytestmark = pytest.mark.gpu

ADAPTER = os.path.join(os.path.dirname(B.LIB_PATH), "gr_adapter")


def test_cpp_adapter_chain_matches_oracle(oracle, templates, tmp_path):
    exe = os.path.join(ADAPTER, "qa_adapter")
    assert os.path.exists(exe), "build gr_adapter first (__graft_entry__.build())"
    t = templates[120]
    L, n, ncalls = 120, 137 * 12, 8
    x, _ = synth.make_record(1, n=ncalls * n + L, nbursts=2, snr_db=25)
    x = (x * 1.8).astype(np.complex64)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as fh:
        fh.write(struct.pack("<iii", L, n, ncalls))
        fh.write(t.tobytes())
        fh.write(x.tobytes())
    r = subprocess.run([exe, str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(fout, "rb").read()
    ntags, nsym = struct.unpack_from("<ii", buf, 0)
    pos = 8

    ce = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
    mk = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    msk_in = np.zeros(0, np.complex64)
    msk_read, seen_tags = 0, 0
    pending = np.zeros(0, dtype=oracle.TAG_DTYPE)
    sym, err, mu = [], [], []
    tag_dt = np.dtype([("offset", "<u8"), ("key", "<i4"), ("port", "<i4"), ("value", "<f8")])
    for call in range(ncalls):
        inbuf = x[call * n:call * n + n + L]
        r0, rc, _, rtags = ce.work(n, inbuf, nitems_written=call * n, two_ports=True)
        out0 = np.frombuffer(buf, np.complex64, n, pos); pos += 8 * n
        out1 = np.frombuffer(buf, np.complex64, n, pos); pos += 8 * n
        assert np.array_equal(out0, r0) and np.array_equal(out1, rc)
        # the adapter adds port-0 tags first, then port-1 (add_item_tag order per port)
        got = np.frombuffer(buf, tag_dt, len(rtags), pos); pos += tag_dt.itemsize * len(rtags)
        want = np.concatenate([rtags[rtags["port"] == 0], rtags[rtags["port"] == 1]])
        for f in ("offset", "key", "port", "value"):
            assert np.array_equal(got[f], want[f]), f
        seen_tags += len(rtags)
        pending = np.concatenate([pending, rtags[rtags["port"] == 0]])
        msk_in = np.concatenate([msk_in, r0])
        o, e, m, c = mk.general_work(len(msk_in), msk_in, pending, nitems_read=msk_read)
        sym.append(o); err.append(e); mu.append(m)
        msk_in = msk_in[c:]
        msk_read += c
    assert seen_tags == ntags and ntags > 0
    sym, err, mu = np.concatenate(sym), np.concatenate(err), np.concatenate(mu)
    assert len(sym) == nsym
    assert np.array_equal(np.frombuffer(buf, np.complex64, nsym, pos), sym); pos += 8 * nsym
    assert np.array_equal(np.frombuffer(buf, np.float32, nsym, pos), err); pos += 4 * nsym
    assert np.array_equal(np.frombuffer(buf, np.float32, nsym, pos), mu); pos += 4 * nsym
    (nvec,) = struct.unpack_from("<i", buf, pos); pos += 4
    hz = np.frombuffer(buf, np.float32, nvec, pos); pos += 4 * nvec
    want_hz, _ = oracle.freqest_work(x[:nvec * 1024].reshape(nvec, 1024))
    assert np.array_equal(hz, want_hz)
    inv = np.frombuffer(buf, np.uint8, 1000, pos); pos += 1000
    src = ((np.arange(1000) * 7 + 3) & 0xFF).astype(np.uint8)
    assert np.array_equal(inv, oracle.invert(src))
    (threw,) = struct.unpack_from("<i", buf, pos); pos += 4
    assert threw == 3      # both constructor range errors surfaced as std::out_of_range
    # gr::ais::pdu_to_nmea through its "to_nmea" message port
    for n in (21, 53):
        (sl,) = struct.unpack_from("<i", buf, pos); pos += 4
        text = bytes(buf[pos:pos + sl]).decode("latin-1"); pos += sl
        pdu = bytes((7 * i + 1) & 0xFF for i in range(n))
        assert text == oracle.pdu_to_nmea(pdu, "B")
