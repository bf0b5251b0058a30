"""CUDA path against the reference's OWN code (oracle/_ref/libais_ref.so = /root/reference/lib/
*_impl.cc compiled unmodified, see tests/test_ref_pin.py): the four stand-alone blocks through
the C-ABI host entry points, and the fused chain / stream against the same schedule run on the
reference's classes.  Bit-exact on every output, quirk cases included.  The library travels to
the GPU box prebuilt; without it these tests fail (they must not silently skip on the box)."""
import os

import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import blocks, synth
from gr_ais_b200.ais_demod import ais_demod, preamble_template
from oracle import ref as R

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TAG = {"corr_start": 0, "phase_est": 1, "time_est": 2, "corr_est": 3}


def same_tags(got, want):
    assert len(got) == len(want)
    for f in ("offset", "key", "port", "value"):
        assert np.array_equal(got[f], want[f], equal_nan=(f == "value")), f


def stream(n, seed, burst_at=(), tmpl=None, amp=1.0):
    rng = np.random.default_rng(seed)
    x = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    for p in burst_at:
        x[p:p + len(tmpl)] += (amp * tmpl).astype(np.complex64)
    return x


def test_the_reference_library_is_present():
    assert R.available(), "oracle/_ref/libais_ref.so must travel to the GPU box (built by __graft_entry__.build())"


@pytest.mark.parametrize("L", [120, 140, 1120])
def test_corr_est_block_equals_reference_class(templates, L):
    t = templates[L]
    blk = blocks.corr_est_cc.make(t, 5.0, 1, 0.9)
    ref = R.CorrEstBlock(t, 5.0, 1, 0.9)
    h = ref.hints()
    assert blk.history() == h["history"] and blk.output_multiple() == h["output_multiple"]
    assert blk.max_noutput_items() == h["max_noutput_items"]
    assert np.array_equal(blk.symbols(), ref.symbols())
    n = ref.nsamples * 3
    x = stream(2 * n + L, 11, burst_at=[L + 50, n + 300, 2 * n - 40], tmpl=t)
    written = 0
    for call in range(2):
        inbuf = x[written:written + n + L]
        out0, out1 = np.zeros((1, n), np.complex64), np.zeros((1, n), np.complex64)
        assert blk.work(n, [inbuf], [out0, out1]) == n
        r0, rc, _, rtags = ref.work(n, inbuf, nitems_written=written, two_ports=True)
        assert np.array_equal(out0[0], r0) and np.array_equal(out1[0], rc)
        same_tags(blk.tags[0], rtags)
        assert len(rtags) > 0
        written += n


def test_corr_est_chunk_edges_and_set_symbols_quirk(templates):
    t = templates[120]
    ns = 137
    x = stream(4 * ns + 120, 5, burst_at=[1, ns, 2 * ns + 17, 4 * ns - 119], tmpl=t, amp=1.1)
    blk, ref = blocks.corr_est_cc.make(t, 5.0, 7, 0.5), R.CorrEstBlock(t, 5.0, 7, 0.5)
    for start, n in ((0, ns), (ns, ns), (2 * ns, 2 * ns)):
        inbuf = x[start:start + n + 120]
        out0, out1 = np.zeros((1, n), np.complex64), np.zeros((1, n), np.complex64)
        blk.work(n, [inbuf], [out0, out1])
        _, rc, _, rtags = ref.work(n, inbuf, nitems_written=start, two_ports=True)
        assert np.array_equal(out1[0], rc)
        same_tags(blk.tags[0], rtags)
    # set_symbols: verbatim taps, threshold and mark_delay kept (corr_est_cc_impl.cc:132-162)
    blk, ref = blocks.corr_est_cc.make(templates[140], 5.0, 130, 0.9), R.CorrEstBlock(templates[140], 5.0, 130, 0.9)
    new = np.conj(templates[120])[::-1].copy()
    blk.set_symbols(new)
    ref.set_symbols(new)
    assert np.array_equal(blk.symbols(), ref.symbols()) and blk.history() == ref.hints()["history"]
    n = 137 * 3
    y = stream(n + 120, 6, burst_at=[200], tmpl=templates[120], amp=1.3)
    out0, out1 = np.zeros((1, n), np.complex64), np.zeros((1, n), np.complex64)
    blk.work(n, [y], [out0, out1])
    _, rc, _, rtags = ref.work(n, y, two_ports=True)
    assert np.array_equal(out1[0], rc)
    same_tags(blk.tags[0], rtags)
    assert len(rtags) > 0


def _tags(items):
    t = np.zeros(len(items), dtype=B.TAG_DTYPE)
    for k, (off, key, val) in enumerate(items):
        t[k] = (off, key, 0, val)
    return t


@pytest.mark.parametrize("osps", [1, 2])
def test_msk_block_equals_reference_class_on_the_quirk_cases(osps):
    rng = np.random.default_rng(4)
    x = synth.gmsk_modulate(rng.integers(0, 2, 800)).astype(np.complex64)
    x += (0.05 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
    tags = _tags([(333, TAG["time_est"], 0.31), (334, TAG["phase_est"], 0.5),
                  (900, TAG["time_est"], -0.42), (1500, TAG["time_est"], np.nan),
                  (1501, TAG["time_est"], 0.07), (2500, TAG["corr_start"], 9.0),
                  (2600, TAG["time_est"], 0.999), (2601, TAG["time_est"], 0.5),
                  (2602, TAG["time_est"], -0.999), (3100, TAG["time_est"], -1e-9)])
    blk = blocks.msk_timing_recovery_cc.make(5.0, 0.04, 0.01, osps)
    ref = R.MskBlock(5.0, 0.04, 0.01, osps)
    assert blk.get_sps() == ref.get_sps() and blk.forecast(100) == ref.forecast(100)
    pos = 0
    for avail, nout in ((800, 1000), (1700, 60), (1702, 1000), (2400, 1000), (len(x), 4000)):
        chunk = x[pos:avail]
        out = np.zeros((1, nout), np.complex64)
        err = np.zeros((1, nout), np.float32)
        mu = np.zeros((1, nout), np.float32)
        k = blk.general_work(nout, [len(chunk)], [chunk], [out, err, mu], tags=[tags])
        r_out, r_err, r_mu, r_cons = ref.general_work(nout, chunk, tags, nitems_read=pos)
        assert k == len(r_out) and blk.consumed[0] == r_cons
        assert np.array_equal(out[0, :k], r_out)
        assert np.array_equal(err[0, :k], r_err)
        assert np.array_equal(mu[0, :k], r_mu)
        pos += r_cons


def test_freqest_and_invert_equal_reference_classes():
    rng = np.random.default_rng(9)
    nvec, n = 6, 1024
    spec = ((rng.standard_normal((nvec, n)) + 1j * rng.standard_normal((nvec, n))) *
            10.0 ** rng.uniform(-6, 6, (nvec, 1))).astype(np.complex64)
    spec[0] = 0
    spec[3] = 0           # maxpos carry-over (freqest_impl.cc:67-68,74)
    blk = blocks.freqest.make(48000.0, 9600, n)
    out = np.zeros((1, nvec), np.float32)
    blk.work(nvec, [spec.reshape(1, -1)], [out])
    assert np.array_equal(out[0], R.freqest_work(spec))
    b = rng.integers(0, 256, 4099).astype(np.uint8)
    o = np.zeros(len(b), np.uint8)
    blocks.invert.make().work(len(b), [b], [o])
    assert np.array_equal(o, R.invert(b))


def test_pdu_to_nmea_equals_reference_class():
    rng = np.random.default_rng(8)
    fmt = blocks.pdu_to_nmea.make("A")
    for n in (1, 2, 5, 21, 42, 43, 53, 84, 85, 168):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert fmt.to_nmea(data) == R.pdu_to_nmea(data, "A"), n


def test_chain_equals_the_chain_run_on_the_reference_blocks(oracle):
    z = np.load(os.path.join(HERE, "golden", "chain_kat.npz"))
    for name in ("l120", "l140"):
        x, tmpl = z[name + "_iq"], z[name + "_template"]
        want = oracle.demod_chain(x, tmpl, blocks=R.blocks())
        d = ais_demod(channels=2, max_samples=len(x), template=tmpl)
        bits, nbits, tags, ntags = d.work(np.stack([x, x]))
        for c in range(2):
            assert np.array_equal(bits[c, :nbits[c]], want["bits"])
            same_tags(tags[c, :ntags[c]], want["tags"])
        d.close()


@pytest.mark.parametrize("kind", ["north_star", "reference"])
def test_chain_on_fresh_records_and_ragged_stream(oracle, kind):
    tmpl = preamble_template(kind)
    C, n = 4, 24000
    x = np.stack([synth.make_record(70 + c, n=n, nbursts=3, snr_db=(20.0, 10.0, 5.0, 2.0)[c],
                                    random_impairments=True)[0] for c in range(C)])
    d = ais_demod(channels=C, max_samples=n, template=tmpl)
    bits, nbits, tags, ntags = d.work(x)
    for c in range(C):
        want = oracle.demod_chain(x[c], tmpl, blocks=R.blocks())
        assert np.array_equal(bits[c, :nbits[c]], want["bits"])
        same_tags(tags[c, :ntags[c]], want["tags"])
    # the same capture as a stream cut at ragged places, against the stream on the reference blocks
    refs = [oracle.DemodStream(tmpl, blocks=R.blocks()) for _ in range(C)]
    d.stream_reset()
    cuts = [0, 1, 5000, 5000, 11111, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        b, nb, tg, nt = d.stream_work(x[:, lo:hi])
        for c in range(C):
            rb, rt = refs[c].work(x[c, lo:hi])
            assert np.array_equal(b[c, :nb[c]], rb)
            same_tags(tg[c, :nt[c]], rt)
    d.close()
