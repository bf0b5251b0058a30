"""GPU parity of the four stand-alone blocks (the reference's operator surface) against the
CPU oracle, through the C-ABI host entry points.  Bit-exact on every output."""
import numpy as np
import pytest

from gr_ais_b200 import binding as B
from gr_ais_b200 import blocks, synth

pytestmark = pytest.mark.gpu


def _stream(n, seed, burst_at=None, tmpl=None, amp=1.0):
    rng = np.random.default_rng(seed)
    x = (0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    if burst_at is not None:
        for p in np.atleast_1d(burst_at):
            x[p:p + len(tmpl)] += (amp * tmpl).astype(np.complex64)
    return x


def _same_tags(got, want):
    assert len(got) == len(want)
    for f in ("offset", "key", "port", "value"):
        assert np.array_equal(got[f], want[f]), f


@pytest.mark.parametrize("L", [120, 140, 1120])
def test_corr_est_work_matches_oracle(oracle, templates, L):
    t = templates[L]
    blk = blocks.corr_est_cc.make(t, 5.0, 1, 0.9)
    ref = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
    assert blk.history() == L + 1 and blk.output_multiple() == ref.nsamples
    assert blk.threshold() == ref.thresh and blk.mark_delay() == ref.mark_delay
    assert np.array_equal(blk.symbols(), ref.symbols())
    n = ref.nsamples * 3
    stream = _stream(n * 2 + L, 11, burst_at=[L + 50, n + 300], tmpl=t)
    written = 0
    for call in range(2):       # two consecutive work() calls, history carried by the caller
        inbuf = stream[written:written + n + L]
        out0 = np.zeros((1, n), np.complex64)
        out1 = np.zeros((1, n), np.complex64)
        assert blk.work(n, [inbuf], [out0, out1]) == n
        r0, rc, rmag, rtags = ref.work(n, inbuf, nitems_written=written, two_ports=True)
        assert np.array_equal(out0[0], r0)
        assert np.array_equal(out1[0], rc), "correlator stream differs"
        _same_tags(blk.tags[0], rtags)
        assert len(rtags) > 0
        written += n
    assert blk.nitems_written() == 2 * n


def test_corr_est_single_output_and_batched_channels(oracle, templates):
    t = templates[120]
    C, n = 5, 137 * 4
    blk = blocks.corr_est_cc.make(t, 5.0, 1, 0.9, channels=C)
    rows = np.stack([_stream(n + 120, 20 + c, burst_at=150 + 7 * c, tmpl=t) for c in range(C)])
    out0 = np.zeros((C, n), np.complex64)
    blk.work(n, [rows], [out0])
    for c in range(C):
        ref = oracle.CorrEstBlock(t, 5.0, 1, 0.9)
        r0, _, _, rtags = ref.work(n, rows[c])
        assert np.array_equal(out0[c], r0)
        _same_tags(blk.tags[c], rtags)       # no port-1 tags without the 2nd output


def test_corr_est_set_symbols_quirk(oracle, templates):
    """set_symbols stores taps verbatim and keeps the threshold (lib/corr_est_cc_impl.cc:132-162)"""
    blk = blocks.corr_est_cc.make(templates[120], 5.0, 1, 0.9)
    ref = oracle.CorrEstBlock(templates[120], 5.0, 1, 0.9)
    new = np.conj(templates[140])[::-1].copy()
    blk.set_symbols(new)
    ref.set_symbols(new)
    assert np.array_equal(blk.symbols(), new) and blk.threshold() == ref.thresh
    assert blk.history() == 141
    n = ref.nsamples * 2
    inbuf = _stream(n + 140, 5, burst_at=300, tmpl=templates[140], amp=1.2)
    out0, out1 = np.zeros((1, n), np.complex64), np.zeros((1, n), np.complex64)
    blk.work(n, [inbuf], [out0, out1])
    r0, rc, _, rtags = ref.work(n, inbuf, two_ports=True)
    assert np.array_equal(out1[0], rc)
    _same_tags(blk.tags[0], rtags)


def test_corr_est_tag_overflow_is_reported(templates):
    t = templates[120]
    blk = blocks.corr_est_cc.make(t, 5.0, 1, 1e-6)     # everything exceeds the threshold
    inbuf = _stream(137 * 4 + 120, 3, burst_at=200, tmpl=t)
    with pytest.raises(B.B200AisError) as e:
        blk.work(137 * 4, [inbuf], [np.zeros((1, 137 * 4), np.complex64)], max_tags=16)
    assert e.value.code == B.E_TAG_OVERFLOW


def _tags(items):
    t = np.zeros(len(items), dtype=B.TAG_DTYPE)
    for k, (off, key, val) in enumerate(items):
        t[k] = (off, key, 0, val)
    return t


def test_msk_general_work_streaming_matches_oracle(oracle):
    rng = np.random.default_rng(4)
    x = synth.gmsk_modulate(rng.integers(0, 2, 700)).astype(np.complex64)
    x += (0.05 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
    tags = _tags([(333, 2, 0.31), (334, 1, 0.5), (900, 2, -0.42), (1500, 2, np.nan), (1501, 2, 0.07),
                  (2500, 0, 9.0), (2600, 2, 0.999)])
    blk = blocks.msk_timing_recovery_cc.make(5.0, 0.04, 0.01, 1)
    ref = oracle.MskBlock(5.0, 0.04, 0.01, 1)
    assert blk.forecast(100) == ref.forecast(100) and blk.get_sps() == 2.5
    pos = 0
    for avail, nout in ((800, 1000), (1700, 60), (2400, 1000), (len(x), 1000)):
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
    assert blk.nitems_read() == pos


def test_msk_osps2_setters_and_batch(oracle):
    rng = np.random.default_rng(6)
    C = 3
    xs = np.stack([synth.gmsk_modulate(rng.integers(0, 2, 300)).astype(np.complex64) for _ in range(C)])
    blk = blocks.msk_timing_recovery_cc.make(5.0, 0.05, 0.1, 2, channels=C)
    blk.set_gain(0.03)
    blk.set_limit(0.02)
    assert blk.get_gain() == pytest.approx(0.03) and blk.get_limit() == pytest.approx(0.02)
    with pytest.raises(IndexError):
        blk.set_gain(-1.0)
    blk.set_gain(0.03)
    nout = 700
    out = np.zeros((C, nout), np.complex64)
    err = np.zeros((C, nout), np.float32)
    k = blk.general_work(nout, [xs.shape[1]], [xs], [out, err])
    for c in range(C):
        ref = oracle.MskBlock(5.0, 0.03, 0.02, 2)
        r_out, r_err, _, r_cons = ref.general_work(nout, xs[c])
        assert k[c] == len(r_out) and blk.consumed[c] == r_cons
        assert np.array_equal(out[c, :k[c]], r_out) and np.array_equal(err[c, :k[c]], r_err)


def test_freqest_work_matches_oracle(oracle):
    rng = np.random.default_rng(9)
    C, nvec, n = 3, 6, 1024
    spec = (rng.standard_normal((C, nvec, n)) + 1j * rng.standard_normal((C, nvec, n))).astype(np.complex64)
    spec[1, 0] = 0           # nothing seen yet: maxpos stays 0 -> -12 kHz
    spec[1, 3] = 0           # silent vector repeats the previous estimate (maxpos carry-over)
    spec[2, :, 700] += 400   # a dominant line
    blk = blocks.freqest.make(48000.0, 9600, n, channels=C)
    out = np.zeros((C, nvec), np.float32)
    assert blk.work(nvec, [spec.reshape(C, -1)], [out]) == nvec
    for c in range(C):
        hz, _ = oracle.freqest_work(spec[c])
        assert np.array_equal(out[c], hz)
    assert out[1, 0] == -12000.0 and out[1, 3] == out[1, 2]
    # a non power-of-two vector length goes through the same entry point
    blk2 = blocks.freqest.make(50000.0, 9600, 1000)
    s2 = (rng.standard_normal((2, 1000)) + 1j * rng.standard_normal((2, 1000))).astype(np.complex64)
    o2 = np.zeros((1, 2), np.float32)
    blk2.work(2, [s2.reshape(1, -1)], [o2])
    hz2, _ = oracle.freqest_work(s2, 50000.0, 9600, 1000)
    assert np.array_equal(o2[0], hz2)


def test_invert_work(oracle):
    rng = np.random.default_rng(10)
    for n in (1, 15, 16, 17, 4099):
        b = rng.integers(0, 256, n).astype(np.uint8)
        out = np.zeros(n, np.uint8)
        assert blocks.invert.make().work(n, [b], [out]) == n
        assert np.array_equal(out, oracle.invert(b))


def test_square_and_fft_sync_stage(oracle):
    """the freq-sync hier-block alone: mixed stream and per-vector estimates"""
    from gr_ais_b200.ais_demod import square_and_fft_sync_cc
    x = np.stack([synth.make_record(c, n=4096, nbursts=1, snr_db=15, cfo_hz=300.0)[0] for c in range(3)])
    blk = square_and_fft_sync_cc(48000.0, 9600.0, 1024, channels=3, max_samples=4096)
    y, fhat = blk.work(x)
    for c in range(3):
        r = oracle.demod_chain(x[c], np.ones(8, np.complex64), oracle.chain_cfg(stages=oracle.STAGE_FREQSYNC),
                               debug=True)
        assert np.array_equal(fhat[c], r["fhat"]) and np.array_equal(y[c], r["mixed"])


def test_square_and_fft_sync_edge_cases(oracle):
    """the spectrum argmax on degenerate and tied spectra: silence and a vanishing level (every
    bin goes through the canonical evaluation), a real-valued record (mirror-symmetric spectrum:
    tied sums), a pure tone (one line, leakage-free) and a record that starts silent"""
    from gr_ais_b200.ais_demod import square_and_fft_sync_cc
    n = 8192
    rng = np.random.default_rng(77)
    base = synth.make_record(5, n=n, nbursts=2, snr_db=12, cfo_hz=-450.0)[0]
    rows = [np.zeros(n, np.complex64),
            (base * 1e-9).astype(np.complex64),
            (base * 3e-7).astype(np.complex64),
            rng.standard_normal(n).astype(np.float32).astype(np.complex64),
            np.exp(2j * np.pi * 37.0 * np.arange(n) / 1024.0).astype(np.complex64),
            np.concatenate([np.zeros(3072, np.complex64), base[:n - 3072]])]
    x = np.stack(rows)
    blk = square_and_fft_sync_cc(48000.0, 9600.0, 1024, channels=len(rows), max_samples=n)
    y, fhat = blk.work(x)
    for c in range(len(rows)):
        r = oracle.demod_chain(x[c], np.ones(8, np.complex64), oracle.chain_cfg(stages=oracle.STAGE_FREQSYNC),
                               debug=True)
        assert np.array_equal(fhat[c], r["fhat"], equal_nan=True), "channel %d" % c
        assert np.array_equal(y[c].view(np.uint32), r["mixed"].view(np.uint32)) or \
            np.array_equal(y[c], r["mixed"], equal_nan=True), "channel %d" % c


@pytest.mark.parametrize("L", [5, 8, 9, 16, 17, 33, 64, 65, 100, 128, 129, 256, 300, 513, 1024, 1025, 2048])
def test_corr_est_every_fft_size(oracle, L):
    """fft sizes 16 .. 4096 (every template instantiation of the overlap-add filter), random
    taps, three work() calls so the filter tail crosses calls and CTA boundaries"""
    rng = np.random.default_rng(100 + L)
    t = np.exp(2j * np.pi * rng.uniform(0, 1, L)).astype(np.complex64)
    blk = blocks.corr_est_cc.make(t, 5.0, 2, 0.5)
    ref = oracle.CorrEstBlock(t, 5.0, 2, 0.5)
    ns = ref.nsamples
    assert blk.output_multiple() == ns
    calls = [max(1, 600 // ns), max(2, 40000 // ns), 1]
    total = sum(calls) * ns
    stream = (0.3 * (rng.standard_normal(total + L) + 1j * rng.standard_normal(total + L))).astype(np.complex64)
    for p in (L + 40, L + total // 2, L + total - L - 3):
        stream[p:p + L] += t
    written = 0
    for k in calls:
        n = k * ns
        inbuf = stream[written:written + n + L]
        out0 = np.zeros((1, n), np.complex64)
        out1 = np.zeros((1, n), np.complex64)
        blk.work(n, [inbuf], [out0, out1], max_tags=8192)
        r0, rc, _, rtags = ref.work(n, inbuf, nitems_written=written, two_ports=True, max_tags=8192)
        assert np.array_equal(out1[0], rc), "correlator stream differs (L=%d)" % L
        assert np.array_equal(out0[0], r0)
        _same_tags(blk.tags[0], rtags)
        written += n
    with pytest.raises(B.B200AisError):
        blk.work(ns + 1, [stream[:ns + 1 + L]], [np.zeros((1, ns + 1), np.complex64)])


def test_corr_est_rejects_unsupported_tap_counts():
    for L in (1, 4, 2049):
        with pytest.raises(B.B200AisError):
            blocks.corr_est_cc.make(np.ones(L, np.complex64), 5.0, 1, 0.9)


def test_agc_gain_division_is_the_ieee_quotient():
    """the AGC kernel's straight-line division (reciprocal estimate, Newton step, quotient,
    remainder, correction: nvcc's own fast path without its range check) against the `/` operator
    on the device: every float divisor in [1e-4, 1e15] -- the range the kernel uses it in -- for
    the reference level 2 and dividends across its allowed range"""
    import ctypes
    import struct

    def bits(x):
        return struct.unpack("<I", struct.pack("<f", x))[0]
    lo, hi = bits(1e-4) - 1, bits(1e15) + 1
    L = B.lib()
    total = 0
    for a in (2.0, 1.0, 0.75, 3.1415927, 1.9999999, 1e-10, 1e10, 123456.7):
        bad = ctypes.c_ulonglong(1)
        B.check(L.b200ais_selftest_div(ctypes.c_float(a), lo, hi, ctypes.byref(bad)))
        assert bad.value == 0, (a, bad.value)
        total += hi - lo + 1
    assert total > 4_000_000_000


@pytest.mark.parametrize("seed", range(12))
def test_msk_random_rates_tags_and_call_sizes(oracle, seed):
    """the timing loop far from the directed cases: rate, gain, limit, osps, tag list (stale,
    off-key, NaN, negative centres) and ragged call sizes drawn at random; with the error / mu
    outputs connected (the per-step kernel) and without them (the straight-line rounds)"""
    rng = np.random.default_rng(9000 + seed)
    sps = float(rng.choice([5.0, 5.0, 5.208, 5.5, 6.0, 8.0]))   # >= 16/3: DESIGN.md section 4
    gain = float(rng.choice([0.01, 0.04, 0.1, 0.175]))
    limit = float(rng.choice([0.005, 0.01, 0.05, 0.2]))
    osps = int(rng.integers(1, 3))
    x = synth.gmsk_modulate(rng.integers(0, 2, 500)).astype(np.complex64)
    x += (0.05 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)
    if sps != 5.0:
        t = np.arange(0, len(x) - 1, 5.0 / sps)
        i = t.astype(int)
        f = (t - i).astype(np.float32)
        x = ((1 - f) * x[i] + f * x[i + 1]).astype(np.complex64)
    n = len(x)
    offs = np.sort(rng.integers(0, n, int(rng.integers(0, 14))))
    tags = _tags([(int(o), 2 if rng.random() < 0.75 else int(rng.integers(0, 4)),
                   float(rng.uniform(-0.999, 0.999)) if rng.random() < 0.9 else float("nan")) for o in offs])
    calls = [(int(rng.integers(0, 900)), int(rng.integers(1, 400))) for _ in range(40)]
    for taps in (True, False):
        blk = blocks.msk_timing_recovery_cc.make(sps, gain, limit, osps)
        ref = oracle.MskBlock(sps, gain, limit, osps)
        pos = 0
        for step, nout in calls:
            chunk = x[pos:min(n, pos + step)]
            out = np.zeros((1, nout), np.complex64)
            err = np.zeros((1, nout), np.float32)
            mu = np.zeros((1, nout), np.float32)
            k = blk.general_work(nout, [len(chunk)], [chunk], [out, err, mu] if taps else [out], tags=[tags])
            r_out, r_err, r_mu, r_cons = ref.general_work(nout, chunk, tags, nitems_read=pos)
            assert k == len(r_out) and blk.consumed[0] == r_cons, (seed, taps, pos)
            assert np.array_equal(out[0, :k].view(np.uint32), r_out.view(np.uint32)), (seed, taps, pos)
            if taps:
                assert np.array_equal(err[0, :k].view(np.uint32), r_err.view(np.uint32)), (seed, pos)
                assert np.array_equal(mu[0, :k].view(np.uint32), r_mu.view(np.uint32)), (seed, pos)
            pos += r_cons
            if pos >= n:
                break
